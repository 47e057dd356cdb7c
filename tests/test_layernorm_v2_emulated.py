"""Thread-level emulation (numpy, CPU) of the second form of the LayerNorm kernels, `ln_fwd_v2_kernel` / `ln_bwd_v2_kernel` in
neurosis_b200/csrc/norm.cu.  Every step of the CUDA code is mirrored with arrays indexed by thread id — column-owner
threads, idle lanes past C / 8, RB rows per batch, warp sums, the [warp] exchange through shared memory, the shifted
single-pass variance, the register accumulation of dgamma / dbeta with one add per block at the end, rows_per_block
partition with a ragged tail — and compared with torch's fp32 layer_norm forward / backward.  The kernels themselves have
not run on a GPU yet; this pins their arithmetic and indexing, not the CUDA code generation."""
import numpy as np
import pytest
import torch


def _bf16(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16).float().numpy()


def emulate_fwd(x, gamma, beta, eps, RB=8, target_blocks=148 * 8):
    rows, C = x.shape
    V = C // 8
    W = (V + 31) // 32
    T = W * 32
    want = -(-rows // target_blocks)
    rpb = max(RB, -(-want // RB) * RB)
    blocks = -(-rows // rpb)
    y = np.zeros((rows, C), np.float32)
    mean, rstd = np.zeros(rows, np.float32), np.zeros(rows, np.float32)
    t = np.arange(T)
    active = t < V
    g = np.zeros((T, 8), np.float32)
    b = np.zeros((T, 8), np.float32)
    g[active] = gamma.reshape(V, 8)
    b[active] = beta.reshape(V, 8)
    inv_c = np.float32(1.0 / C)
    for blk in range(blocks):
        r_begin, r_end = blk * rpb, min(rows, blk * rpb + rpb)
        for r0 in range(r_begin, r_end, RB):
            q = np.zeros((RB, T, 8), np.float32)
            sh = np.zeros(RB, np.float32)
            for i in range(RB):
                row = r0 + i
                if row < r_end:
                    q[i][active] = x[row].reshape(V, 8)
                    sh[i] = x[row, 0]
            red = np.zeros((RB, 2, W), np.float32)
            for i in range(RB):
                d = np.where(active[:, None], q[i] - sh[i], 0.0).astype(np.float32)
                a, c = d.sum(1, dtype=np.float32), (d * d).sum(1, dtype=np.float32)
                red[i, 0] = a.reshape(W, 32).sum(1, dtype=np.float32)   # warp_sum, lane 0 writes red[..][warp]
                red[i, 1] = c.reshape(W, 32).sum(1, dtype=np.float32)
            for i in range(RB):
                row = r0 + i
                if row >= r_end:
                    break
                S, SS = red[i, 0].sum(dtype=np.float32), red[i, 1].sum(dtype=np.float32)
                mu = S * inv_c
                var = max(SS * inv_c - mu * mu, np.float32(0))
                m = sh[i] + mu
                r = np.float32(1.0) / np.sqrt(np.float32(var + eps))
                mean[row], rstd[row] = m, r
                o = (q[i] - m) * r * g + b
                y[row] = o[active].reshape(C)
    return _bf16(y), mean, rstd


def emulate_bwd(dy, x, gamma, mean, rstd, dres, RB=4, target_blocks=148 * 4):
    rows, C = x.shape
    V = C // 8
    W = (V + 31) // 32
    T = W * 32
    want = -(-rows // target_blocks)
    rpb = max(RB, -(-want // RB) * RB)
    blocks = -(-rows // rpb)
    t = np.arange(T)
    active = t < V
    g = np.zeros((T, 8), np.float32)
    g[active] = gamma.reshape(V, 8)
    dx = np.zeros((rows, C), np.float32)
    dgamma, dbeta = np.zeros(C, np.float32), np.zeros(C, np.float32)
    inv_c = np.float32(1.0 / C)
    for blk in range(blocks):
        ag, ab = np.zeros((T, 8), np.float32), np.zeros((T, 8), np.float32)
        r_begin, r_end = blk * rpb, min(rows, blk * rpb + rpb)
        for r0 in range(r_begin, r_end, RB):
            qx, qd = np.zeros((RB, T, 8), np.float32), np.zeros((RB, T, 8), np.float32)
            m, r = np.zeros(RB, np.float32), np.zeros(RB, np.float32)
            for i in range(RB):
                row = r0 + i
                if row < r_end:
                    qx[i][active] = x[row].reshape(V, 8)
                    qd[i][active] = dy[row].reshape(V, 8)
                    m[i], r[i] = mean[row], rstd[row]
            red = np.zeros((RB, 2, W), np.float32)
            for i in range(RB):
                xh = (qx[i] - m[i]) * r[i]
                gd = qd[i] * g
                red[i, 0] = gd.sum(1, dtype=np.float32).reshape(W, 32).sum(1, dtype=np.float32)
                red[i, 1] = (gd * xh).sum(1, dtype=np.float32).reshape(W, 32).sum(1, dtype=np.float32)
                ag += qd[i] * xh
                ab += qd[i]
            for i in range(RB):
                row = r0 + i
                if row >= r_end:
                    break
                S1, S2 = red[i, 0].sum(dtype=np.float32) * inv_c, red[i, 1].sum(dtype=np.float32) * inv_c
                xh = (qx[i] - m[i]) * r[i]
                o = _bf16(r[i] * (qd[i] * g - S1 - xh * S2))
                o = o[active].reshape(C)
                if dres is not None:
                    o = _bf16(o + dres[row])
                dx[row] = o
        dgamma += ag[active].reshape(C)   # one atomicAdd per owned column and block
        dbeta += ab[active].reshape(C)
    return dx, dgamma, dbeta


@pytest.mark.parametrize("rows,C,with_res", [(37, 1280, True), (64, 640, False), (9, 320, True), (5, 2048, False), (3, 64, True),
                                             (1300, 768, False)])
def test_emulated_kernels_match_torch_fp32(rows, C, with_res):
    g = torch.Generator().manual_seed(rows * 7 + C)
    x = (torch.randn(rows, C, generator=g) * 1.7 + 0.3).to(torch.bfloat16).float()
    dy = (torch.randn(rows, C, generator=g) * 0.05).to(torch.bfloat16).float()
    gamma = 1.0 + 0.2 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    dres = (torch.randn(rows, C, generator=g) * 0.05).to(torch.bfloat16).float() if with_res else None
    # small grids so that several blocks / ragged tails are exercised at these sizes
    y, mean, rstd = emulate_fwd(x.numpy(), gamma.numpy(), beta.numpy(), np.float32(1e-5), target_blocks=5)
    dx, dg, db = emulate_bwd(dy.numpy(), x.numpy(), gamma.numpy(), mean, rstd, dres.numpy() if with_res else None, target_blocks=3)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5)
    yr.backward(dy)
    dxr = xr.grad + (dres if with_res else 0)

    def rel(a, b):
        a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
        return float((a - b).norm() / b.norm().clamp_min(1e-20))

    assert rel(mean, x.mean(1)) < 1e-6 and rel(rstd, 1.0 / torch.sqrt(x.var(1, unbiased=False) + 1e-5)) < 1e-5
    assert rel(y, yr) < 3e-3           # bf16 rounding of the output
    assert rel(dx, dxr) < 4e-3         # bf16 rounding (twice with the residual gradient)
    assert rel(dg, gr.grad) < 1e-5 and rel(db, br.grad) < 1e-5
