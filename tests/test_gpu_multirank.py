"""Multi-rank proof on the real kernels over NCCL (needs >= 2 GPUs; skipped on a single-GPU box): reduced gradients of
`BucketedGradReducer` and parameters after `ShardedOptimizerReducer` + fused Adafactor against a single-process run on the
concatenated batch (tools/ddp_equality.py does the work under torchrun)."""
import json
import socket
import subprocess
import sys

import pytest
import torch

from common import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_gradients_and_sharded_optimizer_match_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tools" / "ddp_equality.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(lines[-1])
    print(out)
    assert out["ok"] and out["reduced_grads_identical_on_all_ranks"] and out["params_identical_on_all_ranks_after_sharded_step"]
    assert out["grad_rel_l2_median"] < 2e-2
