"""Index arithmetic of the row-tile-pairing (DUAL) schedule of gemm_tc_kernel, mirrored in Python (CPU).

`decode_tile` (with the grouped raster), the producer's row-tile formula `mt = (2*tc.mt)*CG + rank (+ CG for the second)`,
the epilogue's `local -> (scheduler tile, accumulator)` walk and the host's scheduler-tile counts are restated line by
line from neurosis_b200/csrc/gemm_tc.cu and checked for the property the kernel relies on: over all CTA pairs, ranks
and scheduler tiles, every (128-row tile, N tile, split) of the problem is produced EXACTLY once, by the CTA whose
producer loaded it into the accumulator its epilogue reads, and the tiles past the problem (odd counts) are masked.
The kernels themselves have not run on a GPU yet; this pins the scheduling, not the PTX."""
import itertools
import math

import pytest

CG = 2


def decode_tile(t, tiles_m, tiles_n, nb2, nb1, raster_gm):
    """gemm_tc.cu decode_tile(): -> (mt, nt, b2, b1, split)"""
    if raster_gm > 0:
        per_batch = tiles_m * tiles_n
        tb = t % per_batch
        t //= per_batch
        per_group = raster_gm * tiles_n
        grp = tb // per_group
        r = tb - grp * per_group
        m0 = grp * raster_gm
        gsize = min(raster_gm, tiles_m - m0)
        nt = r // gsize
        mt = m0 + (r - nt * gsize)
    else:
        mt = t % tiles_m
        t //= tiles_m
        nt = t % tiles_n
        t //= tiles_n
    b2 = t % nb2
    t //= nb2
    b1 = t % nb1
    return mt, nt, b2, b1, t // nb1


def host_plan(M, N, BN, splits, nsm=148, dual=True, raster=2):
    """launch_gemm(): tile counts handed to the kernel (CTA pairs)."""
    tiles_m128 = -(-M // 128)
    tiles_m = -(-tiles_m128 // CG)            # pair tiles
    tiles_n = -(-N // BN)
    if dual:
        tiles_m = -(-tiles_m // 2)            # scheduler tiles
    conc = nsm // CG
    side = 1
    while side * side < conc:
        side += 1
    ncols = max(1, min(tiles_n, side))
    gm = -(-conc // ncols)
    gm = 0 if raster == 0 else (1 if raster == 1 else gm)
    raster_gm = min(gm, tiles_m)
    total = tiles_m * tiles_n * splits
    grid_groups = min(total, nsm // 2)
    return dict(tiles_m128=tiles_m128, tiles_m=tiles_m, tiles_n=tiles_n, raster_gm=raster_gm, total=total, groups=grid_groups)


@pytest.mark.parametrize("M,N,BN,splits", [(16384, 1280, 256, 1), (1232, 1280, 256, 1), (640, 96, 32, 1), (4224, 1000, 256, 1),
                                           (1280, 5120, 256, 3), (300, 320, 160, 1), (128 * 9, 256, 256, 2), (65536, 640, 224, 1),
                                           (128 * 5, 64, 64, 1)])
@pytest.mark.parametrize("dual,raster", [(True, 2), (True, 0), (True, 1), (False, 2)])
def test_every_output_tile_is_produced_exactly_once(M, N, BN, splits, dual, raster):
    p = host_plan(M, N, BN, splits, dual=dual, raster=raster)
    produced = {}
    for group, rank in itertools.product(range(p["groups"]), range(CG)):
        # --- producer / MMA view: scheduler tiles of this CTA pair, row tiles loaded into accumulator s
        loaded = []
        for t in range(group, p["total"], p["groups"]):
            mt, nt, b2, b1, split = decode_tile(t, p["tiles_m"], p["tiles_n"], 1, 1, p["raster_gm"])
            first = (2 * mt if dual else mt) * CG + rank
            for s in range(2 if dual else 1):
                loaded.append((first + s * CG, nt, split, s))
        # --- epilogue view: local counts 128-row tiles; DUAL: scheduler tile local >> 1, accumulator local & 1
        local = 0
        seen = []
        while True:
            t = group + ((local >> 1) if dual else local) * p["groups"]
            if t >= p["total"]:
                break
            mt, nt, b2, b1, split = decode_tile(t, p["tiles_m"], p["tiles_n"], 1, 1, p["raster_gm"])
            if dual:
                mt = 2 * mt + (local & 1)
            mt128 = mt * CG + rank
            acc = local & 1
            seen.append((mt128, nt, split, acc if dual else None))
            local += 1
        if dual:
            # the epilogue reads, in order, exactly what the producer / MMA put into accumulator s of each scheduler tile
            assert [(a, b, c, d) for a, b, c, d in loaded] == seen
        for mt128, nt, split, _ in seen:
            if mt128 < p["tiles_m128"]:  # tile_ok / row_ok: tiles past the problem store nothing
                key = (mt128, nt, split)
                produced[key] = produced.get(key, 0) + 1
    want = {(m, n, s) for m in range(p["tiles_m128"]) for n in range(p["tiles_n"]) for s in range(splits)}
    assert set(produced) == want
    assert all(v == 1 for v in produced.values())


def test_accumulator_parities_match_between_mma_and_epilogue():
    """DUAL barrier phases: the MMA thread waits tmem_empty[s] with parity (n & 1) ^ 1 and commits tmem_full[s] for its n-th
    scheduler tile; the epilogue waits tmem_full[acc] with parity (local >> 1) & 1 for local = 2n + s: same n, same buffer."""
    for n in range(7):
        for s in range(2):
            local = 2 * n + s
            assert (local & 1) == s and ((local >> 1) & 1) == (n & 1)


@pytest.mark.parametrize("stages,iters,skew", [(3, 20, 0), (3, 20, 2), (4, 20, 3), (3, 1, 3), (4, 2, 3), (3, 5, 7), (2, 9, 1)])
def test_skewed_issue_order_never_waits_for_a_slot_the_trailing_tile_has_not_freed(stages, iters, skew):
    """the MMA issuer of the DUAL kernels, replayed: row tile 0 at k-iteration j needs slot j % stages loaded, which the
    producer can only do after row tile 1 has consumed k-iteration j - stages.  With sk = min(skew, stages - 1, iters) the
    single issuing thread never blocks on a load that depends on work it has not issued yet (no self-deadlock), every
    k-iteration is consumed exactly once per row tile, and slots are released in order."""
    sk = min(skew, stages - 1, iters)
    consumed0, consumed1, released = [], [], []
    for j in range(iters + sk):
        if j < iters:
            need_free = j - stages            # k-iteration that last occupied this slot
            assert need_free < 0 or need_free in released, (j, released)  # otherwise full_bar[slot] could never complete
            consumed0.append(j)
        if j >= sk:
            j1 = j - sk
            assert j1 in consumed0            # its slot was waited for by row tile 0 already
            consumed1.append(j1)
            released.append(j1)
    assert consumed0 == list(range(iters)) and consumed1 == list(range(iters)) and released == list(range(iters))
