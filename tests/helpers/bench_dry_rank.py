"""one rank of a 2-process CPU dry run of bench.py's default invocation (started by torch.distributed.run from
tests/test_bench_main_dry_run.py): the GPU-only pieces are replaced by the stand-ins of that test module, NCCL by gloo."""
import contextlib
import sys
from pathlib import Path

import torch
import torch.distributed as dist

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import bench  # noqa: E402
import test_bench_main_dry_run as T  # noqa: E402
from neurosis_b200 import graph, ops  # noqa: E402

real_device = torch.device
torch.device = lambda *a, **k: real_device("cpu") if a and a[0] == "cuda" else real_device(*a, **k)
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda: None
torch.cuda.memory_reserved = lambda *a, **k: 0
torch.cuda.Event = T._Event
torch.Tensor.pin_memory = lambda self, *a, **k: self
ops.refresh_weight_copies = lambda force=False: None
ops.invalidate_weight_cache = lambda: None
graph.GraphedTrainStep = T._Graphed
bench.build_engine = lambda dev, seed=42, family="sdxl": T._Engine()
child = ("import json, os; print(json.dumps({'metric': 'm', 'value': 3.0, 'unit': 'images/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', '1')),"
         " 'steps': 2, 'warmup': 3, 'ms_per_step': 10.0, 'config': {'batch_per_gpu': 1, 'workload': os.environ.get('NK_GEMM_DUAL', '?')}}))"
         " if os.environ.get('RANK', '0') == '0' else None")
bench._child_cmd = lambda name, args, world: [sys.executable, "-c", child]
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo", **{kk: v for kk, v in k.items() if kk != "device_id"})

mode = sys.argv[1]
if mode == "variants":
    rank = int(bench.os.environ["RANK"])
    # rank 1's probe rejected the LayerNorm form: it must be dropped on both ranks
    verdict = {"variant": "gemm_row_tile_pairing", "ok": True, "enabled": True, "mode": 1, "min_k_iters": 20, "skew": 0, "classes": 7,
               "speedup": 1.1, "layernorm_column_owner": {"ok": True, "enabled": rank == 0, "mask": 5, "speedup": 1.3},
               "groupnorm_reverse_apply": {"enabled": False}, "epilogue_l2_prefetch": {"enabled": False}, "fused_cross_kv": {"enabled": False}}
    from neurosis_b200 import tune
    tune.autotune = lambda device=0, timeout_s=0, min_speedup=1.01: verdict
    bench._step_guard = lambda args, tuned, world, local, dev: (bench._agree_across_ranks(tuned, world, dev, "x"), bench._apply_tuned(tuned), tuned)[2]
    T._Graphed.variants_faster = True
bench.os.environ["NK_B200_TUNE_CACHE"] = "0"
sys.argv = ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "1"]
bench.main()
