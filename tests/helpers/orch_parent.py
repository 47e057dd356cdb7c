"""stand-in for the parent ranks of a torchrun-launched bench.py: holds its own (idle) process group, like the headline
run does after its measurement, and runs the secondary configurations as child processes."""
import json
import sys
import types
from pathlib import Path

import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

dist.init_process_group("gloo")
bench._child_cmd = lambda name, args, world: [sys.executable, str(Path(__file__).with_name("orch_child.py"))]
out = bench.run_other_configs(types.SimpleNamespace(steps=1, warmup=1), dist.get_world_size(), dist.get_rank(),
                              budget_s=170.0, per_config_s=55.0)
dist.barrier()
if dist.get_rank() == 0:
    print("RESULT " + json.dumps(out), flush=True)
dist.destroy_process_group()
