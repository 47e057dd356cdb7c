"""stand-in for `bench.py --config X` in the orchestration test: joins the children's own rendezvous (env://, the port
the parent chose), all-reduces over gloo, rank 0 prints a bench-shaped JSON line."""
import json
import os

import torch
import torch.distributed as dist

dist.init_process_group("gloo")
t = torch.tensor([float(dist.get_rank() + 1)])
dist.all_reduce(t)
if dist.get_rank() == 0:
    print(json.dumps({"metric": "m", "value": float(t.item()), "unit": "u", "n_gpus": dist.get_world_size(), "steps": 1,
                      "warmup": 1, "ms_per_step": 1.0, "config": {"workload": os.environ["MASTER_PORT"]}}), flush=True)
dist.barrier()
dist.destroy_process_group()
