"""SURVEY.md §8(d) config 4 on the device: SDXL training step over mixed aspect buckets (896x1152 / 1216x832 / 1024^2)
with tag-frequency loss scaling.  One CUDA graph per bucket (shapes are static per bucket), the bucket of every step
drawn per rank, captions from a 50k-tag Zipf(1.1) vocabulary feeding `TagFrequencyHook` (host) whose per-sample weights
enter the graph as a (B,) device buffer.  Prints one JSON line (images/s, per-bucket ms).

    python tools/bucket_bench.py [--batch 4] [--steps 12] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bucket_bench.py

NOT part of the default bench: first written after the round-1 GPU budget was spent (unmeasured so far)."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import bench
    from neurosis_b200 import ops
    from neurosis_b200.ddp import BucketedGradReducer
    from neurosis_b200.graph import GraphedTrainStep
    from neurosis_b200.modules.conditioner import ConcatTimestepEmbedderND
    from neurosis_b200.modules.loss import TagFreqScale, TagFrequencyHook
    from neurosis_b200.synthetic import SDXL_BUCKETS, AspectBucketBatches

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    eng = bench.build_engine(dev)
    params = [p for p in eng.model.parameters() if p.requires_grad]
    reducer = BucketedGradReducer(params, bucket_mb=256.0)
    reducer.attach_as_grad_sink()
    data = AspectBucketBatches(B, rank=rank)
    hook = TagFrequencyHook(alpha=0.2, beta=0.99, strength=1.0,
                            freq_scale=TagFreqScale([[-1, 1.1], [100, 1.0], [1000, 0.95], [40000, 0.8]]))
    fourier = ConcatTimestepEmbedderND(256)

    def vector(batch: dict) -> "torch.Tensor":
        parts = [batch["pooled_emb"].to(dev, non_blocking=True)]
        for k in ("original_size_as_tuple", "crop_coords_top_left", "target_size_as_tuple"):
            parts.append(fourier(torch.tensor(batch[k], dtype=torch.float32).to(dev, non_blocking=True)).float())
        return torch.cat(parts, 1)  # (B, 2816)

    graphs = {}
    for b in range(len(SDXL_BUCKETS)):  # one captured step per bucket shape
        first = data(bucket=b)
        graphs[b] = GraphedTrainStep(eng, reducer, first["image"].to(dev), first["crossattn_emb"].to(dev),
                                     vector(first), warmup=1)
        torch.cuda.synchronize()

    def step() -> int:
        batch = data()
        w = torch.tensor(hook.sample_weights(batch["caption"]), dtype=torch.float32)
        graphs[batch["bucket"]].step(batch["image"].pin_memory(), batch["crossattn_emb"].pin_memory(), vector(batch),
                                     weights=w)
        return batch["bucket"]

    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    drawn = [step() for _ in range(args.steps)]
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"bench": "sdxl_aspect_buckets_tag_frequency", "n_gpus": world, "batch_per_gpu": B,
                          "steps": args.steps, "ms_per_step": float(ms) / args.steps,
                          "images_per_s": world * B * args.steps / (float(ms) * 1e-3),
                          "buckets_drawn_rank0": drawn, "buckets_wh": SDXL_BUCKETS,
                          "host_to_device": "images / conditioning generated on the host every step (pinned copy)",
                          "kernel_launches_per_replay": {b: g.launches_per_replay for b, g in graphs.items()},
                          "launch_counter": ops.LAUNCHES}), flush=True)
    if world > 1:
        graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
