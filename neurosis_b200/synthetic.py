"""Synthetic batches of the benchmark configurations: deterministic, seeded 42 + rank, shaped like
the batches the reference's aspect-bucket loader produces (`dataset/imagefolder/aspect.py:74-98` through
`collate_dict_stack`, `dataset/utils.py:166-191`): stacked image tensor, captions as a list, the SDXL size / crop
conditioning as per-sample tuples.  Host-side only (numpy / torch CPU); used by tools/bucket_bench.py and the tests.
"""
from __future__ import annotations

import numpy as np
import torch

# (W, H) buckets of §8(d) config 4, all present in the reference's SDXLBucketList (dataset/aspect/lists.py:31,36,42)
SDXL_BUCKETS = ((896, 1152), (1216, 832), (1024, 1024))
VOCAB = 50_000


class AspectBucketBatches:
    """one single-bucket batch per call (the reference's batches are single-bucket: aspect.py:160-191)."""

    def __init__(self, batch: int, rank: int = 0, ctx_dim: int = 2048, pooled_dim: int = 1280, buckets=SDXL_BUCKETS,
                 zipf_a: float = 1.1, min_tags: int = 8, max_tags: int = 40, images: bool = True):
        self.batch, self.buckets, self.images = batch, tuple(buckets), images
        self.ctx_dim, self.pooled_dim = ctx_dim, pooled_dim
        self.rs = np.random.RandomState(42 + rank)
        self.gen = torch.Generator().manual_seed(42 + rank)
        # Zipf(a) over a 50k-tag vocabulary: p(k) ~ k^-a, k = 1..VOCAB
        p = np.arange(1, VOCAB + 1, dtype=np.float64) ** -zipf_a
        self.cdf = np.cumsum(p / p.sum())
        self.min_tags, self.max_tags = min_tags, max_tags

    def captions(self) -> list[str]:
        out = []
        for _ in range(self.batch):
            n = int(self.rs.randint(self.min_tags, self.max_tags + 1))
            ids = np.searchsorted(self.cdf, self.rs.random_sample(n))
            out.append(" ".join(f"tag{int(i)}" for i in ids))
        return out

    def draw_bucket(self) -> int:
        return int(self.rs.randint(0, len(self.buckets)))

    def __call__(self, bucket: int | None = None) -> dict:
        b = self.draw_bucket() if bucket is None else bucket
        w, h = self.buckets[b]
        B = self.batch
        batch = {"bucket": b,
                 "caption": self.captions(),
                 "crossattn_emb": torch.randn(B, 77, self.ctx_dim, generator=self.gen),
                 "pooled_emb": torch.randn(B, self.pooled_dim, generator=self.gen),
                 "original_size_as_tuple": [(w, h)] * B,       # (w, h) as in aspect.py:74-85
                 "crop_coords_top_left": [(0, 0)] * B,         # (top, left)
                 "target_size_as_tuple": [(w, h)] * B}
        if self.images:
            batch["image"] = torch.rand(B, 3, h, w, generator=self.gen) * 2 - 1
        return batch
