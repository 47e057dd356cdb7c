"""Drop-in modules of the diffusion training step (same names / ctor signatures / state-dict keys as
neurosis.modules.*), executing on libnk_b200.so."""
from .attention import (BasicTransformerBlock, CrossAttention, FeedForward, GEGLU, MemoryEfficientCrossAttention,
                        SpatialTransformer, TorchSDPCrossAttention)
from .openaimodel import (Downsample, ResBlock, Timestep, TimestepBlock, TimestepEmbedSequential, UNetModel,
                          Upsample)
