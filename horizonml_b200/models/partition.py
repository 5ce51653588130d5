"""Pipeline-stage partitioner.

Same rule as the reference's ``SplitResNet`` (layer_model_parallel_train.py:54-69): the five
atomic blocks are cut into ``num_stages`` contiguous chunks, the first ``5 % num_stages`` stages
get one extra block, more than five stages is an error.  Unlike the reference, one stage is legal
(the reference's world_size=1 path sends to a non-existent rank 1, layer_…:191).
"""
from __future__ import annotations

from typing import List, Tuple

NUM_ATOMIC_BLOCKS = 5
BLOCK_NAMES = ["stem", "layer1", "layer2", "layer3", "layer4+fc"]


def partition_blocks(num_stages: int, num_blocks: int = NUM_ATOMIC_BLOCKS) -> List[Tuple[int, int]]:
    """[(first_block, last_block_inclusive)] per stage."""
    if num_stages < 1:
        raise ValueError("num_stages must be >= 1")
    if num_stages > num_blocks:
        raise ValueError(f"Number of workers ({num_stages}) cannot exceed number of layers ({num_blocks})")
    base, rem = divmod(num_blocks, num_stages)
    out, start = [], 0
    for s in range(num_stages):
        n = base + (1 if s < rem else 0)
        out.append((start, start + n - 1))
        start += n
    return out


def boundary_shape(block_idx: int, batch: int, image_hw: int = 32) -> Tuple[int, int, int, int]:
    """Logical NCHW shape of the activation leaving atomic block ``block_idx``."""
    hw = image_hw // 4            # stem: /2 conv, /2 pool
    chans = [64, 64, 128, 256, 512]
    for i in range(1, block_idx + 1):
        if i >= 2:
            hw = max(hw // 2, 1)
    return (batch, chans[block_idx], hw, hw)
