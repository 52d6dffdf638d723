"""Frame-set sharding across ranks.

Every `process` call is independent given the static tables (the reference even rebuilds its
blender per call, include/ocvstitcher.hpp:1186), so a batch of frame-sets is partitioned by
frame-set index with no collective on the data path.  Tables are replicated per rank at init.
"""
from typing import List, Tuple


def shard_range(num_sets: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the frame-sets rank `rank` owns; remainders go to low ranks."""
    if world < 1 or not (0 <= rank < world) or num_sets < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(num_sets, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_round_robin(num_sets: int, rank: int, world: int) -> List[int]:
    """Frame-set t -> rank t mod world (the streaming order of a live camera rig)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    return list(range(rank, num_sets, world))


def strip_columns(padded_width: int, num_bands: int, world: int) -> List[Tuple[int, int]]:
    """Column strips of the padded destination for the spatial split (config 4): boundaries are
    multiples of 2^num_bands so every pyramid level splits on integer columns."""
    unit = 1 << num_bands
    if padded_width % unit:
        raise ValueError("padded width must be a multiple of 2^num_bands")
    units = padded_width // unit
    out, x = [], 0
    for r in range(world):
        n = units // world + (1 if r < units % world else 0)
        out.append((x * unit, (x + n) * unit))
        x += n
    return out
