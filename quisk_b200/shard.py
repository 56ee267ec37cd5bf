"""quisk_b200/shard.py -- how the hot path is partitioned over the GPUs of one box.

Receivers are independent units: rank r of W owns a contiguous channel range and keeps those
channels' filter state on its GPU for the whole run -- no collective on the data path.  A single
wideband stream is instead cut into time blocks whose starts are multiples of the total decimation
(so every stage's toggle / decim_index at a block start equals the sequential reference's) and
each block is prefixed with a FIR-history halo of H input samples.

Only planning lives here (pure Python, testable on CPU under gloo); the kernels never communicate.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


def channel_range(rank: int, world: int, n_channels: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of the channels rank owns; sizes differ by at most one."""
    if not (0 <= rank < world) or n_channels < 0:
        raise ValueError("bad rank/world/n_channels")
    base, extra = divmod(n_channels, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def cascade_halo(stages: Sequence[Tuple[int, int]]) -> int:
    """Input-rate history a decimating FIR cascade needs before its first exact output:
    sum_k (taps_k - 1) * prod_{j<k} D_j for stages [(taps, decim), ...] (the C1
    decimator 4 x (45, 2) + (98, 2) gives 44*(1+2+4+8) + 97*16 = 2212)."""
    halo, rate = 0, 1
    for taps, decim in stages:
        halo += (taps - 1) * rate
        rate *= decim
    return halo


@dataclass
class TimeBlock:
    rank: int
    start: int          # first input sample whose outputs this rank owns (multiple of total decimation)
    stop: int           # one past the last
    halo_start: int     # first input sample the rank has to read (start - halo rounded up to the decimation, clamped at 0)
    out_start: int      # index of the rank's first output sample in the stream's output
    out_count: int


def time_blocks(n_samples: int, world: int, total_decim: int, halo: int) -> List[TimeBlock]:
    """Split [0, n_samples) into `world` time blocks aligned to `total_decim`.  Output sample m is
    produced on input (m+1)*total_decim - 1 (filter.c:213: an output on every total_decim-th input),
    so the rank owning inputs [start, stop) owns outputs [start/D, stop/D)."""
    if total_decim <= 0 or world <= 0:
        raise ValueError("bad arguments")
    units = n_samples // total_decim            # whole output samples in the stream
    halo = -(-halo // total_decim) * total_decim    # keep every stage's decimation phase: the halo is whole output periods
    blocks = []
    for r in range(world):
        lo, hi = channel_range(r, world, units)
        start, stop = lo * total_decim, hi * total_decim
        if r == world - 1:
            stop = n_samples                    # the ragged tail (no output) stays with the last rank
        blocks.append(TimeBlock(r, start, stop, max(0, start - halo), lo, hi - lo))
    return blocks
