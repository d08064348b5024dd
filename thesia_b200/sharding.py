"""Work split of the analysis path across the GPUs of one box.

Units are (id, ch) channels -- the reference already treats them independently (mod.rs:152-163).
Channels are dealt greedily by sample count; a file longer than a rank's fair share is split by
FRAME RANGE: each part gets the PCM slice its frames touch (a (win - hop)-sample halo plus the
reflected samples at true file ends), so no halo exchange happens on the device.  The only
collective of the path is the 2-float max all-reduce inside thb_minmax_global.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class Unit:
    id: int
    ch: int
    sr: int
    full_len: int        # samples in the file
    frame_begin: int     # first frame of this unit
    frame_count: int     # frames in this unit (whole file: all of them)
    pcm_lo: int          # slice [pcm_lo, pcm_hi) of the file this unit needs
    pcm_hi: int

    @property
    def cost(self) -> int:
        return self.frame_count


def n_frames(length: int, win: int, hop: int) -> int:
    padded = length + 2 * (win // 2)
    return (padded - win) // hop + 1 if padded >= win and hop > 0 else 0


def needed_samples(frame_begin: int, frame_count: int, win: int, hop: int, full_len: int) -> Tuple[int, int]:
    """[lo, hi) of the file that frames [frame_begin, +frame_count) read, reflection included
    (same rule thb_spec_batch validates against)."""
    if frame_count <= 0:
        return 0, 0
    n = full_len
    lo = frame_begin * hop - win // 2
    hi = (frame_begin + frame_count - 1) * hop - win // 2 + win - 1
    need_lo, need_hi = max(lo, 0), min(hi, n - 1)
    if lo < 0:
        need_hi = max(need_hi, min(n - 1, -lo))
    if hi >= n:
        need_lo = min(need_lo, max(0, 2 * (n - 1) - hi))
    if -lo >= n or hi >= 2 * n - 1:
        need_lo, need_hi = 0, n - 1
    return need_lo, need_hi + 1


def split_frames(id: int, ch: int, sr: int, full_len: int, win: int, hop: int, parts: int) -> List[Unit]:
    total = n_frames(full_len, win, hop)
    parts = max(1, min(parts, total)) if total else 1
    out = []
    base, rem = divmod(total, parts)
    begin = 0
    for p in range(parts):
        cnt = base + (1 if p < rem else 0)
        lo, hi = needed_samples(begin, cnt, win, hop, full_len)
        out.append(Unit(id, ch, sr, full_len, begin, cnt, lo, hi))
        begin += cnt
    return out


def plan(channels: Sequence[Tuple[int, int, int, int]], framing, world_size: int) -> List[List[Unit]]:
    """channels: (id, ch, sr, n_samples); framing(sr) -> (hop, win, n_fft).
    Returns, per rank, the units it computes.  Deterministic, identical on every rank."""
    units: List[Unit] = []
    total_frames = 0
    per = []
    for (i, ch, sr, n) in channels:
        hop, win, _ = framing(sr)
        t = n_frames(n, win, hop)
        per.append((i, ch, sr, n, hop, win, t))
        total_frames += t
    fair = max(1, -(-total_frames // world_size))
    for (i, ch, sr, n, hop, win, t) in per:
        # split only files that exceed one rank's fair share (C2: one 1-hour file on N GPUs)
        parts = -(-t // fair) if t > fair else 1
        units.extend(split_frames(i, ch, sr, n, win, hop, parts))
    # longest-processing-time-first onto the least loaded rank
    order = sorted(range(len(units)), key=lambda k: (-units[k].cost, units[k].id, units[k].ch, units[k].frame_begin))
    load = [0] * world_size
    ranks: List[List[Unit]] = [[] for _ in range(world_size)]
    for k in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        ranks[r].append(units[k])
        load[r] += units[k].cost
    for r in ranks:
        r.sort(key=lambda u: (u.id, u.ch, u.frame_begin))
    return ranks
