"""Work split of the analysis path across the GPUs of one box.

Units are (id, ch) channels -- the reference already treats them independently (mod.rs:152-163).
The frames of all channels form one line, cut into one contiguous span per rank; a file that
straddles a cut (C2: one 1-hour file on N GPUs) is split there by FRAME RANGE: each part gets the
PCM slice its frames touch (a (win - hop)-sample halo plus the reflected samples at true file
ends), so no halo exchange happens on the device, and a rank never holds two parts of one file.  The only
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


def frame_weight(n_fft: int) -> int:
    """Relative cost of one frame (the kernels are FFT-bound: ~ n_fft log2 n_fft)."""
    return max(1, n_fft * max(1, n_fft.bit_length() - 1))


def plan(channels: Sequence[Tuple[int, int, int, int]], framing, world_size: int) -> List[List[Unit]]:
    """channels: (id, ch, sr, n_samples); framing(sr) -> (hop, win, n_fft).
    Returns, per rank, the units it computes.  Deterministic, identical on every rank.

    The frames of all channels, in the order given, form one line weighted by their cost; rank r takes the
    r-th of `world_size` equal spans of it.  A channel that straddles a span boundary is split there by frame
    range (at an even frame, so that the frame-pair kernel keeps whole pairs).  Consequences the device store
    relies on (it is keyed by (id, ch)): a rank holds AT MOST ONE unit of any channel, and that unit is one
    contiguous frame range -- two parts of a file can never meet on one rank."""
    per = []
    total = 0
    for (i, ch, sr, n) in channels:
        hop, win, n_fft = framing(sr)
        t = n_frames(n, win, hop)
        w = frame_weight(int(n_fft))
        per.append((i, ch, sr, n, hop, win, t, w))
        total += t * w
    ranks: List[List[Unit]] = [[] for _ in range(world_size)]
    if total == 0:
        return ranks
    # cut points of the weighted line: rank r owns [r * total / W, (r + 1) * total / W)
    cuts = [(r * total) // world_size for r in range(world_size + 1)]
    pos = 0
    r = 0
    for (i, ch, sr, n, hop, win, t, w) in per:
        begin = 0
        while begin < t:
            while r + 1 < world_size and pos >= cuts[r + 1]:
                r += 1
            # frames of this channel that still fit the span of rank r
            room = cuts[r + 1] - pos if r + 1 < world_size else (t - begin) * w
            cnt = min(t - begin, max(1, -(-room // w)))
            if begin + cnt < t:
                # not the channel's tail: keep the cut on an even frame
                cnt = cnt + 1 if (begin + cnt) & 1 else cnt
                cnt = min(cnt, t - begin)
            lo, hi = needed_samples(begin, cnt, win, hop, n)
            ranks[r].append(Unit(i, ch, sr, n, begin, cnt, lo, hi))
            begin += cnt
            pos += cnt * w
    # one contiguous unit per (id, ch) and rank: merge neighbours that the even-frame rounding left on one rank
    for q in range(world_size):
        merged: List[Unit] = []
        for u in ranks[q]:
            m = merged[-1] if merged else None
            if m and (m.id, m.ch) == (u.id, u.ch) and m.frame_begin + m.frame_count == u.frame_begin:
                hop, win, _ = framing(u.sr)
                cnt = m.frame_count + u.frame_count
                lo, hi = needed_samples(m.frame_begin, cnt, win, hop, u.full_len)
                merged[-1] = Unit(u.id, u.ch, u.sr, u.full_len, m.frame_begin, cnt, lo, hi)
            else:
                merged.append(u)
        ranks[q] = merged
    return ranks
