#!/usr/bin/env python
"""Step time of SMALL jobs on one GPU: what a rank of a frame-range shard runs at N = 8 (an eighth of C2's one-hour
file: 84 375 frames, mel 128, hop 256) and C1 (44 s mono, linear).  Queued steps (thb_spec_batch + thb_update_spec_imgs
without a host round trip), CUDA events on the launching stream; then the per-kernel profile of the same step.  The
difference between the two is what the launches cost -- the thing THB_PDL / THB_EDGE_SIDE act on (one process per
setting: both are read once).

    THB_PDL=0 python tools/smallstep.py ; THB_PDL=1 python tools/smallstep.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

import thesia_b200 as thb  # noqa: E402


def run(ctx, stream, name, n, setting, sr, interior, steps=300):
    pcm = torch.empty(n + 4096, dtype=torch.float32, device="cuda")
    ctx.synth_pcm(pcm, sr, 0, 0, 0)
    hop, win, n_fft = setting.calc_framing_params(sr)
    if interior:   # a middle rank's shard: every frame interior, no file edge
        T = thb.n_frames(8 * n, win, hop) // 8 // 2 * 2
        fb = T  # frames [T, 2T) of a file eight times as long
        lo = fb * hop - win // 2 - (n_fft - win) // 2
        tr = dict(pcm=pcm[: T * hop + n_fft], id=0, ch=0, sr=sr, full_len=8 * n, pcm_offset=lo, frame_begin=fb, frame_count=T)
    else:
        T = thb.n_frames(n, win, hop)
        tr = dict(pcm=pcm[:n], id=0, ch=0, sr=sr)
    tracks = ctx.prepare_tracks([tr])
    for _ in range(5):
        ctx.spec_batch(tracks, setting)
        ctx.update_spec_imgs(100.0, 258, sr, wait=False)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(stream)
        for _ in range(steps):
            ctx.spec_batch(tracks, setting)
            ctx.update_spec_imgs(100.0, 258, sr, wait=False)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(20):
        ctx.spec_batch(tracks, setting)
        ctx.update_spec_imgs(100.0, 258, sr, wait=False)
    ctx.synchronize()
    kn = "stft_mel_db" if setting.freq_scale == thb.FreqScale.Mel else "stft_lin_db"
    parts = {k: ctx.profile_get(k)[0] / 20 * 1e3 for k in (kn, kn + "_edges", "minmax_reduce", "spec_to_img")}
    ctx.profile_enable(False)
    ctx.release_all()
    print(f"{name:<44} {T:>7} frames  step {1e3 * best:7.1f} us   kernels: " + ", ".join(f"{k} {v:.1f}" for k, v in parts.items()) +
          f"   sum {sum(parts.values()):.1f} us")


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300   # (a handful under ncu)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = thb.Context(0, stream.cuda_stream)
    print("THB_PDL=%s THB_EDGE_SIDE=%s" % (os.environ.get("THB_PDL", "(default)"), os.environ.get("THB_EDGE_SIDE", "(default)")))
    c2 = thb.SpecSetting(2048 / 48.0, 8, 1, thb.FreqScale.Mel, 128)
    run(ctx, stream, "C2 / 8, a middle rank's shard (no file edge)", 3600 * 48000 // 8, c2, 48000, True, steps)
    run(ctx, stream, "C2 / 8 as a whole file (both file edges)", 3600 * 48000 // 8, c2, 48000, False, steps)
    run(ctx, stream, "C1 (44 s mono, linear 2048/512)", 2113529, thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Linear), 48000, False, steps)
    ctx.close()


if __name__ == "__main__":
    main()
