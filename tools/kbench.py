#!/usr/bin/env python
"""A/B timing of the STFT kernel variants on one GPU (device-resident synthetic PCM, CUDA events through
the library's own per-kernel profile hooks).  Not a bench line: a tuning aid.

    python tools/kbench.py [--channels 32] [--seconds 150] [--hop-overlap 4] [--scale mel|linear]
                           [--variants generic,fast,pair:8,pair:10,pair:12]
"""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

import thesia_b200 as thb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--seconds", type=int, default=150)
    ap.add_argument("--t-overlap", type=int, default=4)
    ap.add_argument("--scale", default="mel")
    ap.add_argument("--n-mel", type=int, default=128)
    ap.add_argument("--win-ms", type=float, default=2048 / 48.0)
    ap.add_argument("--sr", type=int, default=48000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="fast,pair:8,pair:10,pair:12")
    a = ap.parse_args()
    sr = a.sr
    n = sr * a.seconds
    fs = thb.FreqScale.Mel if a.scale == "mel" else thb.FreqScale.Linear
    setting = thb.SpecSetting(a.win_ms, a.t_overlap, 1, fs, a.n_mel if a.scale == "mel" else 0)
    hop, win, n_fft = setting.calc_framing_params(sr)
    T = thb.n_frames(n, win, hop)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)   # (the default stream's handle 0 would make the library create its own stream)
    ctx = thb.Context(0, stream.cuda_stream)
    row = (n + 63) // 64 * 64
    pcm = torch.empty((a.channels, row), dtype=torch.float32, device="cuda")
    for c in range(a.channels):
        ctx.synth_pcm(pcm[c, :n], sr, c // 2, c % 2, 0)
    tracks = [dict(pcm=pcm[c, :n], id=c // 2, ch=c % 2, sr=sr) for c in range(a.channels)]
    kname = "stft_mel_db" if a.scale == "mel" else "stft_lin_db"
    frames = a.channels * T
    print(f"{a.channels} ch x {a.seconds} s @ {sr} Hz, win {win} hop {hop} n_fft {n_fft}, {a.scale}, {frames} frames")
    ref = None
    for var in a.variants.split(","):
        # name[:warps][/ENV=VALUE...]   e.g.  pair/THB_PAIR_PREFETCH=0/THB_MEL4=0
        head, *envs = var.split("/")
        name, _, nw = head.partition(":")
        if name == "auto":   # the library's own choice
            os.environ.pop("THB_STFT_KERNEL", None)
        else:
            os.environ["THB_STFT_KERNEL"] = name
        os.environ.pop("THB_PAIR_WARPS", None)
        for k in ("THB_PAIR_PREFETCH", "THB_MEL4", "THB_PAIR_SHARE", "THB_PAIR_PL", "THB_MEL_DIRECT"):
            os.environ.pop(k, None)
        for kv in envs:
            k, _, v = kv.partition("=")
            os.environ[k] = v
        if nw:
            os.environ["THB_PAIR_WARPS"] = nw
        base = ctx
        ctx = thb.Context(0, stream.cuda_stream)  # a fresh plan cache: some switches act when a plan is built
        ctx.spec_batch(tracks, setting)  # warm-up
        ctx.synchronize()
        ctx.profile_enable(True)
        ctx.profile_reset()
        for _ in range(a.reps):
            ctx.spec_batch(tracks, setting)
        ctx.synchronize()
        ms, launches = ctx.profile_get(kname)
        ms += ctx.profile_get(kname + "_edges")[0]
        ctx.profile_enable(False)
        ms /= a.reps  # the scope covers every launch of one spec_batch call
        out = ctx.spec_read(0, 0)
        ctx.close()
        ctx = base
        if ref is None:
            ref = out
            diff = 0.0
        else:
            import numpy as np
            fin = np.isfinite(ref) & np.isfinite(out)
            diff = float(np.abs(ref[fin] - out[fin]).max()) if fin.any() else 0.0
        clk = ms * 1e-3 * 1.965e9 * 148 / frames
        print(f"  {var:<44} {ms:8.3f} ms  {frames / ms / 1e3:8.1f} Mframes/s  {clk:7.1f} clk/frame/SM   max|dB - first| {diff:.2e}")
    ctx.close()


if __name__ == "__main__":
    main()
