#!/usr/bin/env python
"""Kernel timings of BASELINE.json's five configurations on one GPU (device-resident synthetic PCM, CUDA events
through the library's per-kernel profile hooks).  Not the bench line (bench.py is, on C3): the survey table behind
DESIGN.md's per-config numbers.  One JSON object per line on stdout and in gpurun_out/configs.jsonl.

    python tools/configs_bench.py [--only C1,C2,C4L,C4M,C5] [--reps 3]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

import thesia_b200 as thb  # noqa: E402

FP32_PEAK_TFLOPS = 1965e6 * 148 * 256 / 1e12  # sm_max_mhz x SMs x 128 lanes x 2 (bench.py overrides it with the measured clock)
HBM_GBS = 6552.3
try:
    HBM_GBS = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def synth(ctx, n_ch, n, sr):
    row = (n + 63) // 64 * 64
    pcm = torch.empty((n_ch, row), dtype=torch.float32, device="cuda")
    for c in range(n_ch):
        ctx.synth_pcm(pcm[c, :n], sr, c // 2, c % 2, 0)
    ctx.synchronize()
    return pcm


def stft_case(ctx, name, n_ch, seconds, sr, win_ms, t_overlap, scale, n_mel, reps, out):
    n = int(sr * seconds)
    setting = thb.SpecSetting(win_ms, t_overlap, 1, scale, n_mel)
    hop, win, n_fft = setting.calc_framing_params(sr)
    T = thb.n_frames(n, win, hop)
    B = setting.n_bins(sr)
    pcm = synth(ctx, n_ch, n, sr)
    tracks = [dict(pcm=pcm[c, :n], id=c, ch=0, sr=sr) for c in range(n_ch)]
    kname = "stft_mel_db" if scale == thb.FreqScale.Mel else "stft_lin_db"
    ctx.spec_batch(tracks, setting)  # warm-up
    ctx.synchronize()
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(reps):
        ctx.spec_batch(tracks, setting)
    ctx.synchronize()
    ms = (ctx.profile_get(kname)[0] + ctx.profile_get(kname + "_edges")[0]) / reps
    ctx.profile_enable(False)
    alg = n_ch * (4 * n + 4 * T * B)
    flops = n_ch * T * (2.5 * n_fft * (n_fft.bit_length() - 1) + 4 * (n_fft // 2 + 1) + 4 * (n_fft // 2 + 1) + B)
    rec = {"config": name, "channels": n_ch, "seconds": seconds, "sr": sr, "win": win, "hop": hop, "n_fft": n_fft, "bins": B,
           "frames": n_ch * T, "stft_ms": ms, "audio_hours_per_s": n_ch * seconds / 3600.0 / (ms * 1e-3),
           "Mframes_per_s": n_ch * T / ms / 1e3, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
           "hbm_frac": alg / (ms * 1e-3) / 1e9 / HBM_GBS, "fp32_tflops": flops / (ms * 1e-3) / 1e12,
           "frac_fp32": flops / (ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS}
    # the whole thb_spec_batch call, queued back to back (CUDA events on the launching stream): for a single short file the
    # scalar launches run NEXT TO the packed kernel (side stream), so the call is shorter than the sum of its kernel scopes
    k = max(reps, min(200, int(20.0 / max(ms, 0.02))))
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prepared = ctx.prepare_tracks(tracks)   # (descriptors marshalled once: the Python side must not be what is timed)
    ctx.spec_batch(prepared, setting)
    e0.record(st)
    for _ in range(k):
        ctx.spec_batch(prepared, setting)
    e1.record(st)
    ctx.synchronize()
    rec["spec_batch_ms"] = e0.elapsed_time(e1) / k
    ctx.release_all()
    del pcm
    torch.cuda.empty_cache()
    out(rec)


def envelope_case(ctx, n_ch, seconds, sr, reps, out, levels=range(9, 16), pcm=None):
    n = int(sr * seconds)
    if pcm is None:
        pcm = synth(ctx, n_ch, n, sr)
    wavs = [pcm[c, :n] for c in range(n_ch)]
    for level in levels:
        ctx.waveform_level_batch(wavs, 1, level, want_host=False)
        ctx.synchronize()
        ctx.profile_enable(True)
        ctx.profile_reset()
        for _ in range(reps):
            ctx.waveform_level_batch(wavs, 1, level, want_host=False)
        ctx.synchronize()
        ms = ctx.profile_get("envelope")[0] / reps
        ctx.profile_enable(False)
        bins = -(-n // (1 << level))
        tiles = -(-bins // 1024)
        alg = n_ch * (4 * n + 12 * bins + 24 * tiles)
        out({"config": "C5", "level": level, "columns_per_channel": bins, "channels": n_ch, "samples_per_channel": n,
             "envelope_ms": ms, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / HBM_GBS,
             "audio_hours_per_s": n_ch * seconds / 3600.0 / (ms * 1e-3)})
    # f3: sum of squares + absolute maximum of every channel (one pass over the same PCM)
    ctx.channel_stats(wavs)
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(reps):
        ctx.channel_stats(wavs)
    ms = ctx.profile_get("channel_stats")[0] / reps
    ctx.profile_enable(False)
    out({"config": "f3 channel_stats", "channels": n_ch, "samples_per_channel": n, "stats_ms": ms,
         "algorithmic_GBps": n_ch * 4 * n / (ms * 1e-3) / 1e9, "hbm_frac": n_ch * 4 * n / (ms * 1e-3) / 1e9 / HBM_GBS})
    # f4: gain + guard clipping in place on the device, 64 stereo tracks (4 B read + 4 B written per sample;
    # ReduceGlobalLevel adds the read-only peak pass).  A gain of 4 makes both guards bite.
    from thesia_b200 import _lib
    scratch = torch.empty_like(pcm)
    outs = [scratch[c, :n] for c in range(n_ch)]
    tracks = [dict(wavs=wavs[2 * t:2 * t + 2], outs=outs[2 * t:2 * t + 2], gain=4.0, id=t) for t in range(n_ch // 2)]
    for mode, name in ((_lib.GUARD_CLIP, "clip"), (_lib.GUARD_REDUCE_GLOBAL_LEVEL, "reduce_global_level")):
        ctx.apply_gain(tracks, mode)
        ctx.profile_enable(True)
        ctx.profile_reset()
        for _ in range(reps):
            ctx.apply_gain(tracks, mode)
        ms = ctx.profile_get("gain_apply")[0] / reps
        ms_peak = ctx.profile_get("gain_peak")[0] / reps
        ctx.profile_enable(False)
        alg = n_ch * 8 * n
        out({"config": "f4 apply_gain " + name, "channels": n_ch, "samples_per_channel": n, "gain_apply_ms": ms,
             "gain_peak_ms": ms_peak, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / HBM_GBS,
             "peak_pass_GBps": (n_ch * 4 * n / (ms_peak * 1e-3) / 1e9) if ms_peak else None})
    del pcm, scratch
    torch.cuda.empty_cache()


def tile_case(ctx, n_ch, seconds, sr, reps, out, levels=((0, 0), (2, 0), (4, 1)), pcm=None):
    """f2: every tile of a level for n_ch mel-default spectrogram images (347 x 56 251 at 10 min), levels 0 / 2 / 4 in x."""
    from thesia_b200.analysis import spectrogram_tile_geometry
    n = int(sr * seconds)
    if pcm is None:
        pcm = synth(ctx, n_ch, n, sr)
    setting = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 0)
    ctx.spec_batch([dict(pcm=pcm[c, :n], id=c, ch=0, sr=sr) for c in range(n_ch)], setting, want_host=False)
    ctx.update_spec_imgs(100.0, 258)
    H, W = ctx.img_read(0, 0).shape
    cm = bytes((i * 7 + j * 31) & 255 for i in range(258) for j in range(4))
    for lx, ly in levels:
        g = spectrogram_tile_geometry(H, W, lx, ly, 0, 0)
        tiles_x, tiles_y = -(-g[0] // 512), -(-g[1] // 512)
        reqs = [(c, 0, lx, ly, tx, ty) for c in range(n_ch) for ty in range(tiles_y) for tx in range(tiles_x)]
        bufs = ctx.spectrogram_tiles(cm, 1, reqs, want_bytes=False)
        ctx.profile_enable(True)
        ctx.profile_reset()
        for _ in range(reps):
            ctx.spectrogram_tiles(cm, 1, reqs, want_bytes=False)
        ms = ctx.profile_get("spectrogram_tile")[0] / reps
        ctx.profile_enable(False)
        out_bytes = sum(b.size - 40 for b in bufs)
        alg = n_ch * 2 * H * W + out_bytes
        out({"config": "f2 spectrogram tiles", "level_x": lx, "level_y": ly, "images": n_ch, "image_shape": [H, W], "tiles": len(reqs),
             "tile_kernels_ms": ms, "rgba_bytes": out_bytes, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
             "hbm_frac": alg / (ms * 1e-3) / 1e9 / HBM_GBS, "Mpixels_per_s": out_bytes / 4 / (ms * 1e-3) / 1e6})
    ctx.release_all()
    del pcm
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="C1,C2,C2D,C3D,C4L,C4M,C5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--c4-seconds", type=int, default=3600, help="shorter C4 tracks (profiling runs)")
    a = ap.parse_args()
    only = set(a.only.split(","))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)   # (the default stream's handle 0 would make the library create its own stream)
    ctx = thb.Context(0, stream.cuda_stream)
    outdir = Path("gpurun_out")
    outdir.mkdir(exist_ok=True)
    fh = open(outdir / "configs.jsonl", "a")

    def out(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        fh.write(line + "\n")
        fh.flush()

    Mel, Lin = thb.FreqScale.Mel, thb.FreqScale.Linear
    if "C1" in only:   # one 44 s mono file, win 2048 hop 512, linear dB
        stft_case(ctx, "C1 linear 2048/512, 2 113 529 samples mono", 1, 2113529 / 48000.0, 48000, 2048 / 48.0, 4, Lin, 0, a.reps, out)
    if "C2" in only:   # one 1-hour mono track, hop 256, mel 128
        stft_case(ctx, "C2 mel128 2048/256, 1 h mono", 1, 3600, 48000, 2048 / 48.0, 8, Mel, 128, a.reps, out)
    if "C2D" in only:  # same with the reference's default mel size (347 bands)
        stft_case(ctx, "C2 mel-default(347) 2048/256, 1 h mono", 1, 3600, 48000, 2048 / 48.0, 8, Mel, 0, a.reps, out)
    if "C3D" in only:  # the reference's default setting (40 ms / 4 -> win 1920 hop 480, default mel) on 32 x 10 min
        stft_case(ctx, "C3' default setting 40 ms/4 (1920/480/2048, mel 347), 32 ch x 10 min", 32, 600, 48000, 40.0, 4, Mel, 0, a.reps, out)
    if "C4M" in only:  # large FFT, default mel (1621 bands), 2 of the 8 one-hour 96 kHz tracks
        stft_case(ctx, "C4 mel-default 16384/1024 @96 kHz, 2 tracks", 2, a.c4_seconds, 96000, 16384 / 96.0, 16, Mel, 0, max(1, a.reps - 1), out)
    if "C4L" in only:  # large FFT, linear (8193 bins: 11 GB of f32 per track), 1 track
        stft_case(ctx, "C4 linear 16384/1024 @96 kHz, 1 track", 1, a.c4_seconds, 96000, 16384 / 96.0, 16, Lin, 0, max(1, a.reps - 1), out)
    if "C3S" in only:  # the reference default at 44.1 kHz: hop 441 is odd, every other frame starts on an odd sample
        stft_case(ctx, "C3'' default setting 40 ms/4 @44.1 kHz (1764/441/2048, mel default), 32 ch x 10 min", 32, 600, 44100, 40.0, 4, Mel, 0, a.reps, out)
    if "G96" in only:  # the 96 kHz default (n_fft 4096, mel default) alone: profiling target
        stft_case(ctx, "G default setting @96 kHz (3840/960/4096, mel default), 8 ch x 10 min", 8, 600, 96000, 40.0, 4, Mel, 0, a.reps, out)
    if "G" in only:    # other sample rates with the reference's default setting (40 ms / 4): the general kernel's sizes
        stft_case(ctx, "G default setting @16 kHz (640/160/1024, mel default), 32 ch x 10 min", 32, 600, 16000, 40.0, 4, Mel, 0, a.reps, out)
        stft_case(ctx, "G default setting @8 kHz (320/80/512, mel default), 32 ch x 10 min", 32, 600, 8000, 40.0, 4, Mel, 0, a.reps, out)
        stft_case(ctx, "G default setting @96 kHz (3840/960/4096, mel default), 8 ch x 10 min", 8, 600, 96000, 40.0, 4, Mel, 0, a.reps, out)
        stft_case(ctx, "G default setting @96 kHz linear (3840/960/4096), 8 ch x 10 min", 8, 600, 96000, 40.0, 4, Lin, 0, a.reps, out)
    if "C5" in only:
        envelope_case(ctx, 128, 600, 48000, a.reps, out)
    if "F2" in only:
        tile_case(ctx, 16, 600, 48000, a.reps, out)
    ctx.close()


if __name__ == "__main__":
    main()
