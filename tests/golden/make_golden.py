#!/usr/bin/env python
"""Regenerates tests/golden/*.npz / *.json.  Run from the repo root: python tests/golden/make_golden.py

Two kinds of fixtures:
  reference_kats.json  the known-answer vectors the reference's OWN tests hold for the hot path, transcribed value
                       for value with the file:line of the test they come from (/root/reference is Rust and cannot
                       be built or imported here -- no cargo -- so these literals are the only outputs of the real
                       reference that exist).
  oracle_*.npz         outputs of the CPU oracle (f64 truth leg for the spectrograms) on the deterministic synthetic
                       PCM of thesia_b200/synth.py.  They pin the oracle against drift and give the GPU tests a
                       committed target; they are NOT outputs of the reference itself (end-to-end parity unpinned,
                       see DESIGN.md section 2).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402
from thesia_b200.synth import LOUD, ZERO_GAP, synth_pcm  # noqa: E402

HERE = Path(__file__).resolve().parent

SPEC_CASES = [
    # name, sr, n_samples, win_ms, t_overlap, f_overlap, scale, n_mel, track, flags
    ("lin_2048_512", 48000, 9000, 2048 / 48.0, 4, 1, "linear", 0, 0, 0),
    ("mel128_2048_256", 48000, 20000, 2048 / 48.0, 8, 1, "mel", 128, 5, 0),
    ("default_40ms_44k1", 44100, 22050, 40.0, 4, 1, "mel", 0, 7, 0),
    ("mel128_zero_gap", 8000, 20000, 64.0, 4, 1, "mel", 128, 2, ZERO_GAP),
    ("lin_16384_1024", 96000, 9000, 16384 / 96.0, 16, 1, "linear", 0, 3, 0),
]


def main():
    kats = {
        "stft_works": {"cite": "src-tauri/src/core/spectrogram/stft.rs:173-196", "wav": [0, 0, 1, 0], "win": 4, "hop": 2, "n_fft": 4,
                       "re": [[0, 0, 0], [0.25, -0.25, 0.25], [0.25, -0.25, 0.25]], "im": [[0, 0, 0], [0, 0, 0], [0, 0, 0]]},
        "hann_window_works": {"cite": "src-tauri/src/core/windows.rs:88-91", "size": 4, "symmetric": False, "out": [0, 0.5, 1.0, 0.5]},
        "pad_works_reflect": {"cite": "src-tauri/src/core/utils.rs:166-176", "x": [1, 2, 3], "left": 3, "right": 4,
                              "out": [2, 3, 2, 1, 2, 3, 2, 1, 2, 3]},
        "spectrogram_to_img": {"cite": "src-tauri/src/core/visualize/drawing.rs:41-56", "spec": [[-100.0, -50.0, 0.0], [100.0, -200.0, -25.0]],
                               "i_freq_range": [0, 4], "dB_range": [-100.0, 0.0], "colormap_length": 4,
                               "img": [[16384, 65535], [40960, 0], [65535, 53247], [0, 0]]},
        "waveform_tile": {"cite": "src-tauri/src/core/render_tiles.rs:408-433",
                          "cases": [{"wav": [-1.0, 0.0, 0.5, 1.0], "revision": 3, "level": 1, "tile": 0, "bins": 2,
                                     "first": [-1.0, 0.0, -0.5]},
                                    {"wav_const": 0.25, "wav_len": 1025, "revision": 1, "level": 0, "tile": 1, "bins": 1},
                                    {"wav_ramp_from": -32.0, "wav_len": 64, "revision": 1, "level": 6, "tile": 0, "bins": 1,
                                     "first": [-32.0, 31.0, -0.5]}]},
        "spectrogram_tile": {"cite": "src-tauri/src/core/render_tiles.rs:435-471", "colors": [0, 0, 0, 255, 255, 0, 0, 255],
                             "cases": [{"img": [[0, 65535], [65535, 65535]], "args": [4, 1, 1, 0, 0], "width": 1, "height": 1,
                                        "pixels": [255, 0, 0, 255]},
                                       {"img_const": 65535, "shape": [513, 513], "args": [4, 0, 0, 1, 1], "width": 5, "height": 5,
                                        "origin": [508, 508], "all_pixels": [255, 0, 0, 255]},
                                       {"img": [[0], [65535]], "args": [4, 0, 0, 0, 0], "width": 1, "height": 2,
                                        "pixels": [255, 0, 0, 255, 0, 0, 0, 255]}]},
        "guard_clipping_stats": {"cite": "src-tauri/src/core/dynamics/stats.rs:224-241", "before_clip": [-1.5, -1.0, 0.5, 2.0],
                                 "reduction_cnt": 2, "max_reduction_gain_dB_of_amp": 0.5, "display": "max -6.02 dB, total 2 samples"},
        "normalize_targets": {"cite": "src-tauri/src/core/dynamics/normalize.rs:84-110",
                              "stats": {"lufs": -23.0, "rms_dB": -12.0, "max_peak": 0.5, "max_peak_dB": -6.0},
                              "cases": [["LUFS", -20.0, 3.0], ["RMSdB", -18.0, -6.0], ["PeakdB", -1.0, 5.0]]},
    }
    (HERE / "reference_kats.json").write_text(json.dumps(kats, indent=1) + "\n")

    specs = {}
    for name, sr, n, win_ms, t_ov, f_ov, scale, n_mel, track, flags in SPEC_CASES:
        wav = synth_pcm(n, sr, track, 0, flags)
        an = orc.Analyzer(sr, win_ms, t_ov, f_ov, orc.MEL if scale == "mel" else orc.LINEAR, n_mel)
        truth = an.calc_spec_truth(wav)
        specs[name] = truth.astype(np.float32)   # f32 storage: 4e-6 dB of quantisation against a 1e-3 dB bar
    np.savez_compressed(HERE / "oracle_spec_truth.npz", **specs)

    # quantiser + tiles + gain on small fixed inputs
    rng = np.random.default_rng(2026)
    spec = (rng.random((300, 96)) * 130.0 - 120.0).astype(np.float32)
    spec[5, 7] = -np.inf
    img = orc.spec_to_img(spec, (0, 96), (-100.0, 0.0), 258)
    cm = rng.integers(0, 256, 258 * 4, dtype=np.uint8)
    big = (np.add.outer(np.arange(140) * 401.0, np.arange(1100) * 37.0) % 65536).astype(np.uint16)
    tiles = {f"tile_{lx}_{ly}_{tx}_{ty}": np.frombuffer(orc.encode_spectrogram_tile(big, cm.tobytes(), 7, lx, ly, tx, ty), np.uint8)
             for lx, ly, tx, ty in ((0, 0, 2, 0), (1, 0, 1, 0), (2, 1, 0, 0), (4, 3, 0, 0))}
    wav = synth_pcm(50000, 48000, 9, 0, 0)
    wtile = np.frombuffer(orc.encode_waveform_tile(wav, 5, 4, 1), np.uint8)
    loud = synth_pcm(30000, 48000, 4, 0, LOUD)[None, :] * np.float32(0.06)
    g_clip = orc.apply_gain(loud, 1.7, orc.GUARD_CLIP)
    g_red = orc.apply_gain(loud, 1.7, orc.GUARD_REDUCE_GLOBAL_LEVEL)
    np.savez_compressed(HERE / "oracle_misc.npz", spec=spec, img=img, colormap=cm, big=big, waveform_tile=wtile,
                        gain_in=loud, clip_out=g_clip[0], clip_cnt=np.array([s[1] for s in g_clip[3]], np.uint64),
                        reduce_out=g_red[0], reduce_gain=np.float32(g_red[2]), **tiles)
    real_excerpt()
    print("wrote", sorted(p.name for p in HERE.iterdir()))


def real_excerpt():
    """BASELINE config C1 names a real file.  samples/sample_48k.wav is missing from the reference checkout;
    samples/sample_44k1.wav (the same programme, 16-bit mono 44.1 kHz; its shape is what audio.rs:467-511 tests) stands
    in.  An 8 s excerpt of its i16 samples is committed as a TEST FIXTURE (data, not source; VERDICT r1 #4a) so that the
    GPU box -- which has no /root/reference -- can run real audio through the CUDA path; every 16th row of the oracle's
    f64-truth dB spectrogram for C1's setting (win 2048 hop 512 linear) pins the oracle on it."""
    import wave
    path = Path("/root/reference/samples/sample_44k1.wav")
    if not path.exists():
        print("reference samples not present: keeping the committed excerpt")
        return
    with wave.open(str(path), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 44100)
        pcm = np.frombuffer(w.readframes(w.getnframes()), np.int16)
    sr, start, secs = 44100, 44100 * 20, 8
    ex = pcm[start:start + sr * secs].copy()
    x = ex.astype(np.float32) / np.float32(32768.0)            # the decoder's rule (audio.rs:262-439)
    an = orc.Analyzer(sr, 2048 / sr * 1000.0, 4, 1, orc.LINEAR, 0)
    truth = an.calc_spec_truth(x, n_threads=8)
    np.savez_compressed(HERE / "c1_sample_44k1_excerpt.npz", pcm_i16=ex, sr=np.int64(sr), start=np.int64(start),
                        truth_rows=np.arange(0, truth.shape[0], 16), truth_db=truth[::16].astype(np.float32))


if __name__ == "__main__":
    main()
