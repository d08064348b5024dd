"""Shared by the GPU parity tests: the comparison of a CUDA dB spectrogram with the oracle (f64 truth leg as the arbiter,
reference-like f32 leg as the yardstick of what f32 arithmetic can deliver) under the stated bars, and the record of every
case's measured margins (written to gpurun_out/parity_margins.json at the end of the session)."""
import json
from pathlib import Path

import numpy as np

import thesia_b200 as thb

FLOOR = 1e-5
POW_RTOL = 1e-4
DB_TOL = 1e-3


def _scale(orc, fs):
    return orc.MEL if fs == thb.FreqScale.Mel else orc.LINEAR


MARGINS = {}   # tag -> measured margins of that case; written to gpurun_out/parity_margins.json at session end


def dump_margins():
    if MARGINS:
        out = Path(__file__).resolve().parent.parent / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / "parity_margins.json").write_text(json.dumps(MARGINS, indent=1, sort_keys=True) + "\n")


def check_spec(orc, gpu_db, wav, sr, setting: thb.SpecSetting, tag=""):
    an = orc.Analyzer(sr, setting.win_ms, setting.t_overlap, setting.f_overlap, _scale(orc, setting.freq_scale),
                      setting.n_mel)
    truth_db, truth_amp = an.calc_spec_truth(wav, want_amp=True, n_threads=8)
    f32_db = an.calc_spec(wav, n_threads=8)
    assert gpu_db.shape == truth_db.shape, (tag, gpu_db.shape, truth_db.shape)
    g = gpu_db.astype(np.float64)
    neg = np.isneginf(truth_db)
    assert not np.isnan(g).any(), f"{tag}: NaN in GPU output"
    P = truth_amp ** 2
    # all-zero input frames: every bin must be exactly -inf (0 -> -inf, decibel.rs:193)
    silent = P.max(axis=1) == 0.0
    assert np.all(np.isneginf(g[silent])), f"{tag}: silent frames must be -inf"
    assert not np.isposinf(g).any()
    with np.errstate(over="ignore", invalid="ignore"):
        Pg = np.where(np.isneginf(g), 0.0, 10.0 ** (g / 10.0))
    floor = FLOOR * P.max(axis=1, keepdims=True)
    rel = np.abs(Pg - P) / np.maximum(np.maximum(P, floor), 1e-300)
    rel[silent] = 0.0
    worst_pow = float(rel.max()) if rel.size else 0.0
    above = (P > floor) & ~neg
    with np.errstate(invalid="ignore"):
        ddb = np.where(~neg, np.abs(g - np.where(neg, 0.0, truth_db)), 0.0)
        ddb32 = np.where(~neg, np.abs(f32_db.astype(np.float64) - np.where(neg, 0.0, truth_db)), 0.0)
    worst_db = float(ddb[above].max()) if above.any() else 0.0
    # the margins of this case at the survey's floors too (SURVEY.md section 7 asked 1e-7 of the frame's peak power):
    # worst relative power error of the GPU and of the reference-like f32 oracle above each floor
    with np.errstate(over="ignore", invalid="ignore"):
        Po = np.where(np.isneginf(f32_db), 0.0, 10.0 ** (f32_db.astype(np.float64) / 10.0))
    sweep = {}
    for fl in (1e-5, 1e-6, 1e-7):
        fl_abs = fl * P.max(axis=1, keepdims=True)
        den = np.maximum(np.maximum(P, fl_abs), 1e-300)
        rg, ro = np.abs(Pg - P) / den, np.abs(Po - P) / den
        rg[silent] = 0.0
        ro[silent] = 0.0
        sweep["%g" % fl] = {"gpu": float(rg.max()) if rg.size else 0.0, "f32_oracle": float(ro.max()) if ro.size else 0.0}
    hop_, win_, nfft_ = setting.calc_framing_params(sr)
    MARGINS[tag or "untagged"] = {"n_fft": int(nfft_), "win": int(win_), "hop": int(hop_), "bins": int(gpu_db.shape[1]),
                                  "power_rel_err_above_floor": sweep, "dB_err_above_1e-5_floor": worst_db,
                                  "f32_oracle_dB_err_above_1e-5_floor": float(ddb32[above].max()) if above.any() else 0.0}
    # at the survey's floors the reference-like f32 leg itself passes 1e-4 (measured: up to 4.6e-4 at 1e-7 for linear
    # spectra, profiles/r02_parity_margins.json), so below 1e-5 the bar is the survey's RELATIVE one: the GPU within 2x of
    # what f32 arithmetic in the reference's own operation order delivers.  The additive term covers the cases of a
    # handful of frames, where the maximum over a few hundred bins of two different FFT factorizations is luck
    # (measured on the 3 - 6 frame inputs of test_edges_and_short_inputs / test_extreme_levels: up to 5.5x at 1e-7).
    slack = {"1e-05": 2e-5, "1e-06": 5e-5, "1e-07": 3.5e-4}
    for fl, e in sweep.items():
        assert e["gpu"] <= 2.0 * e["f32_oracle"] + slack[fl], f"{tag}: power rel err above the {fl} floor {e['gpu']:.3g} vs f32 oracle {e['f32_oracle']:.3g}"
    assert worst_pow <= POW_RTOL, f"{tag}: power rel err {worst_pow:.3g}"
    assert worst_db <= DB_TOL, f"{tag}: dB err above floor {worst_db:.3g}"
    # below the floor an f32 FFT's error is ABSOLUTE (set by the frame's energy, not by the bin): compare the
    # amplitude error normalised by the frame's peak amplitude with the reference-like f32 oracle's own worst
    peak = np.sqrt(P.max(axis=1, keepdims=True))
    live = (peak[:, 0] > 0)
    if live.any():
        with np.errstate(over="ignore", invalid="ignore"):
            amp_g = np.where(np.isneginf(g), 0.0, 10.0 ** (g / 20.0))
            amp_o = np.where(np.isneginf(f32_db), 0.0, 10.0 ** (f32_db.astype(np.float64) / 20.0))
        eg = (np.abs(amp_g - truth_amp)[live] / peak[live]).max()
        eo = (np.abs(amp_o - truth_amp)[live] / peak[live]).max()
        MARGINS[tag or "untagged"].update({"amp_err_over_frame_peak": {"gpu": float(eg), "f32_oracle": float(eo)},
                                           "ratio_to_f32_oracle": float(eg / eo) if eo > 0 else None})
        # + 1.5e-6: an f32 dB value near -150 dB is itself quantised to 8.8e-7 relative in amplitude
        assert eg <= 2.0 * eo + 1.5e-6, f"{tag}: amplitude err / frame peak {eg:.3g} vs f32 oracle {eo:.3g}"
    return worst_pow, worst_db


