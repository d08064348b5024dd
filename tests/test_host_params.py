"""CPU tests of the product's host arithmetic through the C ABI (no GPU needed): the library must
load, export every symbol include/thesia_b200.h declares, and reproduce the oracle's integer
parameters and f32 tables exactly."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import thesia_b200 as thb
from thesia_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "thesia_b200.h").read_text()
    declared = set(re.findall(r"\b(thb_[a-z0-9_]+)\s*\(", header))
    declared -= {"thb_status"}
    assert len(declared) >= 35
    l = C.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(thb.ThbError) as e:
        thb.Context()
    assert e.value.code == _lib.THB_ERR_CUDA


@pytest.mark.parametrize("case", [(40.0, 48000, 4, 1), (40.0, 44100, 4, 1), (2048 / 48.0, 48000, 8, 1),
                                  (16384 / 96.0, 96000, 16, 1), (40.0, 22050, 4, 2), (1.0, 8000, 1, 1),
                                  (23.3, 16000, 32, 1), (40.0, 96000, 4, 1)])
def test_framing_params_match_oracle(orc, case):
    win_ms, sr, t, f = case
    s = thb.SpecSetting(win_ms, t, f)
    assert s.calc_framing_params(sr) == orc.framing_params(win_ms, sr, t, f)


def test_n_frames_matches_oracle(orc):
    for n in (2, 3, 100, 1919, 1920, 1921, 48000, 2113529):
        for w, h in ((1920, 480), (2048, 512), (2048, 256), (5, 5), (7, 1), (16384, 1024)):
            assert thb.n_frames(n, w, h) == orc.n_frames(n, w, h)


@pytest.mark.parametrize("win,n_fft", [(4, 4), (1920, 2048), (2048, 2048), (1764, 2048), (16384, 16384), (7, 8)])
def test_window_bit_exact(orc, win, n_fft):
    assert np.array_equal(thb.calc_normalized_win(win, n_fft), orc.normalized_hann(win, n_fft))


@pytest.mark.parametrize("sr,n_fft,n_mel", [(48000, 2048, 128), (24000, 2048, 80), (44100, 2048, 347),
                                            (96000, 16384, 128), (8000, 256, 40), (16000, 512, 1)])
def test_mel_fb_bit_exact(orc, sr, n_fft, n_mel):
    got = thb.calc_mel_fb(sr, n_fft, n_mel)
    want = orc.mel_fb(sr, n_fft, n_mel)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("sr", [400, 8000, 16000, 22050, 24000, 44100, 48000, 88200, 96000])
def test_mel_default_rule_matches_oracle(orc, sr):
    for e in (5, 8, 10, 11, 12, 14):
        n_fft = 2 ** e
        got = thb.calc_mel_fb_default(sr, n_fft)
        assert got.shape[1] == orc.mel_default_n(sr, n_fft), (sr, n_fft)
        if e in (8, 11):
            assert np.array_equal(got, orc.mel_fb_default(sr, n_fft))


def test_mel_default_sizes_of_the_survey():
    assert thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel).n_bins(48000) == 347
    assert thb.SpecSetting(16384 / 96.0, 16, 1, thb.FreqScale.Mel).n_bins(96000) == 1621
    assert thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Linear).n_bins(48000) == 1025
    assert thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128).n_bins(48000) == 128


def test_hz_range_to_idx_matches_oracle(orc):
    for scale in (0, 1):
        for sr in (8000, 22050, 44100, 48000):
            for max_sr in (48000, 96000, 44100):
                for nb in (128, 347, 1025):
                    assert thb.hz_range_to_idx(scale, (0.0, max_sr / 2), sr, nb) == \
                        orc.hz_range_to_idx(scale, 0.0, max_sr / 2, sr, nb)
    assert thb.hz_range_to_idx(0, (100.0, 100.0), 48000, 10) == (0, 0)


def test_waveform_level_bytes():
    l = _lib.lib()
    assert l.thb_waveform_level_bytes(4, 1) == 24 + 12 * 2
    assert l.thb_waveform_level_bytes(1025, 0) == 2 * 24 + 12 * 1025
    assert l.thb_waveform_level_bytes(0, 3) == 0


def test_synth_twin_is_deterministic_and_pcm_like():
    from thesia_b200.synth import synth_pcm, LOUD, ZERO_GAP
    a = synth_pcm(48000 * 3, 48000, 5, 0, ZERO_GAP)
    b = synth_pcm(48000 * 3, 48000, 5, 0, ZERO_GAP)
    assert np.array_equal(a, b)
    assert np.all(a[48000:96000] == 0) and np.any(a[:48000] != 0)
    assert np.all(a * 32768 == np.round(a * 32768))
    assert np.abs(a).max() < 0.45
    c1 = synth_pcm(4800, 48000, 5, 1, 0)
    c0 = synth_pcm(4800, 48000, 5, 0, 0)
    assert np.all(c1[:7] == 0) and np.abs(c1[7:] - 0.8 * c0[:-7]).max() < 2 / 32768
    assert np.abs(synth_pcm(4800, 48000, 2, 0, LOUD)).max() > 1.0


@pytest.mark.parametrize("sr,n_mel", [(48000, 128), (48000, 0), (44100, 128), (44100, 0), (48000, 8), (48000, 40),
                                      (16000, 64), (96000, 256), (48000, 1025), (8000, 0), (24000, 80)])
def test_mel_warp_schedule_is_the_filterbank(sr, n_mel):
    """The bin-major warp schedule of the n_fft 2048 kernels (MelItems, thb_host.cpp) applies every non-zero
    weight of calc_mel_fb exactly once, to the right band, with the reference's f32 value."""
    fb = thb.calc_mel_fb(sr, 2048, n_mel) if n_mel else thb.calc_mel_fb_default(sr, 2048)
    out = np.zeros_like(fb)
    stats = (C.c_uint32 * 4)()
    rc = _lib.lib().thb_mel_schedule_replay(sr, 2048, n_mel, out.ctypes.data_as(C.POINTER(C.c_float)), stats)
    assert rc == 0 and stats[0] == 1
    assert np.array_equal(out, fb)
    # every bin is walked once: steps per frame stay within 2x of the 1025 / 32 lower bound
    assert stats[2] <= (64 if fb.shape[1] <= 512 else 80)
    if n_mel in (0, 128):
        assert stats[3] <= 8  # (almost) bank-conflict free for the headline configurations


@pytest.mark.parametrize("sr,n_fft,n_mel", [(16000, 1024, 0), (22050, 1024, 0), (24000, 1024, 0), (8000, 512, 0), (16000, 1024, 64),
                                            (8000, 512, 40), (16000, 512, 0), (11025, 512, 0)])
def test_mel_schedules_of_the_warp_kernel_are_the_filterbank(sr, n_fft, n_mel):
    """n_fft 1024 / 512 (thb_stft_warp.cu): the bin-major schedule and the band-major one with its rounds PADDED IN PAIRS
    (the packed kernel walks two rounds at a time) both spell calc_mel_fb weight for weight -- the replay also checks
    that padding only ever adds zero weights and that bands past the last one have none."""
    fb = thb.calc_mel_fb(sr, n_fft, n_mel) if n_mel else thb.calc_mel_fb_default(sr, n_fft)
    out = np.zeros_like(fb)
    stats = (C.c_uint32 * 4)()
    rc = _lib.lib().thb_mel_schedule_replay(sr, n_fft, n_mel, out.ctypes.data_as(C.POINTER(C.c_float)), stats)
    assert rc == 0 and stats[0] == 1
    assert np.array_equal(out, fb)


# dynamics/normalize.rs:84-110 through the C ABI's host arithmetic (no device needed)
def test_normalize_gain_matches_reference_tests(orc):
    from thesia_b200.analysis import normalize_gain
    assert normalize_gain(_lib.NORM_OFF, -3.0, -23.0, -12.0, -6.0) == 1.0
    assert abs(normalize_gain(_lib.NORM_LUFS, -20.0, -23.0, -12.0, -6.0) - 10 ** (3.0 / 20.0)) <= 1e-6
    assert abs(normalize_gain(_lib.NORM_RMS_DB, -18.0, -23.0, -12.0, -6.0) - 10 ** (-6.0 / 20.0)) <= 1e-6
    assert abs(normalize_gain(_lib.NORM_PEAK_DB, -1.0, -23.0, -12.0, -6.0) - 10 ** (5.0 / 20.0)) <= 1e-6
    for kind in range(4):
        for tgt, a, b, c in ((-14.0, -26.2033, -31.5, -4.25), (0.0, -3.0, 1.5, 2.0)):
            assert normalize_gain(kind, tgt, a, b, c) == orc.normalize_gain(kind, tgt, a, b, c)


# render_tiles.rs:290-312 through the C ABI (host arithmetic): LOD sizes, gutters, partial and absent tiles
def test_spectrogram_tile_geometry_matches_oracle(orc):
    from thesia_b200.analysis import spectrogram_tile_geometry
    assert spectrogram_tile_geometry(513, 513, 0, 0, 1, 1) == (513, 513, 508, 508, 5, 5)     # render_tiles.rs:447-463
    assert spectrogram_tile_geometry(2, 2, 1, 1, 0, 0) == (1, 1, 0, 0, 1, 1)                 # render_tiles.rs:436-444
    for H, W in ((128, 56251), (1025, 4128), (8193, 337501), (1, 1), (347, 675001)):
        for lx, ly, tx, ty in ((0, 0, 0, 0), (0, 0, 3, 0), (3, 1, 1, 0), (9, 4, 0, 0), (40, 70, 0, 0), (2, 0, 10 ** 6, 0),
                               (0, 0, W // 512, H // 512), (1, 1, (W // 2) // 512, 0), (0, 3, 5, 1)):
            assert spectrogram_tile_geometry(H, W, lx, ly, tx, ty) == orc.spectrogram_tile_geometry(H, W, lx, ly, tx, ty)


# ---- randomized sweeps: the product's host arithmetic (C ABI, no device) against the oracle's restatement ----
def test_random_settings_framing_frames_bins_and_ranges(orc):
    """2 000 random (win_ms, sr, t_overlap, f_overlap) settings: hop / win / n_fft (spectrogram.rs:57-98, round-half-away in
    f64), frame counts, hz_range_to_idx and tile geometry must equal the oracle's integer for integer."""
    from thesia_b200.analysis import spectrogram_tile_geometry
    rng = np.random.default_rng(20261017)
    srs = [8000, 11025, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 176400, 192000]
    checked = 0
    for _ in range(2000):
        sr = int(rng.choice(srs))
        win_ms = float(rng.choice([rng.uniform(0.5, 200.0), rng.integers(1, 200), 2048 / 48.0, 1000.0 * 1024 / sr]))
        t_ov, f_ov = int(rng.choice([1, 2, 3, 4, 8, 16, 32])), int(rng.choice([1, 1, 1, 2, 4]))
        want = orc.framing_params(win_ms, sr, t_ov, f_ov)
        got = thb.SpecSetting(win_ms, t_ov, f_ov).calc_framing_params(sr)
        assert got == want, (win_ms, sr, t_ov, f_ov)
        hop, win, n_fft = got
        if hop == 0 or win < 3:
            continue
        n = int(rng.integers(2, 5_000_000))
        assert thb.n_frames(n, win, hop) == orc.n_frames(n, win, hop)
        nb = int(rng.integers(1, 4000))
        max_sr = int(rng.choice(srs))
        for scale in (0, 1):
            assert thb.hz_range_to_idx(scale, (0.0, max_sr / 2), sr, nb) == orc.hz_range_to_idx(scale, 0.0, max_sr / 2, sr, nb)
        H, W = int(rng.integers(1, 9000)), int(rng.integers(1, 700000))
        lx, ly = int(rng.integers(0, 14)), int(rng.integers(0, 10))
        tx, ty = int(rng.integers(0, 1 + (W >> lx) // 512 + 1)), int(rng.integers(0, 1 + (H >> ly) // 512 + 1))
        assert spectrogram_tile_geometry(H, W, lx, ly, tx, ty) == orc.spectrogram_tile_geometry(H, W, lx, ly, tx, ty)
        checked += 1
    assert checked > 1500


def test_mel_banks_bit_exact_over_sample_rates(orc):
    """calc_mel_fb_default over every sample rate x n_fft the kernels specialise for: the whole bank, bit for bit."""
    for sr in (8000, 16000, 22050, 44100, 48000, 96000, 192000):
        for n_fft in (512, 1024, 2048, 4096, 8192, 16384):
            if n_fft < sr // 100:     # windows shorter than 10 ms never occur with the reference's settings range
                continue
            got = thb.calc_mel_fb_default(sr, n_fft)
            assert np.array_equal(got, orc.mel_fb_default(sr, n_fft)), (sr, n_fft)


def test_integration_doc_declares_every_entry_point():
    """INTEGRATION.md's Rust extern "C" blocks name every function of include/thesia_b200.h, and nothing else."""
    import re
    header = (ROOT / "include" / "thesia_b200.h").read_text()
    declared = set(re.findall(r"\b(thb_[a-z0-9_]+)\s*\(", header))
    bound = set(re.findall(r"pub fn (thb_[a-z0-9_]+)", (ROOT / "INTEGRATION.md").read_text()))
    assert declared == bound, declared ^ bound
