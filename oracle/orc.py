"""ctypes loader + numpy front end for oracle/thesia_oracle.c (TEST INFRASTRUCTURE ONLY).

Every function here is a thin wrapper; the arithmetic lives in the C restatement, which
cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_u64, _u32, _f32, _f64, _int = C.c_uint64, C.c_uint32, C.c_float, C.c_double, C.c_int
_pf32 = C.POINTER(C.c_float)
_pf64 = C.POINTER(C.c_double)
_pu16 = C.POINTER(C.c_uint16)
_pu8 = C.POINTER(C.c_uint8)
_pu64 = C.POINTER(C.c_uint64)


def build(force: bool = False) -> None:
    """Compile the oracle with the recipe in oracle/Makefile (gcc, OpenMP)."""
    so = _HERE / "libthesia_oracle.so"
    src = _HERE / "thesia_oracle.c"
    inc = _HERE / "orc_fft.inc"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, inc.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)


def _cpu_has_avx2_fma() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " avx2" in flags and " fma" in flags


def _load() -> C.CDLL:
    if not (_HERE / "libthesia_oracle.so").exists():
        build()
    name = "libthesia_oracle_avx2.so" if _cpu_has_avx2_fma() and (_HERE / "libthesia_oracle_avx2.so").exists() \
        else "libthesia_oracle.so"
    if os.environ.get("THESIA_ORACLE_BASELINE_ISA"):
        name = "libthesia_oracle.so"
    lib = C.CDLL(str(_HERE / name))
    sig = {
        "orc_framing_params": (None, [_f64, _u32, _u32, _u32, _pu64, _pu64, _pu64]),
        "orc_hann_f32": (None, [_u64, _int, _pf32]),
        "orc_normalized_hann_f32": (None, [_u64, _u64, _pf32]),
        "orc_pad_reflect_f32": (None, [_pf32, _u64, _u64, _u64, _pf32]),
        "orc_stft_frames_f32": (_u64, [_pf32, _u64, _u64, _u64, _pf32]),
        "orc_n_frames": (_u64, [_u64, _u64, _u64]),
        "orc_reflect_index": (C.c_int64, [C.c_int64, C.c_int64]),
        "orc_mel_to_hz_f32": (_f32, [_f32]),
        "orc_mel_from_hz_f32": (_f32, [_f32]),
        "orc_mel_to_hz_f64": (_f64, [_f64]),
        "orc_mel_from_hz_f64": (_f64, [_f64]),
        "orc_mel_fb_f32": (None, [_u32, _u64, _u64, _f32, _int, _f32, _int, _pf32]),
        "orc_mel_fb_f64": (None, [_u32, _u64, _u64, _f64, _int, _f64, _int, _pf64]),
        "orc_mel_fb_default_f32": (_u64, [_u32, _u64, _pf32]),
        "orc_hz_range_to_idx": (None, [_int, _f32, _f32, _u32, _u64, _pu64, _pu64]),
        "orc_dB_scalar_f32": (_f32, [_f32, _f32, _f32, _f32]),
        "orc_dB_from_amp_inplace_f32": (None, [_pf32, _u64, _f32, _f32]),
        "orc_find_min_max_f32": (None, [_pf32, _u64, _pf32, _pf32]),
        "orc_sum_squares_f32": (C.c_float, [_pf32, _u64]),
        "orc_abs_max_f32": (C.c_float, [_pf32, _u64]),
        "orc_audio_stats_f32": (None, [_pf32, _u64, _u64, _pf32]),
        "orc_sum_simd_order_f32": (_f32, [_pf32, _u64, _u32]),
        "orc_normalize_gain": (_f32, [_int, _f32, _f64, _f32, _f32]),
        "orc_apply_gain": (_int, [_pf32, _u64, _u64, _f32, _int, _pf32, _pf32, _pf32, _pf32, _pu64]),
        "orc_encode_waveform_tile": (_u64, [_pf32, _u64, _u64, _u32, _u32, _pu8]),
        "orc_resize_spectrogram_tile": (None, [_pu16, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _pu16]),
        "orc_spectrogram_tile_geometry": (None, [_u64, _u64, _u32, _u32, _u32, _u32, _pu64]),
        "orc_encode_spectrogram_tile": (_u64, [_pu16, _u64, _u64, _pu8, _u64, _u64, _u32, _u32, _u32, _u32, _pu8]),
        "orc_spec_to_img": (None, [_pf32, _u64, _u64, _u64, _u64, _f32, _f32, _int, _u32, _pu16]),
        "orc_clamp_minmax": (None, [_f32, _f32, _f32, _pf32, _pf32]),
        "orc_analyzer_new": (C.c_void_p, [_u32, _f64, _u32, _u32, _int, _u64]),
        "orc_analyzer_free": (None, [C.c_void_p]),
        "orc_analyzer_dims": (None, [C.c_void_p, _pu64, _pu64, _pu64, _pu64]),
        "orc_analyzer_window": (_pf32, [C.c_void_p]),
        "orc_analyzer_mel_fb": (_pf32, [C.c_void_p]),
        "orc_calc_spec_f32": (_u64, [C.c_void_p, _pf32, _u64, _pf32, _pf32, _pf32, _int]),
        "orc_calc_spec_f64": (_u64, [C.c_void_p, _pf32, _u64, _pf64, _pf64, _pf64, _int]),
        "orc_perform_stft_f32": (_u64, [_pf32, _u64, _u64, _u64, _u64, _pf32, _pf32, _pf32]),
        "orc_update_specs_and_imgs": (None, [C.c_void_p, C.POINTER(_pf32), _pu64, _u64, C.POINTER(_pf32),
                                             C.POINTER(_pu16), _f32, _u32, _int, _pf32, _pf32]),
        "orc_synth_pcm": (None, [_pf32, _u64, _u32, _u32, _u32, _u32, _int]),
        "orc_max_threads": (_int, []),
    }
    for name_, (res, args) in sig.items():
        fn = getattr(lib, name_)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _f32c(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32)


LINEAR, MEL = 0, 1


# ---- a1
def framing_params(win_ms: float, sr: int, t_overlap: int, f_overlap: int = 1):
    h, w, n = _u64(), _u64(), _u64()
    lib().orc_framing_params(win_ms, sr, t_overlap, f_overlap, C.byref(h), C.byref(w), C.byref(n))
    return h.value, w.value, n.value


# ---- a2
def hann(size: int, symmetric: bool = False) -> np.ndarray:
    out = np.empty(size, np.float32)
    lib().orc_hann_f32(size, int(symmetric), _p(out, _pf32))
    return out


def normalized_hann(win: int, n_fft: int) -> np.ndarray:
    out = np.empty(win, np.float32)
    lib().orc_normalized_hann_f32(win, n_fft, _p(out, _pf32))
    return out


# ---- a3
def pad_reflect(x, pad_left: int, pad_right: int) -> np.ndarray:
    x = _f32c(x)
    out = np.empty(x.size + pad_left + pad_right, np.float32)
    lib().orc_pad_reflect_f32(_p(x, _pf32), x.size, pad_left, pad_right, _p(out, _pf32))
    return out


# ---- a4
def stft_frames(x, win: int, hop: int) -> np.ndarray:
    """Raw frames (T, win) by the reference's three-piece front/mid/back construction."""
    x = _f32c(x)
    T = lib().orc_stft_frames_f32(_p(x, _pf32), x.size, win, hop, None)
    out = np.empty((T, win), np.float32)
    if T:
        lib().orc_stft_frames_f32(_p(x, _pf32), x.size, win, hop, _p(out, _pf32))
    return out


def n_frames(n: int, win: int, hop: int) -> int:
    return lib().orc_n_frames(n, win, hop)


def reflect_index(s: int, n: int) -> int:
    return lib().orc_reflect_index(s, n)


def perform_stft(x, win: int, hop: int, n_fft: int, window=None) -> np.ndarray:
    x = _f32c(x)
    T = lib().orc_perform_stft_f32(_p(x, _pf32), x.size, win, hop, n_fft, None, None, None)
    F = n_fft // 2 + 1
    re = np.empty((T, F), np.float32)
    im = np.empty((T, F), np.float32)
    w = None if window is None else _f32c(window)
    lib().orc_perform_stft_f32(_p(x, _pf32), x.size, win, hop, n_fft,
                               None if w is None else _p(w, _pf32), _p(re, _pf32), _p(im, _pf32))
    return re + 1j * im.astype(np.complex64)


# ---- a7
def mel_to_hz(mel: float, f64: bool = False) -> float:
    return lib().orc_mel_to_hz_f64(mel) if f64 else lib().orc_mel_to_hz_f32(mel)


def mel_from_hz(hz: float, f64: bool = False) -> float:
    return lib().orc_mel_from_hz_f64(hz) if f64 else lib().orc_mel_from_hz_f32(hz)


def mel_fb(sr: int, n_fft: int, n_mel: int, fmin: float = 0.0, fmax=None, do_norm: bool = True,
           f64: bool = False) -> np.ndarray:
    F = n_fft // 2 + 1
    if f64:
        out = np.empty((F, n_mel), np.float64)
        lib().orc_mel_fb_f64(sr, n_fft, n_mel, fmin, int(fmax is not None), fmax or 0.0, int(do_norm), _p(out, _pf64))
    else:
        out = np.empty((F, n_mel), np.float32)
        lib().orc_mel_fb_f32(sr, n_fft, n_mel, fmin, int(fmax is not None), fmax or 0.0, int(do_norm), _p(out, _pf32))
    return out


def mel_fb_default(sr: int, n_fft: int) -> np.ndarray:
    F = n_fft // 2 + 1
    n_mel = lib().orc_mel_fb_default_f32(sr, n_fft, None)
    out = np.empty((F, n_mel), np.float32)
    lib().orc_mel_fb_default_f32(sr, n_fft, _p(out, _pf32))
    return out


def mel_default_n(sr: int, n_fft: int) -> int:
    return lib().orc_mel_fb_default_f32(sr, n_fft, None)


def hz_range_to_idx(freq_scale: int, hz0: float, hz1: float, sr: int, n_bins: int):
    a, b = _u64(), _u64()
    lib().orc_hz_range_to_idx(freq_scale, hz0, hz1, sr, n_bins, C.byref(a), C.byref(b))
    return a.value, b.value


# ---- a9
def dB_scalar(x: float, ref: float = 1.0, amin: float = 0.0, factor: float = 20.0) -> float:
    return lib().orc_dB_scalar_f32(x, ref, amin, factor)


def dB_from_amp_inplace(x: np.ndarray, ref: float = 1.0, amin: float = 0.0) -> np.ndarray:
    assert x.dtype == np.float32 and x.flags.c_contiguous
    lib().orc_dB_from_amp_inplace_f32(_p(x, _pf32), x.size, ref, amin)
    return x


# ---- a12 / a17
def find_min_max(x):
    x = _f32c(x).ravel()
    a, b = _f32(), _f32()
    lib().orc_find_min_max_f32(_p(x, _pf32), x.size, C.byref(a), C.byref(b))
    return a.value, b.value


def sum_squares(x) -> float:
    """sum_squares_scalar (simd.rs:820-832)."""
    x = _f32c(x).ravel()
    return lib().orc_sum_squares_f32(_p(x, _pf32), x.size)


def abs_max(x) -> float:
    """abs_max_scalar (simd.rs:935-937)."""
    x = _f32c(x).ravel()
    return lib().orc_abs_max_f32(_p(x, _pf32), x.size)


def audio_stats(wavs):
    """StatCalculator::calc without global_lufs (stats.rs:56-85): (mean_squared, rms_dB, max_peak, max_peak_dB)."""
    w = _f32c(wavs)
    if w.ndim == 1:
        w = w[None, :]
    out = (C.c_float * 4)()
    lib().orc_audio_stats_f32(_p(w, _pf32), w.shape[0], w.shape[1], out)
    return tuple(out)


NORM_OFF, NORM_LUFS, NORM_RMS_DB, NORM_PEAK_DB = 0, 1, 2, 3
GUARD_CLIP, GUARD_REDUCE_GLOBAL_LEVEL, GUARD_LIMITER = 0, 1, 2


def normalize_gain(kind: int, target: float, global_lufs: float = 0.0, rms_dB: float = 0.0,
                   max_peak_dB: float = 0.0) -> float:
    """Normalize::normalize_default's gain (dynamics/normalize.rs:23-45)."""
    return lib().orc_normalize_gain(kind, target, global_lufs, rms_dB, max_peak_dB)


def apply_gain(wavs, gain: float, mode: int = GUARD_CLIP):
    """AudioTrack::apply_gain + guard clipping (track.rs:152-171, audio.rs:49-63,134-160).
    Returns (out (n_ch, n), before_clip or None, global_gain, [(max_reduction_gain_dB, reduction_cnt)] per channel)."""
    w = _f32c(wavs)
    if w.ndim == 1:
        w = w[None, :]
    out = np.empty_like(w)
    before = np.empty_like(w)
    gg = _f32()
    dB = np.zeros(w.shape[0], np.float32)
    cnt = np.zeros(w.shape[0], np.uint64)
    rc = lib().orc_apply_gain(_p(w, _pf32), w.shape[0], w.shape[1], gain, mode, _p(out, _pf32), _p(before, _pf32),
                              C.byref(gg), _p(dB, _pf32), _p(cnt, _pu64))
    if rc:
        raise ValueError("limiter mode is not restated")
    clipped = mode == GUARD_CLIP and np.isfinite(gain) and gain != 1.0
    return out, (before if clipped else None), gg.value, list(zip(dB.tolist(), cnt.tolist()))


def sum_simd_order(x, align_elems: int = 0) -> float:
    x = _f32c(x)
    return lib().orc_sum_simd_order_f32(_p(x, _pf32), x.size, align_elems)


def clamp_minmax(mn: float, mx: float, dB_range: float):
    a, b = _f32(), _f32()
    lib().orc_clamp_minmax(mn, mx, dB_range, C.byref(a), C.byref(b))
    return a.value, b.value


# ---- a16
def encode_waveform_tile(wav, revision: int, level: int, tile_index: int) -> bytes:
    wav = _f32c(wav)
    n = lib().orc_encode_waveform_tile(_p(wav, _pf32), wav.size, revision, level, tile_index, None)
    out = np.empty(n, np.uint8)
    lib().orc_encode_waveform_tile(_p(wav, _pf32), wav.size, revision, level, tile_index, _p(out, _pu8))
    return out.tobytes()


# ---- a14
def spectrogram_tile_geometry(H: int, W: int, level_x: int, level_y: int, tile_x: int, tile_y: int):
    """(lod_width, lod_height, origin_x, origin_y, width, height) of render_tiles.rs:290-312."""
    g = (C.c_uint64 * 6)()
    lib().orc_spectrogram_tile_geometry(H, W, level_x, level_y, tile_x, tile_y, g)
    return tuple(g)


def resize_spectrogram_tile(img, lod_w: int, lod_h: int, start_x: int, start_y: int, width: int, height: int):
    img = np.ascontiguousarray(img, np.uint16)
    out = np.empty((height, width), np.uint16)
    lib().orc_resize_spectrogram_tile(_p(img, _pu16), img.shape[0], img.shape[1], lod_w, lod_h, start_x, start_y, width,
                                      height, _p(out, _pu16))
    return out


def encode_spectrogram_tile(img, colormap_rgba, revision: int, level_x: int, level_y: int, tile_x: int, tile_y: int) -> bytes:
    """encode_spectrogram_tile (render_tiles.rs:281-350); img is (H, W) u16, colormap_rgba a byte string of RGBA."""
    img = np.ascontiguousarray(img, np.uint16)
    cm = np.frombuffer(bytes(colormap_rgba), np.uint8)
    n = lib().orc_encode_spectrogram_tile(_p(img, _pu16), img.shape[0], img.shape[1], _p(cm, _pu8), cm.size, revision,
                                          level_x, level_y, tile_x, tile_y, None)
    out = np.empty(n, np.uint8)
    lib().orc_encode_spectrogram_tile(_p(img, _pu16), img.shape[0], img.shape[1], _p(cm, _pu8), cm.size, revision,
                                      level_x, level_y, tile_x, tile_y, _p(out, _pu8))
    return out.tobytes()


def spec_to_img(spec, i_freq_range, dB_range, colormap_length=None) -> np.ndarray:
    spec = _f32c(spec)
    T, B = spec.shape
    i0, i1 = i_freq_range
    out = np.empty((i1 - i0, T), np.uint16)
    lib().orc_spec_to_img(_p(spec, _pf32), T, B, i0, i1, dB_range[0], dB_range[1],
                          int(colormap_length is not None), colormap_length or 0, _p(out, _pu16))
    return out


# ---- a10 analyzer
class Analyzer:
    """SpectrogramAnalyzer for one (sr, setting): window, FFT plan, mel bank (spectrogram.rs:101-212)."""

    def __init__(self, sr: int, win_ms: float, t_overlap: int, f_overlap: int = 1, freq_scale: int = MEL,
                 n_mel: int = 0):
        self.sr = sr
        self.freq_scale = freq_scale
        self._h = lib().orc_analyzer_new(sr, win_ms, t_overlap, f_overlap, freq_scale, n_mel)
        h, w, n, b = _u64(), _u64(), _u64(), _u64()
        lib().orc_analyzer_dims(self._h, C.byref(h), C.byref(w), C.byref(n), C.byref(b))
        self.hop, self.win, self.n_fft, self.n_bins = h.value, w.value, n.value, b.value
        self.n_freq = self.n_fft // 2 + 1

    def __del__(self):
        try:
            if self._h:
                lib().orc_analyzer_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def window(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().orc_analyzer_window(self._h), shape=(self.win,)).copy()

    @property
    def mel_fb(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().orc_analyzer_mel_fb(self._h), shape=(self.n_freq, self.n_bins)).copy()

    def n_frames(self, n: int) -> int:
        return n_frames(n, self.win, self.hop)

    def calc_spec(self, wav, n_threads: int = 1, want_stft: bool = False):
        """f32 reference-like dB spectrogram (T, n_bins)."""
        wav = _f32c(wav)
        T = self.n_frames(wav.size)
        out = np.empty((T, self.n_bins), np.float32)
        if want_stft:
            re = np.empty((T, self.n_freq), np.float32)
            im = np.empty((T, self.n_freq), np.float32)
            lib().orc_calc_spec_f32(self._h, _p(wav, _pf32), wav.size, _p(out, _pf32), _p(re, _pf32), _p(im, _pf32), n_threads)
            return out, re + 1j * im.astype(np.complex64)
        lib().orc_calc_spec_f32(self._h, _p(wav, _pf32), wav.size, _p(out, _pf32), None, None, n_threads)
        return out

    def calc_spec_truth(self, wav, n_threads: int = 1, want_amp: bool = False, want_pow: bool = False):
        """f64 truth: dB (T, n_bins) [+ amplitude (T, n_bins)] [+ power (T, n_freq)]."""
        wav = _f32c(wav)
        T = self.n_frames(wav.size)
        db = np.empty((T, self.n_bins), np.float64)
        amp = np.empty((T, self.n_bins), np.float64) if want_amp else None
        pw = np.empty((T, self.n_freq), np.float64) if want_pow else None
        lib().orc_calc_spec_f64(self._h, _p(wav, _pf32), wav.size, _p(db, _pf64),
                                None if amp is None else _p(amp, _pf64),
                                None if pw is None else _p(pw, _pf64), n_threads)
        res = [db]
        if want_amp:
            res.append(amp)
        if want_pow:
            res.append(pw)
        return res[0] if len(res) == 1 else tuple(res)

    def update_specs_and_imgs(self, wavs, dB_range: float = 100.0, colormap_length: int = 258,
                              n_threads: int = 1, want_imgs: bool = True):
        """TrackManager::update_specs + update_spec_imgs over channels sharing this analyzer."""
        wavs = [_f32c(w) for w in wavs]
        n = len(wavs)
        Ts = [self.n_frames(w.size) for w in wavs]
        specs = [np.empty((t, self.n_bins), np.float32) for t in Ts]
        imgs = [np.empty((self.n_bins, t), np.uint16) for t in Ts] if want_imgs else None
        pcm_arr = (_pf32 * n)(*[_p(w, _pf32) for w in wavs])
        len_arr = (_u64 * n)(*[w.size for w in wavs])
        spec_arr = (_pf32 * n)(*[_p(s, _pf32) for s in specs])
        img_arr = (_pu16 * n)(*[_p(i, _pu16) for i in imgs]) if want_imgs else None
        mn, mx = _f32(), _f32()
        lib().orc_update_specs_and_imgs(self._h, pcm_arr, len_arr, n, spec_arr, img_arr, dB_range,
                                        colormap_length, n_threads, C.byref(mn), C.byref(mx))
        return specs, imgs, mn.value, mx.value


def synth_pcm(length: int, sr: int, track: int, channel: int = 0, flags: int = 0, n_threads: int = 1,
              out: np.ndarray = None) -> np.ndarray:
    if out is None:
        out = np.empty(length, np.float32)
    assert out.dtype == np.float32 and out.size == length and out.flags.c_contiguous
    lib().orc_synth_pcm(_p(out, _pf32), length, sr, track, channel, flags, n_threads)
    return out


def max_threads() -> int:
    return lib().orc_max_threads()
