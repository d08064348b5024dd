"""ctypes binding of libthesia_b200.so -- exactly the declarations of include/thesia_b200.h.

The library is built in-tree by `__graft_entry__.build()` / `make -C thesia_b200/csrc`.  There is
no fallback: if the shared object is missing, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libthesia_b200.so"

THB_OK = 0
THB_ERR_INVALID = -1
THB_ERR_UNSUPPORTED = -2
THB_ERR_CUDA = -3
THB_ERR_NOMEM = -4
THB_ERR_NOT_FOUND = -5
THB_ERR_NCCL = -6
THB_ERR_SMALL_BUFFER = -7

FREQ_LINEAR = 0
FREQ_MEL = 1
SYNTH_LOUD = 1
SYNTH_ZERO_GAP = 2


class Setting(C.Structure):
    """thb_setting == SpecSetting (spectrogram.rs:30-38) + n_mel."""
    _fields_ = [("win_ms", C.c_double), ("t_overlap", C.c_uint32), ("f_overlap", C.c_uint32),
                ("freq_scale", C.c_uint32), ("n_mel", C.c_uint32)]


class Track(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("len", C.c_uint64), ("id", C.c_uint64), ("ch", C.c_uint32),
                ("sr", C.c_uint32), ("full_len", C.c_uint64), ("pcm_offset", C.c_uint64),
                ("frame_begin", C.c_uint64), ("frame_count", C.c_uint64), ("pcm_format", C.c_uint32),
                ("reserved", C.c_uint32)]


PCM_F32, PCM_I16 = 0, 1


class SpecOut(C.Structure):
    _fields_ = [("n_frames", C.c_uint64), ("total_frames", C.c_uint64), ("n_bins", C.c_uint32),
                ("hop", C.c_uint32), ("win", C.c_uint32), ("n_fft", C.c_uint32),
                ("spec_host", C.c_void_p), ("spec_host_cap", C.c_uint64)]


class AudioStats(C.Structure):
    _fields_ = [("mean_squared", C.c_float), ("rms_dB", C.c_float), ("max_peak", C.c_float), ("max_peak_dB", C.c_float)]


class GainChannel(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("len", C.c_uint64), ("id", C.c_uint64), ("pcm_format", C.c_uint32),
                ("gain", C.c_float), ("out", C.c_void_p), ("before_clip", C.c_void_p)]


class GainResult(C.Structure):
    _fields_ = [("global_gain", C.c_float), ("max_reduction_gain_dB", C.c_float), ("reduction_cnt", C.c_uint64),
                ("sum_squares", C.c_float), ("abs_max", C.c_float)]


class SpecTileReq(C.Structure):
    _fields_ = [("id", C.c_uint64), ("ch", C.c_uint32), ("level_x", C.c_uint32), ("level_y", C.c_uint32),
                ("tile_x", C.c_uint32), ("tile_y", C.c_uint32), ("reserved", C.c_uint32), ("out", C.c_void_p),
                ("cap", C.c_size_t), ("written", C.c_size_t)]


NORM_OFF, NORM_LUFS, NORM_RMS_DB, NORM_PEAK_DB = 0, 1, 2, 3
GUARD_CLIP, GUARD_REDUCE_GLOBAL_LEVEL, GUARD_LIMITER = 0, 1, 2

_P = C.POINTER
_vp, _u8p = C.c_void_p, _P(C.c_uint8)
_u64, _u32, _f32, _i = C.c_uint64, C.c_uint32, C.c_float, C.c_int

# name -> (restype, argtypes); one entry per declaration in include/thesia_b200.h
SIGNATURES = {
    "thb_ctx_create": (_i, [_i, _vp, _P(_vp)]),
    "thb_ctx_destroy": (None, [_vp]),
    "thb_set_stream": (_i, [_vp, _vp]),
    "thb_synchronize": (_i, [_vp]),
    "thb_last_error": (C.c_char_p, [_vp]),
    "thb_abi_version": (_i, []),
    "thb_host_alloc": (_i, [C.c_size_t, _P(_vp)]),
    "thb_host_free": (_i, [_vp]),
    "thb_framing_params": (_i, [_P(Setting), _u32, _P(_u64), _P(_u64), _P(_u64)]),
    "thb_n_frames": (_u64, [_u64, _u64, _u64]),
    "thb_n_bins": (_i, [_P(Setting), _u32, _P(_u32)]),
    "thb_hann_window": (_i, [_u64, _u64, _P(_f32)]),
    "thb_mel_fb": (_i, [_u32, _u64, _u32, _P(_f32), _P(_u32)]),
    "thb_channel_stats": (_i, [_vp, _P(Track), C.c_size_t, _P(_f32), _P(_f32)]),
    "thb_audio_stats": (_i, [_P(_f32), _P(_f32), _P(_u64), C.c_size_t, _P(AudioStats)]),
    "thb_normalize_gain": (_f32, [_u32, _f32, C.c_double, _f32, _f32]),
    "thb_apply_gain": (_i, [_vp, _P(GainChannel), C.c_size_t, _u32, _P(GainResult)]),
    "thb_mel_schedule_replay": (_i, [_u32, _u64, _u32, _P(_f32), _P(_u32)]),
    "thb_hz_range_to_idx": (_i, [_u32, _f32, _f32, _u32, _u64, _P(_u64), _P(_u64)]),
    "thb_spec_batch": (_i, [_vp, _P(Track), C.c_size_t, _P(Setting), _P(SpecOut)]),
    "thb_spec_put": (_i, [_vp, _u64, _u32, _u32, _u32, _vp, _u64, _u32]),
    "thb_spec_read": (_i, [_vp, _u64, _u32, _vp, _u64, _P(_u64), _P(_u32)]),
    "thb_spec_device_ptr": (_i, [_vp, _u64, _u32, _P(_vp), _P(_u64), _P(_u32)]),
    "thb_spec_minmax": (_i, [_vp, _u64, _u32, _P(_f32), _P(_f32)]),
    "thb_release": (_i, [_vp, _u64, _u32]),
    "thb_release_all": (_i, [_vp]),
    "thb_plans_prepare": (_i, [_vp, _P(Setting), _P(_u32), C.c_size_t]),
    "thb_plan_kernel": (_i, [_vp, _P(Setting), _u32, _P(_u32), _P(_u32)]),
    "thb_plans_retain": (_i, [_vp, _P(Setting), _P(_u32), C.c_size_t, _P(C.c_size_t)]),
    "thb_minmax_global": (_i, [_vp, _f32, _P(_f32), _P(_f32)]),
    "thb_spec_to_img": (_i, [_vp, _u64, _u32, _u64, _u64, _f32, _f32, _u32, _vp, _u64]),
    "thb_update_spec_imgs": (_i, [_vp, _f32, _u32, _u32, _P(_u64), C.c_size_t, _P(_f32), _P(_f32)]),
    "thb_range_get": (_i, [_vp, _P(_f32), _P(_f32)]),
    "thb_update_spec_imgs_range": (_i, [_vp, _f32, _f32, _u32, _u32, _P(_u64), C.c_size_t]),
    "thb_img_read": (_i, [_vp, _u64, _u32, _vp, _u64, _P(_u64), _P(_u64)]),
    "thb_img_read_batch": (_i, [_vp, C.c_size_t, _P(_u64), _P(_u32), _P(_vp), _P(_u64)]),
    "thb_img_put": (_i, [_vp, _u64, _u32, _vp, _u64, _u64]),
    "thb_img_device_ptr": (_i, [_vp, _u64, _u32, _P(_vp), _P(_u64), _P(_u64), _P(_u64)]),
    "thb_spectrogram_tile_geometry": (_i, [_u64, _u64, _u32, _u32, _u32, _u32, _P(_u64)]),
    "thb_spectrogram_tile": (_i, [_vp, _u64, _u32, _vp, C.c_size_t, _u64, _u32, _u32, _u32, _u32, _vp, C.c_size_t,
                                  _P(C.c_size_t)]),
    "thb_spectrogram_tile_batch": (_i, [_vp, _vp, C.c_size_t, _u64, _P(SpecTileReq), C.c_size_t]),
    "thb_waveform_tile": (_i, [_vp, _vp, _u64, _u64, _u32, _u32, _vp, C.c_size_t, _P(C.c_size_t)]),
    "thb_pcm_cache_stats": (_i, [_vp, _P(_u64), _P(_u64), _P(_u64), _P(_u64)]),
    "thb_pcm_cache_clear": (_i, [_vp]),
    "thb_waveform_level": (_i, [_vp, _vp, _u64, _u64, _u32, _vp, C.c_size_t, _P(C.c_size_t)]),
    "thb_waveform_level_batch": (_i, [_vp, _P(Track), C.c_size_t, _u64, _u32, _P(_vp), _P(C.c_size_t),
                                      _P(C.c_size_t), _P(_vp)]),
    "thb_waveform_level_bytes": (_u64, [_u64, _u32]),
    "thb_comm_unique_id": (_i, [_u8p]),
    "thb_comm_init": (_i, [_vp, _i, _i, _u8p]),
    "thb_comm_destroy": (_i, [_vp]),
    "thb_comm_peer_exchange": (_i, [_vp]),
    "thb_profile_enable": (_i, [_vp, _i]),
    "thb_profile_reset": (_i, [_vp]),
    "thb_profile_get": (_i, [_vp, C.c_char_p, _P(C.c_double), _P(_u64)]),
    "thb_launch_count": (_u64, [_vp]),
    "thb_synth_pcm": (_i, [_vp, _vp, _u64, _u32, _u32, _u32, _u32]),
}


def build(verbose: bool = False) -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (thesia_b200/csrc/Makefile)."""
    r = subprocess.run(["make", "-C", str(_HERE / "csrc"), "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libthesia_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib() -> C.CDLL:
    """Load libthesia_b200.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(thesia_b200 has no CPU fallback)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.thb_abi_version() != 5:
            raise RuntimeError("libthesia_b200.so ABI version mismatch")
        _lib = l
    return _lib


class ThbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"thesia_b200 error {code}: {msg}")
        self.code = code


def check(rc: int, ctx=None) -> None:
    if rc != THB_OK:
        msg = lib().thb_last_error(ctx)
        raise ThbError(rc, msg.decode() if msg else "")
