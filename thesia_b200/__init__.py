"""thesia_b200 -- B200-native analysis hot path of Sytronik/thesia behind a C ABI.

Layout: csrc/ (sm_100a kernels + extern "C" layer), _lib.py (ctypes binding of
include/thesia_b200.h), analysis.py (host-side mirror of the reference's SpecSetting /
SpectrogramAnalyzer / TrackManager / encode_waveform_tile interface), sharding.py (multi-GPU
work split), synth.py (deterministic synthetic PCM).
"""
from ._lib import FREQ_LINEAR, FREQ_MEL, ThbError  # noqa: F401
from .analysis import (Context, FreqScale, SpecSetting, TrackManager, calc_mel_fb, calc_mel_fb_default,  # noqa: F401
                       calc_normalized_win, encode_waveform_tile, hz_range_to_idx, n_frames)

__all__ = ["Context", "FreqScale", "SpecSetting", "TrackManager", "calc_mel_fb", "calc_mel_fb_default",
           "calc_normalized_win", "encode_waveform_tile", "hz_range_to_idx", "n_frames", "ThbError",
           "FREQ_LINEAR", "FREQ_MEL"]
