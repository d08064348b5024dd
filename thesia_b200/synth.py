"""numpy twin of thb_synth_pcm (thesia_b200/csrc/thb_envelope.cu): deterministic synthetic PCM in
integer arithmetic only, so host and device agree bit for bit.

x = 0.25 * tone(55 Hz + 13.75 Hz * (track % 61)) + 0.1 * chirp(50 Hz -> 0.45 sr) + 0.05 * noise,
quantised like 16-bit PCM (v / 32768); channel 1 = channel 0 delayed 7 samples * 0.8;
LOUD (flag 1) multiplies by 32 (> 0 dBFS, exercises the max_dB clamp); ZERO_GAP (flag 2) zeroes
samples [sr, 2 sr).
"""
import numpy as np

LOUD, ZERO_GAP = 1, 2


def _para_sine(phase):
    phase = phase.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    x = (phase >> np.uint64(15)) & np.uint64(0xFFFF)
    y = ((x * (np.uint64(65536) - x)) >> np.uint64(14)).astype(np.int64)
    return np.where((phase >> np.uint64(31)) != 0, -y, y)


def _mix32(h):
    m = np.uint64(0xFFFFFFFF)
    h = h & m
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & m
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE3D)) & m
    h ^= h >> np.uint64(16)
    return h


def _base(n, length, sr, track, flags):
    n = n.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    f0_mhz = 55000 + 13750 * (track % 61)
    inc0 = np.uint64((f0_mhz << 32) // (1000 * sr))
    ph0 = (n * inc0) & m
    inc_a = np.uint64((50 << 32) // sr)
    dinc = np.uint64(1932735283) - inc_a
    n2 = n * n
    two_len = np.uint64(2 * length)
    t_int = n2 // two_len
    t_rem = n2 % two_len
    ph1 = (inc_a * n + dinc * t_int + (dinc * t_rem) // two_len) & m
    h = _mix32(((n & m) * np.uint64(0x9E3779B1) + np.uint64((track * 0x7F4A7C15) & 0xFFFFFFFF) + np.uint64(0x7E51A)) & m)
    a0 = (8192 * _para_sine(ph0)) >> 16
    a1 = (3277 * _para_sine(ph1)) >> 16
    nz = (((h >> np.uint64(16)).astype(np.int64) - 32768) * 1638) >> 15
    v = a0 + a1 + nz
    if flags & ZERO_GAP:
        v = np.where((n >= np.uint64(sr)) & (n < np.uint64(2 * sr)), 0, v)
    if flags & LOUD:
        v = v * 32
    return v


def synth_pcm(length: int, sr: int, track: int, channel: int = 0, flags: int = 0) -> np.ndarray:
    n = np.arange(length, dtype=np.uint64)
    with np.errstate(over="ignore"):
        if channel == 0:
            v = _base(n, length, sr, track, flags)
        else:
            nn = np.where(n >= 7, n - np.uint64(7), 0).astype(np.uint64)
            v = np.where(n >= 7, (_base(nn, length, sr, track, flags) * 26214) >> 15, 0)
    return (v.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
