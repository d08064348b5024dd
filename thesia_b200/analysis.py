"""Host-side mirror of the reference's interface for the analysis path, on top of the C ABI.

Same names and argument meaning as the Rust backend so that tests read like the reference's own:
  SpecSetting            src-tauri/src/core/spectrogram.rs:30-99
  FreqScale              src-common/src/lib.rs:106-160
  calc_normalized_win    src-tauri/src/core/windows.rs:12-28
  calc_mel_fb[_default]  src-common/src/lib.rs:46-103
  TrackList (minimal)    src-tauri/src/core/track.rs:199-437  (only what TrackManager reads)
  TrackManager           src-tauri/src/core/mod.rs:33-231
  encode_waveform_tile   src-tauri/src/core/render_tiles.rs:232-259
Everything numeric happens in libthesia_b200.so (CUDA); this file only marshals pointers.  PCM may be
a numpy float32 array (host) or a torch CUDA tensor (device resident).
"""
from __future__ import annotations

import ctypes as C
import enum
import math
from dataclasses import dataclass, replace
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import PCM_F32, PCM_I16, Setting, SpecOut, Track, check, lib

IdCh = Tuple[int, int]


class FreqScale(enum.IntEnum):
    Linear = _lib.FREQ_LINEAR
    Mel = _lib.FREQ_MEL

    def hz_range_to_idx(self, hz_range: Tuple[float, float], sr: int, n_freqs_or_mels: int) -> Tuple[int, int]:
        return hz_range_to_idx(self, hz_range, sr, n_freqs_or_mels)


@dataclass
class SpecSetting:
    """SpecSetting (spectrogram.rs:30-38); `n_mel` is this library's extension (0 = the
    reference's calc_mel_fb_default rule, which is what TrackManager always uses)."""
    win_ms: float = 40.0
    t_overlap: int = 4
    f_overlap: int = 1
    freq_scale: FreqScale = FreqScale.Mel
    n_mel: int = 0

    def _c(self) -> Setting:
        return Setting(float(self.win_ms), int(self.t_overlap), int(self.f_overlap), int(self.freq_scale),
                       int(self.n_mel))

    def calc_framing_params(self, sr: int) -> Tuple[int, int, int]:
        """(hop_length, win_length, n_fft) -- spectrogram.rs:67-72."""
        h, w, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        s = self._c()
        check(lib().thb_framing_params(C.byref(s), sr, C.byref(h), C.byref(w), C.byref(n)))
        return h.value, w.value, n.value

    def calc_hop_length(self, sr: int) -> int:
        return self.calc_framing_params(sr)[0]

    def calc_win_length(self, sr: int) -> int:
        return self.calc_framing_params(sr)[1]

    def calc_sr_win_nfft(self, sr: int) -> Tuple[int, int, int]:
        _, w, n = self.calc_framing_params(sr)
        return sr, w, n

    def n_bins(self, sr: int) -> int:
        b = C.c_uint32()
        s = self._c()
        check(lib().thb_n_bins(C.byref(s), sr, C.byref(b)))
        return b.value


def n_frames(length: int, win: int, hop: int) -> int:
    return lib().thb_n_frames(length, win, hop)


def calc_normalized_win(win_length: int, n_fft: int) -> np.ndarray:
    out = np.empty(win_length, np.float32)
    check(lib().thb_hann_window(win_length, n_fft, out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def calc_mel_fb(sr: int, n_fft: int, n_mel: int) -> np.ndarray:
    """(n_fft/2+1, n_mel) f32 -- calc_mel_fb(sr, n_fft, n_mel, 0, None, true)."""
    out = np.empty((n_fft // 2 + 1, n_mel), np.float32)
    check(lib().thb_mel_fb(sr, n_fft, n_mel, out.ctypes.data_as(C.POINTER(C.c_float)), None))
    return out


def calc_mel_fb_default(sr: int, n_fft: int) -> np.ndarray:
    n = C.c_uint32()
    check(lib().thb_mel_fb(sr, n_fft, 0, None, C.byref(n)))
    out = np.empty((n_fft // 2 + 1, n.value), np.float32)
    check(lib().thb_mel_fb(sr, n_fft, 0, out.ctypes.data_as(C.POINTER(C.c_float)), None))
    return out


def hz_range_to_idx(freq_scale: int, hz_range: Tuple[float, float], sr: int, n_bins: int) -> Tuple[int, int]:
    a, b = C.c_uint64(), C.c_uint64()
    check(lib().thb_hz_range_to_idx(int(freq_scale), hz_range[0], hz_range[1], sr, n_bins, C.byref(a), C.byref(b)))
    return a.value, b.value


SPECTROGRAM_TILE_SIZE = 512   # render_tiles.rs:15


def spectrogram_tile_geometry(height: int, width: int, level_x: int, level_y: int, tile_x: int, tile_y: int):
    """(lod_width, lod_height, origin_x, origin_y, width, height) of a tile (render_tiles.rs:290-312)."""
    g = (C.c_uint64 * 6)()
    check(lib().thb_spectrogram_tile_geometry(height, width, level_x, level_y, tile_x, tile_y, g))
    return tuple(g)


class PreparedTracks:
    """A thb_track array built once from track dicts (and the buffers it points at, kept alive)."""

    def __init__(self, tracks: Sequence[dict]):
        self.n = len(tracks)
        self.arr = (Track * max(self.n, 1))()
        self.keep = []
        for i, t in enumerate(tracks):
            addr, ln, k, fmt = _ptr_len_fmt(t["pcm"])
            self.keep.append(k)
            self.arr[i] = Track(addr, ln, int(t["id"]), int(t.get("ch", 0)), int(t["sr"]), int(t.get("full_len", 0)),
                                int(t.get("pcm_offset", 0)), int(t.get("frame_begin", 0)), int(t.get("frame_count", 0)), fmt, 0)

    def __len__(self) -> int:
        return self.n


def normalize_gain(kind: int, target: float, global_lufs: float = 0.0, rms_dB: float = 0.0,
                   max_peak_dB: float = 0.0) -> float:
    """Normalize::normalize_default's gain (dynamics/normalize.rs:23-45); kind = _lib.NORM_*."""
    return lib().thb_normalize_gain(kind, target, global_lufs, rms_dB, max_peak_dB)


def _ptr_len_fmt(x) -> Tuple[int, int, object, int]:
    """(address, n_samples, keep-alive, pcm_format) of a 1-D PCM array: int16 arrays / tensors are handed over as
    THB_PCM_I16 (sample = s / 32768), everything else as float32."""
    if isinstance(x, np.ndarray) and x.dtype == np.int16:
        x = np.ascontiguousarray(x).reshape(-1)
        return x.ctypes.data, x.size, x, PCM_I16
    if hasattr(x, "data_ptr") and str(x.dtype) == "torch.int16":
        if not x.is_contiguous() or x.dim() != 1:
            raise ValueError("PCM tensors must be 1-D contiguous")
        return x.data_ptr(), x.numel(), x, PCM_I16
    return _ptr_len(x) + (PCM_F32,)


def _ptr_len(x) -> Tuple[int, int, object]:
    """(address, n_samples, keep-alive) of a 1-D float32 numpy array or torch tensor."""
    if isinstance(x, np.ndarray):
        if x.dtype != np.float32 or not x.flags.c_contiguous or x.ndim != 1:
            x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        return x.ctypes.data, x.size, x
    # torch tensor (host or device) without importing torch here
    if hasattr(x, "data_ptr"):
        if str(x.dtype) != "torch.float32" or not x.is_contiguous() or x.dim() != 1:
            raise ValueError("PCM tensors must be 1-D contiguous float32")
        return x.data_ptr(), x.numel(), x
    arr = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    return arr.ctypes.data, arr.size, arr


class Context:
    """thb_ctx: one per process and device.  `stream` is an optional raw cudaStream_t (int)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        check(lib().thb_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    def close(self) -> None:
        if self._h:
            lib().thb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def set_stream(self, stream: Optional[int]) -> None:
        check(lib().thb_set_stream(self._h, C.c_void_p(stream) if stream else None), self._h)

    def synchronize(self) -> None:
        check(lib().thb_synchronize(self._h), self._h)

    # ---- update_specs seam ----
    @staticmethod
    def prepare_tracks(tracks: Sequence[dict]) -> "PreparedTracks":
        """Marshal a track list into the C array once; pass the result to spec_batch() as often as needed."""
        return PreparedTracks(tracks)

    def spec_batch(self, tracks: Sequence[dict], setting: SpecSetting, want_host: bool = False):
        """tracks: dicts with pcm, id, ch, sr and optional full_len, pcm_offset, frame_begin,
        frame_count.  Returns a list of (n_frames, n_bins[, spec ndarray])."""
        if isinstance(tracks, PreparedTracks):   # descriptors marshalled once (a host program keeps its thb_track array)
            if want_host:
                raise ValueError("prepared track lists are for device-resident results")
            check(lib().thb_spec_batch(self._h, tracks.arr, tracks.n, C.byref(setting._c()), None), self._h)
            return None
        n = len(tracks)
        if n == 0:
            return []
        arr = (Track * n)()
        keep = []
        outs = (SpecOut * n)()
        s = setting._c()
        for i, t in enumerate(tracks):
            addr, ln, k, fmt = _ptr_len_fmt(t["pcm"])
            keep.append(k)
            arr[i] = Track(addr, ln, int(t["id"]), int(t.get("ch", 0)), int(t["sr"]), int(t.get("full_len", 0)),
                           int(t.get("pcm_offset", 0)), int(t.get("frame_begin", 0)), int(t.get("frame_count", 0)), fmt, 0)
        host = []
        if want_host:
            hop, win, n_fft = None, None, None
            for i, t in enumerate(tracks):
                hop, win, n_fft = setting.calc_framing_params(int(t["sr"]))
                full = int(t.get("full_len", 0)) or arr[i].len
                total = n_frames(full, win, hop)
                cnt = int(t.get("frame_count", 0)) or (total - int(t.get("frame_begin", 0)))
                b = setting.n_bins(int(t["sr"]))
                buf = np.empty((max(cnt, 0), b), np.float32)
                host.append(buf)
                outs[i].spec_host = buf.ctypes.data
                outs[i].spec_host_cap = buf.size
        check(lib().thb_spec_batch(self._h, arr, n, C.byref(s), outs), self._h)
        res = []
        for i in range(n):
            if want_host:
                res.append((outs[i].n_frames, outs[i].n_bins, host[i]))
            else:
                res.append((outs[i].n_frames, outs[i].n_bins))
        return res

    def plans_prepare(self, setting: SpecSetting, srs: Iterable[int]) -> None:
        """SpectrogramAnalyzer::prepare (spectrogram.rs:116-154) for the sample rates of the track list."""
        v = [int(x) for x in srs]
        arr = (C.c_uint32 * max(len(v), 1))(*v)
        check(lib().thb_plans_prepare(self._h, C.byref(setting._c()), arr, len(v)), self._h)

    def plans_retain(self, setting: SpecSetting, srs: Iterable[int]) -> int:
        """SpectrogramAnalyzer::retain (spectrogram.rs:156-185); returns the number of cached plans left."""
        v = [int(x) for x in srs]
        arr = (C.c_uint32 * max(len(v), 1))(*v)
        left = C.c_size_t()
        check(lib().thb_plans_retain(self._h, C.byref(setting._c()), arr, len(v), C.byref(left)), self._h)
        return left.value

    def calc_spec(self, wav, sr: int, setting: SpecSetting, id: int = 0, ch: int = 0) -> np.ndarray:
        """SpectrogramAnalyzer::calc_spec (spectrogram.rs:187-212): dB spectrogram (T, B)."""
        return self.spec_batch([dict(pcm=wav, id=id, ch=ch, sr=sr)], setting, want_host=True)[0][2]

    def spec_put(self, id: int, ch: int, sr: int, freq_scale: int, spec: np.ndarray) -> None:
        spec = np.ascontiguousarray(spec, np.float32)
        check(lib().thb_spec_put(self._h, id, ch, sr, int(freq_scale), spec.ctypes.data, spec.shape[0], spec.shape[1]),
              self._h)

    def spec_read(self, id: int, ch: int) -> np.ndarray:
        t, b = C.c_uint64(), C.c_uint32()
        check(lib().thb_spec_read(self._h, id, ch, None, 0, C.byref(t), C.byref(b)), self._h)
        out = np.empty((t.value, b.value), np.float32)
        check(lib().thb_spec_read(self._h, id, ch, out.ctypes.data, out.size, None, None), self._h)
        return out

    def spec_minmax(self, id: int, ch: int) -> Tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        check(lib().thb_spec_minmax(self._h, id, ch, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def release(self, id: int, ch: int) -> None:
        check(lib().thb_release(self._h, id, ch), self._h)

    def release_all(self) -> None:
        check(lib().thb_release_all(self._h), self._h)

    # ---- update_spec_imgs seam ----
    def minmax_global(self, dB_range: float) -> Tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        check(lib().thb_minmax_global(self._h, dB_range, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def spec_to_img(self, id: int, ch: int, i_freq_range: Tuple[int, int], dB_range: Tuple[float, float],
                    colormap_length: int) -> np.ndarray:
        t, b = C.c_uint64(), C.c_uint32()
        check(lib().thb_spec_read(self._h, id, ch, None, 0, C.byref(t), C.byref(b)), self._h)
        i0, i1 = i_freq_range
        out = np.zeros((i1 - i0, t.value), np.uint16)
        check(lib().thb_spec_to_img(self._h, id, ch, i0, i1, dB_range[0], dB_range[1], colormap_length,
                                    out.ctypes.data, out.size), self._h)
        return out

    @staticmethod
    def _ids(only_ids):
        if only_ids is None:
            return None, 0
        lst = list(only_ids)
        return (C.c_uint64 * max(len(lst), 1))(*lst), len(lst)

    def update_spec_imgs(self, dB_range: float, colormap_length: int, max_sr: int = 0,
                         only_ids: Optional[Iterable[int]] = None, wait: bool = True) -> Optional[Tuple[float, float]]:
        """thb_update_spec_imgs (one collective when a communicator is attached).  wait=False only queues the
        work on the stream (no host round trip): read the range later with range_get()."""
        a, b = C.c_float(), C.c_float()
        ids, n_ids = self._ids(only_ids)
        check(lib().thb_update_spec_imgs(self._h, dB_range, colormap_length, max_sr, ids, n_ids,
                                         C.byref(a) if wait else None, C.byref(b) if wait else None), self._h)
        return (a.value, b.value) if wait else None

    def update_spec_imgs_range(self, dB_range: Tuple[float, float], colormap_length: int, max_sr: int = 0,
                               only_ids: Optional[Iterable[int]] = None) -> None:
        """thb_update_spec_imgs_range: the quantise step alone with an already reduced (min_dB, max_dB); no collective."""
        ids, n_ids = self._ids(only_ids)
        check(lib().thb_update_spec_imgs_range(self._h, dB_range[0], dB_range[1], colormap_length, max_sr, ids, n_ids), self._h)

    def plan_kernel(self, setting: SpecSetting, sr: int) -> Tuple[int, int]:
        """thb_plan_kernel: (kernel family, mel schedule) thb_spec_batch uses for (setting, sr)."""
        fam, sch = C.c_uint32(), C.c_uint32()
        check(lib().thb_plan_kernel(self._h, C.byref(setting._c()), sr, C.byref(fam), C.byref(sch)), self._h)
        return fam.value, sch.value

    def comm_peer_exchange(self) -> bool:
        """True when the global dB range is exchanged over the peers' NVLink-mapped memory (thb_comm_peer_exchange)."""
        return bool(lib().thb_comm_peer_exchange(self._h))

    def range_get(self) -> Tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        check(lib().thb_range_get(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def pcm_cache_stats(self) -> dict:
        v = [C.c_uint64() for _ in range(4)]
        check(lib().thb_pcm_cache_stats(self._h, *[C.byref(x) for x in v]), self._h)
        return dict(zip(("entries", "bytes", "hits", "misses"), (x.value for x in v)))

    def pcm_cache_clear(self) -> None:
        check(lib().thb_pcm_cache_clear(self._h), self._h)

    def img_read(self, id: int, ch: int) -> np.ndarray:
        h, w = C.c_uint64(), C.c_uint64()
        check(lib().thb_img_read(self._h, id, ch, None, 0, C.byref(h), C.byref(w)), self._h)
        out = np.empty((h.value, w.value), np.uint16)
        check(lib().thb_img_read(self._h, id, ch, out.ctypes.data, out.size, None, None), self._h)
        return out

    def img_put(self, id: int, ch: int, img: np.ndarray) -> None:
        """thb_img_put: install an (H, T) u16 image as the retained image of (id, ch)."""
        img = np.ascontiguousarray(img, np.uint16)
        check(lib().thb_img_put(self._h, id, ch, img.ctypes.data, img.shape[0], img.shape[1]), self._h)

    def img_read_batch_into(self, id_chs: Sequence[IdCh], addrs: Sequence[int], caps: Sequence[int]) -> None:
        """thb_img_read_batch: all copies queued, one wait."""
        n = len(id_chs)
        ids = (C.c_uint64 * n)(*[int(i) for i, _ in id_chs])
        chs = (C.c_uint32 * n)(*[int(c) for _, c in id_chs])
        outs = (C.c_void_p * n)(*[int(a) for a in addrs])
        cps = (C.c_uint64 * n)(*[int(c) for c in caps])
        check(lib().thb_img_read_batch(self._h, n, ids, chs, outs, cps), self._h)

    def img_read_into(self, id: int, ch: int, addr: int, cap: int) -> Tuple[int, int]:
        h, w = C.c_uint64(), C.c_uint64()
        check(lib().thb_img_read(self._h, id, ch, addr, cap, C.byref(h), C.byref(w)), self._h)
        return h.value, w.value

    # ---- waveform tiles ----
    def waveform_tile(self, wav, revision: int, level: int, tile_index: int) -> bytes:
        addr, ln, _keep = _ptr_len(wav)
        need = C.c_size_t()
        check(lib().thb_waveform_tile(self._h, addr, ln, revision, level, tile_index, None, 0, C.byref(need)), self._h)
        out = np.empty(need.value, np.uint8)
        check(lib().thb_waveform_tile(self._h, addr, ln, revision, level, tile_index, out.ctypes.data, out.size,
                                      C.byref(need)), self._h)
        return out.tobytes()

    def waveform_level(self, wav, revision: int, level: int) -> bytes:
        addr, ln, _keep = _ptr_len(wav)
        nbytes = lib().thb_waveform_level_bytes(ln, level)
        out = np.empty(nbytes, np.uint8)
        wr = C.c_size_t()
        check(lib().thb_waveform_level(self._h, addr, ln, revision, level, out.ctypes.data, out.size, C.byref(wr)), self._h)
        return out.tobytes()

    def waveform_level_batch(self, wavs: Sequence, revision: int, level: int, want_host: bool = True):
        n = len(wavs)
        arr = (Track * n)()
        keep = []
        for i, w in enumerate(wavs):
            addr, ln, k = _ptr_len(w)
            keep.append(k)
            arr[i] = Track(addr, ln, i, 0, 0, 0, 0, 0, 0)
        written = (C.c_size_t * n)()
        dev = (C.c_void_p * n)()
        outs = None
        host_ptrs = None
        caps = None
        if want_host:
            outs = [np.empty(lib().thb_waveform_level_bytes(arr[i].len, level), np.uint8) for i in range(n)]
            host_ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
            caps = (C.c_size_t * n)(*[o.size for o in outs])
        check(lib().thb_waveform_level_batch(self._h, arr, n, revision, level, host_ptrs, caps, written, dev), self._h)
        if want_host:
            return [o.tobytes() for o in outs]
        return [(dev[i], written[i]) for i in range(n)]

    # ---- spectrogram tiles (render_tiles.rs:170-188,281-393) ----
    def spectrogram_tile(self, id: int, ch: int, colormap_rgba: bytes, revision: int, level_x: int, level_y: int,
                         tile_x: int, tile_y: int) -> bytes:
        """encode_spectrogram_tile for the retained image of (id, ch)."""
        return self.spectrogram_tiles(colormap_rgba, revision, [(id, ch, level_x, level_y, tile_x, tile_y)])[0]

    def spectrogram_tiles(self, colormap_rgba: bytes, revision: int, reqs: Sequence[Tuple[int, int, int, int, int, int]],
                          want_bytes: bool = True):
        """n tiles (id, ch, level_x, level_y, tile_x, tile_y) in one pair of launches."""
        n = len(reqs)
        cm = np.frombuffer(bytes(colormap_rgba), np.uint8)
        arr = (_lib.SpecTileReq * n)()
        for i, r in enumerate(reqs):
            arr[i] = _lib.SpecTileReq(int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]), 0, None, 0, 0)
        check(lib().thb_spectrogram_tile_batch(self._h, cm.ctypes.data, cm.size, revision, arr, n), self._h)  # sizes
        bufs = [np.empty(arr[i].written, np.uint8) for i in range(n)]
        for i in range(n):
            arr[i].out = bufs[i].ctypes.data
            arr[i].cap = bufs[i].size
        check(lib().thb_spectrogram_tile_batch(self._h, cm.ctypes.data, cm.size, revision, arr, n), self._h)
        return [b.tobytes() for b in bufs] if want_bytes else bufs

    # ---- level statistics (dynamics/stats.rs:56-85 without the loudness leg) ----
    def channel_stats(self, wavs: Sequence) -> Tuple[np.ndarray, np.ndarray]:
        """(sum_squares[n], abs_max[n]) of n channels (f32 or int16 arrays / tensors, host or device)."""
        n = len(wavs)
        arr = (Track * n)()
        keep = []
        for i, w in enumerate(wavs):
            addr, ln, k, fmt = _ptr_len_fmt(w)
            keep.append(k)
            arr[i] = Track(addr, ln, i, 0, 0, 0, 0, 0, 0, fmt, 0)
        ss, mx = np.zeros(n, np.float32), np.zeros(n, np.float32)
        check(lib().thb_channel_stats(self._h, arr, n, ss.ctypes.data_as(C.POINTER(C.c_float)),
                                      mx.ctypes.data_as(C.POINTER(C.c_float))), self._h)
        return ss, mx

    def calc_stats(self, wavs: Sequence) -> dict:
        """StatCalculator::calc for one track given its channels: rms_dB, max_peak, max_peak_dB (+ mean_squared)."""
        ss, mx = self.channel_stats(wavs)
        lens = (C.c_uint64 * len(wavs))(*[_ptr_len_fmt(w)[1] for w in wavs])
        out = _lib.AudioStats()
        check(lib().thb_audio_stats(ss.ctypes.data_as(C.POINTER(C.c_float)), mx.ctypes.data_as(C.POINTER(C.c_float)), lens,
                                    len(wavs), C.byref(out)), self._h)
        return dict(mean_squared=out.mean_squared, rms_dB=out.rms_dB, max_peak=out.max_peak, max_peak_dB=out.max_peak_dB)

    # ---- gain normalisation + guard clipping (track.rs:152-171, audio.rs:49-63,134-160) ----
    def apply_gain(self, tracks: Sequence[dict], mode: int = _lib.GUARD_CLIP, want_before_clip: bool = False):
        """tracks: [{"wavs": [channel, ...], "gain": g, "id": optional}] -- one entry per Audio; channels are 1-D f32
        (or int16) arrays / tensors, host or device.  Outputs are allocated like the inputs (numpy -> numpy,
        torch -> torch on the same device) unless "outs" is given.  Returns per track a dict(wavs, before_clip,
        global_gain, guard_clip_stats=[(max_reduction_gain_dB, reduction_cnt)], sum_squares, abs_max)."""
        chans = []
        keep = []
        res = []
        for ti, t in enumerate(tracks):
            outs = t.get("outs")
            entry = dict(wavs=[], before_clip=[] if want_before_clip and mode == _lib.GUARD_CLIP else None)
            for ci, w in enumerate(t["wavs"]):
                addr, ln, k, fmt = _ptr_len_fmt(w)
                keep.append(k)

                def alloc():
                    if hasattr(w, "data_ptr"):
                        import torch
                        return torch.empty(ln, dtype=torch.float32, device=w.device)
                    return np.empty(ln, np.float32)
                o = outs[ci] if outs is not None else alloc()
                o_addr = o.data_ptr() if hasattr(o, "data_ptr") else o.ctypes.data
                b_addr = None
                if entry["before_clip"] is not None:
                    b = alloc()
                    entry["before_clip"].append(b)
                    b_addr = b.data_ptr() if hasattr(b, "data_ptr") else b.ctypes.data
                entry["wavs"].append(o)
                chans.append(_lib.GainChannel(addr, ln, int(t.get("id", ti)), fmt, float(t["gain"]), o_addr, b_addr))
            res.append(entry)
        n = len(chans)
        arr = (_lib.GainChannel * n)(*chans)
        out = (_lib.GainResult * n)()
        check(lib().thb_apply_gain(self._h, arr, n, mode, out), self._h)
        k = 0
        for t, entry in zip(tracks, res):
            m = len(t["wavs"])
            rs = out[k:k + m]
            k += m
            entry["global_gain"] = rs[0].global_gain if m else 1.0
            entry["guard_clip_stats"] = [(r.max_reduction_gain_dB, r.reduction_cnt) for r in rs]
            entry["sum_squares"] = np.array([r.sum_squares for r in rs], np.float32)
            entry["abs_max"] = np.array([r.abs_max for r in rs], np.float32)
        return res

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(lib().thb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, n_ranks: int, rank: int, uid: bytes) -> None:
        buf = (C.c_uint8 * 128)(*uid)
        check(lib().thb_comm_init(self._h, n_ranks, rank, buf), self._h)

    def comm_destroy(self) -> None:
        check(lib().thb_comm_destroy(self._h), self._h)

    # ---- measurement ----
    def profile_enable(self, on: bool = True) -> None:
        check(lib().thb_profile_enable(self._h, int(on)), self._h)

    def profile_reset(self) -> None:
        check(lib().thb_profile_reset(self._h), self._h)

    def profile_get(self, kernel: str) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_uint64()
        check(lib().thb_profile_get(self._h, kernel.encode(), C.byref(ms), C.byref(n)), self._h)
        return ms.value, n.value

    def launch_count(self) -> int:
        return lib().thb_launch_count(self._h)

    def synth_pcm(self, dev_tensor, sr: int, track: int, channel: int, flags: int = 0) -> None:
        check(lib().thb_synth_pcm(self._h, dev_tensor.data_ptr(), dev_tensor.numel(), sr, track, channel, flags), self._h)


def encode_waveform_tile(wav, revision: int, level: int, tile_index: int, ctx: Optional[Context] = None) -> bytes:
    """encode_waveform_tile(&[f32], u64, u32, u32) -> Vec<u8> (render_tiles.rs:232)."""
    own = ctx is None
    ctx = ctx or Context()
    try:
        return ctx.waveform_tile(wav, revision, level, tile_index)
    finally:
        if own:
            ctx.close()


class TrackList:
    """The part of TrackList (track.rs:199-437) the analysis path reads: id -> (wavs (n_ch, N), sr)."""

    def __init__(self):
        self._tracks: Dict[int, Tuple[object, int]] = {}

    def add_tracks(self, id_list: Sequence[int], wavs_list: Sequence, sr_list: Sequence[int]) -> List[int]:
        for i, w, sr in zip(id_list, wavs_list, sr_list):
            if isinstance(w, np.ndarray):
                w = np.ascontiguousarray(w, np.float32)
                if w.ndim == 1:
                    w = w[None, :]
            self._tracks[int(i)] = (w, int(sr))
        return list(id_list)

    def remove_tracks(self, id_list: Sequence[int]) -> List[IdCh]:
        removed = []
        for i in id_list:
            if i in self._tracks:
                removed += [(i, ch) for ch in range(self.n_ch(i))]
                del self._tracks[i]
        return removed

    def has(self, id: int) -> bool:
        return id in self._tracks

    def all_ids(self) -> List[int]:
        return sorted(self._tracks)

    def all_id_set(self) -> set:
        return set(self._tracks)

    def n_ch(self, id: int) -> int:
        return self._tracks[id][0].shape[0]

    def sr(self, id: int) -> int:
        return self._tracks[id][1]

    def channel(self, id: int, ch: int):
        return self._tracks[id][0][ch]

    def max_sr(self) -> int:
        return max((sr for _, sr in self._tracks.values()), default=0)

    def id_ch_tuples_from(self, id_list: Sequence[int]) -> List[IdCh]:
        return [(i, ch) for i in id_list if self.has(i) for ch in range(self.n_ch(i))]

    def id_ch_tuples(self) -> List[IdCh]:
        return self.id_ch_tuples_from(self.all_ids())


class TrackManager:
    """TrackManager (mod.rs:33-231) with the same public fields and methods; `specs` and
    `spec_imgs` live on the device and are fetched on demand."""

    def __init__(self, ctx: Optional[Context] = None):
        self.ctx = ctx or Context()
        self.max_dB = -math.inf
        self.min_dB = math.inf
        self.max_sr = 0
        self.setting = SpecSetting()
        self.dB_range = 100.0
        self.colormap_length = 258
        self._spec_keys: set = set()
        self._img_keys: set = set()
        self._no_spec_img_ids: List[int] = []

    # -- mod.rs:62-84
    def add_tracks(self, tracklist: TrackList, added_ids: Sequence[int]) -> None:
        self._update_specs(tracklist, tracklist.id_ch_tuples_from(added_ids))
        self._no_spec_img_ids.extend(added_ids)

    def reload_tracks(self, tracklist: TrackList, reloaded_ids: Sequence[int]) -> None:
        self._update_specs(tracklist, tracklist.id_ch_tuples_from(reloaded_ids))
        self._no_spec_img_ids.extend(reloaded_ids)

    # -- mod.rs:86-100
    def remove_tracks(self, tracklist: TrackList, removed_id_ch_tuples: Sequence[IdCh]) -> None:
        for tup in removed_id_ch_tuples:
            if tup in self._spec_keys:
                self.ctx.release(*tup)
                self._spec_keys.discard(tup)
                self._img_keys.discard(tup)
        # spec_analyzer.retain(construct_all_sr_win_nfft_set(setting), freq_scale)  (mod.rs:96-99)
        self.ctx.plans_retain(self.setting, sorted({tracklist.sr(i) for i in tracklist.all_ids()}))

    # -- mod.rs:102-105
    def apply_track_list_changes(self, tracklist: TrackList):
        s = self._update_spec_imgs(tracklist, False)
        return s, self.max_sr

    # -- mod.rs:107-115
    def set_setting(self, tracklist: TrackList, setting: SpecSetting) -> None:
        self.setting = replace(setting)
        self.ctx.plans_retain(self.setting, sorted({tracklist.sr(i) for i in tracklist.all_ids()}))   # mod.rs:110-112
        self._update_specs(tracklist, tracklist.id_ch_tuples())
        self._update_spec_imgs(tracklist, True)

    def update_all_specs_imgs(self, tracklist: TrackList) -> None:
        self._update_specs(tracklist, tracklist.id_ch_tuples())
        self._update_spec_imgs(tracklist, True)

    # -- mod.rs:123-131
    def set_dB_range(self, tracklist: TrackList, dB_range: float) -> None:
        self.dB_range = dB_range
        self._update_spec_imgs(tracklist, True)

    def set_colormap_length(self, tracklist: TrackList, colormap_length: int) -> None:
        self.colormap_length = colormap_length
        self._update_spec_imgs(tracklist, True)

    # -- mod.rs:133-135
    def get_spectrogram(self, id_ch: IdCh) -> Optional[np.ndarray]:
        if id_ch not in self._img_keys:
            return None
        return self.ctx.img_read(*id_ch)

    def get_spec(self, id_ch: IdCh) -> Optional[np.ndarray]:
        """The private `specs` map of the reference (T, B) f32 dB -- exposed for parity tests."""
        if id_ch not in self._spec_keys:
            return None
        return self.ctx.spec_read(*id_ch)

    @property
    def spec_imgs(self) -> Dict[IdCh, np.ndarray]:
        return {k: self.ctx.img_read(*k) for k in sorted(self._img_keys)}

    # -- mod.rs:137-164
    def _update_specs(self, tracklist: TrackList, id_ch_tuples: Sequence[IdCh]) -> None:
        tracks = [dict(pcm=tracklist.channel(i, ch), id=i, ch=ch, sr=tracklist.sr(i)) for i, ch in id_ch_tuples]
        self.ctx.spec_batch(tracks, self.setting)
        self._spec_keys.update(id_ch_tuples)

    # -- mod.rs:168-230
    def _update_spec_imgs(self, tracklist: TrackList, force_update_all: bool) -> set:
        # The ONE collective of an update (when a communicator is attached): every rank makes it, whatever its own
        # track list needs afterwards; the quantise step below then uses the reduced range and no collective, so
        # ranks whose `ids_need_update` differ cannot fall out of step.
        mn, mx = self.ctx.minmax_global(self.dB_range)
        need_update_all = force_update_all
        if self.max_dB != mx:
            self.max_dB = mx
            need_update_all = True
        if self.min_dB != mn:
            self.min_dB = mn
            need_update_all = True
        max_sr = tracklist.max_sr()
        if self.max_sr != max_sr:
            self.max_sr = max_sr
            need_update_all = True
        if need_update_all:
            self._no_spec_img_ids.clear()
            ids_need_update = tracklist.all_id_set()
        else:
            ids_need_update = set(self._no_spec_img_ids)
            self._no_spec_img_ids.clear()
        if ids_need_update:
            if need_update_all:
                self._img_keys.clear()
            self.ctx.update_spec_imgs_range((self.min_dB, self.max_dB), self.colormap_length, self.max_sr,
                                            None if need_update_all else sorted(ids_need_update))
            self._img_keys.update(k for k in self._spec_keys if k[0] in ids_need_update)
        return ids_need_update
