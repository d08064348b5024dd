#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its wall time (host f32 PCM -> images on the host)."""
import ctypes as C
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import thesia_b200 as thb
from thesia_b200 import _lib

sr, n, nch, nmel = 48000, 48000 * 600, 128, 128
_st = torch.cuda.Stream(); torch.cuda.set_stream(_st)
ctx = thb.Context(0, _st.cuda_stream)
s = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, nmel)
hop, win, _ = s.calc_framing_params(sr)
T = thb.n_frames(n, win, hop)
p = C.c_void_p()
_lib.check(_lib.lib().thb_host_alloc(nch * n * 4, C.byref(p)))
host = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(nch, n))
d = torch.empty(n, dtype=torch.float32, device="cuda")
for c in range(nch):
    ctx.synth_pcm(d, sr, c // 2, c % 2, 0)
    ctx.synchronize()
    torch.from_numpy(host[c]).copy_(d)
q = C.c_void_p()
_lib.check(_lib.lib().thb_host_alloc(nch * nmel * T * 2, C.byref(q)))
tracks = ctx.prepare_tracks([dict(pcm=host[c], id=c // 2, ch=c % 2, sr=sr) for c in range(nch)])
keys = [(c // 2, c % 2) for c in range(nch)]
addrs = [q.value + 2 * c * nmel * T for c in range(nch)]
caps = [nmel * T] * nch
ctx.profile_enable(True)
for it in range(4):
    torch.cuda.synchronize()
    ctx.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    ctx.spec_batch(tracks, s)
    e1.record()
    t1 = time.perf_counter()
    ctx.update_spec_imgs(100.0, 258, sr)
    t2 = time.perf_counter()
    ctx.img_read_batch_into(keys, addrs, caps)
    t3 = time.perf_counter()
    h2d, d2h = nch * n * 4, nch * nmel * T * 2
    print(f"spec_batch (H2D {h2d / 1e9:.2f} GB) {1e3 * (t1 - t0):.1f} ms = {h2d / (t1 - t0) / 1e9:.1f} GB/s | update_spec_imgs {1e3 * (t2 - t1):.2f} ms | "
          f"img_read (D2H {d2h / 1e9:.2f} GB) {1e3 * (t3 - t2):.1f} ms = {d2h / (t3 - t2) / 1e9:.1f} GB/s | total {1e3 * (t3 - t0):.1f} ms")
    print(f"   GPU: all STFT kernels done {e0.elapsed_time(e1):.1f} ms after the start; kernel sums: "
          + ", ".join(f"{k} {ctx.profile_get(k)[0]:.2f} ms / {ctx.profile_get(k)[1]} launches" for k in ("stft_mel_db", "stft_mel_db_edges", "spec_to_img", "minmax_reduce")))
