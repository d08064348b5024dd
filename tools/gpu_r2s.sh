#!/bin/bash
mkdir -p gpurun_out
{
echo "== default"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k spectrogram_tile_parity --tb=short 2>&1 | tail -25
echo "== THB_TILE_CM=l1"; THB_TILE_CM=l1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k spectrogram_tile_parity --tb=short 2>&1 | tail -12
echo "== diff"; timeout 300 python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import thesia_b200 as thb
from oracle import orc
ctx = thb.Context(0)
rng = np.random.default_rng(9)
noisy = rng.integers(0, 65536, (128, 3001), dtype=np.uint16)
smooth = (np.add.outer(np.arange(347) * 90.0, np.arange(5000) * 7.0) % 65536).astype(np.uint16)
for name, img, reqs, ncol in (("long", noisy, [(0, 0, 0, 0), (0, 0, 5, 0), (1, 0, 0, 0)], 2000), ("long-single", noisy, [(0, 0, 0, 0)], 2000), ("1025", noisy, [(0, 0, 0, 0)], 1025), ("1024", noisy, [(0, 0, 0, 0)], 1024),
                              ("mixed", smooth, [(0, 0, 3, 0), (2, 0, 0, 0), (0, 0, 4, 0), (0, 1, 1, 0)], 258)):
    cm = np.random.default_rng(3).integers(0, 256, ncol * 4, dtype=np.uint8).tobytes()
    ctx.spec_put(900, 0, 48000, thb.FreqScale.Mel, np.zeros((img.shape[1], 1), np.float32))
    ctx.img_put(900, 0, img)
    got = ctx.spectrogram_tiles(cm, 4, [(900, 0) + r for r in reqs])
    for r, g in zip(reqs, got):
        want = orc.encode_spectrogram_tile(img, cm, 4, *r)
        a, b = np.frombuffer(g, np.uint8), np.frombuffer(want, np.uint8)
        if len(a) != len(b):
            print(name, r, "LENGTH", len(a), len(b)); continue
        bad = np.nonzero(a != b)[0]
        print(name, r, "ok" if not len(bad) else f"{len(bad)} bytes differ, first at {bad[:6]}, header {a[:40].tolist()} vs {b[:40].tolist()}")
PY
} > gpurun_out/r2s.log 2>&1
tail -60 gpurun_out/r2s.log
