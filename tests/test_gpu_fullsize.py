"""Full-size property tests (BASELINE.json sizes, where the oracle would take minutes): the CUDA path is checked
through size-independent properties of the domain, and against the oracle on randomly drawn frames / bins only.

  C2   one 1-hour 48 kHz mono track, win 2048 hop 256, mel 128 (675 001 frames)
       - frame locality: 96 random frames equal the oracle (f64 truth) computed on the PCM around them alone
       - shard invariance: three frame ranges computed from PCM slices equal the same rows of the whole-file run, bit for bit
       - time-shift: the file advanced by exactly 64 hops reproduces the interior rows 64 frames earlier, bit for bit
       - gain linearity: PCM * 0.5 (exact) lowers every finite dB value by 20 log10(2) within f32 rounding of the log
       - quantiser idempotence: update_spec_imgs twice gives the same image; its range clamp follows mod.rs:179-180
  C5   min / max envelope of that track at levels 9 and 15: bit-exact against numpy over the whole file, and the
       pyramid property (a level-15 bin is the min / max of its 64 level-9 bins)
  C4   one 96 kHz track, win 16384 hop 1024 (large-FFT path): 24 random frames against the oracle + shard invariance
  thread safety: thb_waveform_tile from 8 host threads at once (lib.rs:343-389 serves tiles concurrently)
"""
import math
import threading

import numpy as np
import pytest

import thesia_b200 as thb
from thesia_b200.sharding import split_frames

import os

pytestmark = pytest.mark.gpu
# under compute-sanitizer (tools/gpu_sanitize.sh) the same tests run at a sixth of the size: the tool slows kernels ~50x
SMALL = os.environ.get("THB_TEST_SMALL") == "1"

DB_TOL = 1e-3
FLOOR = 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = thb.Context(0)
    yield c
    c.close()


def _device_track(ctx, n, sr, track, flags=0):
    import torch
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    ctx.synth_pcm(d, sr, track, 0, flags)
    ctx.synchronize()
    return d


def _check_frames_against_oracle(orc, ctx, pcm_dev, n, sr, setting, frames, spec_rows, scale):
    """spec_rows[i] is the GPU's row of frame frames[i]; the oracle sees only the samples that frame touches
    (+ reflection at the true file ends, which the slice bounds reproduce by taking the file ends themselves)."""
    hop, win, n_fft = setting.calc_framing_params(sr)
    an = orc.Analyzer(sr, setting.win_ms, setting.t_overlap, setting.f_overlap, scale, setting.n_mel)
    worst = 0.0
    for f, row in zip(frames, spec_rows):
        lo = f * hop - win // 2
        # an interior frame f of the file is frame k of the slice [lo - k*hop, ...): choose k = 8 so that no
        # reflection of the SLICE touches it
        k = 8
        s_lo = lo - k * hop
        s_hi = lo + win + k * hop
        if s_lo < 0 or s_hi > n:
            continue
        seg = pcm_dev[s_lo:s_hi].cpu().numpy()
        truth = an.calc_spec_truth(seg)[k + (win // 2) // hop]  # slice frame whose taps start at s_lo + k*hop + win/2 - win/2
        got = row.astype(np.float64)
        neg = np.isneginf(truth)
        assert np.array_equal(np.isneginf(got), neg)
        peak = np.where(neg, -np.inf, truth).max()
        above = (~neg) & (truth > peak + 10.0 * math.log10(FLOOR))
        if above.any():
            worst = max(worst, float(np.abs(got[above] - truth[above]).max()))
    assert worst <= DB_TOL, worst
    return worst


def test_c2_full_size_properties(ctx, orc):
    import torch
    sr, n = 48000, 48000 * (600 if SMALL else 3600)
    setting = thb.SpecSetting(2048 / 48.0, 8, 1, thb.FreqScale.Mel, 128)
    hop, win, _ = setting.calc_framing_params(sr)
    pcm = _device_track(ctx, n, sr, 7)
    ctx.spec_batch([dict(pcm=pcm, id=1, ch=0, sr=sr)], setting)
    whole = ctx.spec_read(1, 0)
    T = thb.n_frames(n, win, hop)
    assert whole.shape == (T, 128) and (SMALL or T == 675001)
    assert np.isfinite(whole).all()
    # frame locality against the oracle
    rng = np.random.default_rng(42)
    frames = sorted(rng.integers(64, T - 64, 96).tolist())
    # the slice frame index: slice starts at lo - 8 hops where lo = f*hop - win/2; slice frame j has taps starting at
    # j*hop - win/2 (slice coordinates) -> the file frame is slice frame 8 + (win/2)/hop
    _check_frames_against_oracle(orc, ctx, pcm, n, sr, setting, frames, [whole[f] for f in frames], orc.MEL)
    # shard invariance at full size (frame ranges from PCM slices + halo)
    units = split_frames(1, 0, sr, n, win, hop, 5)
    for u in (units[0], units[2], units[4]):
        tr = dict(pcm=pcm[u.pcm_lo:u.pcm_hi], id=50, ch=0, sr=sr, full_len=n, pcm_offset=u.pcm_lo,
                  frame_begin=u.frame_begin, frame_count=u.frame_count)
        ctx.spec_batch([tr], setting)
        part = ctx.spec_read(50, 0)
        assert np.array_equal(part, whole[u.frame_begin:u.frame_begin + u.frame_count])
    ctx.release(50, 0)
    # time shift by 64 hops: interior rows move up by 64, bit for bit
    sh = 64 * hop
    ctx.spec_batch([dict(pcm=pcm[sh:], id=51, ch=0, sr=sr)], setting)
    shifted = ctx.spec_read(51, 0)
    edge = win // hop  # rows whose taps reach the (different) file starts / ends
    assert np.array_equal(shifted[edge:-edge], whole[64 + edge:64 + shifted.shape[0] - edge])
    ctx.release(51, 0)
    # exact gain of 0.5 -> -20 log10(2) dB everywhere (the FFT is linear in an exact power of two)
    half = pcm * 0.5
    ctx.spec_batch([dict(pcm=half, id=52, ch=0, sr=sr)], setting)
    low = ctx.spec_read(52, 0)
    assert np.abs((whole - low) - 20.0 * math.log10(2.0)).max() <= 2e-5 * np.abs(whole).max()
    del half, low
    ctx.release(52, 0)
    # quantiser: range rule (mod.rs:179-180) and idempotence
    mn, mx = ctx.update_spec_imgs(100.0, 258)
    want_mx = min(np.float32(whole.max()), np.float32(0.0))
    want_mn = max(np.float32(whole.min()), np.float32(want_mx - np.float32(100.0)))
    assert (np.float32(mn), np.float32(mx)) == (want_mn, want_mx)
    img1 = ctx.img_read(1, 0)
    assert img1.shape == (128, T)
    ctx.update_spec_imgs(100.0, 258)
    assert np.array_equal(ctx.img_read(1, 0), img1)
    assert np.array_equal(img1, orc.spec_to_img(whole, (0, 128), (mn, mx), 258))   # the quantiser itself is cheap on the CPU
    ctx.release(1, 0)

    # ---- C5 on the same track: envelope min / max bit-exact over the whole file, and the pyramid property ----
    host = pcm.cpu().numpy()
    lv = {}
    for level in (9, 15):
        raw = ctx.waveform_level(pcm, 3, level)
        spb = 1 << level
        bins = -(-n // spb)
        tiles = -(-bins // 1024)
        vals = []
        off = 0
        for t in range(tiles):
            nb = min(1024, bins - t * 1024)
            hdr = np.frombuffer(raw, np.uint32, 6, off)
            assert hdr[2] == nb and hdr[3] == spb and hdr[4] == t
            vals.append(np.frombuffer(raw, np.float32, nb * 3, off + 24).reshape(nb, 3))
            off += 24 + 12 * nb
        assert off == len(raw)
        v = np.concatenate(vals)
        full = (n // spb) * spb
        blk = host[:full].reshape(-1, spb)
        assert np.array_equal(v[:n // spb, 0], blk.min(axis=1)) and np.array_equal(v[:n // spb, 1], blk.max(axis=1))
        if full < n:
            assert v[-1, 0] == host[full:].min() and v[-1, 1] == host[full:].max()
        mean = blk.mean(axis=1, dtype=np.float64)
        assert np.abs(v[:n // spb, 2] - mean).max() <= 1e-6
        lv[level] = v
    k = lv[15].shape[0] - 1          # whole level-15 bins
    g = lv[9][:k * 64].reshape(k, 64, 3)
    assert np.array_equal(lv[15][:k, 0], g[:, :, 0].min(axis=1)) and np.array_equal(lv[15][:k, 1], g[:, :, 1].max(axis=1))
    del pcm
    torch.cuda.empty_cache()


@pytest.mark.parametrize("scale_name", ["linear", "mel"])
def test_c4_large_fft_full_track(ctx, orc, scale_name):
    import torch
    sr, n = 96000, 96000 * (100 if SMALL else 600)       # 10 min of a C4 track: 56 251 frames of 8193 bins = 1.8 GB of f32 (linear)
    scale = thb.FreqScale.Linear if scale_name == "linear" else thb.FreqScale.Mel
    setting = thb.SpecSetting(16384 / 96.0, 16, 1, scale, 0)
    hop, win, _ = setting.calc_framing_params(sr)
    pcm = _device_track(ctx, n, sr, 3)
    ctx.spec_batch([dict(pcm=pcm, id=2, ch=0, sr=sr)], setting)
    T = thb.n_frames(n, win, hop)
    rng = np.random.default_rng(4)
    frames = sorted(rng.integers(32, T - 32, 24).tolist())
    units = split_frames(2, 0, sr, n, win, hop, 3)
    u = units[1]
    ctx.spec_batch([dict(pcm=pcm[u.pcm_lo:u.pcm_hi], id=60, ch=0, sr=sr, full_len=n, pcm_offset=u.pcm_lo,
                         frame_begin=u.frame_begin, frame_count=u.frame_count)], setting)
    whole = ctx.spec_read(2, 0)
    assert whole.shape[0] == T and np.array_equal(ctx.spec_read(60, 0), whole[u.frame_begin:u.frame_begin + u.frame_count])
    _check_frames_against_oracle(orc, ctx, pcm, n, sr, setting, frames, [whole[f] for f in frames],
                                 orc.LINEAR if scale_name == "linear" else orc.MEL)
    ctx.release(2, 0)
    ctx.release(60, 0)
    del pcm, whole
    torch.cuda.empty_cache()


def test_waveform_tile_is_thread_safe(ctx, orc):
    """get_waveform_tile runs on concurrent IPC threads under read locks (lib.rs:343-367)."""
    from thesia_b200.synth import synth_pcm
    wav = synth_pcm(600000, 48000, 5, 0, 0)
    want = {(lv, t): orc.encode_waveform_tile(wav, 9, lv, t) for lv in (0, 3, 7) for t in range(3)}
    errors = []

    def worker(seed):
        r = np.random.default_rng(seed)
        keys = list(want)
        for _ in range(40):
            lv, t = keys[int(r.integers(len(keys)))]
            got = ctx.waveform_tile(wav, 9, lv, t)
            w = want[(lv, t)]
            a = np.frombuffer(got, np.float32, offset=24).reshape(-1, 3)
            b = np.frombuffer(w, np.float32, offset=24).reshape(-1, 3)
            if got[:24] != w[:24] or not np.array_equal(a[:, :2], b[:, :2]) or np.abs(a[:, 2] - b[:, 2]).max() > 1e-6:
                errors.append((lv, t))
    ths = [threading.Thread(target=worker, args=(s,)) for s in range(8)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors[:5]
