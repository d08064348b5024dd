"""Round-2 seams of the C ABI, on the GPU:

  * thb_spec_batch rejects a batch that holds one (id, ch) twice, and a failed call leaves the store untouched
  * thb_update_spec_imgs queued without a host round trip (min/max pointers NULL) + thb_range_get give the same
    range and images as the waiting call; thb_update_spec_imgs_range (no collective) gives the same images
  * the tile readers run CONCURRENTLY (shared lock, one stream per call): wall clock of 8 threads against 1, and a
    tile call does not wait behind another thread's tile call
  * host PCM handed to thb_waveform_tile is cached on the device by (pointer, length, revision): a second tile of the
    same channel crosses PCIe only for granules it has not seen; a new revision drops the copy
  * spectrogram tiles from concurrent threads equal the oracle byte for byte
"""
import threading
import time

import numpy as np
import pytest

import thesia_b200 as thb
from thesia_b200 import _lib
from thesia_b200.synth import synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = thb.Context(0)
    yield c
    c.close()


def test_duplicate_id_ch_in_one_batch_is_rejected_and_store_untouched(ctx):
    ctx.release_all()
    s = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128)
    w = synth_pcm(60000, 48000, 1, 0, 0)
    ctx.spec_batch([dict(pcm=w, id=5, ch=0, sr=48000)], s)
    before = ctx.spec_read(5, 0)
    T = thb.n_frames(len(w), 2048, 512)
    half = T // 2 // 2 * 2
    parts = [dict(pcm=w, id=5, ch=0, sr=48000, full_len=len(w), pcm_offset=0, frame_begin=0, frame_count=half),
             dict(pcm=w, id=5, ch=0, sr=48000, full_len=len(w), pcm_offset=0, frame_begin=half, frame_count=T - half)]
    with pytest.raises(thb.ThbError) as e:
        ctx.spec_batch(parts, s)
    assert e.value.code == _lib.THB_ERR_INVALID and "twice" in str(e.value)
    # a failed call commits nothing: the retained spectrogram is the one from before, and no new key appeared
    assert np.array_equal(ctx.spec_read(5, 0), before, equal_nan=True)
    with pytest.raises(thb.ThbError):
        ctx.spec_batch([dict(pcm=w, id=77, ch=0, sr=48000), dict(pcm=np.zeros(1, np.float32), id=78, ch=0, sr=48000)], s)
    with pytest.raises(thb.ThbError) as e:
        ctx.spec_read(77, 0)
    assert e.value.code == _lib.THB_ERR_NOT_FOUND
    ctx.release_all()


def test_queued_update_and_range_variant_equal_the_waiting_call(ctx):
    ctx.release_all()
    s = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128)
    wavs = {(t, c): synth_pcm(96000 + 512 * t, 48000, t, c, 1 if t == 1 else 0) for t in range(3) for c in range(2)}
    tracks = [dict(pcm=w, id=t, ch=c, sr=48000) for (t, c), w in wavs.items()]
    ctx.spec_batch(tracks, s)
    rng = ctx.update_spec_imgs(80.0, 258, 48000)
    imgs = {k: ctx.img_read(*k) for k in wavs}
    # queued: no host round trip inside the call, the range is read afterwards
    for _ in range(3):
        ctx.spec_batch(tracks, s)
        assert ctx.update_spec_imgs(80.0, 258, 48000, wait=False) is None
    assert ctx.range_get() == rng
    assert all(np.array_equal(ctx.img_read(*k), imgs[k]) for k in wavs)
    # the quantise step alone, with the caller's range: same pixels; restricted to one id the others stay as they were
    ctx.update_spec_imgs_range((rng[0] + 3.0, rng[1]), 258, 48000)
    changed = {k: ctx.img_read(*k) for k in wavs}
    assert any(not np.array_equal(changed[k], imgs[k]) for k in wavs)
    ctx.update_spec_imgs_range(rng, 258, 48000, only_ids=[1])
    assert all(np.array_equal(ctx.img_read(t, c), imgs[(t, c)] if t == 1 else changed[(t, c)]) for (t, c) in wavs)
    ctx.update_spec_imgs_range(rng, 258, 48000)
    assert all(np.array_equal(ctx.img_read(*k), imgs[k]) for k in wavs)
    ctx.release_all()


def test_quantiser_descriptors_follow_the_store(ctx):
    """thb_update_spec_imgs keeps the quantiser's descriptors on the device between calls and uploads them only when
    they changed.  Walk the store through every kind of change -- same set again, a track released, a track recomputed
    with another length (new buffers), a track added, a subset, another max_sr (image height) -- and compare every image
    with what a fresh context produces for the same set."""
    s = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128)
    wav = {t: synth_pcm(70000 + 4096 * t, 48000, t, 0, 0) for t in range(4)}

    def fresh(ids, lens, max_sr=48000):
        c = thb.Context(0)
        c.spec_batch([dict(pcm=wav[t][:lens[t]], id=t, ch=0, sr=48000) for t in ids], s)
        rng = c.update_spec_imgs(90.0, 258, max_sr)
        out = {t: c.img_read(t, 0) for t in ids}
        c.close()
        return rng, out

    def same(ids, lens, max_sr=48000):
        rng = ctx.update_spec_imgs(90.0, 258, max_sr)
        want_rng, want = fresh(ids, lens, max_sr)
        assert rng == want_rng
        for t in ids:
            got = ctx.img_read(t, 0)
            assert got.shape == want[t].shape and np.array_equal(got, want[t]), t

    ctx.release_all()
    lens = {t: len(wav[t]) for t in wav}
    ctx.spec_batch([dict(pcm=wav[t], id=t, ch=0, sr=48000) for t in (0, 1, 2)], s)
    same((0, 1, 2), lens)
    same((0, 1, 2), lens)                       # unchanged: the cached descriptors
    ctx.release(1, 0)
    same((0, 2), lens)                          # one fewer
    lens[2] = 50000
    ctx.spec_batch([dict(pcm=wav[2][:50000], id=2, ch=0, sr=48000)], s)
    same((0, 2), lens)                          # same keys, other buffers and sizes
    ctx.spec_batch([dict(pcm=wav[3], id=3, ch=0, sr=48000)], s)
    same((0, 2, 3), lens)                       # one more
    same((0, 2, 3), lens, max_sr=24000)         # lower Nyquist limit: fewer image rows
    same((0, 2, 3), lens)
    ctx.release_all()


def test_many_queued_small_steps_equal_one_batch(ctx):
    """Sixty small steps queued back to back without a host round trip -- thb_spec_batch of ONE new short track (file
    edges: the scalar kernel on its side stream next to the packed one) + thb_update_spec_imgs (which re-quantises every
    retained track with the range so far; its small kernels are launched programmatically under each other's tails, its
    descriptors grow by one each step) -- must leave exactly the store that ONE batch of the sixty tracks + ONE update
    leaves: any kernel that started before what it depends on had finished would show up as a differing row."""
    ctx.release_all()
    s = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128)
    n = 60
    wavs = [synth_pcm(30000 + 1024 * (t % 7), 48000, t % 5, t % 2, 0) * np.float32(1.0 / (1 + t % 3)) for t in range(n)]
    for t in range(n):
        ctx.spec_batch([dict(pcm=wavs[t], id=t, ch=0, sr=48000)], s)
        assert ctx.update_spec_imgs(85.0, 258, 48000, wait=False) is None
    rng = ctx.range_get()
    got = [(ctx.spec_read(t, 0), ctx.img_read(t, 0)) for t in range(n)]
    ctx.release_all()
    ctx.spec_batch([dict(pcm=wavs[t], id=t, ch=0, sr=48000) for t in range(n)], s)
    assert ctx.update_spec_imgs(85.0, 258, 48000) == rng
    for t in range(n):
        assert np.array_equal(ctx.spec_read(t, 0), got[t][0], equal_nan=True), t
        assert np.array_equal(ctx.img_read(t, 0), got[t][1]), t
    ctx.release_all()


def _tile_ok(got, want):
    a = np.frombuffer(got, np.float32, offset=24).reshape(-1, 3)
    b = np.frombuffer(want, np.float32, offset=24).reshape(-1, 3)
    return got[:24] == want[:24] and np.array_equal(a[:, :2], b[:, :2]) and np.abs(a[:, 2] - b[:, 2]).max() <= 1e-6


def test_pcm_cache_keeps_channels_on_the_device(ctx, orc):
    ctx.pcm_cache_clear()
    n = 3_000_000
    wav = synth_pcm(n, 48000, 4, 0, 0)
    s0 = ctx.pcm_cache_stats()
    # level 9: 1024 * 512 samples per tile = 8 granules of 65 536 samples
    t0 = ctx.waveform_tile(wav, 7, 9, 0)
    s1 = ctx.pcm_cache_stats()
    assert s1["entries"] == 1 and s1["bytes"] == 4 * n and s1["misses"] == s0["misses"] + 1
    assert _tile_ok(t0, orc.encode_waveform_tile(wav, 7, 9, 0))
    # the same samples at a finer level: nothing crosses PCIe
    for lv, t in ((8, 0), (8, 1), (7, 3), (9, 0)):
        assert _tile_ok(ctx.waveform_tile(wav, 7, lv, t), orc.encode_waveform_tile(wav, 7, lv, t))
    s2 = ctx.pcm_cache_stats()
    assert s2["misses"] == s1["misses"] and s2["hits"] == s1["hits"] + 4
    # a coarser level reaches past what was uploaded: only the new granules are copied, the answer is the oracle's
    assert _tile_ok(ctx.waveform_tile(wav, 7, 12, 0), orc.encode_waveform_tile(wav, 7, 12, 0))
    assert ctx.pcm_cache_stats()["misses"] == s2["misses"] + 1
    assert _tile_ok(ctx.waveform_tile(wav, 7, 12, 0), orc.encode_waveform_tile(wav, 7, 12, 0))
    # the channel changes in place and comes with a new revision (as in the reference): the old copy is dropped
    wav[1000:2000] *= np.float32(0.5)
    assert _tile_ok(ctx.waveform_tile(wav, 8, 9, 0), orc.encode_waveform_tile(wav, 8, 9, 0))
    s3 = ctx.pcm_cache_stats()
    assert s3["entries"] == 1 and s3["bytes"] == 4 * n
    ctx.pcm_cache_clear()
    assert ctx.pcm_cache_stats()["entries"] == 0


def test_tile_readers_run_concurrently(ctx, orc):
    """lib.rs:343-389: tiles are served from concurrent IPC threads under read locks.  Eight Python threads asking for
    level-13 tiles of eight device-resident channels: every answer must equal the oracle's, and the calls must not be
    slower than one after the other.  (The wall-clock PROOF of overlap is tests/cpp/test_host_mirror.cpp
    test_tile_readers_overlap, with std::thread: here the interpreter lock is handed over twice per 50 us call and
    hides what the library does.)"""
    import torch
    n = 1024 * 8192
    chans = []
    for c in range(8):
        d = torch.empty(n, dtype=torch.float32, device="cuda")
        ctx.synth_pcm(d, 48000, c, 0, 0)
        chans.append(d)
    ctx.synchronize()
    want = [orc.encode_waveform_tile(d.cpu().numpy(), 3, 13, 0) for d in chans]
    reps = 40
    bad = []

    def worker(c):
        for _ in range(reps):
            if not _tile_ok(ctx.waveform_tile(chans[c], 3, 13, 0), want[c]):
                bad.append(c)

    worker(0)  # warm-up: lanes, kernel load
    t0 = time.perf_counter()
    worker(0)
    t_one = time.perf_counter() - t0
    ths = [threading.Thread(target=worker, args=(c,)) for c in range(8)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    t_eight = time.perf_counter() - t0
    assert not bad
    import os
    if not os.environ.get("THB_NO_TIMING"):
        assert t_eight < 12.0 * t_one, (t_one, t_eight)
    print(f"one thread {1e6 * t_one / reps:.0f} us/tile; eight threads {1e6 * t_eight / (8 * reps):.0f} us/tile "
          f"({8 * t_one / t_eight:.1f}x overlap)")


def test_spectrogram_tiles_from_concurrent_threads(ctx, orc):
    ctx.release_all()
    s = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 0)
    for t in range(2):
        ctx.spec_batch([dict(pcm=synth_pcm(1_200_000, 48000, t, 0, 0), id=t, ch=0, sr=48000)], s)
    ctx.update_spec_imgs(100.0, 258)
    imgs = {t: ctx.img_read(t, 0) for t in range(2)}
    cm = bytes((i * 7 + j * 31) & 255 for i in range(258) for j in range(4))
    reqs = [(t, lx, ly, tx) for t in range(2) for (lx, ly) in ((0, 0), (1, 0), (2, 1)) for tx in range(2)]
    want = {r: orc.encode_spectrogram_tile(imgs[r[0]], cm, 5, r[1], r[2], r[3], 0) for r in reqs}
    bad = []

    def worker(seed):
        rr = np.random.default_rng(seed)
        for _ in range(12):
            r = reqs[int(rr.integers(len(reqs)))]
            if ctx.spectrogram_tile(r[0], 0, cm, 5, r[1], r[2], r[3], 0) != want[r]:
                bad.append(r)

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not bad, bad[:4]
    ctx.release_all()


def test_every_baseline_configuration_runs_on_its_fast_kernel(ctx):
    """A plan whose tables outgrow a kernel's shared memory silently falls back to a slower kernel (it happened once,
    when a second mel schedule was stored next to the first: 6x).  Pin the family and the mel schedule of BASELINE's
    configurations and of the reference's default setting at its fixture rates."""
    F = thb.FreqScale
    PAIR, BIG, WARP = 2, 3, 4
    cases = [
        (thb.SpecSetting(2048 / 48.0, 4, 1, F.Linear), 48000, PAIR, 0),          # C1
        (thb.SpecSetting(2048 / 48.0, 8, 1, F.Mel, 128), 48000, PAIR, 1),        # C2: wide bands -> bin-major
        (thb.SpecSetting(2048 / 48.0, 8, 1, F.Mel, 0), 48000, PAIR, None),       # C2, default bank
        (thb.SpecSetting(2048 / 48.0, 4, 1, F.Mel, 128), 48000, PAIR, 1),        # C3
        (thb.SpecSetting(), 48000, PAIR, None),                                  # the reference's default setting
        (thb.SpecSetting(), 44100, PAIR, None),
        (thb.SpecSetting(), 24000, WARP, None),
        (thb.SpecSetting(), 22050, WARP, None),
        (thb.SpecSetting(), 16000, WARP, None),
        (thb.SpecSetting(), 8000, WARP, None),
        (thb.SpecSetting(), 96000, BIG, None),
        (thb.SpecSetting(16384 / 96.0, 16, 1, F.Linear), 96000, BIG, 0),         # C4
        (thb.SpecSetting(16384 / 96.0, 16, 1, F.Mel, 0), 96000, BIG, 1),
    ]
    for s, sr, fam, sch in cases:
        got = ctx.plan_kernel(s, sr)
        assert got[0] == fam, (sr, s, got)
        if sch is not None:
            assert got[1] == sch, (sr, s, got)
        elif s.freq_scale == F.Mel:
            assert got[1] in (1, 2)
