"""N > 1 path on CPU: world_size-2 `gloo` processes run the host-side logic of the sharded job
(SURVEY.md section 8e): the deterministic work split (thesia_b200/sharding.py), the single exchange step --
one MAX all-reduce of {max, -min} (mod.rs:169-178 across ranks) -- and the clamp rules (mod.rs:179-180).
The per-shard spectrogram values come from the oracle here (there is no GPU in this tier); the GPU twin of this
test is tests/test_gpu_multi.py, which runs the same flow through thb_comm_init / NCCL.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import thesia_b200 as thb  # noqa: E402
from thesia_b200 import sharding  # noqa: E402
from thesia_b200.analysis import SpecSetting, FreqScale  # noqa: E402
from thesia_b200.synth import LOUD, ZERO_GAP, synth_pcm  # noqa: E402

WORLD = 2
SR = 48000
SETTING = SpecSetting(2048 / 48.0, 4, 1, FreqScale.Mel, 128)
# (id, ch, n_samples, flags): one long file that must be split by frame range, a few short ones
CHANNELS = [(0, 0, 150000, 0), (1, 0, 20000, LOUD), (1, 1, 20000, LOUD), (2, 0, 30001, ZERO_GAP), (3, 0, 9000, 0)]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _wav(idx):
    i, ch, n, fl = CHANNELS[idx]
    return synth_pcm(n, SR, i, ch, fl)


def _plan():
    return sharding.plan([(i, ch, SR, n) for (i, ch, n, _) in CHANNELS], SETTING.calc_framing_params, WORLD)


def _worker(rank: int, port: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist
    from oracle import orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        hop, win, n_fft = SETTING.calc_framing_params(SR)
        an = orc.Analyzer(SR, SETTING.win_ms, SETTING.t_overlap, SETTING.f_overlap, orc.MEL, SETTING.n_mel)
        units = _plan()[rank]
        local_max, local_nmin = -np.inf, -np.inf
        pieces = {}
        for u in units:
            idx = next(k for k, c in enumerate(CHANNELS) if (c[0], c[1]) == (u.id, u.ch))
            wav = _wav(idx)
            # the rank is handed only the slice its frames touch; the oracle works on whole files, so the
            # check is that the slice rule covers every sample the frame range reads (reflection included)
            lo, hi = sharding.needed_samples(u.frame_begin, u.frame_count, win, hop, u.full_len)
            assert (lo, hi) == (u.pcm_lo, u.pcm_hi)
            spec = an.calc_spec(wav, n_threads=2)[u.frame_begin:u.frame_begin + u.frame_count]
            pieces[(u.id, u.ch, u.frame_begin)] = spec
            mn, mx = orc.find_min_max(spec)
            local_max, local_nmin = max(local_max, mx), max(local_nmin, -mn)
        # the path's one collective: MAX over ranks of {max, -min}
        t = torch.tensor([local_max, local_nmin], dtype=torch.float32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mn_db, mx_db = orc.clamp_minmax(-float(t[1]), float(t[0]), 100.0)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), range=np.array([mn_db, mx_db], np.float32),
                 keys=np.array(list(pieces.keys()), np.int64).reshape(-1, 3),
                 **{f"p{k}": v for k, v in enumerate(pieces.values())})
    finally:
        dist.destroy_process_group()


def test_plan_is_deterministic_and_complete():
    hop, win, n_fft = SETTING.calc_framing_params(SR)
    ranks = _plan()
    assert ranks == _plan()
    assert len(ranks) == WORLD and all(ranks)
    # every frame of every channel is computed exactly once
    for (i, ch, n, _) in CHANNELS:
        spans = sorted((u.frame_begin, u.frame_count) for r in ranks for u in r if (u.id, u.ch) == (i, ch))
        pos = 0
        for b, c in spans:
            assert b == pos
            pos += c
        assert pos == sharding.n_frames(n, win, hop)
    # the long file was split and the load is balanced to within one unit
    assert sum(1 for r in ranks for u in r if u.id == 0) >= 2
    loads = [sum(u.cost for u in r) for r in ranks]
    assert max(loads) - min(loads) <= max(u.cost for r in ranks for u in r)


def test_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import orc

    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, port, str(tmp_path))) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    # single-process truth: whole files, one global min/max
    an = orc.Analyzer(SR, SETTING.win_ms, SETTING.t_overlap, SETTING.f_overlap, orc.MEL, SETTING.n_mel)
    whole = {(c[0], c[1]): an.calc_spec(_wav(k), n_threads=2) for k, c in enumerate(CHANNELS)}
    gmn = min(orc.find_min_max(s)[0] for s in whole.values())
    gmx = max(orc.find_min_max(s)[1] for s in whole.values())
    want = orc.clamp_minmax(gmn, gmx, 100.0)
    seen = {k: np.zeros(v.shape[0], bool) for k, v in whole.items()}
    for r in range(WORLD):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert tuple(z["range"]) == tuple(np.float32(want))  # every rank ends with the same global range
        for k, (i, ch, fb) in enumerate(z["keys"]):
            piece = z[f"p{k}"]
            assert np.array_equal(piece, whole[(i, ch)][fb:fb + piece.shape[0]], equal_nan=True)
            assert not seen[(i, ch)][fb:fb + piece.shape[0]].any()
            seen[(i, ch)][fb:fb + piece.shape[0]] = True
    assert all(v.all() for v in seen.values())
    assert want[1] == 0.0  # the LOUD track drives max_dB into the min(max, 0) clamp


def test_random_plans_cover_every_frame_and_carry_every_sample(orc):
    """300 random jobs (channel counts, lengths, settings, world sizes 1..8): every frame is computed exactly once, a unit's
    PCM slice holds every sample its frames read -- checked against the oracle's own reflect index, sample by sample at the
    unit's first and last frame -- the load is balanced to within a frame pair, and no rank holds two parts of one channel (the device
    store is keyed by (id, ch)).  Also the BASELINE shapes: C3 deals 16
    channels to each of 8 ranks, C2's single 1-hour file is cut into 8 frame ranges."""
    rng = np.random.default_rng(77)
    for _ in range(300):
        world = int(rng.integers(1, 9))
        sr = int(rng.choice([8000, 16000, 44100, 48000, 96000]))
        s = thb.SpecSetting(float(rng.choice([10.0, 40.0, 2048 / 48.0, 170.0])), int(rng.choice([1, 2, 4, 8, 16])), 1, thb.FreqScale.Mel)
        hop, win, _ = s.calc_framing_params(sr)
        chans = [(i, c, sr, int(rng.integers(2, 3_000_000))) for i in range(int(rng.integers(1, 7))) for c in range(int(rng.integers(1, 3)))]
        ranks = sharding.plan(chans, lambda _sr: s.calc_framing_params(_sr), world)
        assert len(ranks) == world
        for (i, ch, _, n) in chans:
            units = sorted((u for r in ranks for u in r if (u.id, u.ch) == (i, ch)), key=lambda u: u.frame_begin)
            pos = 0
            for u in units:
                assert u.frame_begin == pos and u.full_len == n and 0 <= u.pcm_lo <= u.pcm_hi <= n
                pos += u.frame_count
                for f in (u.frame_begin, u.frame_begin + u.frame_count - 1) if u.frame_count else ():
                    taps = (f * hop - win // 2, f * hop - win // 2 + win - 1)
                    idx = [orc.reflect_index(t, n) for t in taps] + [orc.reflect_index(t, n) for t in range(taps[0], min(taps[0] + 3, taps[1] + 1))]
                    assert all(u.pcm_lo <= k < u.pcm_hi for k in idx), (n, win, hop, f, idx, u)
            assert pos == sharding.n_frames(n, win, hop) == thb.n_frames(n, win, hop)
        loads = [sum(u.cost for u in r) for r in ranks]
        assert max(loads) - min(loads) <= 4  # contiguous spans: equal shares up to the even-frame rounding of a cut
        for r in ranks:  # the device store is keyed by (id, ch): a rank never holds two parts of one file
            keys = [(u.id, u.ch) for u in r]
            assert len(keys) == len(set(keys)), r
    c3 = thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128)
    ranks = sharding.plan([(t, c, 48000, 28_800_000) for t in range(64) for c in range(2)], lambda sr: c3.calc_framing_params(sr), 8)
    assert [len(r) for r in ranks] == [16] * 8 and all(u.frame_count == 56251 for r in ranks for u in r)
    c2 = thb.SpecSetting(2048 / 48.0, 8, 1, thb.FreqScale.Mel, 128)
    ranks = sharding.plan([(0, 0, 48000, 172_800_000)], lambda sr: c2.calc_framing_params(sr), 8)
    assert [len(r) for r in ranks] == [1] * 8 and sum(u.frame_count for r in ranks for u in r) == 675001
    assert all(u.pcm_hi - u.pcm_lo <= 172_800_000 // 8 + 2 * 2048 for r in ranks for u in r)
