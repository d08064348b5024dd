#!/usr/bin/env python
"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU, NCCL).

Every rank analyses its shard (thesia_b200.sharding.plan: one contiguous span of the frame line per rank, a file that
straddles a cut split by FRAME RANGE with the PCM slice its frames touch), then thb_update_spec_imgs reduces the global dB range with the library's one
ncclAllReduce(max) of {max, -min}.  Rank 0 also runs the whole job alone on its GPU (no communicator) and every
rank compares: identical global range, identical dB values and u16 images for each of its shards.
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import thesia_b200 as thb  # noqa: E402
from thesia_b200 import sharding  # noqa: E402
from thesia_b200.synth import LOUD, ZERO_GAP, synth_pcm  # noqa: E402

SR = 48000
CHANNELS = [(0, 0, 1500000, 0), (1, 0, 200000, LOUD), (1, 1, 200000, LOUD), (2, 0, 300001, ZERO_GAP), (3, 0, 90000, 0),
            (4, 0, 400000, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for setting in (thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Mel, 128), thb.SpecSetting(),
                    thb.SpecSetting(2048 / 48.0, 4, 1, thb.FreqScale.Linear)):
        hop, win, n_fft = setting.calc_framing_params(SR)
        wavs = {(i, ch): synth_pcm(n, SR, i, ch, fl) for (i, ch, n, fl) in CHANNELS}
        plan = sharding.plan([(i, ch, SR, n) for (i, ch, n, _) in CHANNELS], setting.calc_framing_params, world)
        ctx = thb.Context(local)
        uid = [thb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        px = ctx.comm_peer_exchange()   # the global range over NVLink peer memory (or NCCL when IPC is not available)
        units = plan[rank]
        tracks = []
        for k, u in enumerate(units):
            w = wavs[(u.id, u.ch)]
            # (the plan gives a rank at most one unit of any channel, so the file's own id is a valid key)
            tracks.append(dict(pcm=np.ascontiguousarray(w[u.pcm_lo:u.pcm_hi]), id=u.id, ch=u.ch, sr=SR,
                               full_len=u.full_len, pcm_offset=u.pcm_lo, frame_begin=u.frame_begin,
                               frame_count=u.frame_count))
        ctx.spec_batch(tracks, setting)
        rng = ctx.update_spec_imgs(100.0, 258, SR)  # the collective happens in here
        mine = [(u, ctx.spec_read(t["id"], t["ch"]), ctx.img_read(t["id"], t["ch"])) for u, t in zip(units, tracks)]
        ctx.close()

        # one GPU, whole files, no communicator
        solo = thb.Context(local)
        solo.spec_batch([dict(pcm=w, id=i, ch=ch, sr=SR) for (i, ch), w in wavs.items()], setting)
        rng1 = solo.update_spec_imgs(100.0, 258, SR)
        assert rng == rng1, (rank, rng, rng1)
        for u, spec, img in mine:
            whole = solo.spec_read(u.id, u.ch)
            wimg = solo.img_read(u.id, u.ch)
            sl = slice(u.frame_begin, u.frame_begin + u.frame_count)
            assert np.array_equal(spec, whole[sl], equal_nan=True), (rank, u)
            assert np.array_equal(img, wimg[:, sl]), (rank, u)
        solo.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, rng)
        assert all(g == rng for g in gathered)
        if rank == 0:
            print(f"setting win {win} hop {hop}: {world} ranks agree (range over {'peer memory' if px else 'NCCL'}), dB range {rng}, "
                  f"{sum(len(p) for p in plan)} units", flush=True)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
