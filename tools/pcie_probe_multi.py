#!/usr/bin/env python
"""What the box's host side delivers to N GPUs AT ONCE, with no library in the way: plain pinned
cudaMemcpyAsync from every rank at the same moment, in the variants that could lift the e2e leg of bench.py if the
limit were in how the library feeds the GPUs rather than in the box:

  plain        cudaHostAlloc(default), one stream
  bound        the same after binding the rank to its own slice of the cores (the VMs report one NUMA node)
  wc           write-combined pinned memory for the H2D source
  two-streams  the buffer split over two streams (two DMA queues per direction)

Run under torchrun:  python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_multi.py
Rank 0 prints one JSON object per variant: per-rank minimum and the aggregate, H2D and D2H, and both at once."""
import ctypes as C
import json
import os
import time

import torch
import torch.distributed as dist

GB = 1 << 30


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    n = 2 * GB
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        lo, tot = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return float(lo.item()), float(tot.item())

    def host(flags):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
        C.memset(p, 1, n)   # first touch by this (possibly bound) thread
        return p

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        barrier()
        return dt

    H2D, D2H = 1, 2
    results = []

    def variant(name, flags=0, streams=1):
        h, h2 = host(flags), host(0)

        def cp(dst, src, nbytes, kind, st):
            assert rt.cudaMemcpyAsync(dst, src, nbytes, kind, C.c_void_p(st.cuda_stream)) == 0

        def h2d():
            if streams == 1:
                cp(d.data_ptr(), h, n, H2D, s1)
            else:
                cp(d.data_ptr(), h, n // 2, H2D, s1)
                cp(d.data_ptr() + n // 2, h.value + n // 2, n // 2, H2D, s2)

        def d2h():
            if streams == 1:
                cp(h2, d2.data_ptr(), n, D2H, s1)
            else:
                cp(h2, d2.data_ptr(), n // 2, D2H, s1)
                cp(h2.value + n // 2, d2.data_ptr() + n // 2, n // 2, D2H, s2)

        def both():
            cp(d.data_ptr(), h, n, H2D, s1)
            cp(h2, d2.data_ptr(), n, D2H, s2)

        rec = {"variant": name, "ranks": world, "bytes_per_copy": n}
        for key, fn in (("h2d", h2d), ("d2h", d2h), ("both_each_way", both)):
            lo, tot = reduce(n / timed(fn) / 1e9)
            rec[key + "_gbs_per_rank_min"], rec[key + "_gbs_aggregate"] = round(lo, 2), round(tot, 2)
        rt.cudaFreeHost(h)
        rt.cudaFreeHost(h2)
        results.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)

    variant("plain")
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // world)
    os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]) or set(cpus))
    variant("bound to cores %d-%d of %d" % (local * per, (local + 1) * per - 1, len(cpus)))
    variant("bound + write-combined H2D source", flags=0x04)
    variant("bound + two streams per direction", streams=2)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
