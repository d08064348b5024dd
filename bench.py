#!/usr/bin/env python
"""bench.py -- STFT+mel+dB throughput (audio-hours/s) of the analysis hot path on N B200s.

Workload (BASELINE.json configs[2], "C3"): 64 synthetic stereo tracks x 10 min at 48 kHz PER GPU
(128 channels x 28.8 M samples = 14.75 GB f32), win 2048 / hop 512 Hann, 128-band mel, dB, global
min/max (one 2-float NCCL max all-reduce across ranks) and u16 images.  A step = TrackManager's
update_specs + update_spec_imgs over that batch = thb_spec_batch + thb_update_spec_imgs.

  value : whole-job audio-hours/s with the PCM already resident in HBM (torch CUDA events on the
          stream the kernels run on, max over ranks).
  e2e   : the same metric through the C ABI with HOST buffers: pinned PCM is copied host->device and
          the u16 images device->host inside the timed region, every step.
  roofline / cpu_baseline / clocks / gpu_launches: see the task contract; DESIGN.md "Measurement".
  --impl reference: the CPU restatement of the reference (oracle/) on all host threads.

Launch: python bench.py --gpus 1            or
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SR = 48000
N_TRACKS = 64
N_CH = 2
SECONDS = 600
WIN_MS = 2048 / 48.0
T_OVERLAP = 4
N_MEL = 128
DB_RANGE = 100.0
CMAP_LEN = 258
CPU_SAMPLE = (128, 120)  # channels x seconds of the CPU arms' bounded sample: every channel of the batch, 2 of its 10 minutes
METRIC = "STFT+mel+dB throughput"
UNIT = "audio-hours/s"


def workload_name(scale: float) -> str:
    s = "" if scale == 1.0 else f" (scaled x{scale:g})"
    return (f"C3: {N_TRACKS} stereo tracks x {SECONDS // 60} min @48 kHz per GPU, win 2048 hop 512 Hann, "
            f"mel {N_MEL} dB, global min/max + u16 images{s}")


def track_flags(track: int) -> int:
    # one track > 0 dBFS (max_dB clamp), one with a 1 s zero gap (-inf frames)  
    return (1 if track == 1 else 0) | (2 if track == 2 else 0)


# -------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_ev.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_ev.wait(self.period)

    def stop(self) -> dict:
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        loaded = [m for m, u in self.samples if u >= 50] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "source": "fallback"}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def gpu_numa_cpus(torch, index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when the box does not say.  Pinned buffers
    allocated and filled by a thread running there are local to the GPU's PCIe root: with several ranks on a
    two-socket box this is what keeps the host <-> device copies off the inter-socket link."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in (Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else None
    except Exception:
        return None


# -------------------------------------------------------------------------------------------------
def cpu_reference_run(n_ch: int, seconds: int, steps: int, warmup: int, threads: int, want_imgs: bool = True):
    """Times the oracle port (reference-like f32, dense mel product, threaded like mod.rs:152) on
    `n_ch` channels of `seconds` s.  Returns (audio_hours_per_s, ms_per_step)."""
    from oracle import orc
    n = SR * seconds
    an = orc.Analyzer(SR, WIN_MS, T_OVERLAP, 1, orc.MEL, N_MEL)
    wavs = []
    for c in range(n_ch):
        tr, ch = divmod(c, N_CH)
        wavs.append(orc.synth_pcm(n, SR, tr, ch, track_flags(tr), n_threads=threads))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        an.update_specs_and_imgs(wavs, DB_RANGE, CMAP_LEN, n_threads=threads, want_imgs=want_imgs)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    hours = n_ch * seconds / 3600.0 * len(times)
    return hours / total, 1000.0 * total / len(times)


def cpu_scipy_run(n_ch: int, seconds: int, threads: int):
    """A second, independently optimised CPU point: the same per-channel pipeline written with numpy +
    scipy.fft.rfft(workers=cores) and a BLAS matmul for the dense mel product (what ndarray + OpenBLAS does in the
    reference, spectrogram.rs:207).  Not a restatement of the reference's arithmetic order: a yardstick only."""
    import scipy.fft
    from oracle import orc
    from thesia_b200.analysis import calc_mel_fb, calc_normalized_win
    n = SR * seconds
    hop, win, n_fft = (int(v) for v in orc.framing_params(WIN_MS, SR, T_OVERLAP, 1))
    window = calc_normalized_win(win, n_fft)
    fb = calc_mel_fb(SR, n_fft, N_MEL)
    wavs = [orc.synth_pcm(n, SR, c // N_CH, c % N_CH, track_flags(c // N_CH), n_threads=threads) for c in range(n_ch)]
    T = 1 + n // hop

    def one_pass():
        lo, hi = np.float32(np.inf), np.float32(-np.inf)
        for w in wavs:
            padded = np.pad(w, win // 2, mode="reflect")
            frames = np.lib.stride_tricks.as_strided(padded, (T, win), (4 * hop, 4), writeable=False)
            spec = np.abs(scipy.fft.rfft(frames * window, n=n_fft, axis=1, workers=threads)).astype(np.float32, copy=False)
            with np.errstate(divide="ignore"):
                db = np.log10(spec @ fb) * np.float32(20.0)
            lo, hi = min(lo, db.min()), max(hi, db.max())
        return lo, hi

    one_pass()
    t0 = time.perf_counter()
    one_pass()
    dt = time.perf_counter() - t0
    return n_ch * seconds / 3600.0 / dt


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    threads = host_cores()
    # bounded sample of the same workload (CPU_SAMPLE)
    n_ch, seconds = CPU_SAMPLE if args.scale >= 1.0 else (4, 20)
    val, ms = cpu_reference_run(n_ch, seconds, args.steps, args.warmup, threads)
    sample = f"{n_ch} of {N_TRACKS * N_CH} channels x {seconds} s per step (audio-hours/s is duration-invariant)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(1.0), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# -------------------------------------------------------------------------------------------------
def strong_scaling(args, torch, dist, thb, ctx, pcm_weak, n, rank, world, local_rank, dev, stream) -> dict:
    """The fixed jobs dealt over the ranks by thesia_b200.sharding.plan:
       c3: the 128-channel batch of configs[2] by channel (units = (id, ch), mod.rs:152-163);
       c2: the single 1-hour file of configs[1] by FRAME RANGE (stft.rs:100-113 is the reference's frames-parallel leg),
           every rank holding only the PCM slice its frames touch.
    Timed like `value` (queued steps, CUDA events on the launching stream, max over ranks).  In the same run rank 0
    also runs each whole job alone on its GPU (no communicator): `efficiency` = that time / (N x sharded time), and --
    for N > 1 -- every rank checks its shard against that one-GPU result bit for bit (dB rows, u16 images, range)."""
    from thesia_b200 import sharding

    steps, warm = max(args.steps, 5), max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()

    def gather(x: float):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [x]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def timed(c, tracks, setting, max_sr, k, w):
        for _ in range(w):
            c.spec_batch(tracks, setting)
            c.update_spec_imgs(DB_RANGE, CMAP_LEN, max_sr, wait=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            c.spec_batch(tracks, setting)
            c.update_spec_imgs(DB_RANGE, CMAP_LEN, max_sr, wait=False)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    def breakdown(c, tracks, setting, max_sr, kname):
        c.profile_enable(True)
        c.profile_reset()
        for _ in range(3):
            c.spec_batch(tracks, setting)
            c.update_spec_imgs(DB_RANGE, CMAP_LEN, max_sr, wait=False)
        torch.cuda.synchronize()
        # the packed kernel alone, and the scalar launches next to it (rescue list, file-edge frames: the latter run on a
        # side stream NEXT TO the packed kernel when they are one or two CTAs -- their time is then not on the step's path)
        k_ms = c.profile_get(kname)[0] / 3
        edge_ms = c.profile_get(kname + "_edges")[0] / 3
        img_ms = c.profile_get("spec_to_img")[0] / 3
        red_ms = c.profile_get("minmax_reduce")[0] / 3
        c.profile_enable(False)
        return k_ms, edge_ms, img_ms, red_ms

    def job(name, channels, setting, sr, pcm_of, hours):
        """channels: [(id, ch, sr, n_samples)]; pcm_of(id, ch) -> device tensor of the whole channel."""
        framing = setting.calc_framing_params
        plan = sharding.plan(channels, framing, world)
        units = plan[rank]
        tracks = ctx.prepare_tracks([dict(pcm=pcm_of(u.id, u.ch)[u.pcm_lo:u.pcm_hi], id=u.id, ch=u.ch, sr=u.sr, full_len=u.full_len,
                                          pcm_offset=u.pcm_lo, frame_begin=u.frame_begin, frame_count=u.frame_count) for u in units])
        ctx.release_all()
        barrier()
        ms = timed(ctx, tracks, setting, sr, steps, warm)
        barrier()
        ms_all = gather(ms)
        k_ms, edge_ms, img_ms, red_ms = breakdown(ctx, tracks, setting, sr, "stft_mel_db")
        rng = ctx.range_get()
        k_all, edge_all, red_all = gather(k_ms), gather(edge_ms), gather(red_ms)
        # one GPU, whole job, no communicator: rank 0 while the others wait
        solo_ms, solo_rng = (max(ms_all) if world == 1 else None), None
        solo = None
        if rank == 0 and world > 1:
            solo = thb.Context(local_rank, stream.cuda_stream)
            whole = solo.prepare_tracks([dict(pcm=pcm_of(i, ch), id=i, ch=ch, sr=s_) for (i, ch, s_, _) in channels])
            solo_ms = timed(solo, whole, setting, sr, steps, warm)
            solo_rng = solo.range_get()
        check = None
        if world > 1:
            check = shard_check(torch, dist, ctx, solo, units, channels, framing, rank, dev, rng, solo_rng)
        if solo is not None:
            solo.close()
        barrier()
        ctx.release_all()
        worst = max(ms_all)
        rec = {"job": name, "units_per_rank": [len(p) for p in plan], "frames_per_rank": [sum(u.frame_count for u in p) for p in plan],
               "ms_per_step": worst, "value": hours / (worst * 1e-3), "unit": UNIT,
               "rank_ms": {"min": min(ms_all), "max": worst},
               "stft_kernel_ms": {"min": min(k_all), "max": max(k_all), "skew": max(k_all) - min(k_all)},
               "scalar_launches_ms": {"min": min(edge_all), "max": max(edge_all),
                                      "note": "rescue list + file-edge frames; the edge frames of a frame-range shard run beside the packed kernel"},
               "spec_to_img_ms": img_ms, "minmax_allreduce_ms": {"min": min(red_all), "max": max(red_all)},
               "one_gpu_ms_same_run": None, "efficiency": None, "limiter": None, "check": check, "dB_range": list(rng)}
        sm = torch.tensor([solo_ms if solo_ms is not None else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.broadcast(sm, src=0)
        one = float(sm.item())
        rec["one_gpu_ms_same_run"] = one
        rec["efficiency"] = one / (world * worst) if worst > 0 else None
        ideal = one / world
        over = worst - ideal
        rec["limiter"] = (f"{1e3 * over:.0f} us per step over the ideal {1e3 * ideal:.0f} us: all-reduce scope (the wait for the slowest "
                          f"rank + the exchange) {1e3 * max(red_all):.0f} us, packed-kernel skew between ranks {1e3 * (max(k_all) - min(k_all)):.0f} us, "
                          f"packed kernel {1e3 * max(k_all):.0f} us + image kernel {1e3 * img_ms:.0f} us on the slowest rank, "
                          f"the rest launch gaps / the scalar launches")
        return rec

    out = {"note": "fixed total work split over the ranks; `scaling` of the headline stays weak (the driver computes its efficiency)",
           "global_range_exchange": ("one kernel over NVLink peer memory (CUDA IPC)" if ctx.comm_peer_exchange() else
                                     ("ncclAllReduce(max) of 2 floats" if world > 1 else "single rank"))}
    # c3: 128 channels x 10 min (the weak batch of rank 0 IS this job: tracks 0..63)
    setting3 = thb.SpecSetting(WIN_MS, T_OVERLAP, 1, thb.FreqScale.Mel, N_MEL)
    n_ch = pcm_weak.shape[0]
    chans3 = [(c // N_CH, c % N_CH, SR, n) for c in range(n_ch)]
    if rank == 0:
        pcm3 = {(c // N_CH, c % N_CH): pcm_weak[c, :n] for c in range(n_ch)}
    else:
        mine = sharding.plan(chans3, setting3.calc_framing_params, world)[rank]
        buf = torch.empty((len(mine), (n + 63) // 64 * 64), dtype=torch.float32, device=dev)
        pcm3 = {}
        for k, u in enumerate(mine):
            ctx.synth_pcm(buf[k, :n], SR, u.id, u.ch, track_flags(u.id))
            pcm3[(u.id, u.ch)] = buf[k, :n]
        ctx.synchronize()
    out["c3"] = job("C3: 64 stereo tracks x 10 min, mel 128, hop 512 -- sharded by channel", chans3, setting3, SR,
                    lambda i, ch: pcm3[(i, ch)], n_ch * (n / SR) / 3600.0)
    # c2: one 1-hour mono file, hop 256, mel 128 -- sharded by frame range
    n2 = SR * 3600 if args.scale >= 1.0 else SR * 120
    setting2 = thb.SpecSetting(WIN_MS, 8, 1, thb.FreqScale.Mel, N_MEL)
    file2 = torch.empty((n2 + 63) // 64 * 64, dtype=torch.float32, device=dev)
    ctx.synth_pcm(file2[:n2], SR, 7, 0, 0)
    ctx.synchronize()
    out["c2"] = job("C2: one 1-hour 48 kHz mono file, mel 128, hop 256 -- sharded by frame range", [(7, 0, SR, n2)], setting2, SR,
                    lambda i, ch: file2[:n2], (n2 / SR) / 3600.0)
    del file2
    return out


def shard_check(torch, dist, ctx, solo, units, channels, framing, rank, dev, rng, solo_rng) -> str:
    """Every rank compares what it computed for its shard with the one-GPU run of the whole job on rank 0: same dB range,
    and -- for up to two of its units -- the same dB rows and u16 image columns, bit for bit (tools/multi_gpu_check.py's
    assertion, from inside the bench run).  Raises on a mismatch."""
    world = dist.get_world_size()
    r = torch.tensor(list(solo_rng) if rank == 0 else [0.0, 0.0], dtype=torch.float32, device=dev)
    dist.broadcast(r, src=0)
    assert tuple(float(x) for x in r.tolist()) == tuple(np.float32(x) for x in rng), (rank, rng, r.tolist())
    # each rank names the head of its first unit and the tail of its last one (the cuts of a frame-range split);
    # rank 0 serves those rows of the one-GPU result
    W = 30000
    mine = []
    if units:
        a, b = units[0], units[-1]
        mine.append((a.id, a.ch, a.frame_begin, min(a.frame_count, W)))
        tail = (b.id, b.ch, b.frame_begin + max(0, b.frame_count - W), min(b.frame_count, W))
        if tail != mine[0]:
            mine.append(tail)
    names = [None] * world
    dist.all_gather_object(names, mine)
    checked = 0
    cache = {}
    for src_rank, lst in enumerate(names):
        for (i, ch, fb, fc) in lst:
            shape = torch.zeros(2, dtype=torch.int64, device=dev)
            spec = img = None
            if rank == 0:
                if (i, ch) not in cache:
                    cache.clear()
                    cache[(i, ch)] = (solo.spec_read(i, ch), solo.img_read(i, ch))
                spec = cache[(i, ch)][0][fb:fb + fc]
                img = cache[(i, ch)][1][:, fb:fb + fc]
                shape[0], shape[1] = spec.shape[1], img.shape[0]
            dist.broadcast(shape, src=0)
            B, H = int(shape[0]), int(shape[1])
            t_spec = torch.from_numpy(np.ascontiguousarray(spec)).to(dev) if rank == 0 else torch.empty((fc, B), dtype=torch.float32, device=dev)
            t_img = torch.from_numpy(np.ascontiguousarray(img).view(np.int16)).to(dev) if rank == 0 else torch.empty((H, fc), dtype=torch.int16, device=dev)
            dist.broadcast(t_spec, src=0)
            dist.broadcast(t_img.view(torch.uint8), src=0)   # (NCCL has no 16-bit integer type: the same bytes as u8)
            if rank == src_rank:
                u = next(x for x in units if (x.id, x.ch) == (i, ch))
                got = ctx.spec_read(i, ch)[fb - u.frame_begin:fb - u.frame_begin + fc]
                gimg = ctx.img_read(i, ch)[:, fb - u.frame_begin:fb - u.frame_begin + fc]
                assert np.array_equal(got, t_spec.cpu().numpy(), equal_nan=True), f"rank {rank}: dB rows of ({i}, {ch}) differ from one GPU"
                assert np.array_equal(gimg.view(np.int16), t_img.cpu().numpy()), f"rank {rank}: image of ({i}, {ch}) differs from one GPU"
            checked += 1
    return f"{checked} shard units over {world} ranks: dB rows, u16 images and the dB range bit-equal to the one-GPU run"


def other_configs(torch, thb, ctx, pcm, n, peak_fp32, hbm_gbs) -> list:
    """Kernel times of BASELINE.json's other configurations on one GPU (PCM resident; the library's per-kernel CUDA
    events), the rows of DESIGN.md's per-configuration table (tools/design_table.py renders them)."""
    sys.path.insert(0, str(ROOT / "tools"))
    import configs_bench as cb
    cb.FP32_PEAK_TFLOPS, cb.HBM_GBS = peak_fp32, hbm_gbs
    recs = []
    t0 = time.perf_counter()
    Mel, Lin = thb.FreqScale.Mel, thb.FreqScale.Linear
    ctx.release_all()
    cb.stft_case(ctx, "C1 linear 2048/512, 2 113 529 samples mono", 1, 2113529 / 48000.0, 48000, 2048 / 48.0, 4, Lin, 0, 3, recs.append)
    cb.stft_case(ctx, "C2 mel128 2048/256, 1 h mono", 1, 3600, 48000, 2048 / 48.0, 8, Mel, 128, 3, recs.append)
    cb.stft_case(ctx, "C2 mel-default(347) 2048/256, 1 h mono", 1, 3600, 48000, 2048 / 48.0, 8, Mel, 0, 3, recs.append)
    cb.stft_case(ctx, "C3' default setting 40 ms/4 (1920/480/2048, mel 347), 32 ch x 10 min", 32, 600, 48000, 40.0, 4, Mel, 0, 3, recs.append)
    cb.stft_case(ctx, "C4 linear 16384/1024 @96 kHz, one 15-min track", 1, 900, 96000, 16384 / 96.0, 16, Lin, 0, 2, recs.append)
    cb.stft_case(ctx, "C4 mel-default(1621) 16384/1024 @96 kHz, one 15-min track", 1, 900, 96000, 16384 / 96.0, 16, Mel, 0, 2, recs.append)
    cb.stft_case(ctx, "default setting @16 kHz (640/160/1024, mel default), 32 ch x 10 min", 32, 600, 16000, 40.0, 4, Mel, 0, 2, recs.append)
    cb.stft_case(ctx, "default setting @8 kHz (320/80/512, mel default), 32 ch x 10 min", 32, 600, 8000, 40.0, 4, Mel, 0, 2, recs.append)
    # C5 (levels 9 and 15) + f3 + f4 on C3's own PCM, f2 on 16 of its channels
    cb.envelope_case(ctx, pcm.shape[0], n / SR, SR, 3, recs.append, levels=(9, 15), pcm=pcm)
    cb.tile_case(ctx, 16, n / SR, SR, 3, recs.append, levels=((0, 0),), pcm=pcm)
    recs.append({"config": "(time spent on this table)", "seconds": time.perf_counter() - t0})
    return recs


def tile_latency_us(thb, ctx, n) -> dict:
    """get_waveform_tile is one tile per call from the host's channel (lib.rs:343-367): latency of thb_waveform_tile with
    pageable HOST PCM, first call (the tile's samples cross PCIe) and repeated (the channel's device copy is reused),
    against the CPU port's encode_waveform_tile on the same tile."""
    from oracle import orc
    wav = orc.synth_pcm(n, SR, 9, 0, 0)
    out = {}
    for level in (9, 15):
        ctx.pcm_cache_clear()
        tile = 3 if level == 9 else 0
        t0 = time.perf_counter()
        got = ctx.waveform_tile(wav, 1, level, tile)
        cold = time.perf_counter() - t0
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.waveform_tile(wav, 1, level, tile)
        warm = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(5):
            want = orc.encode_waveform_tile(wav, 1, level, tile)
        cpu = (time.perf_counter() - t0) / 5
        a = np.frombuffer(got, np.float32, offset=24).reshape(-1, 3)
        b = np.frombuffer(want, np.float32, offset=24).reshape(-1, 3)
        assert got[:24] == want[:24] and np.array_equal(a[:, :2], b[:, :2])
        out[f"level{level}"] = {"samples_in_tile": min(n, 1024 << level), "gpu_first_call_us": 1e6 * cold, "gpu_repeat_us": 1e6 * warm,
                                "cpu_port_us": 1e6 * cpu, "note": "python ctypes call overhead (two calls: size query + tile) included on the GPU side"}
    ctx.pcm_cache_clear()
    return out


# -------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    import thesia_b200 as thb
    from thesia_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: thesia_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_ch_total = max(2, int(round(N_TRACKS * N_CH * args.scale)) // 2 * 2)
    n = SR * SECONDS if args.scale >= 1.0 else max(SR * 10, int(SR * SECONDS * max(args.scale, 0.02)))
    setting = thb.SpecSetting(WIN_MS, T_OVERLAP, 1, thb.FreqScale.Mel, N_MEL)
    hop, win, n_fft = setting.calc_framing_params(SR)
    T = thb.n_frames(n, win, hop)

    # The library launches on the stream it is given.  torch's default stream has handle 0, which the C ABI reads as
    # "create your own stream" -- and torch.cuda.Event on the default stream would then time nothing the library does.
    # So: one explicit torch stream, made current (synthetic PCM, NCCL ordering, events) and handed to the library.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = thb.Context(local_rank, stream.cuda_stream)
    if world > 1:
        obj = [thb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ctx.comm_init(world, rank, obj[0])

    # ---- synthetic PCM, generated on the device (rows 256-byte aligned) ----
    row = (n + 63) // 64 * 64
    pcm = torch.empty((n_ch_total, row), dtype=torch.float32, device=dev)
    for c in range(n_ch_total):
        tr, ch = divmod(c, N_CH)
        # every rank analyses its own 64 tracks of the N x 64-track job (weak scaling)
        ctx.synth_pcm(pcm[c, :n], SR, tr + rank * N_TRACKS, ch, track_flags(tr))
    ctx.synchronize()
    # the thb_track array is marshalled once, as a host program would keep it (the ctypes marshalling of 128 Python
    # dicts per step otherwise idles the GPU for ~0.3 ms ahead of every launch)
    tracks_dev = ctx.prepare_tracks([dict(pcm=pcm[c, :n], id=c // N_CH, ch=c % N_CH, sr=SR) for c in range(n_ch_total)])

    def step_resident():
        # both calls only queue work on the stream (PCM resident, no host outputs, range read after the loop):
        # the host runs ahead and the device works back to back, as a re-analysis loop of the host program would
        ctx.spec_batch(tracks_dev, setting)
        ctx.update_spec_imgs(DB_RANGE, CMAP_LEN, SR, wait=False)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- value: device-resident ----
    for _ in range(args.warmup):
        step_resident()
    ctx.profile_enable(True)
    ctx.profile_reset()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    torch.cuda.synchronize()
    mn, mx = ctx.range_get()
    clocks = sampler.stop()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0  # kernels of this library launched inside the timed region
    k_ms, k_launches = ctx.profile_get("stft_mel_db")          # the frame-pair STFT kernel alone
    edge_ms, edge_launches = ctx.profile_get("stft_mel_db_edges")  # scalar kernel: file edges + rescue list
    img_ms, img_launches = ctx.profile_get("spec_to_img")
    red_ms, _ = ctx.profile_get("minmax_reduce")
    ctx.profile_enable(False)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    hours_per_step = world * n_ch_total * (n / SR) / 3600.0
    value = hours_per_step * args.steps / (ms_total / 1000.0)

    # ---- roofline of the dominant kernel (pass 1: PCM -> dB spec + min/max) ----
    peaks = measured_peaks()
    alg_bytes = n_ch_total * (4 * n + 4 * T * N_MEL)  # 4 B/sample in + 4 B/bin out
    k_avg_ms = k_ms / max(k_launches, 1)
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    log2n = int(np.log2(n_fft))
    flops_frame = 2.5 * n_fft * log2n + 4 * (n_fft // 2 + 1) + 2 * 2 * (n_fft // 2 + 1) + N_MEL
    alg_flops = n_ch_total * T * flops_frame
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("stft_mel_db_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # FP32 roof of the part at the clock it actually ran at: SMs x 128 lanes x 2 flop x SM clock
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_fp32 = sm_count * 256 * sm_mhz * 1e6 / 1e12
    fp32_tflops = alg_flops / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0
    roofline = {
        "kernel": "stft_mel_db", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"],
        "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the same kernel at full size, read from "
                          "profiles/roofline_traffic.json (captured under ncu in a separate run, not measured by this one)",
        # what ncu says binds the kernel (profiles/): FP32 pipe + issue slots + shared-memory wavefronts, not HBM --
        # `bound`/`frac` above keep the contract's HBM figure, these two give the roof that actually limits it
        "limiter_per_ncu": "fp32 pipe / issue slots (HBM traffic is 1.004 - 1.006x algorithmic and far below the roof)",
        "frac_fp32": fp32_tflops / peak_fp32, "peak_fp32_tflops": peak_fp32, "sm_mhz_for_fp32_peak": sm_mhz,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_avg_ms,
        "achieved_fp32_tflops": fp32_tflops,
        "share_of_step": (k_ms / ms_total) if ms_total > 0 else None,
        "launches_per_step": k_launches / max(args.steps, 1),
        "edge_kernels_ms_per_step": edge_ms / max(args.steps, 1),
        "spec_to_img_avg_ms": img_ms / max(img_launches, 1),
        "spec_to_img_gbs": (n_ch_total * (4 * T * N_MEL + 2 * T * N_MEL)) / (img_ms / max(img_launches, 1) * 1e-3) / 1e9
        if img_ms > 0 else None,
        "minmax_allreduce_avg_ms": red_ms / max(args.steps, 1),
    }

    # ---- strong scaling: the FIXED jobs of BASELINE configs[2] / configs[1] dealt over the N ranks ----
    strong = None
    if not args.no_strong:
        strong = strong_scaling(args, torch, dist, thb, ctx, pcm, n, rank, world, local_rank, dev, stream)

    # ---- the other BASELINE configurations, kernel times only (N == 1) ----
    configs = None
    if rank == 0 and world == 1 and not args.no_configs and args.scale >= 1.0:
        configs = other_configs(torch, thb, ctx, pcm, n, peak_fp32, peaks["hbm_gbs"])

    # ---- the per-tile seam: one thb_waveform_tile call against the CPU port, host PCM ----
    tile_latency = None
    if rank == 0 and world == 1 and not args.no_configs:
        tile_latency = tile_latency_us(thb, ctx, n)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----
    e2e = None
    e2e_i16 = None
    host_affinity = "unbound"
    if not args.no_e2e:
        # the host side of a rank lives next to its GPU: bind before the pinned buffers are allocated and touched
        # (restored before the CPU baseline, which uses every core)
        affinity_before = os.sched_getaffinity(0)
        numa = gpu_numa_cpus(torch, local_rank)
        if numa is not None:
            os.sched_setaffinity(0, numa[1])
            host_affinity = f"NUMA node {numa[0]} of the GPU ({len(numa[1])} CPUs)"
        elif world > 1:
            # sysfs reports one node (a VM that hides the topology): give every rank its own slice of the cores, so that
            # the threads that allocate, touch and feed a rank's pinned buffers at least do not migrate over each other
            cpus = sorted(affinity_before)
            per = max(1, len(cpus) // world)
            mine = set(cpus[local_rank * per:(local_rank + 1) * per]) or set(cpus)
            os.sched_setaffinity(0, mine)
            host_affinity = f"no NUMA information: cores {min(mine)}-{max(mine)} of {len(cpus)} by local rank"
        img_bytes = n_ch_total * N_MEL * T * 2
        img_stride = N_MEL * T
        host_img_p = C.c_void_p()
        _lib.check(_lib.lib().thb_host_alloc(img_bytes, C.byref(host_img_p)))

        def pinned(nbytes, dtype, shape):
            ptr = C.c_void_p()
            _lib.check(_lib.lib().thb_host_alloc(nbytes, C.byref(ptr)))
            ctype = C.c_float if dtype == np.float32 else C.c_int16
            return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=shape)

        def time_e2e(tracks_host):
            def step_e2e():
                ctx.spec_batch(tracks_host, setting)
                r = ctx.update_spec_imgs(DB_RANGE, CMAP_LEN, SR)
                ctx.img_read_batch_into([(c // N_CH, c % N_CH) for c in range(n_ch_total)],
                                        [host_img_p.value + 2 * c * img_stride for c in range(n_ch_total)],
                                        [img_stride] * n_ch_total)
                return r

            e2e_steps = max(1, min(args.steps, args.e2e_steps))
            step_e2e()  # warm-up: the staging pool grows, the pinned pages are touched; the second pass settles the
            step_e2e()  # copy pipeline (tools/e2e_phases.py: 318 ms, then 303 ms from the second pass on)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                mn_e, mx_e = step_e2e()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            dt = float(t_e.item())
            assert (mn_e, mx_e) == (mn, mx), "host-buffer and device-resident runs disagree on the dB range"
            return hours_per_step * e2e_steps / dt, e2e_steps, 1000.0 * dt / e2e_steps

        # (1) the reference's own host layout: f32 channels (Audio.wavs rows)
        h2d_bytes = n_ch_total * n * 4
        host_pcm_p, host_pcm = pinned(h2d_bytes, np.float32, (n_ch_total, n))
        torch.from_numpy(host_pcm).copy_(pcm[:, :n])  # fill the pinned input once (outside the timed region)
        torch.cuda.synchronize()
        tracks_host = [dict(pcm=host_pcm[c], id=c // N_CH, ch=c % N_CH, sr=SR) for c in range(n_ch_total)]
        v, k, ms = time_e2e(tracks_host)
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": img_bytes + 8, "steps": k,
               "ms_per_step": ms, "host_memory": "pinned (thb_host_alloc)", "pcm_format": "f32",
               "pipeline": "H2D in 256 MB stages on a copy stream, overlapped with the STFT kernels"}
        imgs_f32 = np.ctypeslib.as_array(C.cast(host_img_p, C.POINTER(C.c_uint16)), shape=(img_bytes // 2,)).copy() \
            if args.scale < 1.0 else None

        # (2) SURVEY.md 8 f1: 16-bit PCM handed over as i16 (the >0 dBFS test track cannot be 16-bit and stays f32)
        is16 = [track_flags(c // N_CH) & 1 == 0 for c in range(n_ch_total)]
        n16 = sum(is16)
        host_i16_p, host_i16 = pinned(max(n16, 1) * n * 2, np.int16, (max(n16, 1), n))
        row16 = {}
        for c in range(n_ch_total):
            if is16[c]:
                row16[c] = len(row16)
                q = torch.round(pcm[c, :n] * 32768.0).to(torch.int16)
                torch.from_numpy(host_i16[row16[c]]).copy_(q)
        torch.cuda.synchronize()
        tracks_i16 = [dict(pcm=host_i16[row16[c]] if is16[c] else host_pcm[c], id=c // N_CH, ch=c % N_CH, sr=SR)
                      for c in range(n_ch_total)]
        v, k, ms = time_e2e(tracks_i16)
        if imgs_f32 is not None:
            got = np.ctypeslib.as_array(C.cast(host_img_p, C.POINTER(C.c_uint16)), shape=(img_bytes // 2,))
            assert np.array_equal(got, imgs_f32), "i16 ingest changed the images"
        e2e_i16 = {"value": v, "unit": UNIT, "h2d_bytes_per_step": n16 * n * 2 + (n_ch_total - n16) * n * 4,
                   "d2h_bytes_per_step": img_bytes + 8, "steps": k, "ms_per_step": ms, "pcm_format": "i16",
                   "note": f"{n16} of {n_ch_total} channels as 16-bit PCM (sample = s / 32768, exact): same dB range and "
                           "images as the f32 run; the reference-side decoder would hand over i16 instead of f32"}
        _lib.lib().thb_host_free(host_pcm_p)
        _lib.lib().thb_host_free(host_i16_p)
        _lib.lib().thb_host_free(host_img_p)
        # what the box's host side delivers to N GPUs AT ONCE with no library in the way: plain pinned cudaMemcpyAsync
        # (torch) of 1 GiB per rank, every rank at the same time.  e2e cannot beat bytes / this.
        probe_n = 1 << 30
        hp = torch.empty(probe_n, dtype=torch.uint8, pin_memory=True)
        dp = torch.empty(probe_n, dtype=torch.uint8, device=dev)
        hp.fill_(1)
        feed = {}
        for name, fn in (("h2d", lambda: dp.copy_(hp, non_blocking=True)), ("d2h", lambda: hp.copy_(dp, non_blocking=True))):
            fn()
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            gbs = 3 * probe_n / (time.perf_counter() - t0) / 1e9
            barrier()
            g = torch.tensor([gbs], dtype=torch.float64, device=dev)
            lo, tot = g.clone(), g.clone()
            if world > 1:
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            feed[name + "_gbs_per_rank_min"] = float(lo.item())
            feed[name + "_gbs_aggregate"] = float(tot.item())
        del hp, dp
        os.sched_setaffinity(0, affinity_before)
        e2e["host_affinity"] = host_affinity
        e2e["host_feed_probe"] = dict(feed, how=f"{world} rank(s) at once, 3 x 1 GiB pinned cudaMemcpyAsync each way, no library")
        # the floor the host feed sets for one e2e step of this rank: its PCM in, its images out, one after the other
        # (the images need the global range, which needs every spectrogram, which needs every sample)
        floor_ms = 1e3 * (h2d_bytes / (feed["h2d_gbs_per_rank_min"] * 1e9) + img_bytes / (feed["d2h_gbs_per_rank_min"] * 1e9))
        e2e["host_limit_ms_per_step"] = floor_ms
        e2e["host_limit_value"] = hours_per_step / (floor_ms * 1e-3)
        e2e["host_limit_gbs"] = feed["h2d_gbs_aggregate"]

    # ---- CPU baseline (rank 0, N == 1 only): oracle port on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_cores()
        n_ch_cpu, sec_cpu = CPU_SAMPLE if args.scale >= 1.0 else (4, 20)
        v, _ = cpu_reference_run(n_ch_cpu, sec_cpu, 1, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_ch_cpu} of {N_TRACKS * N_CH} channels x {sec_cpu} s, 1 warm-up + 1 timed pass"}
        try:
            n_sp = min(32, n_ch_cpu)
            cpu["scipy_pipeline"] = {"value": cpu_scipy_run(n_sp, sec_cpu, threads), "unit": UNIT,
                                     "sample": f"{n_sp} channels x {sec_cpu} s, numpy + scipy.fft.rfft(workers={threads}) + BLAS mel product "
                                               "(STFT -> |X| -> mel -> dB -> min/max; no u16 images)"}
        except Exception as e:  # a yardstick: its absence must not void the bench line
            cpu["scipy_pipeline"] = {"unavailable": repr(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.scale), "channels_per_gpu": n_ch_total, "samples_per_channel": n,
                       "frames_per_channel": T, "hop": hop, "win": win, "n_fft": n_fft, "n_mel": N_MEL,
                       "parallelism": f"channels sharded, {world} rank(s), one 2-float NCCL max all-reduce",
                       "l2": "inputs (14.75 GB per GPU) are far larger than the 126 MB L2; no flush needed",
                       "dB_range": [mn, mx]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_i16": e2e_i16, "clocks": clocks,
            "gpu_launches": int(launches), "strong": strong, "configs": configs, "tile_latency_us": tile_latency,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="<1 shrinks the workload (debug only; invalid as a result)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and args.scale >= 1.0:
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
