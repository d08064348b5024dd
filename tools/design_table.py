#!/usr/bin/env python
"""Renders DESIGN.md's per-configuration table from a bench line (BENCH_rNN.json's `parsed`, or a file holding the JSON
line bench.py printed): the numbers in the document are the driver-run ones, not a lease log.

    python tools/design_table.py BENCH_r02.json        # or gpurun_out/r2b_bench.json
"""
import json
import sys


def load(path):
    txt = open(path).read().strip()
    try:
        d = json.loads(txt)
        return d.get("parsed", d)
    except json.JSONDecodeError:
        return json.loads(txt.split("\n")[-1])


def main():
    d = load(sys.argv[1])
    r = d["roofline"]
    print("| config | kernel time | audio-h/s | alg. GB/s (of %.0f) | FP32 TF/s (of %.1f) |" % (r["peak"], r["peak_fp32_tflops"]))
    print("|---|---|---|---|---|")
    print("| %s | %.2f ms (step %.2f ms) | %.0f | %.0f (%.1f %%) | %.1f (%.1f %%) |" % (
        "C3 mel 128, hop 512, 128 ch x 10 min (the bench workload)", r["avg_launch_ms"], d["ms_per_step"], d["value"], r["achieved"],
        100 * r["frac"], r["achieved_fp32_tflops"], 100 * r["frac_fp32"]))
    for c in d.get("configs") or []:
        name = c["config"]
        if "stft_ms" in c:
            call = c.get("spec_batch_ms")
            t = "%.3f ms" % c["stft_ms"] + (" (the call, queued: %.3f ms)" % call if call and abs(call - c["stft_ms"]) > 0.05 * c["stft_ms"] else "")
            print("| %s | %s | %.0f | %.0f (%.1f %%) | %.1f (%.1f %%) |" % (name, t, c["audio_hours_per_s"], c["algorithmic_GBps"],
                                                                       100 * c["hbm_frac"], c["fp32_tflops"], 100 * c.get("frac_fp32", 0)))
        elif "envelope_ms" in c:
            print("| C5 envelope, level %d (%d columns / channel) | %.2f ms | %.0f | %.0f (%.0f %%) | -- |" % (
                c["level"], c["columns_per_channel"], c["envelope_ms"], c["audio_hours_per_s"], c["algorithmic_GBps"], 100 * c["hbm_frac"]))
        elif "stats_ms" in c:
            print("| f3 level statistics, C3's PCM | %.2f ms | -- | %.0f (%.0f %%) | -- |" % (c["stats_ms"], c["algorithmic_GBps"], 100 * c["hbm_frac"]))
        elif "gain_apply_ms" in c:
            print("| %s | %.2f ms | -- | %.0f (%.0f %%) | -- |" % (name, c["gain_apply_ms"], c["algorithmic_GBps"], 100 * c["hbm_frac"]))
        elif "tile_kernels_ms" in c:
            print("| f2 tiles, level (%d, %d), %d x %s | %.2f ms | -- | %.0f (%.0f %%) | -- |" % (
                c["level_x"], c["level_y"], c["images"], "x".join(map(str, c["image_shape"])), c["tile_kernels_ms"], c["algorithmic_GBps"], 100 * c["hbm_frac"]))
    s = d.get("strong")
    if s:
        print()
        print("| strong scaling (N = %d) | ms / step | audio-h/s | one GPU, same run | efficiency | all-reduce scope | kernel skew |" % d["n_gpus"])
        print("|---|---|---|---|---|---|---|")
        for k in ("c3", "c2"):
            j = s[k]
            print("| %s | %.3f | %.0f | %.3f ms | %.3f | %.0f us | %.0f us |" % (j["job"], j["ms_per_step"], j["value"], j["one_gpu_ms_same_run"],
                                                                          j["efficiency"], 1e3 * j["minmax_allreduce_ms"]["max"], 1e3 * j["stft_kernel_ms"]["skew"]))
    t = d.get("tile_latency_us")
    if t:
        print()
        print("| get_waveform_tile, one call, host PCM | GPU first call | GPU repeated (PCM cached on device) | CPU port |")
        print("|---|---|---|---|")
        for k, v in t.items():
            print("| %s (%d samples) | %.0f us | %.0f us | %.0f us |" % (k, v["samples_in_tile"], v["gpu_first_call_us"], v["gpu_repeat_us"], v["cpu_port_us"]))


if __name__ == "__main__":
    main()
