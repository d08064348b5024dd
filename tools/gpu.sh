#!/bin/bash
# build (fail loudly), then one gpu_cycle.sh on the GPU box.  usage: tools/gpu.sh <tag> [kernel-regex]
set -e
cd "$(dirname "$0")/.."
make -C thesia_b200/csrc -j8 > /tmp/thb_build.log 2>&1 || { grep -i "error" /tmp/thb_build.log | head; echo BUILD FAILED; exit 1; }
TAG=${1:-cycle}; KRE=${2:-stft2048_pair}
/usr/local/graft/bin/gpurun --timeout 1500 -- "bash tools/gpu_cycle.sh $TAG $KRE" > /tmp/gpurun_$TAG.log 2>&1 || true
tail -3 /tmp/gpurun_$TAG.log | cut -c1-200
python - "$TAG" <<'PY'
import json, re, sys
t = open(f'gpurun_out/{sys.argv[1]}.log').read()
print(t[:160].replace("\n", " | "))
m = re.search(r'^\{.*\}$', t, re.M)
if m:
    d = json.loads(m.group(0))
    r = d['roofline']
    print(f"value {d['value']:.1f} step {d['ms_per_step']:.2f} ms  stft {r['avg_launch_ms']:.2f} ms frac {r['frac']:.4f}  img {r['spec_to_img_avg_ms']:.2f} ms  e2e {d['e2e']['value'] if d['e2e'] else None}  cpu {d['cpu_baseline']['value'] if d['cpu_baseline'] else None}")
PY
