#!/usr/bin/env python
"""Aggregate an ncu SASS source page by the kernel's own CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all build/thb_stft_fast.o ; nvdisasm -gi -c x.cubin > x.disasm
    python tools/ncu_by_line.py sass.csv x.disasm <kernel-substring> <file.cu> [units]

Every SASS instruction is attributed to the OUTERMOST line of <file.cu> in its inline chain (so the
cost of an inlined dft32() lands on the call site), then instructions executed, shared-memory
wavefronts and stall samples are summed per line and printed per `units` (e.g. frames)."""
import csv
import re
import sys
from collections import defaultdict


def parse_disasm(path, kernel_sub, cu_name):
    lines = open(path).read().split("\n")
    in_k = False
    cur = None
    pend = []
    out = []  # (offset, line, opcode)
    for ln in lines:
        if ln.startswith(".text."):
            in_k = kernel_sub in ln
            continue
        if not in_k:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pend.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            if pend:
                # the last annotation of a block is the outermost frame
                own = [p for p in pend if p[0].endswith(cu_name)]
                cur = own[-1][1] if own else cur
                pend = []
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    sass_csv, disasm, ksub, cu = sys.argv[1:5]
    units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    ins = parse_disasm(disasm, ksub, cu)
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ci = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[h + 1:] if len(r) >= len(hdr) - 2]
    if len(body) != len(ins):
        print(f"warning: {len(body)} profiled instructions vs {len(ins)} disassembled", file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    ops = defaultdict(float)
    tot = [0.0, 0.0, 0.0, 0.0]
    for r, (_, line, op) in zip(body, ins):
        def f(name):
            try:
                return float(r[ci[name]] or 0)
            except (KeyError, ValueError):
                return 0.0
        v = [f("Instructions Executed"), f("L1 Wavefronts Shared"), f("# Samples"), f("L1 Wavefronts Shared Excessive")]
        for i in range(4):
            agg[line][i] += v[i]
            tot[i] += v[i]
        opname = op.split()[1] if op.startswith("@") else op.split()[0]
        ops[opname.split(".")[0]] += v[0]
    src = open(cu if "/" in cu else f"thesia_b200/csrc/{cu}").read().split("\n")
    print(f"total: instr {tot[0] / units:.1f}  smem wavefronts {tot[1] / units:.1f} (excess {tot[3] / units:.1f})  samples {tot[2]:.0f}")
    print(f"{'line':>5} {'instr':>9} {'smem_wf':>8} {'excess':>7} {'samp%':>6}  source")
    for line in sorted(agg, key=lambda k: (k is None, k)):
        a = agg[line]
        if a[0] / units < 0.5 and a[2] / max(tot[2], 1) < 0.002:
            continue
        text = src[line - 1].strip()[:90] if line and line <= len(src) else "?"
        print(f"{line!s:>5} {a[0] / units:9.1f} {a[1] / units:8.1f} {a[3] / units:7.1f} {100 * a[2] / max(tot[2], 1):6.1f}  {text}")
    print("opcodes (executed per unit):")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {k:<10} {v / units:9.1f}")


if __name__ == "__main__":
    main()
