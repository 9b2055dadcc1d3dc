#!/usr/bin/env python
"""Instruction working set of a kernel from an ncu capture: how many bytes of SASS produce 50 / 80 / 90 / 95 / 99 / 99.9 % of
the executed instructions, by instruction and by 128-byte instruction-cache line, and the share of every KB of the kernel.

Why: the warp-queue render kernel ran with 91 % / 94 % instruction-cache hits (`sm__icc_request_hit_rate`) because its 99.9 %
set was 34 KB / 33 KB against a 32 KB cache; this is the tool that showed it.  Runs without a GPU.

  ncu -i gpurun_out/r2_wq_rgbbox.ncu-rep --page source --csv --print-source sass > /tmp/k.csv
  python tools/ncu_code_footprint.py /tmp/k.csv
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
    ins = []
    for r in rows[2:]:
        if len(r) > ie:
            try:
                ins.append((int(r[ia], 16), int(r[ie])))
            except ValueError:
                pass
    base, tot = ins[0][0], sum(e for _, e in ins)
    print(f"{len(ins)} SASS instructions = {len(ins) * 16 / 1024:.1f} KB, {tot} executed (warp level)")
    marks = (0.5, 0.8, 0.9, 0.95, 0.99, 0.999)

    def report(counts, unit_bytes, what):
        acc, mi = 0, 0
        for n, e in enumerate(sorted(counts, reverse=True), 1):
            acc += e
            while mi < len(marks) and acc >= marks[mi] * tot:
                print(f"  {marks[mi] * 100:5.1f} % of the executed instructions come from the hottest {n} {what} = {n * unit_bytes / 1024:.1f} KB")
                mi += 1

    report([e for _, e in ins], 16, "instructions")
    lines = {}
    for a, e in ins:
        lines[(a - base) // 128] = lines.get((a - base) // 128, 0) + e
    report(list(lines.values()), 128, "128-byte lines")
    print("share of the executed instructions per KB of the kernel (KB offset: %):")
    per_kb = {}
    for a, e in ins:
        per_kb[(a - base) // 1024] = per_kb.get((a - base) // 1024, 0) + e
    print(" ".join(f"{k}:{100 * e / tot:.1f}" for k, e in sorted(per_kb.items())))


if __name__ == "__main__":
    main()
