#!/usr/bin/env python
"""Extracts the issue-side counters of the committed ncu captures (gpurun_out/r2_wq_<scene>.ncu-rep, written by
tools/ncu_refresh.sh) into profiles/issue_counters.json (read by bench.py's roofline_issue block) and
profiles/dram_traffic.json (roofline.traffic).  Runs here, without a GPU: `python tools/ncu_issue_counters.py`."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration_ms",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_inst",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio": "not_predicated_off_threads_per_warp_inst",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_lsu_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflict_wavefronts",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "sm__icc_request_hit_rate.pct": "icache_hit_pct",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed": "gpc_icache_requests_pct_of_peak",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    out, traffic = {}, {}
    for scene in ("rgbbox", "irreg", "random"):
        rep = os.path.join(ROOT, "gpurun_out", f"r2_wq_{scene}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                d["kernel"] = v
            if h in WANT and v not in ("", "no data"):
                x = float(v.replace(",", ""))
                if WANT[h].startswith("dram_"):
                    x *= UNIT.get(u, 1.0)
                d[WANT[h]] = x
        log = os.path.join(ROOT, "gpurun_out", f"r2_ncu_{scene}.log")
        if os.path.exists(log):
            m = re.search(r"^WORK (\{.*\})$", open(log).read(), re.M)
            if m:
                w = json.loads(m.group(1))
                d["work"] = w
                d["warp_instructions_per_32_node_steps"] = round(d["warp_instructions"] * 32 / w["node_steps"], 1)
                d["thread_instructions_per_segment"] = round(d["warp_instructions"] * 32 / w["segments"], 1)
        out[scene] = d
        traffic[scene] = int(d.get("dram_read", 0) + d.get("dram_write", 0))
    if not out:
        print("no captures found", file=sys.stderr)
        return 1
    with open(os.path.join(ROOT, "profiles", "issue_counters.json"), "w") as f:
        json.dump({"source": "ncu --set full, tools/ncu_refresh.sh, one launch per scene (cold-cache, serialised)", "kernels": out}, f, indent=1)
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update({k: v for k, v in traffic.items() if k in ("rgbbox", "irreg")})
    old["random1M_2spp"] = traffic.get("random", old.get("random1M_2spp"))
    old["source"] = ("ncu --set full, render_warpqueue_kernel as planned by default, 1000x1000 frames at 64 spp and the 1 M-sphere scene at 2 spp "
                     "(tools/ncu_refresh.sh, profiles/r2_warpqueue_*_details.txt): dram__bytes_read.sum + dram__bytes_write.sum per launch. The scene is "
                     "shared-memory / L2 resident; the frame (4 MB) is written once; the rest is spill of the L2-resident sample-colour buffer.")
    json.dump(old, open(tp, "w"), indent=1)
    print(json.dumps(out, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
