#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into a small tracked text file.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME.txt [--traffic-key config2_f32]

Writes the key speed-of-light / memory / scheduler / stall metrics per captured
launch, and (with --traffic-key) records dram read+write bytes per launch in
profiles/step_kernel_traffic.json, which bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu --set full --clock-control none summary of {os.path.basename(rep)}",
             "# (per-launch values are cold-cache and serialised under the profiler; never a bench number)"]
    traffic = []
    for r in data:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"\n== launch ID {r[0]}: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:75s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False \
                    and h.endswith(".ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append("top warp stall reasons (cycles per issued instruction):")
        for v, h in stalls[:6]:
            lines.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''):60s} {v:8.2f}")

        def val(k):
            i = hdr.index(k)
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            return float(r[i]) * mult
        try:
            traffic.append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
        except Exception:
            pass
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if traffic_key and traffic:
        path = os.path.join(os.path.dirname(out), "step_kernel_traffic.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[traffic_key] = sum(traffic) / len(traffic)
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
