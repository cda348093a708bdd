#!/usr/bin/env python
"""Summarises an `ncu --set full` report into a small JSON under profiles/ (read here, on the CPU box).

    python tools/ncu_summary.py gpurun_out/prof_trmm_r01d.ncu-rep profiles/ncu_trmm_r01.json
"""
import csv
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_of_peak_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_of_peak_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for row in rows[2:]:
        k = {"kernel": row[hdr.index("Kernel Name")], "block": row[hdr.index("Block Size")], "grid": row[hdr.index("Grid Size")]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in KEEP and row[i] != "":
                v = float(row[i]) * UNIT_SCALE.get(units[i], 1.0)
                name = KEEP[h]
                if name == "duration":
                    name, v = "duration_us", v * 1e6
                elif name.startswith("dram_") and name != "dram_throughput_pct":
                    name += "_bytes"
                k[name] = v
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and row[i]:
                if float(row[i]) >= 0.2:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(row[i]), 3)
        k["warp_stall_cycles_per_issued_instruction"] = stalls
        if "dram_read_bytes" in k:
            k["dram_traffic_bytes"] = k["dram_read_bytes"] + k.get("dram_write_bytes", 0.0)
        kernels.append(k)
    json.dump({"source": rep, "tool": "ncu --set full --clock-control none (per-launch, cold-cache, serialised)", "kernels": kernels},
              open(out, "w"), indent=1)
    for k in kernels:
        print(k["kernel"][:60], {a: b for a, b in k.items() if a not in ("kernel", "warp_stall_cycles_per_issued_instruction")})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
