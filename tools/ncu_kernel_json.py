#!/usr/bin/env python
"""The numbers of one ncu capture of the bench kernel that bench.py's roofline block quotes, as a tracked JSON file
(profiles/bench_kernel_ncu.json): executed warp instructions (-> dynamic instructions per edge update), DRAM bytes
(-> roofline.traffic), pipe utilisations.  The capture is `ncu --set full` of tools/ab_kernel.py --frames F --reps 1 aot.
    python tools/ncu_kernel_json.py gpurun_out/x.ncu-rep FRAMES MAXITER EDGES Z > profiles/bench_kernel_ncu.json"""
import csv
import json
import subprocess
import sys


def main():
    rep, frames, maxiter, edges, Z = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    d = dict(zip(hdr, rows[2]))
    num = lambda k: float(d[k].replace(",", ""))
    unit = lambda k: rows[1][hdr.index(k)]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    warp_edge_updates = frames * maxiter * edges * Z / 32.0
    res = {"source": rep.split("/")[-1], "kernel": d.get("Kernel Name"), "frames": frames, "maxiter": maxiter, "edges": edges, "Z": Z,
           "duration_ms": num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit("gpu__time_duration.sum")],
           "registers_per_thread": num("launch__registers_per_thread"),
           "warp_instructions_executed": num("smsp__inst_executed.sum"),
           "instructions_per_edge_update_dynamic": num("smsp__inst_executed.sum") / warp_edge_updates,
           "dram_bytes_read": num("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")],
           "dram_bytes_write": num("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")],
           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
           "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
           "shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
           "shared_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
           "shared_wavefronts_per_edge_update": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / warp_edge_updates,
           "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
           "sm_clock_ghz": num("sm__cycles_elapsed.avg.per_second")}
    res["dram_bytes_per_frame"] = (res["dram_bytes_read"] + res["dram_bytes_write"]) / frames
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
