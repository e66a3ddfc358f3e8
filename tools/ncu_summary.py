#!/usr/bin/env python
"""Summarise an .ncu-rep into a short markdown table (run here; ncu reads reports without a GPU).
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv, subprocess, sys, io

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tcgen05 bf16 tensor pipe % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long_scoreboard %"),
]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path}` (`--set full --clock-control none`)\n")
    for r in rows[2:]:
        print(f"## {r[idx['Kernel Name']]}\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                print(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        print()

if __name__ == "__main__":
    main(sys.argv[1])
