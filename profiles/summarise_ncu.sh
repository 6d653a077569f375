#!/bin/bash
# profiles/summarise_ncu.sh <report.ncu-rep> <out.txt>: the text summary kept under profiles/ for one `ncu --set full` capture
# (details page + the raw metrics DESIGN.md quotes).  Run where ncu is installed; needs no GPU.
set -e
rep="$1"; out="$2"
{
  echo "# ncu --set full --clock-control none --import-source on : $(basename "$rep")"
  ncu -i "$rep" --page details 2>/dev/null | grep -v "^==PROF==" | grep -E "^\s+(void |[A-Za-z_]+<|k_)|Duration|Elapsed Cycles|SM Frequency|DRAM Frequency|Throughput|Executed Ipc|Issue Slots Busy|Issued Ipc|SM Busy|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|Mem Pipes Busy|One or More Eligible|No Eligible|Active Warps Per Scheduler|Eligible Warps Per Scheduler|Warp Cycles Per|Avg\. Active Threads|Executed Instructions|Issued Instructions|Registers Per Thread|Shared Memory|Block Limit|Theoretical Occupancy|Achieved Occupancy|Achieved Active Warps|Grid Size|Block Size|Waves Per SM|bank conflict|fused|Local Speedup|Est\. Speedup|stalled|uncoalesced|excessive" || true
  echo
  echo "# raw metrics"
  ncu -i "$rep" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:120])
    for w in want:
        if w in hdr:
            print("  %-70s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
'
} > "$out"
echo "wrote $out"
