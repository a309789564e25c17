"""Summarise an .ncu-rep: key raw metrics per launch + hottest source lines (needs -lineinfo).
usage: python tools/ncu_summary.py REPORT.ncu-rep [kernel-id-filter]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__cycles_active.avg"]
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


for r in rows[2:]:
    for wname in want:
        if wname in hdr:
            i = hdr.index(wname)
            print(f"{wname} = {r[i]} {units[i]}")
    # sector-per-request efficiency of the global loads (32-byte sectors fetched per warp-level load request;
    # 4 = a fully coalesced 128-byte line per request for 4-byte accesses, 32 = one sector per lane)
    a, b = "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"
    if a in hdr and b in hdr:
        sa, sb = num(r[hdr.index(a)]), num(r[hdr.index(b)])
        if sa and sb:
            print(f"derived: global-load sectors per request = {sa / sb:.2f}")
    a, b = "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"
    if a in hdr and b in hdr:
        sa, sb = num(r[hdr.index(a)]), num(r[hdr.index(b)])
        if sa and sb:
            print(f"derived: global-store sectors per request = {sa / sb:.2f}")
    print("----")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] +
                     (["--kernel-id", sys.argv[2]] if len(sys.argv) > 2 else []), capture_output=True, text=True).stdout
cur, agg = None, {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[7]), int(r[4])
    except ValueError:
        continue
    k = (cur, ln)
    a = agg.setdefault(k, [0, 0, r[1].strip()[:100]])
    a[0] += inst
    a[1] += samp
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp-instructions", tot, "stall samples", tots)
for (f, ln), (inst, samp, text) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{inst / tot * 100:5.1f}% inst {samp / tots * 100:5.1f}% samples  {f}:{ln}  {text}")
