"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + hottest SASS instructions.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.md]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct"]
for r in rows[2:]:
    print("## kernel:", r[hdr.index("Kernel Name")], file=out)
    print("| metric | unit | value |\n|---|---|---|", file=out)
    for k in KEYS:
        if k in hdr:
            print("| %s | %s | %s |" % (k, units[hdr.index(k)], r[hdr.index(k)]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
start = [i for i, l in enumerate(lines) if l and l[0] == "Address"]
if start:
    h = lines[start[0]]
    body = [l for l in lines[start[0] + 1:] if len(l) == len(h) and l[0].startswith("0x")]
    ie, te, smp = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    tot_i = sum(int(l[ie]) for l in body); tot_t = sum(int(l[te]) for l in body); tot_s = sum(int(l[smp]) for l in body)
    print("\nSASS instructions: %d, warp-instructions executed: %d, avg active threads %.2f, stall samples %d" % (
        len(body), tot_i, tot_t / max(1, tot_i), tot_s), file=out)
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(l[h.index(c)]) for l in body) for c in stalls}
    print("stall samples by reason: " + ", ".join("%s=%d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v), file=out)
    print("\n### hottest instructions by stall samples\n| # samples | executed | avg thr | top stall | SASS |\n|---|---|---|---|---|", file=out)
    for l in sorted(body, key=lambda l: -int(l[smp]))[:40]:
        top = max(stalls, key=lambda c: int(l[h.index(c)]))
        print("| %s | %s | %s | %s | `%s` |" % (l[smp], l[ie], l[h.index("Avg. Threads Executed")], top[6:], l[1].strip()), file=out)
