"""Summarises an ncu --set full capture: per kernel the headline counters and, for kernels compiled with -lineinfo, the hottest
source lines (stall samples and warp instructions).   python tools/profile_summary.py gpurun_out/prof.ncu-rep [robot_steps] [topN]"""
import collections, csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'launch__local_mem_per_thread' if False else 'smsp__inst_executed_op_local_ld.sum']


def ncu(rep, args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    robots = float(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(ncu(rep, ["--page", "raw", "--csv"]).splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print("##", v[h.index('Kernel Name')][:90])
        for n in WANT:
            if n in h:
                i = h.index(n); print("   %-88s %s %s" % (n, v[i], u[i]))
        if robots and 'smsp__inst_executed.sum' in h:
            print("   warp instructions per robot-step: %.0f" % (float(v[h.index('smsp__inst_executed.sum')]) / robots))
    rows = list(csv.reader(ncu(rep, ["--page", "source", "--print-source", "cuda,sass", "--csv"]).splitlines()))
    ker = fname = ie = None; data = collections.defaultdict(list)
    for r in rows:
        if r and r[0] == 'Function Name': ker = r[1].split('(')[0]; continue
        if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
        if 'Instructions Executed' in r: ie = r.index('Instructions Executed'); smp = r.index('# Samples'); continue
        if ie is not None and len(r) > ie and r[0] not in ('', 'Line No') and r[2] == '-':
            try: n = int(r[ie]); s = int(r[smp]); ln = int(r[0])
            except ValueError: continue
            data[ker].append((s, n, fname, ln, r[1].strip()[:120]))
    for k, v in data.items():
        ts = sum(x[0] for x in v) or 1; ti = sum(x[1] for x in v) or 1
        print("== %s: %d stall samples, %d warp instructions (source-attributed)" % (k, ts, ti))
        for x in sorted(v, reverse=True)[:top]:
            print("   %5.1f%% %5.1f%%  %s:%d  %s" % (100.0 * x[0] / ts, 100.0 * x[1] / ti, x[2], x[3], x[4]))


if __name__ == "__main__":
    main()
