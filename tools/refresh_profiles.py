"""Rebuilds profiles/r01_* from the last gpurun capture (gpurun_out/): ncu raw summary, DRAM traffic, source hotspots."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REP = os.path.join(ROOT, "gpurun_out", "prof_view_c4_r01_final.ncu-rep")
ROBOTS = 25600
def ncu(args):
    return subprocess.run(["ncu", "-i", REP] + args, capture_output=True, text=True).stdout
rows = list(csv.reader(ncu(["--page", "raw", "--csv"]).splitlines()))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
out, tr = [], {}
sc = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
for v in rows[2:]:
    name = v[h.index('Kernel Name')]
    out.append("%s, bench.py default workload c4 (128 scenes x 200 robots + 200 ervoscene peds, 7333^2 grid), ncu --set full --clock-control none (cold, serialised)" % name)
    for n in want:
        i = h.index(n); out.append("%-90s %s %s" % (n, v[i], u[i]))
    r, w = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
    tr[name.split('(')[0]] = float(v[r]) * sc[u[r]] + float(v[w]) * sc[u[w]]
    out.append("")
open(os.path.join(ROOT, "profiles", "r01_k_view_c4_ncu_summary.txt"), "w").write("\n".join(out))
tot = sum(tr.values())
json.dump({"c4": {"bytes_per_robot_step": tot / ROBOTS, "note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_view launch plus the concurrent k_ped_obs launch "
           "(128 scenes x 200 robots = 25600 robot-steps), ncu --set full, profiles/r01_k_view_c4_ncu_summary.txt", "per_kernel": tr}},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(l for l in out if "time_duration" in l or "inst_executed.sum" in l or "issue_active" in l), tot / ROBOTS)
# source hotspots
rows = list(csv.reader(ncu(["--page", "source", "--print-source", "cuda,sass", "--csv"]).splitlines()))
src = open(os.path.join(ROOT, "img_env_b200", "csrc", "view.cuh")).read().splitlines()
def find(pat):
    for i, l in enumerate(src):
        if pat in l: return i + 1
    raise KeyError(pat)
marks = [("prologue: pose transforms, static tables -> smem", find("extern __shared__ __align__(16) unsigned char smem_raw[];")),
         ("A: collision code over the footprint lattice", find("// ---- Phase A: collision code")),
         ("B: inverse (world->view) rasterisation incl. FOV-edge pixels", find("// ---- Phase B: egocentric occupancy raster")),
         ("B': forward tile path (lasers off; unused here) + barrier after B", find("int* n_active = &sh->red[2];")),
         ("C: ray first hits from boundary-cell lists, laser output, hit prefix", find("// ---- Phase C: first occupied cell of every laser ray")),
         ("D/E/F: laser_map reconstruction + cubic 400->48 + f16", find("// D/E: the final view_map_ value of a pixel")),
         ("G: state vector + bookkeeping", find("// ---- Phase G: state vector")),
         ("end", find("// Pedestrian observation of every robot"))]
rt0 = find("__device__ __forceinline__ int ray_touch(")
ker = fname = ie = None; data = collections.defaultdict(list)
for r in rows:
    if r and r[0] == 'Function Name': ker = r[1].split('(')[0]; continue
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if 'Instructions Executed' in r: ie = r.index('Instructions Executed'); smp = r.index('# Samples'); continue
    if ie is not None and len(r) > ie and r[0] not in ('', 'Line No') and r[2] == '-':
        try: n = int(r[ie]); s = int(r[smp]); ln = int(r[0])
        except ValueError: continue
        data[ker].append((s, n, fname, ln, r[1].strip()[:110]))
out = []
for k, v in data.items():
    ts = sum(x[0] for x in v); ti = sum(x[1] for x in v)
    out.append("== %s: %d stall samples, %d warp instructions (source-attributed), c4 128 scenes x 200 robots" % (k, ts, ti))
    if 'k_view' in k:
        b = collections.Counter(); bi = collections.Counter()
        for s_, n_, f_, ln, _ in v:
            key = "helpers inlined from other headers (tfmath, state, intrinsics, atomics)"
            if f_ == 'view.cuh':
                key = "view.cuh helpers (global_value, push_cell, ...)"
                for (nm, a), (_, bb) in zip(marks[:-1], marks[1:]):
                    if a <= ln < bb: key = nm
                if rt0 <= ln < rt0 + 26: key = "ray_touch (closed-form line walk test; phases C and D)"
            b[key] += s_; bi[key] += n_
        out.append("   share of stall samples / warp instructions by phase:")
        for kk, _ in sorted(b.items(), key=lambda x: -x[1]): out.append("   %5.1f%% %5.1f%%  %s" % (100 * b[kk] / ts, 100 * bi[kk] / ti, kk))
    out.append("   top source lines by stall samples: (samples, warp-instr, file:line, source)")
    for x in sorted(v, reverse=True)[:20]: out.append("   %6d %10d %s:%d  %s" % x)
    out.append("")
open(os.path.join(ROOT, "profiles", "r01_k_view_c4_source_hotspots.txt"), "w").write("\n".join(out))
for f, t in [("r01_launches_c4.csv", "r01_launches_c4.csv"), ("bench_default.json", "r01_bench_c4.json"), ("bench_ref.json", "r01_bench_c4_reference.json"),
             ("bench_c1.json", "r01_bench_c1.json"), ("bench_c2.json", "r01_bench_c2.json"), ("bench_c3.json", "r01_bench_c3.json"), ("bench_c5.json", "r01_bench_c5.json")]:
    p = os.path.join(ROOT, "gpurun_out", f)
    if os.path.exists(p) and os.path.getsize(p) > 0: shutil.copy(p, os.path.join(ROOT, "profiles", t))
