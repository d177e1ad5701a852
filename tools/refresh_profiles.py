"""Rebuilds profiles/r02_* from the last tools/capture_profiles.sh run (gpurun_out/): bench lines, ncu launch list, per-kernel ncu
summaries + hottest source lines, phase breakdown of k_view, DRAM traffic per robot-step (keyed to the source hash), SASS opcodes."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
ROBOTS = {"c1": 8192, "c3": 8192, "c4": 25600, "c5": 32768}


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main():
    from bench import source_sha
    for f in os.listdir(G):
        if f.startswith("r02_bench_") and f.endswith(".json") and os.path.getsize(os.path.join(G, f)) > 0:
            shutil.copy(os.path.join(G, f), os.path.join(P, f))
    for f in ("r02_launches_c4.csv", "r02_loop_rate.txt"):
        if os.path.exists(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, f))
    traffic = {}
    for w, robots in ROBOTS.items():
        rep = os.path.join(G, "r02_prof_%s.ncu-rep" % w)
        if not os.path.exists(rep):
            continue
        txt = run([sys.executable, os.path.join(ROOT, "tools", "profile_summary.py"), rep, str(robots), "30"])
        if w != "c4":
            txt = txt.split("== ")[0]
        open(os.path.join(P, "r02_%s_ncu_summary.txt" % w), "w").write(
            "bench.py --workload %s, one B200, ncu --set full --clock-control none (cold caches, serialised launches)\n" % w + txt)
        rows = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
        h, u = rows[0], rows[1]
        sc = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
        per = {}
        for v in rows[2:]:
            name = v[h.index('Kernel Name')].split('(')[0].replace("void ", "")
            r_, w_ = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
            per[name] = float(v[r_]) * sc[u[r_]] + float(v[w_]) * sc[u[w_]]
        obs = sum(b for k, b in per.items() if k.startswith("k_view") or k.startswith("k_ped_obs"))
        traffic[w] = {"bytes_per_robot_step": obs / robots, "source_sha": source_sha(), "per_kernel_bytes": per,
                      "note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_view launch plus the k_ped_obs launch that runs beside it, "
                              "ncu --set full, profiles/r02_%s_ncu_summary.txt; valid for the sources with this hash only" % w}
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    # phase breakdown of k_view (instruction / stall share by source range)
    rep = os.path.join(G, "r02_prof_c4.ncu-rep")
    if os.path.exists(rep):
        src = open(os.path.join(ROOT, "img_env_b200", "csrc", "view.cuh")).read().splitlines()
        def find(p):
            for i, l in enumerate(src):
                if p in l: return i + 1
            raise KeyError(p)
        marks = [("prologue: mbarriers + bulk copies issued, ray table initialised", find("extern __shared__ __align__(16) unsigned char smem_raw[];")),
                 ("gather: footprint records near the FOV / the robot, static block list", find("// ---- Gather: footprint records")),
                 ("A: collision lattice (lambda; runs for robots with something under them)", find("// ---- Phase A: collision code")),
                 ("B/C: ray update of a found cell (push_cell, heavy-cell list)", find("// ---- Phase B: egocentric occupancy raster")),
                 ("B: forward rasteriser (not used by this variant)", find("const int n_trow = (vh + 31) >> 5")),
                 ("B: FOV-edge pixels (lambda)", find("// FOV-edge pixels (and the laser origin): forward")),
                 ("B: light loops (edge pixels, lattice)", find("// the light, even parts of the phase")),
                 ("B: candidate cell -> view pixels (float pre-test, exact forward check)", find("auto candidate = [&](int cX, int cY) {")),
                 ("B: chunk decode + block prefix sum", find("for (int c0 = 0; c0 < n_items; c0 += CAND_CHUNK) {")),
                 ("B: even split, bisection, bit walk", find("const int per = (total + VIEW_THREADS - 1) / VIEW_THREADS;")),
                 ("post-B barrier, code published, heavy cells", find("if (!DEBUG_FULL && need_A) {")),
                 ("C: laser ranges out, hit bits, per-block nearest / farthest hit", find("// laser ranges out; one bit per ray")),
                 ("E: value of a source pixel (pixel_code lambda)", find("// ---- Phase D/E/F: laser_map reconstruction")),
                 ("debug raster (not used by this variant)", find("// whole 400x400 raster for the tests")),
                 ("D1: segments of 8 outputs", find("uint16_t* o_img = d.o_sensor")),
                 ("D2: outputs of the listed segments (hit-free / all-shadow / dirty)", find("// (2) the outputs of the listed segments")),
                 ("E/F: dirty outputs (cubic resize + f16)", find("const int n4 = sh->n_dirty * 4;")),
                 ("end", find("// Pedestrian observation of every robot"))]
        rows = list(csv.reader(run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]).splitlines()))
        ker = fname = ie = None; b = collections.Counter(); bi = collections.Counter()
        for r in rows:
            if r and r[0] == 'Function Name': ker = r[1].split('(')[0]; continue
            if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
            if 'Instructions Executed' in r: ie = r.index('Instructions Executed'); smp = r.index('# Samples'); continue
            if ie is not None and len(r) > ie and r[0] not in ('', 'Line No') and r[2] == '-' and ker and 'k_view' in ker:
                try: n = int(r[ie]); s_ = int(r[smp]); ln = int(r[0])
                except ValueError: continue
                key = "inlined from other headers (tfmath, foot, intrinsics, atomics)"
                if fname == 'view.cuh':
                    key = "view.cuh helpers (ray_touch, nth_set_bit, exact_cell, cell_rays_inline)"
                    for (nm, a), (_, bb) in zip(marks[:-1], marks[1:]):
                        if a <= ln < bb: key = nm
                b[key] += s_; bi[key] += n
        ts, ti = sum(b.values()) or 1, sum(bi.values()) or 1
        out = ["k_view<0,0>, C4 (128 scenes x 200 robots): share of stall samples / warp instructions by phase (ncu source page, -lineinfo)",
               "total warp instructions (source-attributed): %d = %.0f per robot-step" % (ti, ti / 25600.0)]
        for k, _ in sorted(bi.items(), key=lambda x: -x[1]):
            out.append("  %5.1f%% stalls  %5.1f%% instr  %s" % (100.0 * b[k] / ts, 100.0 * bi[k] / ti, k))
        open(os.path.join(P, "r02_k_view_c4_phases.txt"), "w").write("\n".join(out) + "\n")
    # SASS opcode histogram of the built library
    so = os.path.join(ROOT, "img_env_b200", "libimgenv_b200.so")
    sass = run(["cuobjdump", "-sass", so])
    fn = None; hist = collections.defaultdict(collections.Counter)
    for l in sass.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m: fn = m.group(1); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m and fn: hist[fn][m.group(1).split(".")[0]] += 1
    out = ["cuobjdump -sass img_env_b200/libimgenv_b200.so: arch " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))) +
           "; opcode histogram per kernel (UBLKCP + SYNCS = cp.async.bulk staging + mbarriers in k_view, REDUX = redux.sync; "
           "no UTCMMA/LDTM: nothing on this path is a contraction, BASELINE.json north_star)"]
    for f, c in sorted(hist.items(), key=lambda x: -sum(x[1].values())):
        tot = sum(c.values())
        special = ", ".join("%s %d" % (k, c[k]) for k in ("UBLKCP", "SYNCS", "REDUX", "ATOMS", "ATOMG", "RED") if c[k])
        out.append("%s: %d instructions (%.1f KB)  " % (f, tot, tot * 16 / 1024.0) + ", ".join("%s %d" % kv for kv in c.most_common(14)) + (" | " + special if special else ""))
    open(os.path.join(P, "r02_sass_opcodes.txt"), "w").write("\n".join(out) + "\n")
    print("profiles refreshed;", {k: round(v["bytes_per_robot_step"]) for k, v in traffic.items()})


if __name__ == "__main__":
    main()
