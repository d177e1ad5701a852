#!/usr/bin/env python
"""bench.py — robot-steps/s of the img_env hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c4] [--scenes S]
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one imgenv_step over S scenes x R robots of synthetic input (full State: 48x48
sensor_maps, 3x48x48 ped_maps, lasers, vector states, codes).  `value` = device-resident
throughput (inputs already in HBM), `e2e` = the same through the host-buffer C-ABI call with the
H2D copy of the actions and a D2H read of the per-robot results inside the timed region.
One process per GPU; scenes are independent so ranks share nothing (weak scaling, no collective in
the step); timing = CUDA events, max over ranks.  The reference arm times the UNMODIFIED reference
node (oracle/_ref) + its Python post-processing on the host cores, one scene per process.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "robot-steps/s (48x48 sensor+ped maps, lasers)"
UNIT = "robot-steps/s"

from img_env_b200.scenarios import WORKLOADS, make_cfg, make_resets, random_actions, workload_map  # noqa: E402


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline (TEST INFRASTRUCTURE: the only place bench.py touches oracle/)
# --------------------------------------------------------------------------------------------
def _ref_worker(args):
    wname, seed, warmup, steps, barrier_dir, nproc, rank, target_s = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import cv2
    cv2.setNumThreads(1)
    from img_env_b200.spec import build_spec
    from oracle.pyref import RefEnv, PyPost
    w = WORKLOADS[wname]
    spec = build_spec(make_cfg(w))
    rs = make_resets(spec, w, 1, seed)[0]
    ref = RefEnv(spec); post = PyPost(spec)
    post.on_reset(); post.get_states(ref.reset(rs))
    rng = np.random.default_rng(seed + 99)
    R = spec["R"]
    alive = np.ones(R, np.uint8)
    rb, pd = ref.get_internal()

    core = [0.0]

    def one_step():
        # keep every robot alive (same rule as the GPU arm: clear the collision/arrival flags)
        rb_, pd_ = ref.get_internal()
        rb_[:, 12] = 0; rb_[:, 13] = 0
        ref.set_internal(rb_, None)
        acts = random_actions(R, rng)
        tc = time.time()
        st = ref.step(acts, alive)                 # (A) the C++ node core: ImgEnv::_step incl. get_states
        core[0] += time.time() - tc
        post.get_states(st)                        # (B) adds the restated yaml_env.py post-processing
    for _ in range(warmup):
        one_step()
    # file barrier so all workers time the same interval
    open(os.path.join(barrier_dir, "ready%d" % rank), "w").close()
    t_wait = time.time()
    while len([f for f in os.listdir(barrier_dir) if f.startswith("ready")]) < nproc and time.time() - t_wait < 600:
        time.sleep(0.01)
    core[0] = 0.0
    t0 = time.time()
    done = 0
    while done < steps or (target_s > 0 and time.time() - t0 < target_s):     # target_s: sample sized by time, not by steps
        one_step(); done += 1
    t1 = time.time()
    return t0, t1, R * done, core[0]


def run_reference(wname, steps, warmup, procs=None, budget_s=200.0, target_s=0.0, hard_timeout=None):
    """Times oracle/_ref (+ restated Python post-processing), one scene per process."""
    import multiprocessing as mp
    import tempfile
    from oracle.pyref import have_ref
    if not have_ref():
        return None
    w = WORKLOADS[wname]
    ncores = os.cpu_count() or 1
    if procs is None:
        procs = ncores
        # each reference robot keeps a private full-map clone (agent.h:83): bound host memory
        cells = (w["map_px"] * w["gres"] / 0.015) ** 2
        per_proc = cells * (w["R"] + 8) * 1.1 + 1.5e9
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
            procs = max(1, min(procs, int(0.5 * avail / per_proc)))
        except Exception:
            procs = max(1, min(procs, 4))
    # bound the wall time: estimated seconds per reference scene-step
    est = w["R"] * (0.012 + 1.2e-9 * (w["map_px"] * w["gres"] / 0.015) ** 2 * 0.35) + 0.05
    steps_eff = max(1, min(steps, int(budget_s / est) - warmup))
    warm_eff = max(0, min(warmup, max(0, int(0.25 * budget_s / est))))
    if w["scene"] == "pedscene" and w["R"] >= 9:
        # libpedsim's quadtree recurses without bound once 9 agents share a point: the node's robots are all constructed at
        # (0,0) (pedscene.h:53-56), so the reference crashes on this configuration
        return dict(unavailable="the reference node cannot run pedscene with >= 9 robots (quadtree stack overflow)")
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as bd:
        pool = ctx.Pool(procs)
        try:      # a crashed worker would make a plain map() wait forever
            res = pool.map_async(_ref_worker, [(wname, 1000 + i, warm_eff, steps_eff, bd, procs, i, target_s) for i in range(procs)]).get(
                timeout=hard_timeout or (2.0 * budget_s + 120.0))
        except Exception as e:
            pool.terminate()
            return dict(unavailable="reference workers failed or timed out: %s" % type(e).__name__)
        finally:
            pool.terminate()
    t0 = min(r[0] for r in res); t1 = max(r[1] for r in res)
    total = sum(r[2] for r in res)
    core_s = sum(r[3] for r in res) / len(res)          # mean seconds a worker spent inside the node during the timed steps
    return dict(value=total / (t1 - t0), seconds=t1 - t0, procs=procs, steps=int(round(total / max(1, sum(1 for _ in res)) / w["R"])),
                warmup=warm_eff, robot_steps=total, core_value=total / core_s if core_s > 0 else None)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def measure_gather(sim, step, world, S, R, kg, dist, torch):
    """sim + delivery of all nine State tensors of every rank to rank 0; device-timed, max over ranks."""
    from img_env_b200.parallel import ObservationGatherer
    gat = ObservationGatherer(sim.out, dst=0)
    step(0); gat(sim.out)
    torch.cuda.synchronize(); dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for t in range(kg):
        step(t); gat(sim.out)
    g1.record()
    torch.cuda.synchronize()
    tg = torch.tensor([g0.elapsed_time(g1)], device="cuda"); dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    out = {"value": world * S * R * kg / (float(tg.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(tg.item()) / kg,
           "bytes_to_learner_per_step": int(gat.bytes_to_learner),
           "note": "sim + point-to-point delivery of all nine State tensors of every rank to rank 0 (NCCL isend/irecv "
                   "over NVLink); not part of the headline metric"}
    del gat
    return out


def measure_c5(args, world, rank, local_rank, dist, torch):
    """BASELINE config 5 beside the headline line: device-timed sim-only rate and sim + gather to the learner GPU."""
    from img_env_b200.spec import build_spec
    from img_env_b200.lib import BatchedSim
    w = WORKLOADS["c5"]
    S = w["scenes"]
    spec = build_spec(make_cfg(w, args.synthetic_map))
    R = spec["R"]
    sim = BatchedSim(spec, num_scenes=S, device=local_rank, seed=99 + rank, ped_yaw_mode=1)
    sim.reset(make_resets(spec, w, S, 4321 + rank * 100003))
    rng = np.random.default_rng(77 + rank)
    acts = torch.from_numpy(np.stack([np.stack([random_actions(R, rng) for _ in range(S)]) for _ in range(4)])).cuda()
    alive = torch.ones(S, R, dtype=torch.uint8, device="cuda")

    def step(t):
        sim.revive(); sim.step(acts[t % 4], alive)
    for t in range(3):
        step(t)
    torch.cuda.synchronize(); dist.barrier()
    k5 = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(k5):
        step(t)
    e1.record()
    torch.cuda.synchronize()
    tm = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    out = {"workload": "c5: " + w["desc"], "scenes_per_gpu": S, "scenes_total": S * world, "value": world * S * R * k5 / (float(tm.item()) * 1e-3),
           "unit": UNIT, "ms_per_step": float(tm.item()) / k5, "steps": k5,
           "gather_to_learner": measure_gather(sim, step, world, S, R, 5, dist, torch)}
    sim.close()
    return out


def source_sha():
    """Hash of the CUDA sources the library is built from: keys measured-traffic records to the code they were taken on."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "img_env_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def run_b200(args):
    import torch
    from img_env_b200.build import build
    from img_env_b200.spec import build_spec
    from img_env_b200.lib import BatchedSim
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if rank == 0:
        build()
    if dist:
        dist.barrier()
    w = WORKLOADS[args.workload]
    S = args.scenes or w["scenes"]
    spec = build_spec(make_cfg(w, args.synthetic_map))
    R = spec["R"]
    sim = BatchedSim(spec, num_scenes=S, device=local_rank, seed=1234 + rank, ped_yaw_mode=1)
    sim.reset(make_resets(spec, w, S, 1234 + rank * 100003))
    rng = np.random.default_rng(4321 + rank)
    T = 8
    acts_np = np.stack([np.stack([random_actions(R, rng) for _ in range(S)]) for _ in range(T)])   # [T,S,R,3]
    acts = torch.from_numpy(acts_np).cuda()
    alive = torch.ones(S, R, dtype=torch.uint8, device="cuda")
    K, W = args.steps, max(args.warmup, 3)
    launches_per_step = sim.launches_per_step + 1     # + k_revive

    def step(t):
        sim.revive()                                   # all robots alive each timed step
        sim.step(acts[t % T], alive)

    for t in range(W):
        step(t)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sim.profile_begin(K)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(K):
        step(W + t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nprof, kms = sim.profile_end()
    # an ORCA agent that found neither a table slot nor a pool slab would keep only its nearest obstacle edges (the reference
    # keeps all): a number measured with such a truncation is not a parity-grade number
    overflows = sim.debug_counters()[0]
    if overflows:
        raise RuntimeError("bench: %d ORCA table overflows during the timed run" % overflows)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        tms = torch.tensor([ms], device="cuda"); dist.all_reduce(tms, op=dist.ReduceOp.MAX); ms = float(tms.item())
        dist.barrier()

    # ---- e2e: host buffers through the C-ABI (imgenv_step_host), per-step D2H of the per-robot results ----
    pin_acts = torch.from_numpy(acts_np).pin_memory()
    pin_alive = torch.ones(S, R, dtype=torch.uint8).pin_memory()
    small = ["vector_states", "is_collisions", "is_arrives", "step_ds", "ped_min_dists"]
    host_out = {k: torch.empty_like(sim.out[k], device="cpu").pin_memory() for k in small}
    h2d = pin_acts[0].numel() * 4 + pin_alive.numel()
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    def step_e2e(t):
        sim.revive()
        sim.step_host_ptr(pin_acts[t % T].data_ptr(), pin_alive.data_ptr())
        for k in small:
            host_out[k].copy_(sim.out[k], non_blocking=True)
        torch.cuda.synchronize()                       # the host consumes dones/rewards before the next action
    for t in range(3):
        step_e2e(t)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for t in range(K):
        step_e2e(t)
    e2e_s = time.perf_counter() - t0
    if dist:
        tt = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_s = float(tt.item())
    # full State to pinned host memory as well (what a CPU-side learner would need): reported, not the headline
    full_host = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in sim.out.items()}
    d2h_full = sum(v.numel() * v.element_size() for v in full_host.values())

    def step_full(t):
        sim.revive()
        sim.step_host_ptr(pin_acts[t % T].data_ptr(), pin_alive.data_ptr())
        for k, v in full_host.items():
            v.copy_(sim.out[k], non_blocking=True)
        torch.cuda.synchronize()
    step_full(0)
    kf = max(2, min(K, 10))
    t0 = time.perf_counter()
    for t in range(kf):
        step_full(t)
    full_s = time.perf_counter() - t0
    if dist:
        tt = torch.tensor([full_s], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); full_s = float(tt.item())

    # ---- N > 1: the same steps with the State of every rank delivered to the learner GPU (rank 0) over NVLink,
    #      on this workload and on BASELINE config 5 (4096 SFM scenes x (64 + 64) sharded over 8 GPUs = 512 per GPU)
    gather = None
    c5 = None
    B, range_total, pvs_len, ped_img = sim.bytes_per_robot_step, sim.range_total, sim.pvs_len, sim.spec["ped_image_size"][0]
    if dist:
        gather = measure_gather(sim, step, world, S, R, max(2, min(K, 10)), dist, torch)
        if args.workload != "c5" and not args.no_c5:
            sim.close(); del sim
            torch.cuda.empty_cache()
            c5 = measure_c5(args, world, rank, local_rank, dist, torch)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    total_robot_steps = world * S * R * K
    value = total_robot_steps / (ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    view_ms = kms["k_view"]
    achieved = (B * S * R) / (view_ms * 1e-3) / 1e9 if view_ms > 0 else None
    # bytes k_view itself moves: everything except the pedestrian observation, which k_ped_obs writes concurrently
    B_ped = 3 * ped_img ** 2 * 4 + 4 * pvs_len + 4
    traffic = None
    try:   # measured per robot-step by ncu (tools/profile_summary.py --traffic): only valid for the sources it was taken on
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tj and tj.get("source_sha") == source_sha() and not args.synthetic_map:
            traffic = float(tj["bytes_per_robot_step"]) * S * R
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "scenes_per_gpu": S, "robots_per_scene": R, "peds_per_scene": spec["P"],
                   "grid": list(spec["grid"].shape), "view": "400x400@0.015", "range_total": range_total,
                   "map": workload_map(w, args.synthetic_map)[1],
                   "l2_policy": "working set (State outputs + footprint records, %.2f GB) larger than the 126 MB L2" % (S * R * B / 1e9),
                   "all_robots_alive": True, "parallelism": "scenes sharded, %d per GPU, no collective in the step" % S},
        "e2e": {"value": world * S * R * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "note": "imgenv_step_host from pinned host actions; vector_states/codes/step_ds/ped_min_dists read back and "
                        "synchronised every step; sensor_maps/ped_maps/lasers stay device-resident for the learner"},
        "e2e_full_state_to_host": {"value": world * S * R * kf / full_s, "unit": UNIT, "d2h_bytes_per_step": int(d2h_full),
                                   "note": "the same call with ALL nine State tensors copied to pinned host memory every step (what a "
                                           "CPU-side learner would need); PCIe-bound"},
        "gpu_launches": int(K * launches_per_step),
        "solver_table_overflows": int(overflows),
        "kernel_ms": {k: round(v, 4) for k, v in kms.items()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "kernel": "k_view", "bytes_per_robot_step": B, "robot_steps_per_launch": S * R,
                     "kernel_own_bytes_per_robot_step": B - B_ped,
                     "step_frac": (B * S * R) / (ms / K * 1e-3) / 1e9 / peak,
                     "note": "achieved = SURVEY 8(d) bytes per robot-step x robots per launch / k_view's CUDA-event time; the pedestrian "
                             "part of those bytes is written by k_ped_obs, which runs concurrently on the side stream inside that window; "
                             "step_frac = the same bytes over the whole step time (all kernels)",
                     "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s"},
        "clocks": clocks,
    }
    if gather:
        out["gather_to_learner"] = gather
    if c5:
        out["c5"] = c5
    if world == 1 and not args.no_cpu_baseline:
        try:
            ref = run_reference(args.workload, steps=2, warmup=1, budget_s=25.0 * 8, target_s=12.0, hard_timeout=240.0)
            if ref and "unavailable" in ref:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": ref["unavailable"]}
            elif ref:
                out["cpu_baseline"] = {"value": ref["value"], "unit": UNIT, "cores": ref["procs"], "kind": "reference",
                                       "sample": "%d scene(s) of this workload (one per process, oracle/_ref = unmodified reference node + "
                                                 "restated Python post-processing), %d timed step(s), %.1f s" %
                                                 (ref["procs"], ref["steps"], ref["seconds"]),
                                       "node_core_value": ref["core_value"]}
        except Exception as e:   # the baseline is informative; never lose the GPU line
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    ref = run_reference(args.workload, steps=args.steps, warmup=args.warmup, budget_s=200.0)
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libimgenv_ref.so is not built"}))
        return
    if "unavailable" in ref:
        print(json.dumps({"impl": "reference", "unavailable": ref["unavailable"]}))
        return
    out = {"impl": "reference", "metric": METRIC, "value": ref["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": ref["steps"],
           "steps_requested": args.steps, "warmup": ref["warmup"], "ms_per_step": 1e3 * ref["seconds"] / ref["steps"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload + ": " + w["desc"], "scenes": ref["procs"], "robots_per_scene": w["R"],
                      "peds_per_scene": w["P"], "note": "bounded sample: one scene per host process (the reference's own parallelism)"},
           "cpu_baseline": {"value": ref["value"], "unit": UNIT, "cores": ref["procs"], "kind": "reference",
                            "sample": "%d scene(s), %d timed step(s), %.1f s" % (ref["procs"], ref["steps"], ref["seconds"]),
                            "node_core_value": ref["core_value"],
                            "note": "value = end-to-end State (C++ node + restated yaml_env.py post-processing, BASELINE.md 3B); "
                                    "node_core_value = the C++ node alone (3A)"},
           "e2e": {"value": ref["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--scenes", type=int, default=0, help="scenes per GPU (default: workload-specific)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true", help="(kept for compatibility: N > 1 always reports sim + delivery of the State to rank 0)")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the BASELINE config 5 sub-record")
    ap.add_argument("--synthetic-map", action="store_true", help="round-1 synthetic room with random blocks instead of the reference's PNG")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
