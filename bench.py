#!/usr/bin/env python
"""bench.py -- ADMM iterations per second on the synthetic tetrahedralised cube (BASELINE.json configs[4]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cube 55] [--impl reference]

One "step" is one System::step() frame = `admm_iters` (10) ADMM iterations of the hot path (local step for
every force + global step) on the N^3 Kuhn cube with NeoHookean tets (mu = lambda = 1e5, max_iterations 5,
density-weighted mass 1000, gravity, dt 0.04, x1.3 stretch excitation) -- SURVEY.md 8(d) input 5.
  value   whole-job ADMM iterations/s with x, v resident in HBM (admmb_step_resident), CUDA events on the
          library's stream, max over ranks;
  e2e     the same through the reference-facing call admmb_step() with HOST x/v buffers (pinned staging,
          H2D + D2H inside the timed region), host wall clock;
  N > 1   scene ensemble: every rank steps its own copy of the scene on its own GPU, no data-path collective
          ("scaling": "weak"); torch.distributed (NCCL) only carries the barrier and the max-over-ranks.
  --impl reference   the UNMODIFIED reference (oracle/_ref, Eigen + OpenMP on the host cores) on a bounded sample
          of the same workload (a smaller cube), scaled to the metric's unit by tet count (see "sample").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "admm-elastic-sca_b200", "pyhost"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "admm_iterations_per_s_cube_1M_tets"
UNIT = "ADMM iterations/s"
ADMM_ITERS = 10
FULL_TETS = 998250  # N = 55
# The reference's line search only reaches its steady regime (maxfev evaluations in almost every tet) after ~20 frames of
# the x1.3 stretch excitation: these frames are ALWAYS run, untimed, before the --warmup frames, on every arm, so that
# `value` does not depend on --warmup.
CONDITION_FRAMES = 20
SHIPPED = ("bunnyexpand", "windyflag", "poordillo", "plinkopony")
LOCAL_BYTES_PER_TET = 608.0   # idx 16 + x gather 96 + B 96 + u r/w 144 + z w 72 + state r/w 64 + weights 24 + P w 96


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The sampler is started
    before the warm-up (nvidia-smi needs a few hundred ms to come up -- longer than a short timed region) and every row
    carries nvidia-smi's own timestamp, so only the rows that fall inside [begin(), end()] are used."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.time() + 5.0          # nvidia-smi takes a few hundred ms to deliver its first row
            while not self.rows and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def begin(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)   # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        parsed = []
        for r in self.rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
                parsed.append((ts, float(r[1]), float(r[2]), r))
            except Exception:
                continue
        pad = datetime.timedelta(milliseconds=30)
        inside = [p for p in parsed if self.t0 is not None and self.t1 is not None and self.t0 - pad <= p[0] <= self.t1 + pad]
        where = "timed region"
        if not inside and parsed and self.t1 is not None:
            # region shorter than the sampling period: the samples closest to it (the GPU is under the same load in the warm-up)
            inside = sorted(parsed, key=lambda p: abs((p[0] - self.t1).total_seconds()))[:3]
            where = "nearest samples (region shorter than the 50 ms sampling period)"
        sm = [p[1] for p in inside]
        mx = inside[-1][2] if inside else None
        reasons = set()
        for p in inside:
            r = p[3]
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": where}


def make_scene(N):
    import scenes
    return scenes.cube_scene(N, kind=scenes.TET_NH, mu=1e5, lam=1e5, maxit=5, mass=1000.0, dt=0.04, iters=ADMM_ITERS, stretch=1.3)


def load_workload(name):
    """(scene dict, label) for 'cubeN' or one of the reference's shipped scenes (fixtures exported by the reference's own
    scene layer: tests/golden/shipped_*.scene.npz, oracle/scene_export.cpp)."""
    import scenes
    if name.startswith("cube"):
        N = int(name[4:])
        sc = make_scene(N)
        return sc, f"cube N={N} ({sc['batches'][0]['idx'].shape[0]} tets) NeoHookean mu=lambda=1e5 max_iterations=5, {sc['iters']} ADMM iterations per step, x1.3 stretch"
    sc = scenes.load_scene(os.path.join(ROOT, "tests", "golden", f"shipped_{name}.scene.npz"))
    nel = sum(len(b["idx"]) for b in sc["batches"] if "idx" in b and b["type"] in ("tets", "tris"))
    return sc, f"samples/{name} ({nel} elements, {sc['x'].shape[0]} nodes), {sc['iters']} ADMM iterations per step, dt {sc['dt']}, no user interaction"


def reference_rate(sc, frames, cond, threads=None):
    """ADMM iterations/s of the UNMODIFIED reference (oracle/_ref: Eigen + OpenMP on all host cores) on scene `sc`:
    `cond` untimed frames, then `frames` timed with std::chrono inside the shim (System::step only)."""
    from oracle import ref
    cores = threads or os.cpu_count() or 1
    ref.lib().ref_set_omp_threads(cores)
    t0 = time.perf_counter()
    sim = ref.RefSystem(sc, probe=False)
    t_init = time.perf_counter() - t0
    if "x_after_init" in sc:
        sim.set_x(sc["x_after_init"])
    for _ in range(cond):
        sim.step()
    sec = sim.step_timed(frames)
    stats = sim.stats()
    sim.close()
    return {"it_s": frames * sc["iters"] / sec, "ms_per_frame": 1e3 * sec / frames, "init_s": t_init, "cores": cores, "L_nnz": stats["L_nnz"],
            "frames": frames, "conditioning_frames": cond}


def device_rate(sc, device, frames, cond):
    """ADMM iterations/s of the CUDA path on scene `sc`: resident (CUDA events on the library's stream) and end to end
    through admmb_step with page-locked host x / v (host clock), after `cond` untimed frames."""
    import admm_b200
    import torch
    t0 = time.perf_counter()
    sim = admm_b200.System(sc, device=device, pin_host=True)
    t_setup = time.perf_counter() - t0
    if "x_after_init" in sc:
        sim.set_x(sc["x_after_init"])
    sim.upload()
    sim.step_resident(frames=cond)
    l0 = sim.info()["launches_total"]
    sim.step_resident(frames=frames)
    ms = sim.last_region_ms()
    l1 = sim.info()["launches_total"]
    sim.download()
    sim.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(frames):
        sim.step()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    n3 = sim.n3
    sim.close()
    return {"it_s": frames * sc["iters"] / (ms * 1e-3), "ms_per_frame": ms / frames, "e2e_it_s": frames * sc["iters"] / sec, "e2e_ms_per_frame": 1e3 * sec / frames,
            "launches_per_frame": (l1 - l0) / frames, "setup_s": t_setup, "frames": frames, "conditioning_frames": cond,
            "h2d_bytes_per_step": 2 * n3 * 8, "d2h_bytes_per_step": 2 * n3 * 8}


PAIR_WORKLOADS = SHIPPED + ("cube30",)


def pair_frames(name):
    return (8, CONDITION_FRAMES) if name.startswith("cube") else (60, 10)


def same_config_pairs(device, which):
    """Same workload, same frame counts, both implementations, in this run: the reference on the host cores and the CUDA
    path on the device.  These are the ratios that are NOT extrapolated (the reference cannot even initialise the 1 M-tet
    headline mesh in bench time: ~44 min)."""
    from oracle import ref
    out = {}
    for name in which:
        sc, label = load_workload(name)
        frames, cond = pair_frames(name)
        d = device_rate(sc, device, frames, cond)
        row = {"workload": label, "b200_it_s": d["it_s"], "b200_e2e_it_s": d["e2e_it_s"], "b200_launches_per_frame": d["launches_per_frame"],
               "frames": frames, "conditioning_frames": cond}
        if ref.available():
            r = reference_rate(sc, frames, cond)
            row.update({"reference_it_s": r["it_s"], "reference_cores": r["cores"], "reference_init_s": r["init_s"], "b200_setup_s": d["setup_s"],
                        "ratio_resident": d["it_s"] / r["it_s"], "ratio_e2e": d["e2e_it_s"] / r["it_s"], "same_config": True})
        out[name] = row
    return out


def run_reference(args, rank, world):
    """Times the unmodified reference on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libadmm_ref.so missing (build it where /root/reference exists)"}))
        return
    if args.scene:
        # same workload as the b200 arm of --scene: a true same-config line
        sc, label = load_workload(args.scene)
        r = reference_rate(sc, args.steps, max(args.warmup, 10))
        line = {"impl": "reference", "metric": f"admm_iterations_per_s_{args.scene}", "value": r["it_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "shipped scene" if args.scene in SHIPPED else "synthetic", "config": {"workload": label},
                "cpu_baseline": {"value": r["it_s"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": f"the whole workload, {args.steps} frames; initialize() {r['init_s']:.2f} s not included"},
                "e2e": {"value": r["it_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    N = args.ref_cube
    sc = make_scene(N)
    ntets = sc["batches"][0]["idx"].shape[0]
    r = reference_rate(sc, args.steps, CONDITION_FRAMES + args.warmup)
    its = r["it_s"]
    value = its * ntets / FULL_TETS  # scaled by tet count to the 1M-tet unit (generous: the reference's cost grows superlinearly)
    sample = (f"cube N={N} ({ntets} tets) stepped {args.steps} frames x {ADMM_ITERS} iterations after {CONDITION_FRAMES + args.warmup} untimed frames = {its:.2f} it/s measured, "
              f"scaled x{ntets}/{FULL_TETS} to the 1M-tet unit (the reference needs ~44 min to initialise the 1M-tet mesh itself); initialize() {r['init_s']:.1f} s not included; L nnz {r['L_nnz']}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cube N={args.cube} NeoHookean, {ADMM_ITERS} ADMM iterations per step (reference timed on N={N}, EXTRAPOLATED by tet count; "
                                   f"measured same-config pairs: `same_config_pairs` of the b200 line)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "measured_same_config": {f"cube{N}": {"workload": load_workload(f"cube{N}")[1], "reference_it_s": its, "frames": args.steps,
                                                  "conditioning_frames": CONDITION_FRAMES + args.warmup, "cores": r["cores"]}}}
    if not args.no_pairs:
        for name in SHIPPED:
            scs, label = load_workload(name)
            frames, cond = pair_frames(name)
            rr = reference_rate(scs, frames, cond)
            line["measured_same_config"][name] = {"workload": label, "reference_it_s": rr["it_s"], "frames": frames, "conditioning_frames": cond, "cores": rr["cores"]}
    print(json.dumps(line))


def committed_traffic(cube):
    """DRAM bytes per launch (local-step kernel) and per solve (its 26 launches) from the ncu --set full captures of the
    round's final build committed under profiles/ (r2_local.txt, r2_solve.txt: cube N=55).  Not measured in this run -- ncu
    cannot run inside the timed bench -- so the values carry their source, and any other workload gets None."""
    out = {"local": None, "solve": None}
    if cube != 55:
        return out
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    try:
        rd = wr = None
        for ln in open(os.path.join(root, "r2_local.txt")):
            t = ln.split()
            if len(t) == 2 and t[0] == "dram__bytes_read.sum":
                rd = float(t[1])
            if len(t) == 2 and t[0] == "dram__bytes_write.sum":
                wr = float(t[1])
        if rd is not None and wr is not None:
            out["local"] = (rd + wr) * 1e6
    except (OSError, ValueError):
        pass
    try:
        import re
        m = re.search(r"DRAM traffic ([0-9.]+) MB", open(os.path.join(root, "r2_solve.txt")).read())
        if m:
            out["solve"] = float(m.group(1)) * 1e6
    except (OSError, ValueError):
        pass
    return out


def cpu_baseline_leg(N):
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    sc = make_scene(N)
    ntets = sc["batches"][0]["idx"].shape[0]
    frames = 8
    r = reference_rate(sc, frames, CONDITION_FRAMES)
    # SURVEY 8(d) asks for OMP_NUM_THREADS=1 beside all cores: a smaller sample of the same workload keeps it to a few seconds
    N1 = min(N, 12)
    sc1 = make_scene(N1)
    ntets1 = sc1["batches"][0]["idx"].shape[0]
    r1 = reference_rate(sc1, frames, CONDITION_FRAMES, threads=1)
    return {"value": r["it_s"] * ntets / FULL_TETS, "unit": UNIT, "cores": r["cores"], "kind": "reference",
            "sample": f"unmodified reference (Eigen + OpenMP, {r['cores']} threads) on cube N={N} ({ntets} tets): {r['it_s']:.2f} it/s over {frames} frames after "
                      f"{CONDITION_FRAMES} untimed frames, scaled x{ntets}/{FULL_TETS}; initialize() {r['init_s']:.1f} s excluded",
            "one_thread": {"value": r1["it_s"] * ntets1 / FULL_TETS, "unit": UNIT, "cores": 1,
                           "sample": f"the same with 1 thread on cube N={N1} ({ntets1} tets): {r1['it_s']:.2f} it/s over {frames} frames, scaled x{ntets1}/{FULL_TETS}"}}


def algorithmic_flops_per_tet(sim, sc, sample=4096, seed=12345):
    """Algorithmic FP64 flops per tet and ADMM iteration of the local step IN THE REGIME JUST TIMED: a random sample of the
    run's tets -- current positions, duals and optimiser state downloaded from the device -- goes through the kernels'
    per-force body compiled for the host with flop counters (csrc/flopcount.cpp; add / mul / div / sqrt / log = 1)."""
    import admm_b200
    b = sc["batches"][0]
    T = b["idx"].shape[0]
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(T, size=min(sample, T), replace=False))
    sim.download()
    u = sim.u.reshape(T, 9)[pick]
    st = sim.prox_state()[pick]
    flops, evals, its = admm_b200.flopcount_hyper_tets(int(b["kind"]), sc["x"], b["idx"][pick], b["p0"], b["p1"], b["maxit"], sc["dt"], sim.m_x, u, st)
    return {"flops_per_tet_iteration": flops / len(pick), "objective_evaluations_per_tet_iteration": evals / len(pick),
            "lbfgs_iterations_per_tet_iteration": its / len(pick), "sampled_tets": int(len(pick))}


def run_ensemble(args, rank, world, local_rank, torch, dist, admm_b200):
    """Scene ensemble: independent scenes, one context (and stream, and CUDA graph) each, round-robin over the ranks, no
    data-path collective.  All scenes of a rank are enqueued before any is waited for, so they overlap on the GPU."""
    import ensemble
    import scenes
    mine = ensemble.scene_shard(args.ensemble, rank, world)
    sims = []
    t0 = time.perf_counter()
    for sidx in mine:
        # per-copy seed: every scene starts from its own perturbed stretch ("per-copy seed")
        sc = scenes.cube_scene(args.ens_cube, kind=scenes.TET_NH, mu=1e5, lam=1e5, maxit=5, mass=1000.0, dt=0.04, iters=ADMM_ITERS,
                               stretch=1.3, seed=1000 + sidx)
        sim = admm_b200.System(sc, device=local_rank, pin_host=True)
        sim.set_x(sc["x_after_init"])
        sim.upload()
        sims.append(sim)
    t_setup = time.perf_counter() - t0
    ntets = sims[0].scene["batches"][0]["idx"].shape[0] if sims else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(frames):
        for s in sims:
            s.step_resident_async(frames=frames)
        for s in sims:
            s.sync()

    sampler = ClockSampler(local_rank)
    sampler.start()
    run(args.warmup)
    barrier()
    l0 = sum(s.info()["launches_total"] for s in sims)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record()                       # idle stream: completes at once -> a device timestamp of "now"
    run(args.steps)                    # the timed region: K frames of every scene of this rank
    ev1.record()
    torch.cuda.synchronize()
    sampler.end()
    ms_region = ev0.elapsed_time(ev1)
    l1 = sum(s.info()["launches_total"] for s in sims)
    clocks = sampler.stop()
    barrier()
    # end to end: every scene through admmb_step_async with (page-locked) host buffers: x, v in and out every frame, all
    # scenes of the rank in flight, collected with admmb_sync before the next frame
    for s in sims:
        s.download()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for s in sims:
            s.step_async()
        for s in sims:
            s.sync()
    torch.cuda.synchronize()
    sec_e2e = time.perf_counter() - t0
    barrier()
    ms_max, frames_total = ensemble.reduce_job(ms_region, args.steps * len(sims))
    ms_e2e_max, _ = ensemble.reduce_job(sec_e2e * 1e3, args.steps * len(sims))
    if rank == 0:
        nverts = sims[0].n3 // 3
        line = {
            "metric": "ensemble_scene_frames_per_s", "value": frames_total / (ms_max * 1e-3), "unit": "scene-frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.ensemble} independent scenes: cube N={args.ens_cube} ({ntets} tets, {nverts} nodes) NeoHookean mu=lambda=1e5 "
                                   f"max_iterations=5, {ADMM_ITERS} ADMM iterations per frame, per-copy seed",
                       "parallelism": f"scenes round-robin over {world} rank(s), {len(sims)} scenes in flight per GPU on their own streams, no collective",
                       "l2": "64 scenes x (factor + force arrays) exceed the L2 together; scenes interleave, no flush"},
            "admm_iterations_per_s": frames_total * ADMM_ITERS / (ms_max * 1e-3),
            "e2e": {"value": frames_total / (ms_e2e_max * 1e-3), "unit": "scene-frames/s", "h2d_bytes_per_step": len(sims) * 2 * 3 * nverts * 8,
                    "d2h_bytes_per_step": len(sims) * 2 * 3 * nverts * 8, "note": "admmb_step_async + admmb_sync per scene and frame, host x / v in and out every frame (page-locked), scenes overlap"},
            "gpu_launches": int(l1 - l0), "clocks": clocks, "setup": {"seconds": t_setup, "scenes_on_rank0": len(sims)},
        }
        print(json.dumps(line))
    for s in sims:
        s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5, help="untimed frames after the fixed 20 conditioning frames")
    ap.add_argument("--cube", type=int, default=55, help="cube resolution N (N=55: 998,250 tets)")
    ap.add_argument("--scene", default=None, choices=list(SHIPPED),
                    help="time one of the reference's shipped scenes instead of the cube (both --impl arms run the SAME workload)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-cube", type=int, default=30, help="cube resolution of the reference arm's bounded sample")
    ap.add_argument("--cpu-cube", type=int, default=20, help="cube resolution of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-region", action="store_true",
                    help="bracket the timed region with cudaProfilerStart / cudaProfilerStop: `ncu --profile-from-start off` then only sees (and only "
                         "serialises) the kernels of the timed region, so the conditioning frames run at full speed")
    ap.add_argument("--no-pairs", action="store_true", help="skip the same-config reference / b200 pairs (shipped scenes, cube N=30)")
    ap.add_argument("--pairs", default=",".join(PAIR_WORKLOADS), help="comma-separated workloads of the same-config pairs")
    ap.add_argument("--ensemble", type=int, default=0,
                    help="scene-ensemble mode (BASELINE configs[4] '64-scene ensemble'): this many independent scenes (cube --ens-cube) dealt "
                         "round-robin to the ranks, all scenes of a rank in flight at once on their own streams; prints scene-frames/s")
    ap.add_argument("--ens-cube", type=int, default=20, help="cube resolution of the ensemble scenes (N=20: 48,000 tets)")
    ap.add_argument("--solver", default="direct", choices=["direct", "pcg"])
    ap.add_argument("--partition", default="auto", choices=["auto", "off", "only", "pcg"],
                    help="N > 1: besides the scene ensemble (`value`), time ONE mesh partitioned over the ranks (local step partitioned, right-hand "
                         "side all-gathered, direct solve sharded by elimination subtrees) and report it as `partition` (strong scaling); "
                         "'pcg' = partitioned Jacobi-PCG rows instead of the replicated direct solve")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import admm_b200
    if world > 1:
        # every rank factors its own scene on the host: share the cores instead of oversubscribing them (torchrun
        # exports OMP_NUM_THREADS=1, which would leave all but one core per rank idle during setup)
        admm_b200.lib().admmb_set_host_threads(max(1, (os.cpu_count() or 1) // world))
    if args.ensemble > 0:
        run_ensemble(args, rank, world, local_rank, torch, dist, admm_b200)
        return
    if args.scene:
        sc, label = load_workload(args.scene)
        metric = f"admm_iterations_per_s_{args.scene}"
    else:
        sc, label = load_workload(f"cube{args.cube}")
        metric = METRIC
    iters_per_frame = int(sc["iters"])
    hyper = (not args.scene)
    ntets = sc["batches"][0]["idx"].shape[0]
    nverts = sc["x"].shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(dist_arg=None, solver=None):
        t0 = time.perf_counter()
        sim = admm_b200.System(sc, device=local_rank, solver=solver if solver is not None else (admm_b200.SOLVER_DIRECT if args.solver == "direct" else admm_b200.SOLVER_PCG),
                               cg_tol=1e-10, dist=dist_arg, pin_host=True)
        t = time.perf_counter() - t0
        if "x_after_init" in sc:
            sim.set_x(sc["x_after_init"])
        sim.upload()
        return sim, t

    def timed_resident(sim, sampler=None):
        """CONDITION_FRAMES + --warmup untimed frames, then EXACTLY --steps frames between barriers; CUDA events on the library's stream."""
        sim.step_resident(frames=CONDITION_FRAMES)
        sim.step_resident(frames=max(args.warmup, 3))
        barrier()
        l0 = sim.info()["launches_total"]
        cg0 = sim.info()["cg_iters_total"]
        if sampler:
            sampler.begin()
        if args.profile_region:
            torch.cuda.cudart().cudaProfilerStart()
        sim.step_resident(frames=args.steps)
        if args.profile_region:
            torch.cuda.cudart().cudaProfilerStop()
        if sampler:
            sampler.end()
        ms = sim.last_region_ms()
        l1 = sim.info()["launches_total"]
        timed_resident.cg_per_iteration = (sim.info()["cg_iters_total"] - cg0) / float(args.steps * iters_per_frame)
        barrier()
        return ms, l1 - l0

    import ensemble  # whole-job figures: MAX over ranks of the timed region, SUM over ranks of the units

    # ---- resident path: `value` (N > 1: one replica of the scene per GPU, no data-path collective) ------------------------
    sim, t_setup = build()
    info0 = sim.info()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_region, launches = timed_resident(sim, sampler)
    cg_per_iteration = timed_resident.cg_per_iteration
    clocks = sampler.stop()
    # per-phase split, same regime (the optimiser reached its steady state in the conditioning frames): CUDA events around
    # the three phases of every iteration, as many frames as the timed region
    sim.timing(True)
    sim.timing_read(reset=True)
    cg_before_phases = sim.info()["cg_iters_total"]
    sim.step_resident(frames=args.steps)
    phases = sim.timing_read(reset=True)
    cg_phases_pass = (sim.info()["cg_iters_total"] - cg_before_phases) / float(max(phases["iters"], 1))
    sim.timing(False)
    flops = algorithmic_flops_per_tet(sim, sc) if (hyper and rank == 0) else None

    # ---- end to end through admmb_step with host buffers: `e2e` ---------------------------------------------------
    sim.download()
    for _ in range(3):
        sim.step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.step()
    torch.cuda.synchronize()
    sec_e2e = time.perf_counter() - t0
    barrier()
    ms_region_max, total_iters = ensemble.reduce_job(ms_region, args.steps * iters_per_frame)
    ms_e2e_max, _ = ensemble.reduce_job(sec_e2e * 1e3, args.steps * iters_per_frame)
    value = total_iters / (ms_region_max * 1e-3)
    e2e_value = total_iters / (ms_e2e_max * 1e-3)
    sim.close()

    # ---- N > 1: the same mesh ONCE, partitioned over the ranks (strong scaling) ------------------------------------------
    partition = None
    if world > 1 and args.partition != "off" and not args.scene:
        holder = [admm_b200.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        psolver = admm_b200.SOLVER_PCG if args.partition == "pcg" else admm_b200.SOLVER_DIRECT
        psim, pt_setup = build(dist_arg=(rank, world, holder[0]), solver=psolver)
        pms, planches = timed_resident(psim)
        psim.timing(True)
        psim.timing_read(reset=True)
        psim.step_resident(frames=max(2, args.steps // 2))
        pph = psim.timing_read(reset=True)
        psim.timing(False)
        pms_max, _ = ensemble.reduce_job(pms, 0)
        pit = max(pph["iters"], 1)
        partition = {"value": args.steps * iters_per_frame / (pms_max * 1e-3), "unit": UNIT, "scaling": "strong", "ms_per_step": pms_max / args.steps,
                     "mode": "local step partitioned by elements; " + ("Jacobi-PCG rows partitioned, NCCL all-gather + all-reduce per CG iteration" if args.partition == "pcg"
                                                                      else "right-hand side all-gathered (ncclAllGather of 3n doubles per ADMM iteration); direct solve sharded by subtrees of the elimination tree "
                                                                           "(each rank streams its subtrees' factor tiles + the replicated top separators; one small all-reduce of the top rows, one all-reduce of x)"),
                     "phases_ms_per_iteration_rank0": {"local": pph["local_ms"] / pit, "rhs_and_allgather": pph["rhs_ms"] / pit, "solve_and_allreduce": pph["solve_ms"] / pit},
                     "factor_bytes_rank0": psim.info()["factor_bytes"], "factor_bytes_single_gpu": info0["factor_bytes"],
                     "gpu_launches_rank0": int(planches), "replica_it_s_per_gpu": value / world}
        psim.close()

    if rank == 0:
        pk = peaks()
        hbm_peak = (pk or {}).get("hbm_gbs", 6650.0)
        peak_src = "MEASURED_PEAKS.json hbm_gbs (copy, of measured)" if pk else "fallback 6650 GB/s (B200_PROFILING.md, of fallback)"
        iters = max(phases["iters"], 1)
        solve_ms = phases["solve_ms"] / iters
        local_ms = phases["local_ms"] / iters
        rhs_ms = phases["rhs_ms"] / iters
        phase_sum = local_ms + rhs_ms + solve_ms
        step_ms = ms_region_max / args.steps / iters_per_frame
        traffic = committed_traffic(args.cube if (not args.scene and args.solver == "direct") else None)
        # global step: k_solve_level streams the packed factor once forward and once backward
        solve_bytes = info0["factor_bytes"] + 9 * nverts * 8 * 3  # factor tiles + b, y, x vectors read/written
        solve_gbs = solve_bytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        roof_global = {"bound": "hbm", "kernel": f"{'k_solve_level_pf' if os.environ.get('ADMMB_SOLVE_MODE') == '3' else 'k_solve_level_tma'} ({2 * info0['n_levels']} launches per solve, programmatic dependent launch chain)", "achieved": solve_gbs, "peak": hbm_peak, "unit": "GB/s",
                       "frac": solve_gbs / hbm_peak, "traffic": traffic["solve"],
                       "traffic_note": "bytes per solve (all its launches), dram__bytes_read + write of the committed ncu --set full capture profiles/r2_solve.txt (final build of the round, same workload); algorithmic bytes per solve: %d" % int(solve_bytes) if traffic["solve"] else None,
                       "peak_source": peak_src, "share_of_step": solve_ms / phase_sum,
                       "note": f"algorithmic bytes per solve = packed factor, both copies ({info0['factor_bytes']} B) + 9 vector passes of 3n doubles; {info0['n_levels']} levels; "
                               "ncu dram traffic of the same launches: profiles/ (1.03x the algorithmic bytes at N=55)"}
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_region_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic" if not args.scene else "shipped scene",
            "config": {"workload": label + f"; solver={args.solver}",
                       "conditioning": f"{CONDITION_FRAMES} untimed frames before the --warmup frames on every arm (the reference's line search needs ~20 frames to reach "
                                       "its steady regime: maxfev evaluations in almost every tet)",
                       "cg_iterations_per_admm_iteration": cg_per_iteration if args.solver == "pcg" else None,
                       "parallelism": f"ensemble x{world} (one scene per GPU, no collective); the single mesh partitioned over the same GPUs is `partition`" if world > 1 else "single GPU",
                       "l2": "working set (factor %.2f GB + force arrays %.2f GB) exceeds the 126 MB L2, no flush needed" % (
                           info0["factor_bytes"] / 1e9, LOCAL_BYTES_PER_TET * ntets / 1e9) if not args.scene else
                             "small scene: the working set fits the L2 by nature (what the reference's users run); no flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * 3 * nverts * 8, "d2h_bytes_per_step": 2 * 3 * nverts * 8,
                    "ms_per_step": ms_e2e_max / args.steps,
                    "note": "admmb_step(iters, m_x, m_v): host x and v in and out every step, buffers page-locked once with "
                            "admmb_register_host_buffer (pinned host memory, as the contract asks)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "phases_ms_per_iteration": {"local": local_ms, "rhs": rhs_ms, "solve": solve_ms, "sum": phase_sum, "step": step_ms,
                                        "note": "phases: CUDA events around each phase in a second pass of --steps frames in the same (conditioned) regime; "
                                                "step = ms_per_step / iterations of the timed region itself"
                                                + ("; PCG: the CG iteration count falls as the scene settles, so the solve phase of the (later) phase pass is "
                                                   "cheaper than the timed region's -- cg_iterations_per_admm_iteration gives both" if args.solver == "pcg" else ""),
                                        "cg_iterations_per_admm_iteration": cg_phases_pass if args.solver == "pcg" else None},
            "setup": {"seconds": t_setup, "factor_seconds": info0["factor_seconds"], "nnz_L": info0["nnz_L"],
                      "supernodes": info0["n_supernodes"], "levels": info0["n_levels"], "factor_bytes": info0["factor_bytes"]},
        }
        if hyper:
            # The dominant kernel is the hyperelastic local step: FP64-pipe bound, not HBM or tensor bound.  Its roofline uses
            # MEASURED denominators (admmb_probe_fp64 on this device, now) and COUNTED algorithmic flops (this run's tets).
            fp = admm_b200.probe_fp64(local_rank)
            peak_fma = 2.0 * fp["dfma_tinst_s"]                       # Tflop/s if every instruction were a fused multiply-add
            peak_nofma = 0.5 * (fp["dadd_tinst_s"] + fp["dmul_tinst_s"])  # Tflop/s of unfused adds / multiplies: all the bit-exact path may issue
            alg = flops["flops_per_tet_iteration"] * ntets / (local_ms * 1e-3) / 1e12 if local_ms > 0 else 0.0
            line["roofline"] = {
                "bound": "fp64", "kernel": "k_local_tets_hyper<NHModel,5> (1 launch per ADMM iteration)", "achieved": alg, "peak": peak_fma, "unit": "TFLOP/s",
                "frac": alg / peak_fma, "traffic": traffic["local"],
                "traffic_note": "bytes per launch, dram__bytes_read + write of the committed ncu --set full capture profiles/r2_local.txt (final build of the round, same workload); algorithmic bytes per launch: %d" % (LOCAL_BYTES_PER_TET * ntets) if traffic["local"] else None,
                "share_of_step": local_ms / phase_sum,
                "peak_source": "admmb_probe_fp64 on this device in this run: independent DFMA chains, 2 flops each (of measured)",
                "frac_of_unfused_peak": alg / peak_nofma, "unfused_peak": peak_nofma,
                "algorithmic_flops_per_tet_iteration": flops["flops_per_tet_iteration"], "flop_count": flops,
                "algorithmic_GBps": LOCAL_BYTES_PER_TET * ntets / (local_ms * 1e-3) / 1e9 if local_ms > 0 else 0.0,
                "fp64_probe": fp,
                "note": "achieved = counted algorithmic flops (add / mul / div / sqrt / log = 1 each, host restatement of the kernel body on a sample of this run's "
                        "tets) x tets / measured kernel time.  The kernel restates x86 arithmetic bit for bit, so no multiply-add may be fused (-fmad=false) and every "
                        "division / sqrt / log expands into 9-27 FP64 instructions: frac_of_unfused_peak is the fraction of what unfused FP64 issue can deliver; "
                        "pipe utilisation itself is in profiles/ (ncu)"}
            line["roofline_global"] = roof_global
        else:
            line["roofline"] = roof_global
        if partition is not None:
            line["partition"] = partition
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args.cpu_cube) if not args.scene else None
            if args.scene:
                from oracle import ref
                if ref.available():
                    r = reference_rate(sc, args.steps, max(args.warmup, 10))
                    line["cpu_baseline"] = {"value": r["it_s"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": f"the whole workload ({label}), {args.steps} frames"}
        if world == 1 and not args.no_pairs and not args.scene:
            line["same_config_pairs"] = same_config_pairs(local_rank, [w for w in args.pairs.split(",") if w])
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
