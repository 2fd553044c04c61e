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
EXEC_FP64_FLOPS_PER_TET_ITER = 9301      # profiles/r1e_local.txt (steady state): (1184.0 + 1325.7 + 2 x 1477.5) flops/clk x 1.6991 Mclk / 998250 tets
SOLVE_DRAM_BYTES_NCU = 1.7221e9          # profiles/r1e_solve.txt: dram bytes read+written by the 26 launches of one solve (cube N=55)


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


def run_reference(args, rank, world):
    """Times the unmodified reference on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libadmm_ref.so missing (build it where /root/reference exists)"}))
        return
    N = args.ref_cube
    sc = make_scene(N)
    ntets = sc["batches"][0]["idx"].shape[0]
    cores = os.cpu_count() or 1
    ref.lib().ref_set_omp_threads(cores)
    t0 = time.perf_counter()
    sim = ref.RefSystem(sc, probe=False)
    t_init = time.perf_counter() - t0
    sim.set_x(sc["x_after_init"])
    for _ in range(args.warmup):
        sim.step()
    sec = sim.step_timed(args.steps)
    stats = sim.stats()
    sim.close()
    its = args.steps * ADMM_ITERS / sec
    value = its * ntets / FULL_TETS  # scaled by tet count to the 1M-tet unit (generous: the reference's cost grows superlinearly)
    sample = (f"cube N={N} ({ntets} tets) stepped {args.steps} frames x {ADMM_ITERS} iterations = {its:.2f} it/s measured, "
              f"scaled x{ntets}/{FULL_TETS} to the 1M-tet unit; initialize() {t_init:.1f} s not included; L nnz {stats['L_nnz']}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cube N={args.cube} NeoHookean, {ADMM_ITERS} ADMM iterations per step (reference timed on N={N})"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_leg(N):
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    sc = make_scene(N)
    ntets = sc["batches"][0]["idx"].shape[0]
    cores = os.cpu_count() or 1
    ref.lib().ref_set_omp_threads(cores)
    t0 = time.perf_counter()
    sim = ref.RefSystem(sc, probe=False)
    t_init = time.perf_counter() - t0
    sim.set_x(sc["x_after_init"])
    sim.step()
    frames = 8
    sec = sim.step_timed(frames)
    sim.close()
    its = frames * ADMM_ITERS / sec
    return {"value": its * ntets / FULL_TETS, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"unmodified reference (Eigen + OpenMP, {cores} threads) on cube N={N} ({ntets} tets): {its:.2f} it/s over {frames} frames, "
                      f"scaled x{ntets}/{FULL_TETS}; initialize() {t_init:.1f} s excluded"}


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
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--cube", type=int, default=55, help="cube resolution N (N=55: 998,250 tets)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-cube", type=int, default=30, help="cube resolution of the reference arm's bounded sample")
    ap.add_argument("--cpu-cube", type=int, default=20, help="cube resolution of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ensemble", type=int, default=0,
                    help="scene-ensemble mode (BASELINE configs[4] '64-scene ensemble'): this many independent scenes (cube --ens-cube) dealt "
                         "round-robin to the ranks, all scenes of a rank in flight at once on their own streams; prints scene-frames/s")
    ap.add_argument("--ens-cube", type=int, default=20, help="cube resolution of the ensemble scenes (N=20: 48,000 tets)")
    ap.add_argument("--solver", default="direct", choices=["direct", "pcg"])
    ap.add_argument("--partition", action="store_true",
                    help="N > 1: ONE mesh partitioned over the ranks (PCG rows + local step, NCCL all-gather / all-reduce) instead of "
                         "the default scene ensemble; strong scaling")
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
    sc = make_scene(args.cube)
    ntets = sc["batches"][0]["idx"].shape[0]
    nverts = sc["x"].shape[0]
    dist_arg = None
    if args.partition and world > 1:
        args.solver = "pcg"   # sparse triangular solves do not shard (replicas only)
        holder = [admm_b200.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        dist_arg = (rank, world, holder[0])
    t0 = time.perf_counter()
    sim = admm_b200.System(sc, device=local_rank, solver=admm_b200.SOLVER_DIRECT if args.solver == "direct" else admm_b200.SOLVER_PCG,
                           cg_tol=1e-10, dist=dist_arg, pin_host=True)
    t_setup = time.perf_counter() - t0
    info0 = sim.info()
    sim.set_x(sc["x_after_init"])
    sim.upload()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident path: `value` -----------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    sim.step_resident(frames=args.warmup)
    barrier()
    l0 = sim.info()["launches_total"]
    sampler.begin()
    sim.step_resident(frames=args.steps)      # the timed region: K frames, CUDA events on the library's stream
    sampler.end()
    ms_region = sim.last_region_ms()
    l1 = sim.info()["launches_total"]
    barrier()
    clocks = sampler.stop()
    # second pass with per-phase events (not part of `value`): local / rhs / solve split
    sim.timing(True)
    sim.timing_read(reset=True)
    sim.step_resident(frames=max(2, args.steps // 2))
    phases = sim.timing_read(reset=True)
    sim.timing(False)

    # ---- end to end through admmb_step with host buffers: `e2e` ---------------------------------------------------
    sim.download()
    for _ in range(2):
        sim.step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.step()
    torch.cuda.synchronize()
    sec_e2e = time.perf_counter() - t0
    barrier()

    import ensemble  # whole-job figures: MAX over ranks of the timed region, SUM over ranks of the units
    ms_region_max, total_iters = ensemble.reduce_job(ms_region, args.steps * ADMM_ITERS)
    ms_e2e_max, _ = ensemble.reduce_job(sec_e2e * 1e3, args.steps * ADMM_ITERS)
    if dist_arg is not None:
        total_iters = args.steps * ADMM_ITERS   # one mesh: the ranks share the same iterations
    value = total_iters / (ms_region_max * 1e-3)
    e2e_value = total_iters / (ms_e2e_max * 1e-3)

    if rank == 0:
        pk = peaks()
        hbm_peak = (pk or {}).get("hbm_gbs", 6650.0)
        peak_src = "MEASURED_PEAKS.json hbm_gbs (copy, of measured)" if pk else "fallback 6650 GB/s (B200_PROFILING.md, of fallback)"
        iters = max(phases["iters"], 1)
        solve_ms = phases["solve_ms"] / iters
        local_ms = phases["local_ms"] / iters
        rhs_ms = phases["rhs_ms"] / iters
        # dominant kernel of the global step: k_solve_level streams the packed factor once forward and once backward
        solve_bytes = info0["factor_bytes"] + 9 * nverts * 8 * 3  # factor tiles + b, y, x vectors read/written
        solve_gbs = solve_bytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        # local step: algorithmic bytes per tet-iteration: idx 16 + x gather 96 + B 96 + u r/w 144 + z w 72
        # + state r/w 64 + weights 24 + P w 96 = 608 B
        local_bytes = 608.0 * ntets
        local_gbs = local_bytes / (local_ms * 1e-3) / 1e9 if local_ms > 0 else 0.0
        # Two kernels carry the step.  k_solve_level (global step) is HBM bound and is the `roofline` object;
        # k_local_tets_hyper (local step) is the larger share of the time but is FP64-pipe bound, which the
        # hbm|tensor schema cannot express, so it is reported beside it as `roofline_local`.
        step_ms = ms_region_max / args.steps / ADMM_ITERS
        roof = {"bound": "hbm", "kernel": f"k_solve_level_pf ({2 * info0['n_levels']} launches per solve)", "achieved": solve_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": solve_gbs / hbm_peak, "traffic": SOLVE_DRAM_BYTES_NCU if args.cube == 55 else None, "peak_source": peak_src,
                "share_of_step": solve_ms / (local_ms + rhs_ms + solve_ms),
                "note": f"algorithmic bytes per solve = packed factor, both copies ({info0['factor_bytes']} B) + 9 vector passes of 3n doubles; "
                        f"{info0['n_levels']} levels; traffic = dram bytes of these launches summed, ncu profiles/r1e_solve.txt (cube N=55 only)"}
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        fp64_peak = 148 * 64 * 2 * sm_clock / 1e12   # 64 DFMA lanes per SM per clock (ncu: sm__sass_thread_inst_executed_op_dfma peak)
        local_tflops = EXEC_FP64_FLOPS_PER_TET_ITER * ntets / (local_ms * 1e-3) / 1e12 if local_ms > 0 else 0.0
        roof_local = {"bound": "fp64", "kernel": "k_local_tets_hyper<NHModel,5>", "achieved": local_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": local_tflops / fp64_peak, "share_of_step": local_ms / (local_ms + rhs_ms + solve_ms),
                      "fp64_pipe_active_pct_ncu": 57.9, "algorithmic_GBps": local_gbs,
                      "note": "executed FP64 flops per tet-iteration in steady state (DADD + DMUL + 2 DFMA, ncu profiles/r1e_local.txt: "
                              f"{EXEC_FP64_FLOPS_PER_TET_ITER}); peak = 148 SMs x 64 DFMA/clk x 2 at the sampled SM clock; the kernel is compiled "
                              "with -fmad=false to stay bit-exact with the reference, so no multiply-add is fused"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_region_max / args.steps, "higher_is_better": True, "scaling": "strong" if dist_arg is not None else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cube N={args.cube} ({ntets} tets, {nverts} nodes) NeoHookean mu=lambda=1e5 max_iterations=5, "
                                   f"{ADMM_ITERS} ADMM iterations per step (System::step), dt 0.04, x1.3 stretch; solver={args.solver}",
                       "parallelism": (f"one mesh partitioned over {world} ranks: PCG rows + local step, NCCL all-gather/all-reduce" if dist_arg is not None
                                       else f"ensemble x{world} (one scene per GPU, no collective)") if world > 1 else "single GPU",
                       "cg_iterations_per_admm_iteration": (sim.info()["cg_iters_total"] / max(1, sim.info()["elapsed_s"] / sc["dt"] * ADMM_ITERS)) if args.solver == "pcg" else None,
                       "l2": "working set (factor %.2f GB + force arrays %.2f GB) exceeds the 126 MB L2, no flush needed" % (
                           info0["factor_bytes"] / 1e9, 608.0 * ntets / 1e9)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * 3 * nverts * 8, "d2h_bytes_per_step": 2 * 3 * nverts * 8,
                    "ms_per_step": ms_e2e_max / args.steps,
                    "note": "admmb_step(iters, m_x, m_v): host x and v in and out every step, buffers page-locked once with "
                            "admmb_register_host_buffer (pinned host memory, as the contract asks)"},
            "gpu_launches": int(l1 - l0),
            "clocks": clocks,
            "roofline": roof,
            "roofline_local": roof_local,
            "phases_ms_per_iteration": {"local": local_ms, "rhs": rhs_ms, "solve": solve_ms,
                                        "local_GBps_algorithmic": local_gbs, "solve_GBps": solve_gbs},
            "setup": {"seconds": t_setup, "factor_seconds": info0["factor_seconds"], "nnz_L": info0["nnz_L"],
                      "supernodes": info0["n_supernodes"], "levels": info0["n_levels"], "factor_bytes": info0["factor_bytes"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args.cpu_cube)
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
