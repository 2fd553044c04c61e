"""ONE mesh partitioned over the ranks of a torchrun job (strong scaling), without bench.py's replica ensemble: for meshes whose
set-up is too long to build twice per rank (8 M tets).  Prints one JSON line from rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/strong_scaling.py --cube 110
A single process (no torchrun) gives the one-GPU figure of the same regime."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "admm-elastic-sca_b200", "pyhost"), os.path.join(ROOT, "tests")]
import torch
import admm_b200, scenes, ensemble

ap = argparse.ArgumentParser()
ap.add_argument("--cube", type=int, default=110)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--condition", type=int, default=20)
ap.add_argument("--solver", default="direct", choices=["direct", "pcg"])
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist_arg = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    holder = [admm_b200.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(holder, src=0)
    dist_arg = (rank, world, holder[0])
admm_b200.lib().admmb_set_host_threads(max(1, (os.cpu_count() or 1) // world))   # torchrun exports OMP_NUM_THREADS=1: give the set-up its share of the cores
sc = scenes.cube_scene(args.cube, kind=scenes.TET_NH, mu=1e5, lam=1e5, maxit=5, mass=1000.0, dt=0.04, iters=10, stretch=1.3)
t0 = time.perf_counter()
sim = admm_b200.System(sc, device=local, solver=admm_b200.SOLVER_DIRECT if args.solver == "direct" else admm_b200.SOLVER_PCG, cg_tol=1e-10, dist=dist_arg)
t_setup = time.perf_counter() - t0
sim.set_x(sc["x_after_init"]); sim.upload()
sim.step_resident(frames=args.condition + args.warmup)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
cg0 = sim.info()["cg_iters_total"]
sim.step_resident(frames=args.steps)
ms = sim.last_region_ms()
cg1 = sim.info()["cg_iters_total"]
sim.timing(True); sim.timing_read(reset=True)
sim.step_resident(frames=max(2, args.steps // 2))
ph = sim.timing_read(reset=True)
cg2 = sim.info()["cg_iters_total"]
sim.timing(False)
ms_max, _ = ensemble.reduce_job(ms, 0)
it = max(ph["iters"], 1)
info = sim.info()
if rank == 0:
    print(json.dumps({"cube": args.cube, "tets": int(sc["batches"][0]["idx"].shape[0]), "n_gpus": world, "solver": args.solver,
                      "value": args.steps * int(sc["iters"]) / (ms_max * 1e-3), "unit": "admm_iterations/s", "ms_per_iteration": ms_max / (args.steps * int(sc["iters"])),
                      "phases_ms_per_iteration_rank0": {"local": ph["local_ms"] / it, "rhs": ph["rhs_ms"] / it, "solve": ph["solve_ms"] / it},
                      "timed_pass_ms_per_iteration_rank0": ph["step_ms"] / it,
                      "cg_iterations_per_admm_iteration": [(cg1 - cg0) / float(args.steps * int(sc["iters"])), (cg2 - cg1) / float(it)],
                      "factor_bytes_rank0": info["factor_bytes"], "setup_s_rank0": t_setup}), flush=True)
sim.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
