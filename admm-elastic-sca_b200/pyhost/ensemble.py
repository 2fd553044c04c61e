"""Scene-ensemble sharding: independent scenes are dealt round-robin to the ranks (one process and
one context per GPU) and no data-path collective is needed; torch.distributed only carries the barrier and the
reduction of the per-rank timings / counts.  Backend-agnostic (NCCL on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def scene_shard(num_scenes, rank, world):
    """Indices of the scenes rank `rank` steps: round-robin, so ranks differ by at most one scene."""
    return list(range(rank, num_scenes, world))


def reduce_job(ms_local, units_local, device=None):
    """Whole-job figures: time = MAX over ranks of the device-timed region, units = SUM over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local), float(units_local)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def throughput(ms_job, units_job):
    return units_job / (ms_job * 1e-3)
