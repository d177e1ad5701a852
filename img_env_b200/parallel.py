"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on GPUs, gloo in the CPU
tests).  Scenes are fully independent (one ROS node per scene in the reference,
create_launch.py:25-34), so the step itself has NO collective: each rank owns a contiguous block of
scenes.  The only exchange is the optional delivery of observations to a learner rank
(SURVEY.md §8e), and the max-over-ranks reduction of device timings for bench.py."""
import torch
import torch.distributed as dist


def shard_scenes(total_scenes, world, rank):
    """Contiguous block partition: returns (first_scene, n_scenes) of `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total_scenes), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def owner_of(scene, total_scenes, world):
    base, rem = divmod(int(total_scenes), int(world))
    cut = rem * (base + 1)
    return scene // (base + 1) if scene < cut else rem + (scene - cut) // max(base, 1)


def reduce_max(value, device=None, group=None):
    """Max over ranks of a python float (device timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_observations(out, dst=0, group=None):
    """Delivers every rank's State tensors ([S_local, R, ...]) to rank `dst`, concatenated along the scene
    axis in rank order.  Returns the dict on `dst`, None elsewhere.  Uses all_gather over NVLink/NVSwitch
    when the backend is NCCL (equal shard sizes required), gather on gloo."""
    if not (dist.is_available() and dist.is_initialized()):
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    res = {}
    for k, v in out.items():
        v = v.contiguous()
        if dist.get_backend(group) == "nccl":
            buf = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(buf, v, group=group)
            if rank == dst:
                res[k] = torch.cat(buf, 0)
        else:
            buf = [torch.empty_like(v) for _ in range(world)] if rank == dst else None
            dist.gather(v, buf, dst=dst, group=group)
            if rank == dst:
                res[k] = torch.cat(buf, 0)
    return res if rank == dst else None
