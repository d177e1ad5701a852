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


class ObservationGatherer:
    """Delivers every rank's State tensors ([S_local, R, ...]) to the learner rank `dst`, concatenated along the
    scene axis in rank order.  Point-to-point sends straight into slices of persistent buffers on
    `dst` (one batched isend/irecv group per call): only the learner's NVLink ingress carries data, nothing is
    gathered to ranks that do not need it and there is no concatenation copy.  Works on NCCL (device tensors over
    NVLink/NVSwitch) and gloo (CPU tests); shard sizes may differ per rank."""

    def __init__(self, out, dst=0, group=None):
        self.dst, self.group = dst, group
        self.active = dist.is_available() and dist.is_initialized()
        if not self.active:
            return
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n_local = next(iter(out.values())).shape[0]
        sizes = [None] * self.world
        dist.all_gather_object(sizes, int(n_local), group=group)
        self.sizes = sizes
        self.offsets = [sum(sizes[:r]) for r in range(self.world)]
        self.buf = None
        if self.rank == dst:
            self.buf = {k: torch.empty((sum(sizes),) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device) for k, v in out.items()}
        self.bytes_to_learner = sum(v[:1].numel() * v.element_size() for v in out.values()) * (sum(sizes) - sizes[dst])

    def __call__(self, out):
        """-> dict of gathered tensors on `dst` (persistent buffers, overwritten by the next call), None elsewhere."""
        if not self.active:
            return out
        ops = []
        if self.rank == self.dst:
            for k, v in out.items():
                for src in range(self.world):
                    sl = self.buf[k][self.offsets[src]: self.offsets[src] + self.sizes[src]]
                    if src == self.rank:
                        sl.copy_(v)
                    elif self.sizes[src]:
                        ops.append(dist.P2POp(dist.irecv, sl, src, group=self.group))
        elif self.sizes[self.rank]:
            ops = [dist.P2POp(dist.isend, v.contiguous(), self.dst, group=self.group) for v in out.values()]
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        return self.buf if self.rank == self.dst else None


def gather_observations(out, dst=0, group=None):
    """One-shot form of ObservationGatherer (allocates the destination buffers on every call)."""
    if not (dist.is_available() and dist.is_initialized()):
        return out
    return ObservationGatherer(out, dst=dst, group=group)(out)
