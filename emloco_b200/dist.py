"""Data parallelism of the rollout: one process per GPU, envs are rank-local, no data-path collective.

Reference: Horovod through rl_games (`pacer/pacer/run.py:57-72`: `rank = hvd.rank()`, seed += rank; rl_games `self.hvd`
in `learning/common_agent.py:165-180,308-328`).  Environments never exchange state (own collision group, no inter-env
observations), so the only collectives of the hot path are the bookkeeping ones below; the gradient all-reduce of the
reference lives in the optimiser step, outside this path.  Backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa_node(local_rank):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off (read from sysfs through the GPU's PCI address), so
    that pinned host buffers allocated afterwards are first-touched on that node and host<->device copies do not cross the
    socket interconnect.  With every rank left on the launcher's default affinity (all ranks on node 0) the end-to-end path of
    8 ranks ran at 0.46 efficiency (SCALE_r01).  Best effort: returns the node (or None when it cannot be determined)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank])
                                              if os.environ.get("CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit() else local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                                       # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0)) or cpus
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def rank_seed(seed, rank):
    """run.py:65: every rank simulates its own envs with its own random stream."""
    return int(seed) + int(rank)


def shard_envs(total_envs, rank, world):
    """[lo, hi) of the envs owned by `rank` when a global env count is split (weak scaling keeps envs per rank fixed)."""
    per, rem = divmod(int(total_envs), int(world))
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """Slowest rank's time: the number every multi-GPU throughput is quoted on."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def average_scalar(value, device="cpu"):
    """hvd.average_value (common_agent.py:308-310,327-328): e.g. the KL divergence across ranks."""
    _, _, world = env_rank()
    return sum_over_ranks(value, device) / max(world if dist.is_initialized() else 1, 1)


class FlatGrads:
    """The one data-path collective of the reference's multi-GPU training: the gradient average Horovod's
    `optimizer.synchronize()` performs before every optimiser step (learning/amp_continuous_value.py:381-388, SURVEY 8e:
    11.2 M fp32 parameters = 44.8 MB).  All `.grad` tensors are views of ONE flat buffer, so the average is a single
    all-reduce over NVLink with no pack / unpack copies (Horovod fuses tensors into a buffer and copies both ways)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("FlatGrads expects fp32 parameters (mixed_precision is False in the reference config)")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        """`for param in self.model.parameters(): param.grad = None` (:371-372) without dropping the views."""
        self.flat.zero_()

    def average(self):
        """Average over ranks, in place; a no-op on one rank.  Returns the async work handle's completion (blocking)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        return self.flat

    def nbytes(self):
        return self.flat.numel() * 4


class BucketedAllReduce:
    """The gradient all-reduce of the update step in two buckets over ONE flat buffer: `start_first()` launches the reduction
    of [0, split) as soon as that part of the backward pass has been queued (asynchronously: the kernels issued afterwards
    overlap it), `finish()` reduces [split, n) and waits for both.  Sums, does not average - the 1 / world factor is folded
    into the optimiser kernel.  World size 1 or an uninitialised process group: no-ops."""

    def __init__(self, flat, split, overlap=True, world=None):
        self.flat, self.split, self.overlap = flat, int(split), bool(overlap)
        # world = 1 forces a purely local step even inside an initialised process group (single-process reference runs)
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else int(world)
        self._pending = None

    def start_first(self):
        if self.world > 1 and self.overlap and 0 < self.split < self.flat.numel():
            self._pending = dist.all_reduce(self.flat[:self.split], op=dist.ReduceOp.SUM, async_op=True)

    def finish(self):
        if self.world <= 1:
            return
        if self._pending is not None:
            dist.all_reduce(self.flat[self.split:], op=dist.ReduceOp.SUM)
            self._pending.wait()
            self._pending = None
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def finalize():
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
