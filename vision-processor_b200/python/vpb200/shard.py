"""Multi-GPU plumbing of the detection path: the work shards by camera / by frame with NO data-path collective
(the reference runs one OS process per camera, README.md:71; cameras exchange only small UDP datagrams).

One process per GPU (torchrun); `torch.distributed` is used for exactly two things: a barrier around the timed region
and a MAX-reduction of the per-rank elapsed time.  Backend "nccl" on GPUs, "gloo" in the CPU tests."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    start: int
    stop: int

    def __len__(self) -> int:
        return self.stop - self.start

    def indices(self) -> range:
        return range(self.start, self.stop)


def partition(n_items: int, world: int, rank: int) -> Shard:
    """Contiguous slice of `n_items` frames (or cameras) for `rank`: sizes differ by at most one, earlier ranks get
    the larger slices, every item is owned by exactly one rank."""
    if world < 1 or not 0 <= rank < world or n_items < 0:
        raise ValueError(f"bad shard request: n_items={n_items} world={world} rank={rank}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return Shard(rank, world, start, start + base + (1 if rank < extra else 0))


def camera_of_rank(rank: int, n_cameras: int) -> int:
    """BASELINE config 3: camera c -> GPU c mod G (here: the camera a rank serves)."""
    return rank % max(n_cameras, 1)


def env_rank_world() -> tuple:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the job is as slow as its slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(items_this_rank: int, elapsed_s_this_rank: float, device=None) -> float:
    """Whole-job throughput: items of all ranks / the slowest rank's time."""
    total = sum_over_ranks(float(items_this_rank), device)
    worst = max_over_ranks(float(elapsed_s_this_rank), device)
    return total / worst if worst > 0 else float("inf")


def bind_to_gpu_numa_node(gpu_index: int) -> list:
    """Pin this process to the CPUs next to its GPU (NVML's affinity mask) BEFORE it allocates pinned host memory, so that
    the frame ring of camera c lives on the socket GPU c hangs off -- with eight cameras uploading 5 MB frames at once the
    host-to-device path is otherwise limited by cross-socket traffic.  The reference gets the same effect from running one
    OS process per camera.  Returns the CPU list, or [] when NVML is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []
