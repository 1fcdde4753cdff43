"""SNP-range sharding across the GPUs of one box.

The reference's only parallelism is data parallelism over markers through
worker processes and queues (``genetest/analysis.py:601-633``).  Here it is
one process per GPU (``torchrun``); the batches of markers are dealt to the
ranks in turn, nothing is exchanged while computing, and the fixed-size result
blocks of every round are gathered on rank 0 (NCCL on GPUs; gloo in the CPU
tests), whose sinks write them in marker order while the next round runs.
No statistic spans markers, so no other collective exists.
"""

import os
import sys

import numpy as np

__all__ = ["rank_world", "is_primary", "local_device", "shard_range",
           "gather_results", "gather_arrays", "gather_round",
           "ensure_initialized", "bind_to_device_cpus"]


def ensure_initialized():
    """Under ``torchrun`` (WORLD_SIZE > 1 in the environment) join the process
    group before anything is shared out: NCCL with this rank's own device when
    CUDA is there (``torch.cuda.set_device(LOCAL_RANK)``, ``device_id`` so that
    collectives and object gathers never land on device 0), gloo otherwise (CPU
    tests).  A multi-rank launch that is not initialised would make every rank
    run the whole analysis on cuda:0 and write the same files."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return False
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        if torch.cuda.is_available() and dist.get_backend() == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        return True
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        bind_to_device_cpus(local)
    else:
        dist.init_process_group("gloo")
    return True


def _dist():
    # a process that joined a process group has imported torch already;
    # single-process runs never pay for the import
    if "torch" not in sys.modules:
        return None
    try:
        import torch.distributed as dist
    except ImportError:         # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def rank_world():
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def is_primary():
    return rank_world()[0] == 0


def local_device():
    return int(os.environ.get("LOCAL_RANK", "0")) if _dist() else 0


def bind_to_device_cpus(device):
    """Pin this process to the CPU cores NVML reports as local to ``device``
    (its NUMA node), so that the pinned host buffers of a rank are first
    touched next to its GPU: with eight ranks streaming genotypes over PCIe,
    cross-socket traffic otherwise caps the end-to-end rate.  Returns the
    number of cores bound to, or 0 when NVML / affinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words)
                for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:           # no NVML, no permission: run unbound
        return 0


def shard_range(n, world, rank):
    """Contiguous range [lo, hi) of ``n`` markers owned by ``rank``:
    ``ceil(n / world)`` markers per rank, the last ranks may own fewer."""
    per = -(-n // world) if world > 0 else n
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def gather_arrays(arrays):
    """Gather a list of ``[fields, n_local]`` numpy arrays on rank 0.

    Every rank passes arrays with the same leading dimensions and dtypes;
    ``n_local`` may differ.  Returns, on rank 0, the list of arrays
    concatenated over ranks along the last axis (rank order), elsewhere None.
    """
    dist = _dist()
    if dist is None:
        return arrays
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    nccl = dist.get_backend() == "nccl"
    dev = torch.device("cuda", local_device()) if nccl else torch.device("cpu")
    n_local = arrays[0].shape[-1]
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = n_local
    dist.all_reduce(sizes)
    sizes = [int(s) for s in sizes.cpu()]
    n_max = max(sizes)
    out = []
    for a in arrays:
        t = torch.zeros(a.shape[:-1] + (n_max,),
                        dtype=torch.from_numpy(a[..., :0]).dtype, device=dev)
        if n_local:
            t[..., :n_local] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        bucket = [torch.empty_like(t) for _ in range(world)] \
            if rank == 0 else None
        dist.gather(t, bucket, dst=0)
        if rank == 0:
            out.append(np.concatenate(
                [b[..., :s].cpu().numpy() for b, s in zip(bucket, sizes)],
                axis=-1))
    return out if rank == 0 else None


def gather_round(block, sizes, n_fields, n_params):
    """One round of the packed path: ``block`` = (Results, file MAF or None) of
    this rank's batch (None when it has none this round), ``sizes[r]`` the
    number of markers of rank r's batch (every rank knows them: the plan is
    deterministic).  The blocks travel as ONE float64 tensor per rank --
    [n_fields + NUM_IFIELDS + 1, max(sizes)]: result fields, the integer
    fields, the file-level MAF -- gathered on rank 0 over NCCL (gloo on CPU).
    Returns the list of blocks in rank order on rank 0, None elsewhere."""
    dist = _dist()
    from .engine import Results, NUM_IFIELDS
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    nccl = dist.get_backend() == "nccl"
    dev = torch.device("cuda", local_device()) if nccl else torch.device("cpu")
    n_max = max(sizes)
    rows = n_fields + NUM_IFIELDS + 1
    host = np.zeros((rows, n_max), dtype=np.float64)
    if block is not None:
        res, fmaf = block
        m = res.out.shape[1]
        host[:n_fields, :m] = res.out
        host[n_fields:n_fields + NUM_IFIELDS, :m] = res.iout      # int32 values are exact in float64
        host[-1, :m] = np.nan if fmaf is None else fmaf
    t = torch.from_numpy(host).to(dev)
    bucket = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bucket, dst=0)
    if rank != 0:
        return None
    out = []
    for r, m in enumerate(sizes):
        if m == 0:
            out.append(None)
            continue
        a = bucket[r][:, :m].cpu().numpy()
        res = Results(np.ascontiguousarray(a[:n_fields]),
                      np.ascontiguousarray(a[n_fields:n_fields + NUM_IFIELDS]).astype(np.int32),
                      n_params)
        fm = np.ascontiguousarray(a[-1])
        out.append((res, fm))
    return out


def gather_results(res, meta):
    """Result blocks + marker metadata of every rank -> rank 0."""
    dist = _dist()
    if dist is None:
        return res, meta
    from .engine import Results
    rank, world = dist.get_rank(), dist.get_world_size()
    gathered = gather_arrays([res.out, res.iout])
    metas = [None] * world if rank == 0 else None
    dist.gather_object(meta, metas, dst=0)
    if rank != 0:
        return None, None
    return (Results(gathered[0], gathered[1], res.n_params),
            [m for part in metas for m in part])
