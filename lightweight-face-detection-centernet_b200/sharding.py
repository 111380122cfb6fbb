"""Multi-GPU plumbing: images are independent, weights replicate (1.3 M parameters), so a batch is
split into contiguous per-rank shards and the ONLY exchange is one all-gather of the final
fixed-size box list.  One process per GPU; ``torch.distributed`` with the NCCL
backend over NVLink on the GPU box, ``gloo`` in the CPU tests."""
from __future__ import annotations


def bind_host_to_gpu(device_index):
    """Pin the calling process to the CPUs NVML reports as local to GPU ``device_index`` (its NUMA node), so that the pinned
    host buffers allocated afterwards -- the per-step H2D source of the end-to-end path -- sit behind the GPU's own PCIe root
    complex instead of across the socket interconnect.  (Measured on the 8-GPU box of this pool: no change -- eight ranks x 39 MB
    per 2.37 ms step = 133 GB/s is that box's aggregate host-to-device rate with or without the binding; kept because placement is
    not guaranteed on other hosts.)  Returns (previous affinity, new affinity), or (None, None) when NVML, the affinity call or
    the container's cpuset do not allow it: purely an optimisation."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if visible:  # NVML enumerates all devices of the box, CUDA only the visible ones
                tok = visible.split(",")[device_index].strip()
                if tok.isdigit():
                    idx = int(tok)
                else:
                    idx = None
                    h = pynvml.nvmlDeviceGetHandleByUUID(tok.encode() if hasattr(tok, "encode") else tok)
            if idx is not None:
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        old = os.sched_getaffinity(0)
        new = local & old
        if not new or new == old:
            return (old, old) if new else (None, None)
        os.sched_setaffinity(0, new)
        return old, new
    except Exception:
        return None, None


def shard_range(n_items, rank, world):
    """Contiguous shard [lo,hi) of ``n_items`` for ``rank``; the first ``n_items % world`` ranks get
    one extra item so ragged batches are covered."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} / world {world}")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(dets, n_total=None, group=None):
    """All-gather per-rank ``dets`` [b_r, K, C] (fixed K, C) into [sum b_r, K, C] on every rank.
    Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to the largest shard
    and trimmed after the gather."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dets
    world = dist.get_world_size(group)
    b = dets.shape[0]
    if n_total is None:
        sizes = torch.tensor([b], device=dets.device, dtype=torch.int64)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes, group=group)
        counts = [int(s.item()) for s in all_sizes]
    else:
        counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    bmax = max(counts)
    if b < bmax:
        pad = torch.zeros((bmax - b,) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
        dets = torch.cat([dets, pad], 0)
    out = torch.empty((world * bmax,) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    dist.all_gather_into_tensor(out, dets.contiguous(), group=group)
    if all(c == bmax for c in counts):
        return out
    return torch.cat([out[r * bmax:r * bmax + counts[r]] for r in range(world)], 0)


def gather_variable(dets, counts, group=None):
    """Paths A/B: ``dets`` [b_r, cap, C] padded boxes + ``counts`` [b_r] -> gathered (dets, counts)."""
    return gather_detections(dets, group=group), gather_detections(counts.reshape(-1, 1, 1), group=group).reshape(-1)
