"""Multi-GPU plumbing of the VAD hot path: one process per GPU, streams split into contiguous
blocks by rank, NO collective on the data path -- the reference has none either (every stream is an independent
ORT call, e.g. FireRedVAD/Inference_FireRed_ONNX.py:567-572) -- and one final gather of the per-stream results:
`seg_count [S]` and the `(start, end)` pairs `segments [S, max, 2]` the device post-processors emit.

    init()                       -> (rank, world, device) from the torchrun environment (nccl on GPUs, gloo on CPU)
    shard_streams(n, rank, world)-> this rank's contiguous block of stream indices
    gather_segments(cnt, seg)    -> on every rank: counts [S_total] and pairs [S_total, max, 2] in global stream order
    bind_to_local_numa_node(dev) -> pin this process (and therefore its pinned host buffers, first-touch) to the CPU
                                    cores next to its GPU, so that 8 ranks do not push their H2D traffic through one
                                    socket's memory controller

The gather is two all_gathers: the per-rank block sizes are a pure function of (n_streams, world), so only the padded
[block, max, 2] pair arrays and the counts travel; ranks whose block is shorter pad with zero-count rows that are cut off
again.  Works on CUDA tensors over NCCL (NVLink / NVSwitch) and on CPU tensors over gloo (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import os


def shard_streams(n_streams: int, rank: int, world: int) -> range:
    """Contiguous block of streams owned by `rank`: ceil(n/world) per rank, the last ranks may be short or empty."""
    if world < 1 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError(f"shard_streams: bad arguments n_streams={n_streams} rank={rank} world={world}")
    per = (n_streams + world - 1) // world
    return range(min(n_streams, rank * per), min(n_streams, (rank + 1) * per))


def block_size(n_streams: int, world: int) -> int:
    return (n_streams + world - 1) // world


def weighted_blocks(n_streams: int, weights) -> list:
    """Block sizes proportional to per-rank `weights` (largest-remainder rounding; they sum to n_streams).  Used when the
    ranks are NOT equally fast at the step that bounds them -- on an 8-GPU box the pinned host->device rate differs per GPU
    (PCIe switch sharing: 23 vs 35 GB/s measured, profiles/r02_h2d_ceiling.json), so equal blocks leave the well-fed GPUs
    idle while the others copy."""
    w = [max(float(x), 0.0) for x in weights]
    if not w or sum(w) <= 0 or n_streams < 0:
        raise ValueError("weighted_blocks: need positive weights")
    exact = [n_streams * x / sum(w) for x in w]
    sizes = [int(e) for e in exact]
    order = sorted(range(len(w)), key=lambda i: exact[i] - sizes[i], reverse=True)
    for i in order[:n_streams - sum(sizes)]:
        sizes[i] += 1
    return sizes


def block_of(sizes, rank: int) -> range:
    """This rank's contiguous block under explicit block sizes (weighted_blocks)."""
    lo = sum(sizes[:rank])
    return range(lo, lo + sizes[rank])


def measure_h2d_rates(device, mb: float = 64.0, reps: int = 6, group=None) -> list:
    """Pinned host -> device GB/s of EVERY rank while all ranks copy at the same time (the rate that matters for a
    host-fed pipeline).  Collective: every rank must call it.  Returns the per-rank list on every rank."""
    import torch
    import torch.distributed as dist
    n = int(mb * 1e6) // 2
    h = torch.empty((n,), dtype=torch.int16).pin_memory()
    h.zero_()
    d = torch.empty((n,), dtype=torch.int16, device=device)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(device)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        dist.barrier(group=group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(device)
    gbs = torch.tensor([2.0 * n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float64, device=device)
    if not multi:
        return [float(gbs.item())]
    allv = [torch.zeros_like(gbs) for _ in range(dist.get_world_size(group))]
    dist.all_gather(allv, gbs, group=group)
    return [float(v.item()) for v in allv]


def init(backend: str | None = None):
    """Process-group setup from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, world, device).  With WORLD_SIZE == 1 no process group is created."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
    else:
        dev = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if dev.type == "cuda" else "gloo")
        if backend == "nccl":
            dist.init_process_group(backend, device_id=dev)
        else:
            dist.init_process_group(backend)
    return rank, world, dev


def gather_segments(seg_count, segments, n_streams_total: int | None = None, group=None, sizes=None):
    """All ranks receive every rank's post-processor output in global stream order.
    sizes: explicit per-rank block sizes (weighted_blocks) instead of the equal blocks of shard_streams.

    seg_count: int32 [S_local]; segments: int [S_local, max_segments, 2] (same max_segments and dtype on every rank; CUDA
    for nccl, CPU for gloo).  n_streams_total: the global stream count the blocks came from (shard_streams); default =
    world * S_local (equal blocks).  Returns (counts [S_total], pairs [S_total, max_segments, 2])."""
    import torch
    import torch.distributed as dist
    if seg_count.dim() != 1 or segments.dim() != 3 or segments.shape[0] != seg_count.shape[0] or segments.shape[2] != 2:
        raise ValueError("gather_segments: expected seg_count [S] and segments [S, max_segments, 2]")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if n_streams_total is not None and n_streams_total != seg_count.shape[0]:
            raise ValueError("gather_segments: single rank but n_streams_total != local stream count")
        return seg_count, segments
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    S_local, max_seg = seg_count.shape[0], segments.shape[1]
    if sizes is not None:
        if len(sizes) != world or sizes[rank] != S_local:
            raise ValueError(f"gather_segments: rank {rank} holds {S_local} streams, sizes say {list(sizes)}")
        per = max(sizes)
        cnt, seg = seg_count.contiguous(), segments.contiguous()
        if S_local < per:
            cnt = torch.cat([cnt, cnt.new_zeros((per - S_local,))])
            seg = torch.cat([seg, seg.new_zeros((per - S_local, max_seg, 2))])
        all_cnt = cnt.new_empty((world * per,))
        all_seg = seg.new_empty((world * per, max_seg, 2))
        dist.all_gather_into_tensor(all_cnt, cnt, group=group)
        dist.all_gather_into_tensor(all_seg, seg, group=group)
        if all(sz == per for sz in sizes):
            return all_cnt, all_seg
        keep = torch.cat([torch.arange(r * per, r * per + sizes[r], device=cnt.device) for r in range(world)])
        return all_cnt.index_select(0, keep), all_seg.index_select(0, keep)
    if n_streams_total is None:
        n_streams_total = world * S_local
    per = block_size(n_streams_total, world)
    if S_local != len(shard_streams(n_streams_total, rank, world)):
        raise ValueError(f"gather_segments: rank {rank} holds {S_local} streams, its block of {n_streams_total} has "
                         f"{len(shard_streams(n_streams_total, rank, world))}")
    cnt = seg_count.contiguous()
    seg = segments.contiguous()
    if S_local < per:       # short (or empty) last blocks: pad to the common block size, cut off after the gather
        cnt = torch.cat([cnt, cnt.new_zeros((per - S_local,))])
        seg = torch.cat([seg, seg.new_zeros((per - S_local, max_seg, 2))])
    all_cnt = cnt.new_empty((world * per,))
    all_seg = seg.new_empty((world * per, max_seg, 2))
    dist.all_gather_into_tensor(all_cnt, cnt, group=group)
    dist.all_gather_into_tensor(all_seg, seg, group=group)
    return all_cnt[:n_streams_total], all_seg[:n_streams_total]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """The timing rule of every multi-GPU number: device time as the maximum over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def _cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def local_numa_cpus(device_index: int):
    """CPU cores next to the GPU: NVML's ideal-affinity mask for the device (the driver's own topology answer), else the
    sysfs NUMA node of its PCI function; None when the platform does not say (single-socket boxes report node -1)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        if cpus and len(cpus) < n_cpu:
            return cpus
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(os.path.join(path, "numa_node")).read().strip())
        if node >= 0:
            return _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) or None
    except Exception:
        pass
    return None


def _physical_index(device_index: int) -> int:
    """NVML enumerates every GPU of the box; CUDA_VISIBLE_DEVICES may renumber them for this process."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if device_index < len(ids) and ids[device_index].isdigit():
            return int(ids[device_index])
    return device_index


def bind_to_local_numa_node(device_index: int) -> dict:
    """Restrict this process to the cores next to its GPU BEFORE allocating pinned host buffers: pages are placed on the
    node of the first-touching thread, so the H2D copies of a rank then read local memory.  Returns what was done
    ({"node_cpus": n, "bound": bool}); a no-op where the topology is unknown or the affinity call is refused."""
    cpus = local_numa_cpus(device_index)
    if not cpus:
        return {"node_cpus": 0, "bound": False}
    try:
        allowed = os.sched_getaffinity(0)
        want = set(cpus) & allowed
        if not want:
            return {"node_cpus": len(cpus), "bound": False}
        os.sched_setaffinity(0, want)
        return {"node_cpus": len(want), "bound": True}
    except OSError:
        return {"node_cpus": len(cpus), "bound": False}
