"""Concurrent pinned host -> device copy ceiling at 1/2/4/8 ranks (one process per GPU), with and without binding each
rank to the NUMA node of its GPU.  The end-to-end arm of bench.py moves 262 MB per rank per step through this path, so
this is its roofline.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/h2d_microbench.py [--mb 262] [--reps 20] [--no-bind]
Rank 0 prints one JSON line: per-rank GB/s (min / mean), aggregate GB/s, topology notes."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=262.144)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-bind", action="store_true")
    ap.add_argument("--chunks", type=int, default=1, help="split every copy into this many cudaMemcpyAsync calls")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import vadx  # noqa: F401
    from vadx import distributed as D

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = {"bound": False, "node_cpus": 0} if a.no_bind else D.bind_to_local_numa_node(local_rank)
    rank, world, dev = D.init("nccl")
    n = int(a.mb * 1e6) // 2
    host = [torch.empty((n,), dtype=torch.int16).pin_memory() for _ in range(2)]
    for h in host:
        h.fill_(1)                                  # first touch on this rank's (possibly bound) cores
    d = [torch.empty((n,), dtype=torch.int16, device=dev) for _ in range(2)]
    st = torch.cuda.Stream(device=dev)
    step = (n + a.chunks - 1) // a.chunks

    def copy(i):
        for lo in range(0, n, step):
            d[i % 2][lo:lo + step].copy_(host[i % 2][lo:lo + step], non_blocking=True)

    with torch.cuda.stream(st):
        for i in range(3):
            copy(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for i in range(a.reps):
            copy(i)
        e1.record(st)
    torch.cuda.synchronize()
    gbs = 2.0 * n * a.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        vals = [float(v.item()) for v in allv]
    else:
        vals = [gbs]
    if rank == 0:
        print(json.dumps({"h2d_pinned_gbs_per_rank_min": min(vals), "h2d_pinned_gbs_per_rank_mean": sum(vals) / len(vals),
                          "aggregate_gbs": sum(vals), "ranks": world, "mb_per_copy": a.mb, "reps": a.reps,
                          "chunks_per_copy": a.chunks, "numa_bind": numa, "per_rank": [round(v, 2) for v in vals],
                          "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
