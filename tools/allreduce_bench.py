"""Times the two dense collectives of the multi-GPU path in isolation (NCCL through torch.distributed):
the all-reduce of the uint32 count grid (201^3 and 297^3 cells) and the all-gather of the 44-byte slab
summaries.  Run under torchrun; NCCL_* environment variables select algorithm / protocol variants.

    torchrun --nproc-per-node 8 tools/allreduce_bench.py [label]
"""
import os
import sys

import torch
import torch.distributed as dist


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "default"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = []
    for name, n in (("counts 201^3", 201 ** 3), ("counts 297^3", 297 ** 3)):
        t = torch.ones(n, dtype=torch.int32, device="cuda")
        for _ in range(5):
            dist.all_reduce(t)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            dist.all_reduce(t)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out.append("%s: %.3f ms (%.0f GB/s algbw)" % (name, ms, n * 4 / ms / 1e6))
    s = torch.ones((11, 30000), dtype=torch.int32, device="cuda")
    g = torch.empty((world, 11, 30000), dtype=torch.int32, device="cuda")
    for _ in range(5):
        dist.all_gather_into_tensor(g, s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dist.all_gather_into_tensor(g, s)
    b.record()
    torch.cuda.synchronize()
    out.append("slab summaries all-gather (1.3 MB per rank): %.3f ms" % (a.elapsed_time(b) / 20))
    if rank == 0:
        print("[%s, %d ranks] %s" % (label, world, "; ".join(out)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
