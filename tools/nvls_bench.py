"""The count-grid all-reduce in isolation on N GPUs (run under torchrun): NCCL vs the library's two kernels on
symmetric memory (in-switch multimem.ld_reduce; peer loads + multicast store), each between two cross-rank barriers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from hop_b200 import exchange, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 201 ** 3
t = exchange.symm_empty((n,), torch.int32, torch.device("cuda", lr))
entry = exchange.symm_entry(t)
assert entry is not None, "no multicast support"
hdl, mc, _ = entry
off = t.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])
peers = [int(p) + off for p in hdl.buffer_ptrs]
ref = torch.randint(0, 1000, (n,), dtype=torch.int32, device="cuda", generator=torch.Generator("cuda").manual_seed(7 + rank))
want = ref.clone()
dist.all_reduce(want)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def nvls():
    hdl.barrier(channel=0); K.nvls_allreduce_u32(t, mc, rank, world); hdl.barrier(channel=0)


def p2p():
    hdl.barrier(channel=0); K.p2p_allreduce_u32(t, peers, mc, rank, world); hdl.barrier(channel=0)


for name, fn in (("nvls", nvls), ("p2p", p2p)):
    t.copy_(ref); torch.cuda.synchronize(); dist.barrier()
    fn(); torch.cuda.synchronize()
    ok = bool(torch.equal(t, want))
    us = timeit(fn)
    if rank == 0:
        print("%-5s %d words, %d ranks: correct=%s  %.1f us (barrier + kernel + barrier)" % (name, n, world, ok, us), flush=True)
x = ref.clone()
us = timeit(lambda: dist.all_reduce(x))
us_b = timeit(lambda: hdl.barrier(channel=0))
if rank == 0:
    print("nccl  all_reduce %d words: %.1f us" % (n, us), flush=True)
    print("barrier alone: %.1f us" % us_b, flush=True)
# the small messages: slab summaries (all-gather of 11 x N int32 per rank) and the count matrix (1024^2 int32)
N = 30000
slab = torch.randint(0, 1000, (11, N), dtype=torch.int32, device="cuda")
out = exchange.symm_empty((world, 11, N), torch.int32, torch.device("cuda", lr))
e2 = exchange.symm_entry(out)
out2 = torch.empty((world * 11, N), dtype=torch.int32, device="cuda")


def ag():
    e2[0].barrier(channel=1); K.nvls_allgather(slab, out, e2[1], rank, world); e2[0].barrier(channel=1)


ag(); torch.cuda.synchronize()
dist.all_gather_into_tensor(out2, slab); torch.cuda.synchronize()
ok = bool(torch.equal(out.view(world * 11, N), out2))
us_a, us_n = timeit(ag), timeit(lambda: dist.all_gather_into_tensor(out2, slab))
if rank == 0:
    print("all-gather 11 x %d int32 per rank: multicast store correct=%s %.1f us, nccl %.1f us" % (N, ok, us_a, us_n), flush=True)
m = exchange.symm_empty((1024, 1024), torch.int32, torch.device("cuda", lr))
e3 = exchange.symm_entry(m)
off3 = m.data_ptr() - int(e3[0].buffer_ptrs[rank])
peers3 = [int(p) + off3 for p in e3[0].buffer_ptrs]
m2 = torch.ones((1024, 1024), dtype=torch.int32, device="cuda")
us_p = timeit(lambda: (e3[0].barrier(channel=2), K.p2p_allreduce_u32(m, peers3, e3[1], rank, world), e3[0].barrier(channel=2)))
us_v = timeit(lambda: (e3[0].barrier(channel=2), K.nvls_allreduce_u32(m, e3[1], rank, world), e3[0].barrier(channel=2)))
us_n = timeit(lambda: dist.all_reduce(m2))
if rank == 0:
    print("count matrix 1024^2: p2p %.1f us, nvls %.1f us, nccl %.1f us" % (us_p, us_v, us_n), flush=True)
dist.barrier()
dist.destroy_process_group()
