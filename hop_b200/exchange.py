"""The exchange steps of the frame-sharded pipeline.

One process per GPU.  Nothing here touches the coordinate stream: the three messages are the count
grid, one 44-byte summary per atom and rank, and the transition-count matrix.  On an NVSwitch box the
buffers of the three messages live in symmetric memory (`symm_empty`) and every exchange is one kernel
of this library on the pipeline's own stream -- `multimem.ld_reduce` / `multimem.st` through the
switch, csrc/exchange.cu -- between two cross-rank barriers of the symmetric-memory handle; without
multicast support (or with HOPB_NVLS=0) they go through torch.distributed: NCCL on GPUs, gloo in the
CPU tests of this module.
"""
from __future__ import annotations

import os

_SYMM = {}          # data_ptr of a symmetric buffer -> (handle, multicast address of the tensor, the tensor)
_CHANNELS = {"counts": 0, "slabs": 1, "matrix": 2}
_BARRIER_TIMEOUT_MS = 20000     # a rank that never arrives traps the kernel instead of hanging the GPU


def nvls_wanted():
    """True when the exchange buffers should be symmetric memory: several ranks on CUDA devices over NCCL and not
    switched off by HOPB_NVLS=0."""
    d = dist()
    if d is None or os.environ.get("HOPB_NVLS", "1") == "0":
        return False
    try:
        return d.get_backend() == "nccl"
    except Exception:  # noqa: BLE001
        return False


def symm_empty(shape, dtype, device):
    """A buffer in symmetric memory, mapped on every rank and registered with its multicast address.  COLLECTIVE:
    every rank must call it at the same point with the same shape.  Falls back to an ordinary tensor (the exchange
    then uses torch.distributed) when the box has no multicast support."""
    import torch

    d = dist()
    if not nvls_wanted() or _SYMM.get("unsupported"):
        return torch.empty(shape, dtype=dtype, device=device)
    try:
        import torch.distributed._symmetric_memory as symm

        t = symm.empty(*[int(v) for v in shape], dtype=dtype, device=device)
        hdl = symm.rendezvous(t, d.group.WORLD.group_name)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
    except Exception:  # noqa: BLE001 - no symmetric memory on this box / build
        mc, t, hdl = 0, None, None
    # all ranks take the same branch
    flag = torch.tensor([1 if mc else 0], dtype=torch.int32, device=device)
    d.all_reduce(flag, op=d.ReduceOp.MIN)
    if int(flag.item()) == 0:
        _SYMM["unsupported"] = True
        return torch.empty(shape, dtype=dtype, device=device)
    _SYMM[t.data_ptr()] = (hdl, mc + (t.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])), t)
    return t


def symm_entry(t):
    return None if t is None else _SYMM.get(t.data_ptr())


def _symm_allreduce(entry, t, channel, in_switch):
    """Sum of the ranks' copies of `t`, in place, by one kernel of this library between two cross-rank barriers:
    in_switch -- multimem.ld_reduce (8 bytes per instruction for integers: the small count matrix); else 16-byte
    peer loads added on the SM and one multicast store (the count grid).  Measured on 8 B200s
    (tools/nvls_bench.py, profiles/r2_nvls_exchange_8gpu.txt): 32 MB grid 123 us against NCCL's 149 us; 4 MB
    matrix 30 against 57 us."""
    from . import kernels as K

    hdl, mc, _ = entry
    hdl.barrier(channel=channel, timeout_ms=_BARRIER_TIMEOUT_MS)      # every rank's copy is complete
    if in_switch:
        K.nvls_allreduce_u32(t, mc, hdl.rank, hdl.world_size)
    else:
        off = t.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])
        K.p2p_allreduce_u32(t, [int(p) + off for p in hdl.buffer_ptrs], mc, hdl.rank, hdl.world_size)
    hdl.barrier(channel=channel, timeout_ms=_BARRIER_TIMEOUT_MS)      # every slice has reached every copy


# The 32 MB count grid: on two ranks NCCL's all-reduce (80 us) beats the library's kernel (99 us: the two barriers
# are a fifth of it); from four ranks on the kernel wins.
_COUNTS_MIN_WORLD = int(os.environ.get("HOPB_NVLS_COUNTS_MIN_WORLD", "4"))


def dist():
    """torch.distributed when a process group with more than one rank is active, else None."""
    import torch.distributed as d

    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        return d
    return None


def slab_bounds(n_frames_total, world, rank):
    """Frame range [start, stop) of a rank: contiguous slabs, remainder frames go to the first ranks."""
    base, rem = divmod(int(n_frames_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_counts(counts):
    """Sum the per-rank histograms in place (uint32 counts carried as int32: the sum is the same
    modulo 2^32).  Every rank then labels the identical grid."""
    d = dist()
    if d is not None:
        entry = symm_entry(counts)
        if entry is not None and d.get_world_size() >= _COUNTS_MIN_WORLD:
            _symm_allreduce(entry, counts, _CHANNELS["counts"], in_switch=False)
        else:
            d.all_reduce(counts)
    return counts


def gather_slab_summaries(slab, out=None):
    """all-gather of the per-rank slab summary [11][N] -> [world][11][N] in rank (= time) order."""
    import torch

    d = dist()
    if d is None:
        return slab.unsqueeze(0)
    world = d.get_world_size()
    entry = symm_entry(out)
    if entry is not None and (slab.numel() * slab.element_size()) % 16 == 0:
        from . import kernels as K

        hdl, mc, _ = entry
        slab = slab.contiguous()
        hdl.barrier(channel=_CHANNELS["slabs"], timeout_ms=_BARRIER_TIMEOUT_MS)   # nobody still reads the previous gather
        K.nvls_allgather(slab, out, mc, hdl.rank, world)
        hdl.barrier(channel=_CHANNELS["slabs"], timeout_ms=_BARRIER_TIMEOUT_MS)
        return out
    if out is None:
        out = torch.empty((world,) + tuple(slab.shape), dtype=slab.dtype, device=slab.device)
    # concatenated layout along dim 0: accepted by both NCCL and gloo
    d.all_gather_into_tensor(out.view((world * slab.shape[0],) + tuple(slab.shape[1:])), slab.contiguous())
    return out


def reduce_count_matrix(matrix):
    """Sum the dense (S+2)x(S+2) transition-count matrices of the ranks in place."""
    d = dist()
    if d is not None:
        entry = symm_entry(matrix)
        if entry is not None:
            _symm_allreduce(entry, matrix, _CHANNELS["matrix"], in_switch=True)
        else:
            d.all_reduce(matrix)
    return matrix


def bind_host_to_gpu(device_index=None):
    """Restrict this process to the CPU cores next to its GPU (sysfs `local_cpulist` of the GPU's PCI
    function), so that the pinned host buffers it allocates afterwards (first touch) and the threads that
    feed the copy engines sit on the GPU's NUMA node.  With one process per GPU and 6 GB crossing PCIe
    per step and rank, staging through the far socket is what limits the end-to-end rate on a
    two-socket 8-GPU box.  Returns a short description of what was done (for logs); never raises."""
    import os

    try:
        import torch

        i = torch.cuda.current_device() if device_index is None else int(device_index)
        p = torch.cuda.get_device_properties(i)
        bdf = "%04x:%02x:%02x.0" % (int(getattr(p, "pci_domain_id", 0)), int(p.pci_bus_id), int(p.pci_device_id))
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            return "gpu %s: numa node %d, affinity unchanged (%d cpus)" % (bdf, node, len(allowed))
        os.sched_setaffinity(0, cpus)
        return "gpu %s: numa node %d, bound to %d of %d cpus" % (bdf, node, len(cpus), len(allowed))
    except Exception as exc:  # noqa: BLE001 - an optimisation, never a reason to fail
        return "not bound (%s)" % (str(exc)[:120],)
