"""The exchange steps of the frame-sharded pipeline, on torch.distributed.

One process per GPU; NCCL over NVLink on the GPUs, gloo in the CPU tests of this module.  Nothing
here touches the coordinate stream: the three messages are the count grid, one 44-byte summary per
atom and rank, and the transition-count matrix.
"""
from __future__ import annotations


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
        d.all_reduce(counts)
    return counts


def gather_slab_summaries(slab, out=None):
    """all-gather of the per-rank slab summary [11][N] -> [world][11][N] in rank (= time) order."""
    import torch

    d = dist()
    if d is None:
        return slab.unsqueeze(0)
    world = d.get_world_size()
    if out is None:
        out = torch.empty((world,) + tuple(slab.shape), dtype=slab.dtype, device=slab.device)
    # concatenated layout along dim 0: accepted by both NCCL and gloo
    d.all_gather_into_tensor(out.view((world * slab.shape[0],) + tuple(slab.shape[1:])), slab.contiguous())
    return out


def reduce_count_matrix(matrix):
    """Sum the dense (S+2)x(S+2) transition-count matrices of the ranks in place."""
    d = dist()
    if d is not None:
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
