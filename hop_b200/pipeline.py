"""Device-resident density -> site map -> hoptraj -> transition pipeline, frames sharded over ranks.

This is the hot path the classes in density.py / sitemap.py / trajectory.py / graph.py drive one
stage at a time, put together for whole-trajectory runs (bench.py, smoke, multi-GPU).  Each rank
holds one frame slab [frame0, frame0 + F_local) of the same atoms.

Exchange steps, all through torch.distributed (NCCL on GPUs, gloo in CPU tests of
the host logic):
  * counts: all-reduce(sum) of the uint32 histogram -> every rank labels the identical grid;
  * slab summaries: all-gather of one 44-byte summary per atom and rank; every rank replays the
    slabs before its own to obtain the exact state its slab starts from;
  * transition counts: all-reduce of the dense (S+2)^2 count matrix (COO gather when too large).
No collective touches the coordinate stream.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import kernels as K
from .constants import density_conversion_factor


@dataclass
class PipelineConfig:
    solvent_threshold: float = math.e
    bulk_threshold: float = math.exp(-0.5)
    density_unit: str = "water"
    minsite: int = 1
    with_bulk: bool = True          # bulk site from the same histogram at bulk_threshold (no solute)
    write_hoptraj: bool = True
    annotate: bool = True           # site_properties of the final site map
    n_chunks: int | None = None
    event_capacity: int | None = None


@dataclass
class PipelineResult:
    counts: torch.Tensor = None
    density: torch.Tensor = None
    own: torch.Tensor = None
    member: torch.Tensor = None
    map16: torch.Tensor = None
    n_labels: int = 0                 # len(sites): interstitial + bulk + hydration sites
    hop_x: torch.Tensor = None
    hop_y: torch.Tensor = None
    events: torch.Tensor = None       # this rank's events, sorted by (edge, frame, iatom)
    edge_s0: torch.Tensor = None
    edge_s: torch.Tensor = None
    edge_start: torch.Tensor = None
    edge_count: torch.Tensor = None   # this rank's COO counts
    count_matrix: torch.Tensor = None  # global dense [(S+2),(S+2)] int32 (row/col = label+1), if small enough
    state: torch.Tensor = None        # [4][N] machine state after this rank's slab
    final_state: torch.Tensor = None  # state after the last slab (node records: hop/graph.py:259-274)
    site_props: tuple = None
    n_events_local: int = 0
    timings: dict = field(default_factory=dict)


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class HopPipeline:
    """Runs the four stages on one frame slab held in HBM.

    edges: the three float64 bin-edge arrays (grid geometry comes from frame 0 of the whole
    trajectory, hop/density.py:271-278 -- identical on every rank).
    """

    def __init__(self, edges, n_frames_total, config: PipelineConfig | None = None, device=None):
        self.cfg = config or PipelineConfig()
        self.grid = K.GridTables(edges, device=device)
        self.n_frames_total = int(n_frames_total)
        self.device = self.grid.device
        self._events = None
        self._bufs = {}
        self.num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    # -- buffers reused across runs (steady-state runs do not allocate) --------------------------
    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def _event_buffer(self, capacity):
        if self._events is None or self._events.capacity < capacity:
            self._events = K.EventBuffer(capacity, self.device)
        return self._events

    # -- stages -----------------------------------------------------------------------------------
    def histogram(self, xyz):
        counts = self._buf("counts", self.grid.shape, torch.int32)
        counts.zero_()
        K.hist_accumulate(self.grid, xyz, counts)
        d = _dist()
        if d is not None:
            d.all_reduce(counts)  # int32 sum == uint32 sum modulo 2^32
        return counts

    def site_map(self, counts):
        cfg = self.cfg
        factor = density_conversion_factor("Angstrom^{-3}", cfg.density_unit)
        density = K.density_from_counts(self.grid, counts, self.n_frames_total, factor,
                                        out=self._buf("density", self.grid.shape, torch.float64))
        own, n_sites, _ = K.label_sites(density, cfg.solvent_threshold, cfg.minsite)
        if cfg.with_bulk:
            bulk_labels, n_bulk, _ = K.label_sites(density, cfg.bulk_threshold, cfg.minsite)
            member, map16 = K.insert_bulk(own, bulk_labels if n_bulk > 0 else None, 1)
            n_labels = n_sites + 2
        else:
            member, map16 = None, K.labels_to_map16(own)
            n_labels = n_sites + 1
        props = None
        if cfg.annotate:
            props = K.site_properties(self.grid, density, own, member, n_labels)
        return density, own, member, map16, n_labels, props

    def hoptraj(self, xyz, map16, frame0, n_labels):
        cfg = self.cfg
        F, N = int(xyz.shape[0]), int(xyz.shape[1])
        if cfg.n_chunks is None:
            n_chunks, chunk_len = K.plan_chunks(F, N, self.num_sms)
        else:
            chunk_len = -(-F // cfg.n_chunks)
            n_chunks = -(-F // chunk_len)
        cap = cfg.event_capacity or max(1 << 16, (F * N) // 8)
        events = self._event_buffer(cap)
        events.reset()
        hop_x = hop_y = None
        if cfg.write_hoptraj:
            hop_x = self._buf("hop_x", (F, N), torch.float32)
            hop_y = self._buf("hop_y", (F, N), torch.float32)
        summaries = self._buf("summaries", (n_chunks, K.SUMMARY_INTS, N), torch.int32)
        K.hoptraj(self.grid, xyz, map16, frame0, n_chunks, chunk_len, events, hop_x, hop_y, summaries)

        # state at slab entry: replay the slabs of the ranks before this one
        state = self._buf("state", (K.STATE_INTS, N), torch.int32)
        d = _dist()
        final_state = None
        if d is None:
            K.hop_advance(None, state, init=True)
        else:
            rank, world = d.get_rank(), d.get_world_size()
            slab = K.hop_combine(summaries, chunk_len, F)
            gathered = self._buf("slabs", (world, K.SUMMARY_INTS, N), torch.int32)
            d.all_gather_into_tensor(gathered, slab)
            K.hop_advance(gathered[:rank] if rank > 0 else None, state, init=True)
            final_state = self._buf("final_state", (K.STATE_INTS, N), torch.int32)
            K.hop_advance(gathered, final_state, init=True)
        K.hop_fixup(summaries, chunk_len, F, state, hop_y, events)
        if final_state is None:
            final_state = state
        n_ev = events.n()
        if n_ev > events.capacity:
            raise _capacity_error(n_ev, events.capacity)
        return hop_x, hop_y, events, n_ev, state, final_state

    def transitions(self, events, n_ev, n_atoms, n_labels):
        ev, e_s0, e_s, e_start, e_count = K.events_finalize(events, n_ev, n_atoms, self.n_frames_total, n_labels)
        W = n_labels + 1
        matrix = None
        if W * W <= (1 << 24):
            matrix = self._buf("count_matrix", (W, W), torch.int32)
            matrix.zero_()
            if e_s0.numel():
                idx = (e_s0.long() + 1) * W + (e_s.long() + 1)
                matrix.view(-1).index_put_((idx,), e_count.to(torch.int32))
            d = _dist()
            if d is not None:
                d.all_reduce(matrix)
        return ev, e_s0, e_s, e_start, e_count, matrix

    def run(self, xyz, frame0=0):
        """xyz: CUDA float32 [F_local][N][3] (this rank's slab, starting at global frame `frame0`)."""
        res = PipelineResult()
        res.counts = self.histogram(xyz)
        res.density, res.own, res.member, res.map16, res.n_labels, res.site_props = self.site_map(res.counts)
        res.hop_x, res.hop_y, events, n_ev, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0, res.n_labels)
        res.n_events_local = n_ev
        (res.events, res.edge_s0, res.edge_s, res.edge_start, res.edge_count,
         res.count_matrix) = self.transitions(events, n_ev, int(xyz.shape[1]), res.n_labels)
        return res


def _capacity_error(n, cap):
    from ._lib import HopbError, HOPB_E_CAPACITY

    return HopbError(HOPB_E_CAPACITY, "event buffer too small: %d events, capacity %d" % (n, cap))


def grid_edges_from_frame0(coords0, delta=1.0, padding=2.0):
    """Bin edges exactly as DensityCreator derives them (hop/density.py:119-123, 271-278,
    hop/utilities.py:373-389, numpy.histogramdd's linspace)."""
    from .utilities import fixedwidth_bins

    coords0 = np.asarray(coords0)
    smin = np.min(coords0, axis=0) - padding
    smax = np.max(coords0, axis=0) + padding
    BINS = fixedwidth_bins(delta, smin, smax)
    return [np.linspace(lo, hi, int(n) + 1) for lo, hi, n in zip(BINS["min"], BINS["max"], BINS["Nbins"])]
