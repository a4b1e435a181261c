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

from . import exchange
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
    cache_cells: bool | None = None  # stage 1 also writes the packed hoptraj cell per atom-frame (4 B)
    coarse_map: bool = True          # block map in shared memory for stage 3 (False: always gather)
    block_map: str = "auto"          # "auto" (fine when it fits, else coarse), "fine", "coarse", "none"


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


_dist = exchange.dist


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
    def _use_cells(self, xyz):
        if self.cfg.cache_cells is not None:
            return bool(self.cfg.cache_cells)
        need = xyz.shape[0] * xyz.shape[1] * 4
        if "cells" in self._bufs and self._bufs["cells"].numel() * 4 == need:
            return True
        free, _ = torch.cuda.mem_get_info(self.device)
        return need < 0.5 * free

    def histogram(self, xyz):
        counts = self._buf("counts", self.grid.shape, torch.int32)
        counts.zero_()
        self.cells = None
        if self._use_cells(xyz):
            self.cells = self._buf("cells", (xyz.shape[0], xyz.shape[1]), torch.int32)
        K.hist_accumulate(self.grid, xyz, counts, self.cells)
        return exchange.reduce_counts(counts)

    MAX_LABELS = K.HOPB_MAX_SITES + 2

    def site_map(self, counts):
        """density -> labels (solvent and bulk threshold) -> bulk insertion -> annotation -> block map.
        Everything is enqueued without a host round trip; the number of sites stays on the device
        (self.n_sites_dev) until sync_scalars()."""
        cfg = self.cfg
        shape = self.grid.shape
        factor = density_conversion_factor("Angstrom^{-3}", cfg.density_unit)
        density = K.density_from_counts(self.grid, counts, self.n_frames_total, factor,
                                        out=self._buf("density", shape, torch.float64))
        self.n_sites_dev = self._buf("n_sites_dev", (2,), torch.int32)
        main = torch.cuda.current_stream(self.device)
        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream(device=self.device)
        own_buf = self._buf("own", shape, torch.int32)
        if cfg.with_bulk:
            # the two labellings are independent: the (sparse, cheap) hydration-site map runs on a
            # side stream with its own scratch while the bulk map occupies the main stream
            bulk_buf = self._buf("bulk_labels", shape, torch.int32)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                own = K.label_sites_async(density, cfg.solvent_threshold, cfg.minsite, own_buf, self.n_sites_dev[0:1], lane=1)
            bulk_labels = K.label_sites_async(density, cfg.bulk_threshold, cfg.minsite, bulk_buf, self.n_sites_dev[1:2])
            main.wait_stream(side)
            member, map16 = K.insert_bulk(own, bulk_labels, 1, self._buf("member", shape, torch.uint8),
                                          self._buf("map16", shape, torch.int16))
        else:
            own = K.label_sites_async(density, cfg.solvent_threshold, cfg.minsite, own_buf, self.n_sites_dev[0:1])
            member, map16 = None, K.labels_to_map16(own, self._buf("map16", shape, torch.int16))
        nb, words = K.brick_shape(shape)
        brick_buf = self._buf("brick", (nb * 64,), torch.int16)
        coarse_buf = self._buf("coarse", (words,), torch.int32)
        props = None
        if cfg.annotate:
            # the per-site reductions are not needed by stage 3: they run beside the brick map
            L = self.MAX_LABELS
            out = (self._buf("p_count", (L,), torch.int64), self._buf("p_mean", (L,), torch.float64),
                   self._buf("p_std", (L,), torch.float64), self._buf("p_center", (L, 3), torch.float64))
            side.wait_stream(main)
            with torch.cuda.stream(side):
                props = K.site_properties(self.grid, density, own, member, L, out=out)
            self._props_done = torch.cuda.Event()
            self._props_done.record(side)
        fine_buf = self._buf("fine", (K.fine_map_bytes(shape),), torch.uint8)
        self.brick = K.brick_map(map16, brick_buf, coarse_buf, fine_buf, self._buf("cell2", (nb * 4,), torch.int32))
        return density, own, member, map16, props

    def sync_scalars(self, events):
        """The one host synchronisation of a run: number of sites and number of events."""
        ev = getattr(self, "_props_done", None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)   # annotation ran beside stage 3
            self._props_done = None
        host = getattr(self, "_scalars_host", None)
        if host is None:
            host = self._scalars_host = torch.zeros(4, dtype=torch.int64).pin_memory()
        host[0:2].copy_(self.n_sites_dev, non_blocking=True)
        host[2:3].copy_(events.count, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        n_sites, n_bulk, n_ev = int(host[0]), int(host[1]), int(host[2])
        if n_sites > K.HOPB_MAX_SITES or n_bulk > K.HOPB_MAX_SITES:
            from ._lib import TooManySitesError

            raise TooManySitesError("map_sites: %d sites (bulk map: %d); hop's int16 site map holds at most %d"
                                    % (n_sites, n_bulk, K.HOPB_MAX_SITES))
        n_labels = n_sites + (2 if self.cfg.with_bulk else 1)
        return n_labels, n_ev

    def hoptraj(self, xyz, map16, frame0):
        cfg = self.cfg
        F, N = int(xyz.shape[0]), int(xyz.shape[1])
        if cfg.n_chunks is None:
            n_chunks, chunk_len = K.plan_chunks(F, N, self.num_sms)
        else:
            chunk_len = -(-F // cfg.n_chunks)
            n_chunks = -(-F // chunk_len)
        cap = cfg.event_capacity or max(1 << 20, (F * N) // 64)
        if self._events is not None:
            cap = max(cap, self._events.capacity)
        events = self._event_buffer(cap)
        events.reset()
        hop_x = hop_y = None
        if cfg.write_hoptraj:
            hop_x = self._buf("hop_x", (F, N), torch.float32)
            hop_y = self._buf("hop_y", (F, N), torch.float32)
        summaries = self._buf("summaries", (n_chunks, K.SUMMARY_INTS, N), torch.int32)
        K.hoptraj(self.grid, xyz, self.brick, frame0, n_chunks, chunk_len, events, hop_x, hop_y, summaries,
                  cells=getattr(self, "cells", None), block_map=cfg.block_map if cfg.coarse_map else "none")

        # state at slab entry: replay the slabs of the ranks before this one
        state = self._buf("state", (K.STATE_INTS, N), torch.int32)
        d = _dist()
        final_state = None
        if d is None:
            K.hop_advance(None, state, init=True)
        else:
            rank, world = d.get_rank(), d.get_world_size()
            slab = K.hop_combine(summaries, chunk_len, F)
            gathered = exchange.gather_slab_summaries(slab, self._buf("slabs", (world, K.SUMMARY_INTS, N), torch.int32))
            K.hop_advance(gathered[:rank] if rank > 0 else None, state, init=True)
            final_state = self._buf("final_state", (K.STATE_INTS, N), torch.int32)
            K.hop_advance(gathered, final_state, init=True)
        K.hop_fixup(summaries, chunk_len, F, state, hop_y, events)
        if final_state is None:
            final_state = state
        return hop_x, hop_y, events, state, final_state

    def transitions(self, events, n_ev, n_atoms, n_labels):
        W = n_labels + 1
        matrix = None
        if W * W <= (1 << 24):
            matrix = self._buf("count_matrix", (W, W), torch.int32)
            matrix.zero_()
        cap = events.capacity
        out = (self._buf("ev_sorted", (max(cap, 1), 6), torch.int32), self._buf("e_s0", (max(cap, 1),), torch.int32),
               self._buf("e_s", (max(cap, 1),), torch.int32), self._buf("e_start", (max(cap, 1),), torch.int64),
               self._buf("e_count", (max(cap, 1),), torch.int64))
        ev, e_s0, e_s, e_start, e_count = K.events_finalize(events, n_ev, n_atoms, self.n_frames_total, n_labels,
                                                            count_matrix=matrix, out=out)
        if matrix is not None:
            exchange.reduce_count_matrix(matrix)
        return ev, e_s0, e_s, e_start, e_count, matrix

    def _mark(self, marks, name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    def run(self, xyz, frame0=0, marks=None):
        """xyz: CUDA float32 [F_local][N][3] (this rank's slab, starting at global frame `frame0`).
        `marks` (a list) receives (stage name, CUDA event) pairs recorded on the current stream at
        every stage boundary, for per-stage device timing."""
        res = PipelineResult()
        self._mark(marks, "start")
        res.counts = self.histogram(xyz)
        self._mark(marks, "S1_histogram")
        res.density, res.own, res.member, res.map16, res.site_props = self.site_map(res.counts)
        self._mark(marks, "S2_sitemap")
        res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
        res.n_labels, n_ev = self.sync_scalars(events)
        if n_ev > events.capacity:      # rare: more transitions than budgeted -- grow the buffer and redo stage 3
            self._event_buffer(n_ev + n_ev // 8)
            res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
            res.n_labels, n_ev = self.sync_scalars(events)
        self._mark(marks, "S3S4_hoptraj")
        res.n_events_local = n_ev
        (res.events, res.edge_s0, res.edge_s, res.edge_start, res.edge_count,
         res.count_matrix) = self.transitions(events, n_ev, int(xyz.shape[1]), res.n_labels)
        self._mark(marks, "S4_transitions")
        return res

    def run_host(self, xyz_host, frame0=0, out_x=None, out_y=None, block_bytes=256 << 20):
        """End-to-end run from HOST coordinates to HOST results (what the Python classes do for a
        user): xyz_host is a pinned CPU float32 tensor [F_local][N][3].  Returns
        (PipelineResult, host dict); see run_host_begin."""
        return self.run_host_begin(xyz_host, frame0, out_x, out_y, block_bytes).finish()

    def run_host_begin(self, xyz_host, frame0=0, out_x=None, out_y=None, block_bytes=256 << 20):
        """Start an end-to-end run and return a HostRun; HostRun.finish() waits for the hopping
        trajectory to be in host memory and returns (PipelineResult, host dict).

        Coordinates go to HBM in blocks on an upload stream while the histogram kernel consumes the
        blocks already there; they stay resident for stage 3.  The events are sorted and the small
        results (sorted events, COO transition counts, final per-atom state) copied to the host, then
        the hopping trajectory planes stream back into pinned host tensors (out_x / out_y,
        [F_local][N] float32) on a download stream.  Uploads and downloads use different streams
        (and copy engines), so a caller that begins run k+1 before finishing run k -- with a second
        pair of out_x / out_y -- overlaps the download of k with the upload of k+1; PCIe is full
        duplex.  The device buffers are shared between runs: stage 3 of the next run waits for the
        download of this one."""
        F, N = int(xyz_host.shape[0]), int(xyz_host.shape[1])
        dev = self.device
        xyz = self._buf("xyz_resident", (F, N, 3), torch.float32)
        counts = self._buf("counts", self.grid.shape, torch.int32)
        counts.zero_()
        self.cells = self._buf("cells", (F, N), torch.int32) if self._use_cells(xyz) else None
        if getattr(self, "_h2d_stream", None) is None:
            self._h2d_stream = torch.cuda.Stream(device=dev)
            self._d2h_stream = torch.cuda.Stream(device=dev)
            self._d2h_done = None
        h2d, d2h = self._h2d_stream, self._d2h_stream
        main = torch.cuda.current_stream(dev)
        h2d.wait_stream(main)
        block = max(1, block_bytes // (N * 12))
        for f0 in range(0, F, block):
            f1 = min(F, f0 + block)
            with torch.cuda.stream(h2d):
                xyz[f0:f1].copy_(xyz_host[f0:f1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d)
            main.wait_event(ev)
            K.hist_accumulate(self.grid, xyz[f0:f1], counts, None if self.cells is None else self.cells[f0:f1])
        exchange.reduce_counts(counts)
        res = PipelineResult()
        res.counts = counts
        res.density, res.own, res.member, res.map16, res.site_props = self.site_map(counts)
        if self._d2h_done is not None:
            main.wait_event(self._d2h_done)      # the previous run's planes are still being downloaded
        res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
        res.n_labels, n_ev = self.sync_scalars(events)
        if n_ev > events.capacity:
            self._event_buffer(n_ev + n_ev // 8)
            res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
            res.n_labels, n_ev = self.sync_scalars(events)
        res.n_events_local = n_ev
        if out_x is None:
            out_x = torch.empty((F, N), dtype=torch.float32).pin_memory()
            out_y = torch.empty((F, N), dtype=torch.float32).pin_memory()
        (res.events, res.edge_s0, res.edge_s, res.edge_start, res.edge_count,
         res.count_matrix) = self.transitions(events, n_ev, N, res.n_labels)
        # small results first (these copies synchronise), then the planes: 8 B per atom-frame
        host = dict(hop_x=out_x, hop_y=out_y, events=res.events.cpu(), edge_s0=res.edge_s0.cpu(),
                    edge_s=res.edge_s.cpu(), edge_count=res.edge_count.cpu(), final_state=res.final_state.cpu())
        d2h.wait_stream(main)
        with torch.cuda.stream(d2h):
            out_x.copy_(res.hop_x, non_blocking=True)
            out_y.copy_(res.hop_y, non_blocking=True)
            done = torch.cuda.Event()
            done.record(d2h)
        self._d2h_done = done
        return HostRun(res, host, done)


class HostRun(object):
    """A run_host_begin in flight: finish() returns (PipelineResult, host dict) once the hopping
    trajectory planes are in host memory."""

    def __init__(self, result, host, done):
        self.result, self.host, self._done = result, host, done

    def finish(self):
        self._done.synchronize()
        return self.result, self.host


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
