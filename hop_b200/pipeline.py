"""Device-resident density -> site map -> hoptraj -> transition pipeline, frames sharded over ranks.

This is the hot path the classes in density.py / sitemap.py / trajectory.py / graph.py drive one
stage at a time, put together for whole-trajectory runs (bench.py, smoke, multi-GPU).  Each rank
holds one frame slab [frame0, frame0 + F_local) of the same atoms.

Exchange steps (hop_b200/exchange.py: kernels of the library on NVSwitch multicast
memory where the box has it, else torch.distributed -- NCCL on GPUs, gloo in the CPU tests of the
host logic):
  * counts: all-reduce(sum) of the uint32 histogram -> every rank labels the identical grid;
  * slab summaries: all-gather of one 44-byte summary per atom and rank; every rank replays the
    slabs before its own to obtain the exact state its slab starts from;
  * transition counts: all-reduce of the dense (S+2)^2 count matrix (COO gather when too large).
No collective touches the coordinate stream.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

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
    async_results: bool = True       # from the second run on: no host synchronisation inside run()
    label_from_counts: bool = True   # threshold the counts instead of the density (identical labels)
    fused_site_map: bool = True      # stage 2 as one cooperative launch (hopb_site_map_fused) where the grid allows
    periodic_box: tuple | None = None  # (origin[3], length[3]): wrap coordinates before binning; None = hop's behaviour
    plane_dtype: str = "int16"       # element type of hop_x / hop_y in HBM: "int16" (lossless: labels are values of hop's
                                     # int16 site map or -1; stage 3 writes 4 B per atom-frame) or "float32" (the DCD
                                     # record type, 8 B per atom-frame); files and the class API always see float32

    def plane_torch_dtype(self):
        if self.plane_dtype not in ("int16", "float32"):
            raise ValueError("plane_dtype must be 'int16' or 'float32'")
        return torch.int16 if self.plane_dtype == "int16" else torch.float32


class PipelineResult(object):
    """What a run leaves behind.  The tensors are the pipeline's own buffers: they stay valid until the
    next run.  The fields that depend on counts only the device knows when run() returns -- n_labels,
    n_events_local, events, edge_*, count_matrix, site_props -- are resolved on first access (one
    wait for the finalisation stream), so that back-to-back runs never stall the host."""

    _LAZY = ("n_labels", "n_events_local", "events", "edge_s0", "edge_s", "edge_start", "edge_count", "count_matrix",
             "site_props")

    def __init__(self):
        self.counts = None
        self.density = None
        self.own = None
        self.member = None
        self.map16 = None
        self.hop_x = None
        self.hop_y = None
        self.state = None         # [4][N] machine state after this rank's slab
        self.final_state = None   # state after the last slab (node records: hop/graph.py:259-274)
        self.timings = {}
        self._pending = None
        # lazily resolved (see class docstring):
        #   n_labels        len(sites): interstitial + bulk + hydration sites
        #   events          this rank's events, sorted by (edge, frame, iatom)
        #   edge_s0, edge_s, edge_start, edge_count   this rank's COO counts
        #   count_matrix    global dense [(S+2),(S+2)] int32 (row/col = label+1), if small enough
        #   site_props, n_events_local

    def __getattr__(self, name):
        # only reached when the attribute has not been set yet
        if name in PipelineResult._LAZY:
            pending = self.__dict__.get("_pending")
            if pending is not None:
                self.__dict__["_pending"] = None
                pending.resolve(self)
                return self.__dict__[name]
        raise AttributeError(name)

    def wait(self):
        """Resolve the lazy fields now."""
        return self.n_labels


class _PendingFinalize(object):
    """The finalisation of one run in flight on the pipeline's side stream."""

    def __init__(self, pipe, done, host, events, out, matrix, props, props_done, redo):
        self.pipe, self.done, self.host, self.events, self.out = pipe, done, host, events, out
        self.matrix, self.props, self.props_done, self.redo = matrix, props, props_done, redo
        self.run_index = pipe._run_index

    def resolve(self, res):
        from . import _lib

        with _lib.lanes(self.pipe.lane_base):
            self._resolve(res)

    def _resolve(self, res):
        if self.pipe._run_index > self.run_index + 2:
            # outputs alternate between two buffer sets: after two more runs this result's buffers hold another run's
            raise RuntimeError("PipelineResult resolved too late: its buffers were reused by run %d (resolve a result, or "
                               "call wait(), before queuing the run after next)" % (self.run_index + 2))
        self.done.synchronize()
        n_all, n_edges, n_labels, status = (int(v) for v in self.host.tolist())
        pipe = self.pipe
        pipe._fin_hosts.append(self.host)
        if status & K.FIN_TOO_MANY_SITES:
            from ._lib import TooManySitesError

            raise TooManySitesError("map_sites: more than %d sites; hop's int16 site map cannot hold them"
                                    % K.HOPB_MAX_SITES)
        if status & (K.FIN_OVERFLOW | K.FIN_MATRIX_SKIPPED):
            # rare: more transitions than budgeted, or more sites than the count matrix was sized for
            # -- redo the tail of the run synchronously with what is known now
            pipe._w_known = n_labels + 1
            self.redo(res, n_all if status & K.FIN_OVERFLOW else None, n_all)
            return
        sorted_ev, e_s0, e_s, e_start, e_count = self.out
        res.events = sorted_ev[:n_all]
        res.edge_s0, res.edge_s = e_s0[:n_edges], e_s[:n_edges]
        res.edge_start, res.edge_count = e_start[:n_edges], e_count[:n_edges]
        W = n_labels + 1
        res.count_matrix = None if self.matrix is None else self.matrix[:W, :W]
        res.n_labels, res.n_events_local = n_labels, n_all
        if self.props_done is not None:
            self.props_done.synchronize()
        res.site_props = self.props
        if _dist() is None:       # with several ranks the matrix size must come from something all ranks share: the
            pipe._w_known = W     # first (synchronous) run's site count, not from whichever rank looks at a result


_dist = exchange.dist


class HopPipeline:
    """Runs the four stages on one frame slab held in HBM.

    edges: the three float64 bin-edge arrays (grid geometry comes from frame 0 of the whole
    trajectory, hop/density.py:271-278 -- identical on every rank).
    """

    def __init__(self, edges, n_frames_total, config: PipelineConfig | None = None, device=None, lane_base=0):
        self.cfg = config or PipelineConfig()
        self.lane_base = int(lane_base)   # handles (scratch memory) of its own: see PipelinePair
        self.grid = K.GridTables(edges, device=device)
        if self.cfg.periodic_box is not None:
            self.grid.set_periodic_box(*self.cfg.periodic_box)
        self.n_frames_total = int(n_frames_total)
        self.device = self.grid.device
        self._events = None
        self._bufs = {}
        self._w_known = None      # n_labels + 1 of the last resolved run: sizes the count matrix of the next
        self._s4_done = None      # finalisation of the previous run (reads the event buffer)
        self._run_index = 0
        self.num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    def set_edges(self, edges):
        """Another grid geometry for the next run (a batch of trajectories: every one spans its own grid in frame 0,
        hop/density.py:271-278).  Keeps the buffers when the shape stays the same."""
        if all(len(a) == len(b) and np.array_equal(a, b) for a, b in zip(edges, self.grid.edges)):
            return
        # grids seen before are kept: dropping one would free device tables that queued kernels may still read
        cache = self.__dict__.setdefault("_grids", {})
        key = b"".join(np.ascontiguousarray(e, dtype=np.float64).tobytes() for e in edges)
        cache.setdefault(b"".join(e.tobytes() for e in self.grid.edges), self.grid)
        grid = cache.get(key)
        if grid is None:
            grid = cache[key] = K.GridTables(edges, device=self.device.index)
            if self.cfg.periodic_box is not None:
                grid.set_periodic_box(*self.cfg.periodic_box)
        self.grid = grid

    # -- buffers reused across runs (steady-state runs do not allocate) --------------------------
    def _buf(self, name, shape, dtype, symm=False):
        """symm: a buffer of one of the three exchanges -- symmetric memory when several ranks run on an NVSwitch
        box (allocation is then collective: all ranks reach it at the same point of the same run)."""
        t = self._bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            if symm and exchange.nvls_wanted():
                t = exchange.symm_empty(shape, dtype, self.device)
            else:
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

    def _zeroed_counts(self):
        """The count grid of this run, already zeroed: two grids alternate, and the one of the NEXT run is cleared on a
        side stream while this run's histogram is busy (a 32 MB memset otherwise sits in front of every histogram).
        The counts of a result stay valid until the next run starts, as before."""
        par = self._run_index & 1
        main = torch.cuda.current_stream(self.device)
        zs = getattr(self, "_zero_stream", None)
        if zs is None:
            zs = self._zero_stream = torch.cuda.Stream(device=self.device)
            self._zeroed = {}
        counts = self._buf("counts%d" % par, self.grid.shape, torch.int32, symm=True)
        ready = self._zeroed.pop(par, None)
        if ready is not None and ready[0] is counts:
            main.wait_event(ready[1])
        else:
            counts.zero_()
        other = self._buf("counts%d" % (1 - par), self.grid.shape, torch.int32, symm=True)
        zs.wait_stream(main)                  # everything queued so far that reads the other grid (the previous run) is done
        with torch.cuda.stream(zs):
            other.zero_()
            ev = torch.cuda.Event()
            ev.record(zs)
        self._zeroed[1 - par] = (other, ev)
        return counts

    def histogram(self, xyz):
        counts = self._zeroed_counts()
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
        if cfg.fused_site_map and cfg.label_from_counts and K.site_map_fused_ok(self.grid):
            return self._site_map_fused(counts, factor)
        # density / labels / membership alternate between two buffer sets: the annotation of a run is
        # queued behind its stage 3 and may execute as late as the end of the next run's histogram, so it
        # must not hold up the next run's stage 2; only the run after that reuses its inputs
        par = "%d" % (self._run_index & 1)
        busy = self.__dict__.setdefault("_props_busy", {}).pop(par, None)
        if busy is not None:
            torch.cuda.current_stream(self.device).wait_event(busy)
        # two buffers, alternating: the finalisation of the previous run may still read its own
        self.n_sites_dev = self._buf("n_sites_dev%d" % (self._run_index & 1), (2,), torch.int32)
        main = torch.cuda.current_stream(self.device)
        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream(device=self.device)
            self._density_stream = torch.cuda.Stream(device=self.device)
        density_buf = self._buf("density" + par, shape, torch.float64)
        from_counts = cfg.label_from_counts and self.grid.width_classes > 0
        if from_counts:
            # the labelling thresholds the counts directly (identical labels), so the density -- four
            # double divisions per cell -- is computed beside it instead of in front of it
            dstream = self._density_stream
            dstream.wait_stream(main)
            with torch.cuda.stream(dstream):
                # two CTAs per SM: it trickles along on spare FP64 throughput instead of filling every SM
                # in front of the labelling kernels
                density = K.density_from_counts(self.grid, counts, self.n_frames_total, factor, out=density_buf,
                                                ctas_per_sm=2)
                density_done = torch.cuda.Event()
                density_done.record(dstream)
            nf = self.n_frames_total

            def label(thr, out, n_dev, lane=0):
                return K.label_sites_counts_async(self.grid, counts, nf, factor, thr, cfg.minsite, out, n_dev, lane=lane)
        else:
            density = K.density_from_counts(self.grid, counts, self.n_frames_total, factor, out=density_buf)
            density_done = None

            def label(thr, out, n_dev, lane=0):
                return K.label_sites_async(density, thr, cfg.minsite, out, n_dev, lane=lane)
        own_buf = self._buf("own" + par, shape, torch.int32)
        if cfg.with_bulk:
            # the two labellings are independent: the (sparse, cheap) hydration-site map runs on a
            # side stream with its own scratch while the bulk map occupies the main stream
            bulk_buf = self._buf("bulk_labels", shape, torch.int32)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                own = label(cfg.solvent_threshold, own_buf, self.n_sites_dev[0:1], lane=1)
            bulk_labels = label(cfg.bulk_threshold, bulk_buf, self.n_sites_dev[1:2])
            main.wait_stream(side)
            member, map16 = K.insert_bulk(own, bulk_labels, 1, self._buf("member" + par, shape, torch.uint8),
                                          self._buf("map16", shape, torch.int16))
        else:
            own = label(cfg.solvent_threshold, own_buf, self.n_sites_dev[0:1])
            member, map16 = None, K.labels_to_map16(own, self._buf("map16", shape, torch.int16))
        if density_done is not None:
            main.wait_event(density_done)   # long finished; keeps res.density ordered on the main stream
        nb, words = K.brick_shape(shape)
        brick_buf = self._buf("brick", (nb * 64,), torch.int16)
        coarse_buf = self._buf("coarse", (words,), torch.int32)
        # the per-site reductions are not needed by stage 3; annotate() queues them behind it
        self._annotate_args = (density, own, member)
        fine_buf = self._buf("fine", (K.fine_map_bytes(shape),), torch.uint8)
        self.brick = K.brick_map(map16, brick_buf, coarse_buf, fine_buf, self._buf("cell2", (nb * 4,), torch.int32))
        return density, own, member, map16, None

    def _site_map_fused(self, counts, factor):
        """Stage 2 as one launch: density, both labellings, bulk insertion, painted and bricked map, block
        maps and the annotation (csrc/sitemap_fused.cu)."""
        cfg = self.cfg
        shape = self.grid.shape
        par = "%d" % (self._run_index & 1)     # results of a run stay intact while the next one is queued
        self.n_sites_dev = self._buf("n_sites_dev" + par, (2,), torch.int32)
        density = self._buf("density" + par, shape, torch.float64)
        own = self._buf("own" + par, shape, torch.int32)
        member = self._buf("member" + par, shape, torch.uint8) if cfg.with_bulk else None
        map16 = self._buf("map16", shape, torch.int16)
        nb, words = K.brick_shape(shape)
        self.brick = K.BrickMap(self._buf("brick", (nb * 64,), torch.int16), self._buf("coarse", (words,), torch.int32),
                                self._buf("fine", (K.fine_map_bytes(shape),), torch.uint8),
                                self._buf("cell2", (nb * 4,), torch.int32))
        props = None
        if cfg.annotate:
            L = self.MAX_LABELS
            props = (self._buf("p_count" + par, (L,), torch.int64), self._buf("p_mean" + par, (L,), torch.float64),
                     self._buf("p_std" + par, (L,), torch.float64), self._buf("p_center" + par, (L, 3), torch.float64))
        K.site_map_fused(self.grid, counts, self.n_frames_total, factor, cfg.solvent_threshold, cfg.bulk_threshold,
                         cfg.with_bulk, cfg.minsite, density, own, member, map16, self.brick, self.n_sites_dev, props=props)
        self._annotate_args = None
        self._props = props
        return density, own, member, map16, None

    def annotate(self):
        """site_properties of the site map just built (volume, occupancy, centre per site), on the side
        stream.  Called after stage 3 has been launched: the hoptraj kernel needs whole SMs (1024
        threads x 64 registers), so reductions already running there would hold it up, while queued behind
        it they run beside the fix-up, the event ordering and the next run's histogram."""
        if self._annotate_args is None:      # the one-launch site map has annotated already
            return
        self._props = None
        if not self.cfg.annotate:
            return
        density, own, member = self._annotate_args
        side = self._side_stream
        ready = torch.cuda.Event()      # behind everything stage 3 has queued on the main stream
        ready.record(torch.cuda.current_stream(self.device))
        L = self.MAX_LABELS
        out = (self._buf("p_count", (L,), torch.int64), self._buf("p_mean", (L,), torch.float64),
               self._buf("p_std", (L,), torch.float64), self._buf("p_center", (L, 3), torch.float64))
        side.wait_event(ready)
        with torch.cuda.stream(side):
            self._props = K.site_properties(self.grid, density, own, member, L, out=out)
        self._props_done = torch.cuda.Event()
        self._props_done.record(side)
        self._props_busy["%d" % (self._run_index & 1)] = self._props_done

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
        if not self.cfg.with_bulk:
            n_bulk = 0          # the second counter is only written by the bulk labelling
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
        if self._s4_done is not None:     # the previous run's finalisation reads the event buffer
            torch.cuda.current_stream(self.device).wait_event(self._s4_done)
            self._s4_done = None
        if getattr(self, "_events_clean", None) is events:
            self._events_clean = None     # the previous run's finalisation left the counter at zero (on its own stream)
        else:
            events.reset()
        hop_x = hop_y = None
        if cfg.write_hoptraj:
            hop_x = self._buf("hop_x", (F, N), cfg.plane_torch_dtype())
            hop_y = self._buf("hop_y", (F, N), cfg.plane_torch_dtype())
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
            gathered = exchange.gather_slab_summaries(slab, self._buf("slabs", (world, K.SUMMARY_INTS, N), torch.int32, symm=True))
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

    def _finalize_sync(self, res, events, xyz, frame0, n_atoms=None):
        """Tail of a run with the host in the loop: wait for stage 3, read the counts, sort."""
        res.n_labels, n_ev = self.sync_scalars(events)
        n_local = n_ev
        d = _dist()
        if d is not None:
            # the redo below runs collectives (slab summaries, count matrix): every rank must take the same branch,
            # so the decision is made on the largest event count of any rank
            t = torch.tensor([n_ev, events.capacity], dtype=torch.int64, device=self.device)
            d.all_reduce(t, op=d.ReduceOp.MAX)
            n_ev = max(n_ev, int(t[0].item())) if int(t[0].item()) > events.capacity else n_ev
        if n_ev > events.capacity and xyz is None:
            self._event_buffer(n_ev + n_ev // 8)     # the next attempt fits
            raise _capacity_error(n_ev, events.capacity)
        if n_ev > events.capacity:      # rare: more transitions than budgeted -- grow the buffer and redo stage 3
            self._event_buffer(n_ev + n_ev // 8)
            res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
            res.n_labels, n_local = self.sync_scalars(events)
        n_ev = n_local
        res.n_events_local = n_ev
        (res.events, res.edge_s0, res.edge_s, res.edge_start, res.edge_count,
         res.count_matrix) = self.transitions(events, n_ev, int(xyz.shape[1]) if n_atoms is None else n_atoms, res.n_labels)
        res.site_props = self._props
        res._pending = None
        self._w_known = res.n_labels + 1

    def _finalize_async(self, res, events, xyz, frame0):
        """Tail of a run without the host: the event ordering is queued on a side stream with its sizes
        read on the device (csrc/events.cu), the four result scalars come back through pinned memory
        and are looked at when a lazy field of the result is first used.  The next run may start right
        away; only its stage 3 waits for this finalisation (it reuses the event buffer)."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        s4 = getattr(self, "_s4_stream", None)
        if s4 is None:
            # high priority: its cooperative kernel needs one block on every SM at once and would
            # otherwise wait until the next run's histogram (tens of thousands of short-lived CTAs) drains
            s4 = self._s4_stream = torch.cuda.Stream(device=dev, priority=-1)
            self._fin_hosts = []      # pool of pinned 4-word result buffers; one per unresolved result
        props_done = getattr(self, "_props_done", None)   # the next run's stage 2 waits for it (site_map)
        self._props_done = None
        w_cap = 256
        while w_cap < self._w_known + self._w_known // 4 + 8:
            w_cap *= 2
        # two sets of outputs, alternating: the results of a run stay intact while the next one is finalised
        par = "%d" % (self._run_index & 1)
        matrix = self._buf("count_matrix_cap" + par, (w_cap, w_cap), torch.int32, symm=True) if w_cap * w_cap <= (1 << 24) else None
        cap = max(events.capacity, 1)
        out = (self._buf("ev_sorted" + par, (cap, 6), torch.int32), self._buf("e_s0" + par, (cap,), torch.int32),
               self._buf("e_s" + par, (cap,), torch.int32), self._buf("e_start" + par, (cap,), torch.int64),
               self._buf("e_count" + par, (cap,), torch.int64))
        result_dev = self._buf("fin_result", (4,), torch.int64)
        host = self._fin_hosts.pop() if self._fin_hosts else torch.zeros(4, dtype=torch.int64).pin_memory()
        s4.wait_stream(main)
        with torch.cuda.stream(s4):
            K.events_finalize_dev(events, int(xyz.shape[1]), self.n_frames_total, self.n_sites_dev[0:1],
                                  2 if self.cfg.with_bulk else 1, out, matrix, result_dev)
            if matrix is not None:
                exchange.reduce_count_matrix(matrix)
            if _dist() is not None:
                # every rank learns whether ANY rank ran out of event space: a rank-local recovery would issue
                # collectives the other ranks never match
                flags = self._buf("fin_flags" + par, (2,), torch.int64)
                flags[0] = result_dev[3] & K.FIN_OVERFLOW
                flags[1] = result_dev[0]
                _dist().all_reduce(flags, op=_dist().ReduceOp.MAX)
                result_dev[3] |= flags[0]
                host_any = self._host_buf("fin_any" + par, (2,), torch.int64)
                host_any.copy_(flags, non_blocking=True)
            host.copy_(result_dev, non_blocking=True)
            events.reset()        # for the next run's stage 3, which waits for `done` anyway: one launch less in front of it
            done = torch.cuda.Event()
            done.record(s4)
        self._s4_done = done
        self._events_clean = events

        run_index = self._run_index

        distributed = _dist() is not None

        def redo(res_, n_needed, n_all=0):
            if distributed and n_needed is not None:
                # more transitions than budgeted on some rank (the flag is all-reduced, so every rank that looks at
                # its result gets here): results are resolved lazily and not in lockstep, so the collective redo is
                # left to the caller -- the buffers are grown for the next run
                need = int(self._host_bufs["fin_any%d" % (run_index & 1)][1].item())
                self._event_buffer(need + need // 8)
                raise _capacity_error(need, events.capacity)
            if self._run_index != run_index + 1:
                if n_needed is not None:
                    self._event_buffer(n_needed + n_needed // 8)   # at least the next runs will fit
                raise _capacity_error(n_needed or 0, events.capacity) if n_needed is not None else RuntimeError(
                    "the count matrix of an earlier run was sized for fewer sites and its buffers have been reused")
            if n_needed is not None:
                self._event_buffer(n_needed + n_needed // 8)
                res_.hop_x, res_.hop_y, ev2, res_.state, res_.final_state = self.hoptraj(xyz, res_.map16, frame0)
            else:
                ev2 = events
                events.count.fill_(n_all)     # the finalisation left the counter at zero for the next run; the records are intact
                self._events_clean = None
            self._finalize_sync(res_, ev2, xyz, frame0)

        res._pending = _PendingFinalize(self, done, host, events, out, matrix, self._props, props_done, redo)

    def _finalize(self, res, events, xyz, frame0):
        if self.cfg.async_results and self._w_known is not None:
            self._finalize_async(res, events, xyz, frame0)
        else:
            self._finalize_sync(res, events, xyz, frame0)
        self._run_index += 1

    def _mark(self, marks, name, stream=None):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
            marks.append((name, ev))

    def run(self, xyz, frame0=0, marks=None):
        """xyz: CUDA float32 [F_local][N][3] (this rank's slab, starting at global frame `frame0`).
        `marks` (a list) receives (stage name, CUDA event) pairs recorded at every stage boundary, for
        per-stage device timing (the last one on the finalisation stream when that runs beside the
        next run)."""
        from . import _lib

        with _lib.lanes(self.lane_base):
            return self._run(xyz, frame0, marks)

    def _run(self, xyz, frame0, marks):
        res = PipelineResult()
        self._mark(marks, "start")
        res.counts = self.histogram(xyz)
        self._mark(marks, "S1_histogram")
        res.density, res.own, res.member, res.map16, _ = self.site_map(res.counts)
        self._mark(marks, "S2_sitemap")
        res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
        self.annotate()
        self._mark(marks, "S3S4_hoptraj")
        self._finalize(res, events, xyz, frame0)
        self._mark(marks, "S4_transitions", getattr(self, "_s4_stream", None) if res._pending is not None else None)
        return res

    def run_streamed(self, source, n_frames, chunk_frames, sink=None, cache_cells=None, marks=None):
        """The four stages over a trajectory that does not fit into HBM, in two passes over frame chunks
        (hop reads the trajectory twice as well: DensityCreator.create, hop/density.py:292-304, then
        HoppingTrajectory.map_dcd, hop/trajectory.py:379-394).

        source(f0, f1) -> CUDA float32 [f1 - f0][N][3]: the coordinates of frames [f0, f1) (a reader, an upload
        stage, a generator); called once per chunk and pass, unless the 4-byte brick indices of pass 1 fit into
        HBM (`cache_cells`, default: when they take less than 40 % of the free memory), in which case pass 2
        reads those instead of the coordinates.
        sink(f0, hop_x, hop_y): receives every chunk of the hopping trajectory ([m][N] views (PipelineConfig.plane_dtype) of a
        buffer that is reused for the next chunk -- write it out or reduce it before returning).

        Pass 1 accumulates the histogram chunk by chunk; stage 2 runs once; pass 2 carries the per-atom
        machine state from chunk to chunk (`hop_fixup`), so tau / tbarrier / orbit are those of one run over
        the whole trajectory.  Single rank.  Returns a PipelineResult without hop_x / hop_y."""
        from . import _lib

        with _lib.lanes(self.lane_base):
            return self._run_streamed(source, int(n_frames), int(chunk_frames), sink, cache_cells, marks)

    def _run_streamed(self, source, F, chunk, sink, cache_cells, marks):
        if _dist() is not None:
            raise RuntimeError("run_streamed is a single-rank path (shard the frames over ranks with run())")
        if F != self.n_frames_total:
            raise ValueError("n_frames must be the trajectory length the pipeline was created for")
        cfg = self.cfg
        dev = self.device
        bounds = [(f0, min(F, f0 + chunk)) for f0 in range(0, F, chunk)]
        self._mark(marks, "start")
        counts = self._buf("counts", self.grid.shape, torch.int32)
        counts.zero_()
        first = source(*bounds[0])
        N = int(first.shape[1])
        if cache_cells is None:
            free, _ = torch.cuda.mem_get_info(dev)
            cache_cells = F * N * 4 < 0.4 * free
        cells_all = self._buf("cells_all", (F, N), torch.int32) if cache_cells else None
        # ---- pass 1: histogram -------------------------------------------------------------------------
        for k, (f0, f1) in enumerate(bounds):
            xyz = first if k == 0 else source(f0, f1)
            K.hist_accumulate(self.grid, xyz, counts, None if cells_all is None else cells_all[f0:f1])
            del xyz
        del first
        self._mark(marks, "S1_histogram")
        res = PipelineResult()
        res.counts = counts
        res.density, res.own, res.member, res.map16, _ = self.site_map(counts)
        self.annotate()
        self._mark(marks, "S2_sitemap")
        # ---- pass 2: hopping trajectory + transitions ----------------------------------------------------
        m_max = max(f1 - f0 for f0, f1 in bounds)
        hx_buf = self._buf("hop_x_chunk", (m_max, N), cfg.plane_torch_dtype())
        hy_buf = self._buf("hop_y_chunk", (m_max, N), cfg.plane_torch_dtype())
        cap = cfg.event_capacity or max(1 << 20, (F * N) // 64)
        events = self._event_buffer(max(cap, self._events.capacity if self._events is not None else 0))
        if self._s4_done is not None:
            torch.cuda.current_stream(dev).wait_event(self._s4_done)
            self._s4_done = None
        self._events_clean = None
        events.reset()
        state = self._buf("state", (K.STATE_INTS, N), torch.int32)
        K.hop_advance(None, state, init=True)
        for f0, f1 in bounds:
            m = f1 - f0
            n_chunks, chunk_len = K.plan_chunks(m, N, self.num_sms) if cfg.n_chunks is None else \
                (-(-m // (-(-m // cfg.n_chunks))), -(-m // cfg.n_chunks))
            xyz = None if cells_all is not None else source(f0, f1)
            hx, hy = hx_buf[:m], hy_buf[:m]
            summaries = self._buf("summaries_%d" % n_chunks, (n_chunks, K.SUMMARY_INTS, N), torch.int32)
            K.hoptraj(self.grid, xyz, self.brick, f0, n_chunks, chunk_len, events, hx, hy, summaries,
                      cells=None if cells_all is None else cells_all[f0:f1],
                      block_map=cfg.block_map if cfg.coarse_map else "none")
            K.hop_fixup(summaries, chunk_len, m, state, hy, events)
            if sink is not None:
                sink(f0, hx, hy)
            del xyz
        self._mark(marks, "S3S4_hoptraj")
        res.state = res.final_state = state
        self._finalize_sync(res, events, None, 0, n_atoms=N)
        self._mark(marks, "S4_transitions")
        self._run_index += 1
        return res

    def run_host(self, xyz_host, frame0=0, out_x=None, out_y=None, block_bytes=256 << 20):
        """End-to-end run from HOST coordinates to HOST results (what the Python classes do for a
        user): xyz_host is a pinned CPU float32 tensor [F_local][N][3].  Returns
        (PipelineResult, host dict); see run_host_begin."""
        return self.run_host_begin(xyz_host, frame0, out_x, out_y, block_bytes).finish()

    def run_host_begin(self, xyz_host, frame0=0, out_x=None, out_y=None, block_bytes=256 << 20):
        from . import _lib

        with _lib.lanes(self.lane_base):
            return self._run_host_begin(xyz_host, frame0, out_x, out_y, block_bytes)

    def _run_host_begin(self, xyz_host, frame0, out_x, out_y, block_bytes):
        """Start an end-to-end run and return a HostRun; HostRun.finish() waits for the hopping
        trajectory to be in host memory and returns (PipelineResult, host dict).

        out_x / out_y: pinned host tensors [F_local][N], float32 (the DCD records) or int16 (the same labels, half the
        download; HostRun.planes_float32 widens them).

        Coordinates go to HBM in blocks on an upload stream while the histogram kernel consumes the
        blocks already there; they stay resident for stage 3.  The hopping trajectory planes stream
        back into pinned host tensors (out_x / out_y, [F_local][N] float32) on a download stream; the
        sorted events, COO transition counts and final per-atom state follow at finish().  From the
        second run on nothing in here waits for the device (PipelineConfig.async_results), uploads
        and downloads use different streams and copy engines, and the resident coordinates are double
        buffered: a caller that begins run k+1 before finishing run k -- with a second pair of out_x /
        out_y -- overlaps the upload of k+1 with the stage 2-4 tail and the download of k (PCIe is full
        duplex)."""
        F, N = int(xyz_host.shape[0]), int(xyz_host.shape[1])
        dev = self.device
        par = self._run_index & 1
        xyz = self._buf("xyz_resident%d" % par, (F, N, 3), torch.float32)
        counts = self._buf("counts", self.grid.shape, torch.int32, symm=True)
        counts.zero_()
        self.cells = self._buf("cells%d" % par, (F, N), torch.int32) if self._use_cells(xyz) else None
        if getattr(self, "_h2d_stream", None) is None:
            self._h2d_stream = torch.cuda.Stream(device=dev)
            self._d2h_stream = torch.cuda.Stream(device=dev)
            self._small_stream = torch.cuda.Stream(device=dev)
            self._d2h_done = None
            self._xyz_free = [None, None]
        h2d, d2h = self._h2d_stream, self._d2h_stream
        main = torch.cuda.current_stream(dev)
        if self._xyz_free[par] is not None:
            h2d.wait_event(self._xyz_free[par])       # stage 3 of the run before last read this buffer
        elif self._run_index < 2:
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
        res.density, res.own, res.member, res.map16, _ = self.site_map(counts)
        if self._d2h_done is not None:
            main.wait_event(self._d2h_done)      # the previous run's planes are still being downloaded
        res.hop_x, res.hop_y, events, res.state, res.final_state = self.hoptraj(xyz, res.map16, frame0)
        self.annotate()
        free = torch.cuda.Event()
        free.record(main)
        self._xyz_free[par] = free
        if out_x is None:
            out_x = torch.empty((F, N), dtype=torch.float32).pin_memory()
            out_y = torch.empty((F, N), dtype=torch.float32).pin_memory()
        # int16 host planes (lossless: labels are int16 map values or -1) halve the download; the float32 of the DCD
        # record is then restored on the host by whoever writes a file (HostRun.planes_float32)
        if out_x.dtype == res.hop_x.dtype:
            src_x, src_y = res.hop_x, res.hop_y        # int16 planes straight from stage 3 (the default), or float32 both ways
        elif out_x.dtype == torch.int16:
            src_x = K.labels_to_int16(res.hop_x, self._buf("hop_x16_%d" % par, (F, N), torch.int16))
            src_y = K.labels_to_int16(res.hop_y, self._buf("hop_y16_%d" % par, (F, N), torch.int16))
        else:
            src_x = self._buf("hop_x32_%d" % par, (F, N), torch.float32).copy_(res.hop_x)
            src_y = self._buf("hop_y32_%d" % par, (F, N), torch.float32).copy_(res.hop_y)
        d2h.wait_stream(main)
        with torch.cuda.stream(d2h):
            out_x.copy_(src_x, non_blocking=True)
            out_y.copy_(src_y, non_blocking=True)
            # fixed-size small result; the buffer is rewritten by the next run's stage 3
            state_host = self._host_buf("final_state%d" % par, tuple(res.final_state.shape), torch.int32)
            state_host.copy_(res.final_state, non_blocking=True)
            done = torch.cuda.Event()
            done.record(d2h)
        self._d2h_done = done
        self._finalize(res, events, xyz, frame0)
        return HostRun(self, res, dict(hop_x=out_x, hop_y=out_y, final_state=state_host), done)

    def _host_buf(self, name, shape, dtype):
        bufs = self.__dict__.setdefault("_host_bufs", {})
        t = bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = bufs[name] = torch.empty(shape, dtype=dtype).pin_memory()
        return t


class HostRun(object):
    """A run_host_begin in flight: finish() returns (PipelineResult, host dict) once the hopping
    trajectory planes and the sorted events / transition counts are in host memory."""

    def __init__(self, pipe, result, host, done):
        self.pipe, self.result, self.host, self._done = pipe, result, host, done

    def planes_float32(self):
        """hop_x / hop_y of the host result as float32 [F][N] (the X and Y records of the hoptraj DCD), whatever
        dtype they were downloaded in."""
        self._done.synchronize()
        return self.host["hop_x"].to(torch.float32), self.host["hop_y"].to(torch.float32)

    def finish(self):
        res, host = self.result, self.host
        if "events" not in host:
            res.wait()                                   # finalisation of this run (not of later ones)
            with torch.cuda.stream(self.pipe._small_stream):   # a stream of its own: nothing queued ahead
                host.update(events=res.events.cpu(), edge_s0=res.edge_s0.cpu(), edge_s=res.edge_s.cpu(),
                            edge_count=res.edge_count.cpu())
        self._done.synchronize()
        return res, host


class PipelinePool(object):
    """Several HopPipelines on streams of their own, used round robin -- for independent trajectories on
    the same grid size (BASELINE config 5: 8 trajectories per GPU).  The stages of one run depend on each
    other and stage 2 is a chain of small kernels; with a second run in flight its histogram fills the
    SMs that stage leaves idle (3.43 -> 3.22 ms per run at the bench size on B200).  Every pipeline has
    its own buffers and its own library handles (scratch memory)."""

    def __init__(self, edges, n_frames_total, config: PipelineConfig | None = None, device=None, n=2):
        self.pipes = [HopPipeline(edges, n_frames_total, config, device, lane_base=4 * i) for i in range(n)]
        dev = self.pipes[0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
        self._next = 0

    def run(self, xyz, frame0=0, marks=None, edges=None):
        """Queue one run on the next pipeline and return its PipelineResult without waiting; the
        result's lazy fields, or synchronize(), wait for it.  `edges`: the grid of this trajectory when it differs
        from the previous one's (same shape or not)."""
        i = self._next
        self._next = (i + 1) % len(self.pipes)
        stream = self.streams[i]
        stream.wait_stream(torch.cuda.current_stream(self.pipes[i].device))   # xyz was produced there
        with torch.cuda.stream(stream):
            if edges is not None:
                self.pipes[i].set_edges(edges)
            return self.pipes[i].run(xyz, frame0, marks)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()
        for p in self.pipes:
            s4 = getattr(p, "_s4_stream", None)
            if s4 is not None:
                s4.synchronize()


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
