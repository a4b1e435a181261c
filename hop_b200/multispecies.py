"""Several selections (species) of one trajectory through the four stages with ONE read of the coordinates.

hop treats every selection on its own -- one DensityCreator, one HoppingTrajectory, one TransportNetwork per
species, each re-reading the trajectory (scripts/hop-generate-many-densities.py:98-125 loops over trajectories
and selections; hop/density.py:271-304) -- and every species gets a grid of its own, spanned by THAT
selection in frame 0 (hop/density.py:271-278).  BASELINE.json configs[3] is the case: Na+ / K+ / Cl- in one
box over 10^6 frames.  Here

  stage 1   one launch bins all selected atoms, the species of an atom choosing grid, count array and brick
            numbering (hopb_hist_accumulate_multi); it also leaves the 4-byte brick index of every atom-frame;
  stage 2   one site map per species (hopb_site_map_fused), on streams of their own;
  stage 3+4 one launch per species on its column range of the shared brick-index array
            (hopb_hoptraj_strided): the coordinates are not touched again; events, summaries and transition
            counts stay per species (atom numbers are those inside the selection, like a per-species run).

Every species' result is bit-identical to a HopPipeline run on that selection's coordinates alone
(tests/test_gpu_kernels.py::test_multispecies_equals_independent_runs).
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from .pipeline import HopPipeline, PipelineConfig, PipelineResult, grid_edges_from_frame0


class MultiSpeciesPipeline(object):
    """columns: list of (start, stop) atom ranges of the coordinate array, one per species (contiguous and in order:
    concatenate the selections species by species); edges: the three bin-edge arrays of every species' grid."""

    def __init__(self, edges_list, columns, n_frames_total, config: PipelineConfig | None = None, device=None):
        if len(edges_list) != len(columns) or not columns:
            raise ValueError("one grid per species")
        self.columns = [(int(a), int(b)) for a, b in columns]
        for (a, b), (c, d) in zip(self.columns[:-1], self.columns[1:]):
            if b != c or a >= b:
                raise ValueError("species must own consecutive, non-empty column ranges")
        self.n_atoms = self.columns[-1][1] - self.columns[0][0]
        if self.columns[0][0] != 0:
            raise ValueError("the first species starts at column 0")
        self.cfg = config or PipelineConfig()
        # one pipeline object per species for stages 2-4 (buffers, scratch handles and streams of its own)
        self.pipes = [HopPipeline(e, n_frames_total, PipelineConfig(**{**self.cfg.__dict__, "async_results": False}),
                                  device, lane_base=4 * k) for k, e in enumerate(edges_list)]
        self.device = self.pipes[0].device
        sp = np.zeros(self.n_atoms, dtype=np.uint8)
        for k, (a, b) in enumerate(self.columns):
            sp[a:b] = k
        self.species = torch.from_numpy(sp).to(self.device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.pipes]
        self._bufs = {}

    @classmethod
    def from_frame0(cls, coords0, columns, n_frames_total, delta=1.0, padding=2.0, config=None, device=None):
        """Grids as DensityCreator derives them, per selection (hop/density.py:271-278)."""
        coords0 = np.asarray(coords0)
        return cls([grid_edges_from_frame0(coords0[a:b], delta=delta, padding=padding) for a, b in columns], columns,
                   n_frames_total, config, device)

    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = self._bufs[name] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def run(self, xyz, frame0=0):
        """xyz: CUDA float32 [F][N][3], all species.  Returns one PipelineResult per species; hop_x / hop_y of a
        species are column ranges of planes shared by all species ([F][N], PipelineConfig.plane_dtype, `self.hop_x`, `self.hop_y`)."""
        from . import _lib

        F, N = int(xyz.shape[0]), int(xyz.shape[1])
        if N != self.n_atoms:
            raise ValueError("xyz holds %d atoms, the species cover %d" % (N, self.n_atoms))
        main = torch.cuda.current_stream(self.device)
        counts = []
        for p in self.pipes:
            c = p._buf("counts", p.grid.shape, torch.int32)
            c.zero_()
            counts.append(c)
        cells = self._buf("cells", (F, N), torch.int32)
        K.hist_accumulate_multi([p.grid for p in self.pipes], self.species, xyz, counts, cells)
        plane_dtype = self.pipes[0].cfg.plane_torch_dtype()
        self.hop_x = self._buf("hop_x", (F, N), plane_dtype)
        self.hop_y = self._buf("hop_y", (F, N), plane_dtype)
        results, pending = [], []
        for k, (p, (a, b)) in enumerate(zip(self.pipes, self.columns)):      # queue stages 2 and 3 of every species ...
            st = self.streams[k]
            st.wait_stream(main)
            with torch.cuda.stream(st), _lib.lanes(p.lane_base):
                res = PipelineResult()
                res.counts = counts[k]
                res.density, res.own, res.member, res.map16, _ = p.site_map(counts[k])
                p.annotate()
                # stage 3 of this species: its columns of the shared arrays
                p.cells = cells[:, a:b]
                p._bufs["hop_x"] = self.hop_x[:, a:b]
                p._bufs["hop_y"] = self.hop_y[:, a:b]
                view = xyz[:, a:b]
                res.hop_x, res.hop_y, events, res.state, res.final_state = p.hoptraj(view, res.map16, frame0)
            results.append(res)
            pending.append((events, view))
        for k, (p, res) in enumerate(zip(self.pipes, results)):               # ... then order their events
            with torch.cuda.stream(self.streams[k]), _lib.lanes(p.lane_base):
                p._finalize_sync(res, pending[k][0], pending[k][1], frame0)
                p._run_index += 1
        for st in self.streams:
            main.wait_stream(st)
        return results
