"""Stage 4: transitions between sites -- drop-in for hop.graph.TransportNetwork.

``TransportNetwork._analyze_hops`` (hop/graph.py:156-288) runs a per-atom state machine over the x
column of the hopping trajectory with two nested Python loops; here the same machine runs on the
GPU (csrc/hoptraj.cu, csrc/hopb_machine.h), the events are sorted on the device into the
reference's emission order (csrc/events.cu) and come back as the ``graph_properties`` dictionary:
edge tuple ``(i, j)`` or node ``i`` -> ``{'tau', 'tbarrier', 'frame', 'iatom'}`` arrays, times in ps.
The transition-count matrix is ``len(tau)`` per edge.

``HoppingGraph`` takes the network's results (graph, properties, trjdata, site_properties) and puts
a rate constant on every edge the way hop.graph.HoppingGraph does (hop/graph.py:1677-1748): a
least-squares fit of the survival function of the waiting times, on the host with scipy like the
reference, fed by the event arrays the GPU produced.
"""
from __future__ import annotations

import logging
import os
import pickle

import numpy

from . import _ratefit, sitemap, trajectory
from .constants import SITELABEL

logger = logging.getLogger("MDAnalysis.app.hop.graph")


def _torch():
    import torch

    return torch


class TransportNetwork(object):
    """A framework for computing graphs from hopping trajectories; the unit of time is ps
    (hop/graph.py:72-105)."""

    def __init__(self, traj, density=None, sitelabels=None):
        import networkx as NX

        self._cache = dict()
        if not isinstance(traj, trajectory.HoppingTrajectory):
            errmsg = "'traj' must be a <hop.trajectory.HoppingTrajectory> instance."
            logger.fatal(errmsg)
            raise TypeError(errmsg)
        self.traj = traj
        if traj.hoptraj is not None:
            self.n_atoms = traj.hoptraj.n_atoms
            self.dt = traj.hoptraj.dt
        else:
            self.n_atoms = len(traj.tgroup)
            self.dt = traj.traj.dt
        self.totaltime = self.traj.totaltime
        self.graph = NX.DiGraph(name='Transport network between sites')
        if density is not None and not isinstance(density, sitemap.Density):
            raise TypeError('density must be a <hop.sitemap.Density> instance.')
        self.density = density
        if sitelabels:
            self._cache['sitelabels'] = sitelabels

    # ---- labels on the device ---------------------------------------------------------------------------
    def _hop_x_device(self):
        """x column of the hoptraj as CUDA float32 [F][N]."""
        t = _torch()
        if self.traj.hoptraj is not None:
            x = self.traj.hoptraj.plane_array(0)
        else:
            x, _ = self.traj.hop_arrays()
        return t.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32)).cuda()

    def _n_labels(self, hx):
        try:
            n = self.density._n_labels
            if n is not None:
                return int(n)
        except AttributeError:
            pass
        if hx is None:        # the caller fetches the label plane and asks again
            return None
        return int(hx.max().item()) + 1

    def graph_alltransitions(self):
        """Graph with one edge for each (label[t-1] -> label[t]) pair observed, interstitial and
        outliers included (hop/graph.py:107-135)."""
        import networkx as NX

        from . import kernels as K

        self.graph = NX.DiGraph(name="Transitions between all sites, including interstitial and outliers")
        hx = self._hop_x_device()
        pairs = K.all_transitions(hx, self._n_labels(hx))
        self.graph.add_edges_from((int(a), int(b)) for a, b in pairs)
        self._cache['sitelabels'] = sorted(self.graph.nodes())

    def HoppingGraph(self, verbosity=3):
        """Compute the HoppingGraph from the data and return it."""
        self.compute_residency_times(verbosity)
        return self.hopgraph

    def compute_residency_times(self, verbosity=3):
        self._analyze_hops(verbosity=verbosity)
        try:
            site_properties = self.density.site_properties
        except AttributeError:
            site_properties = None
        self.hopgraph = HoppingGraph(self.graph, self.graph_properties,
                                     trjdata={'dt': self.dt, 'totaltime': self.totaltime, 'time_unit': 'ps'},
                                     site_properties=site_properties)

    def _analyze_hops(self, verbosity=0):
        """Waiting times tau_ji for hops i->j (hop/graph.py:156-288).

        self.graph: directed graph of all transitions; self.graph_properties: per edge (i,j) the
        arrays tau, tbarrier (ps), frame and iatom in emission order, plus per node the residual
        residency of atoms still assigned to it at the last frame (tbarrier None)."""
        import networkx as NX

        from . import kernels as K

        try:
            if len(self._cache['sitelabels']) < 3:
                errmsg = 'TransportNetwork: Need at least 2 sites to analyze transitions: N=%d' % \
                         len(self._cache['sitelabels'])
                logger.error(errmsg)
                raise ValueError(errmsg)
        except KeyError:
            pass
        self.graph = NX.DiGraph()
        self.graph.name = "Transitions between sites"
        self.graph_properties = dict()

        cached = getattr(self.traj, "_events_cache", None)
        if cached is not None:
            events, n_ev, state = cached       # produced by the fused hoptraj pass
            F = self.traj.n_frames
            n_labels = self._n_labels(None) if self.density is not None else None
            if n_labels is None:
                hx = self._hop_x_device()
                n_labels = self._n_labels(hx)
        else:
            hx = self._hop_x_device()
            F, N = int(hx.shape[0]), int(hx.shape[1])
            n_labels = self._n_labels(hx)
            n_chunks, chunk_len = K.plan_chunks(F, N)
            events = K.EventBuffer(max(1 << 16, (F * N) // 8), hx.device)
            for attempt in range(2):
                s = K.hopscan_labels(hx, 0, n_chunks, chunk_len, events)
                state = K.new_state(N, hx.device)
                K.hop_advance(None, state, init=True)
                K.hop_fixup(s, chunk_len, F, state, None, events)
                n_ev = events.n()
                if n_ev <= events.capacity:
                    break
                events = K.EventBuffer(n_ev, hx.device)
        ev, e_s0, e_s, e_start, e_count = K.events_finalize(events, n_ev, self.n_atoms, F, n_labels)
        ev = K.events_to_numpy(ev)
        e_s0 = e_s0.cpu().numpy()
        e_s = e_s.cpu().numpy()
        e_start = e_start.cpu().numpy()
        e_count = e_count.cpu().numpy()
        # edges enter the graph in order of first occurrence, like the reference's add_edge calls
        first = numpy.array([(ev['frame'][s], ev['iatom'][s]) for s in e_start], dtype=numpy.int64).reshape(-1, 2)
        order = numpy.lexsort((first[:, 1], first[:, 0])) if len(first) else []
        for k in order:
            a, b, s, c = int(e_s0[k]), int(e_s[k]), int(e_start[k]), int(e_count[k])
            seg = ev[s:s + c]
            self.graph.add_edge(a, b)
            self.graph_properties[(a, b)] = {
                'tau': self.dt * seg['tau'].astype(numpy.int64),
                'tbarrier': self.dt * seg['tbarrier'].astype(numpy.int64),
                'frame': seg['frame'].astype(numpy.int64),
                'iatom': seg['iatom'].astype(numpy.int64)}
        # mop up: atoms still assigned to a site at the last frame (hop/graph.py:259-274)
        st = state.cpu().numpy()
        last = F - 1
        ia = numpy.nonzero(st[0] != SITELABEL['interstitial'])[0]
        node_vals, first_idx = numpy.unique(st[0][ia], return_index=True) if len(ia) else ([], [])
        for node in [node_vals[k] for k in numpy.argsort(first_idx)]:   # reference order: first atom that holds it
            sel = ia[st[0][ia] == node]
            self.graph.add_node(int(node))
            self.graph_properties[int(node)] = {
                'tau': self.dt * (last - st[1][sel]).astype(numpy.int64),
                'tbarrier': numpy.array([None] * len(sel)),
                'frame': numpy.full(len(sel), last, dtype=numpy.int64),
                'iatom': sel.astype(numpy.int64)}
        if 'sitelabels' not in self._cache:
            self._cache['sitelabels'] = sorted(self.graph.nodes())

    def transition_counts(self):
        """{(i, j): N_ij} -- the transition-count matrix in sparse form (len(tau) per edge)."""
        return {k: len(v['tau']) for k, v in self.graph_properties.items() if isinstance(k, tuple)}


class HoppingGraph(object):
    """Directed graph of sites with rate constants on the edges (hop/graph.py:492-640, 1677-1748).

    ``graph`` is a networkx.DiGraph whose edges carry ``k`` (rate constant in 1/ps), ``N`` (number of
    transitions = the transition count) and ``fit`` (a :class:`fit_func`); ``properties`` holds the
    per-edge / per-node event arrays of TransportNetwork; ``trjdata`` and ``site_properties`` travel
    along.  As in the reference the rate of an edge comes from a least-squares fit (scipy's
    ``leastsq``, i.e. MINPACK, on the host) of a double or single exponential to the survival
    function of the waiting times; the survival function is evaluated by sorting instead of the
    reference's O(N * T) sum of step functions -- the values are identical."""

    #: how the rate constants of a new graph are fitted: "gpu" -- every edge at once on the device (csrc/ratefit.cu:
    #: the survival functions by histogram + suffix sum, MINPACK's lmdif restated with one warp per edge; it takes
    #: the optimiser's path of the reference -- same nfev / info -- and agrees with it to ~1e-7, not always to 1e-6:
    #: lmdif differentiates by forward differences, so rounding in exp() reaches the result); "minpack" -- scipy's
    #: leastsq on the host, the very library call of the reference (hop/graph.py:2369), bit-compatible with it and
    #: what `_rate` for a single edge always uses; "auto" -- "gpu", then the fits that a second, exp()-perturbed run
    #: moves by more than 1e-8 (not determined by the data to the parity bar) once more with "minpack".
    rate_method = os.environ.get("HOPB_RATE_METHOD", "auto")

    def __init__(self, graph=None, properties=None, filename=None, trjdata=None, site_properties=None, workers=None,
                 rate_method=None):
        import networkx as NX

        self.trjdata = trjdata if isinstance(trjdata, dict) else None
        self.site_properties = site_properties
        self._filename = filename
        if not (graph is None or properties is None):
            self.graph = NX.DiGraph(name='Transitions between sites')
            self.properties = properties
            self.graph.add_nodes_from(graph.nodes)
            edges = list(graph.edges)
            fits = _fit_edges([numpy.asarray(properties[e]['tau']) for e in edges], rate_method or self.rate_method, workers)
            ebunch = [(e[0], e[1], _rate_record(f)) for e, f in zip(edges, fits)]
            self.graph.add_edges_from(ebunch)
        elif filename is not None:
            self.load(filename)
        else:
            raise ValueError('provide edges and properties or a filename')

    # ---- access to the data of an edge (hop/graph.py:617-700) ------------------------------------
    def rate(self, edge):
        """Fastest rate on an edge, in 1/ns."""
        f = self.waitingtime_fit(edge)
        if len(f.parameters) == 1:
            k = f.parameters[0]
        elif len(f.parameters) == 3:
            a, k1, k2 = f.parameters
            k = k1 if a > 0.5 else k2
        else:
            raise ValueError('Unknown waitingtime fit, ' + str(f))
        return 1000 * k

    def waitingtime_fit(self, edge):
        """The fit_func instance of an edge (from_site, to_site[, data])."""
        return self.graph[edge[0]][edge[1]]['fit']

    def number_of_hops(self, *edge):
        """Number of transitions recorded on the edge = the transition count N_ij."""
        if len(edge) == 1:
            edge = edge[0]
        return self.graph[edge[0]][edge[1]]['N']

    def from_site(self, edge):
        return edge[0]

    def to_site(self, edge):
        return edge[1]

    def waitingtimes(self, i, j):
        return self.properties[(i, j)]['tau']

    def barriertimes(self, i, j):
        return self.properties[(i, j)]['tbarrier']

    def rates(self, n, use_filtered_graph=False):
        """{(from, to): k} of all edges entering or leaving site n, in 1/ps."""
        out = {}
        for a, b, d in self.graph.in_edges(n, data=True):
            out[(a, b)] = d['k']
        for a, b, d in self.graph.out_edges(n, data=True):
            out[(a, b)] = d['k']
        return out

    def _rate(self, taus, method='survivalfunction', block_w=200):
        """Rate i --> j from the hopping times tau_ji: dict(k, N, fit)  (hop/graph.py:1677-1748).

        Fit a*exp(-k1 t) + (1-a)*exp(-k2 t) to the survival function S(t) when there are more than 10
        events, fall back to exp(-k t) when the double exponential is very asymmetric
        (|0.5 - a| > 0.49) or has a negative rate."""
        if method != 'survivalfunction':
            raise NotImplementedError('Method "%s" is not implemented.' % method)
        return _rate_record(_ratefit.fit_one(taus))

    def filename(self, filename=None, ext=None, set_default=False, use_my_ext=False):
        from .utilities import filename_function

        return filename_function(self, filename, ext, set_default, use_my_ext)

    def save(self, filename=None):
        """Save HoppingGraph as a pickled python object, like hop does (hop/graph.py:760-767): the
        instance itself, so that hop's own load() -- which copies `graph`, `properties`, `trjdata`, ... from
        the unpickled object and reads its `site_properties` -- accepts the file as well."""
        with open(self.filename(filename, 'pickle', set_default=True), 'wb') as fh:
            pickle.dump(self, fh, pickle.HIGHEST_PROTOCOL)

    def load(self, filename=None):
        """Reinstantiate from a pickled HoppingGraph (hop/graph.py:769-784): files written by save()
        here, by hop's HoppingGraph.save(), or the attribute dictionary earlier versions of this class
        wrote."""
        from .utilities import load_pickle

        with open(self.filename(filename, 'pickle', set_default=True), 'rb') as fh:
            h = load_pickle(fh)
        data = h if isinstance(h, dict) else h.__dict__
        for attr in ('graph', 'properties', 'filtered_graph', 'trjdata', 'node_properties', 'occupancy_avg',
                     'occupancy_std'):
            if attr in data:
                self.__dict__[attr] = data[attr]
        # hop keeps site_properties behind a property (name-mangled attribute)
        self.site_properties = data.get('site_properties', data.get('_HoppingGraph__site_properties'))


class fit_func(object):
    """Fit a function f to data (x, y) by least squares (hop/graph.py:2354-2397); ``parameters``
    holds the fitted parameters."""

    _kind = None

    def __init__(self, x, y):
        self.parameters, self.message = _ratefit.leastsq_fit(self._kind, x, y)

    @classmethod
    def _from_fit(cls, parameters, message):
        """An instance around parameters that were fitted elsewhere (a worker process)."""
        self = cls.__new__(cls)
        self.parameters, self.message = parameters, message
        return self

    def f_factory(self):
        return _ratefit.FUNCTIONS[self._kind]

    def initial_values(self):
        return list(_ratefit.INITIAL[self._kind])

    def fit(self, x):
        """Applies the fit to all x values"""
        return self.f_factory()(self.parameters, numpy.asarray(x))


class fitExp(fit_func):
    """y = f(x) = exp(-p[0]*x)"""

    _kind = 'exp'

    def __repr__(self):
        return "<fitExp " + str(self.parameters) + ">"


class fitExp2(fit_func):
    """y = f(x) = p[0]*exp(-p[1]*x) + (1-p[0])*exp(-p[2]*x)"""

    _kind = 'exp2'

    def __repr__(self):
        return "<fitExp2" + str(self.parameters) + ">"


class fitlin(fit_func):
    """y = f(x) = p[0]*x + p[1]"""

    def __init__(self, x, y):
        import scipy.optimize

        p, self.message = scipy.optimize.leastsq(lambda p, x, y: p[0] * x + p[1] - y, [1.0, 0.0],
                                                 args=(numpy.asarray(x), numpy.asarray(y)))
        self.parameters = p

    def f_factory(self):
        return lambda p, x: p[0] * x + p[1]

    def initial_values(self):
        return [1.0, 0.0]

    def __repr__(self):
        return "<fitlin" + str(self.parameters) + ">"


def _fit_edges(taus_list, method, workers=None):
    """(kind, parameters, message, k, N) per edge, by the chosen method (HoppingGraph.rate_method)."""
    if method == "minpack" or not taus_list:
        return _ratefit.fit_all(taus_list, workers=workers)
    if method not in ("gpu", "auto"):
        raise ValueError("rate_method must be 'gpu', 'minpack' or 'auto'")
    from . import kernels as K

    out, _, _ = K.rate_fit(taus_list)
    messages = {1: "Both actual and predicted relative reductions in the sum of squares are at most 0.000000",
                2: "The relative error between two consecutive iterates is at most 0.000000",
                3: "Both actual and predicted relative reductions in the sum of squares are at most 0.000000 and the "
                   "relative error between two consecutive iterates is at most 0.000000",
                4: "The cosine of the angle between func(x) and any column of the Jacobian is at most 0.000000 in "
                   "absolute value",
                5: "Number of calls to function has reached maxfev = %d."}

    def record(o):
        kind = 'exp2' if o[0] == 1.0 else 'exp'
        p = numpy.array(o[1:4] if kind == 'exp2' else o[1:2], dtype=float)
        info = int(o[5])
        msg = messages.get(info, "info = %d" % info)
        if info == 5:
            msg = msg % (800 if kind == 'exp2' else 400)
        return kind, p, msg, float(o[4]), int(o[7])

    fits = [record(o) for o in out]
    if method == "auto":
        second, _, _ = K.rate_fit(taus_list, noise=1)
        moved = (out[:, 0] != second[:, 0]) | (numpy.abs(out[:, 4] - second[:, 4]) > 1e-8 * numpy.abs(out[:, 4])) | (out[:, 5] >= 5)
        idx = numpy.nonzero(moved)[0]
        if len(idx):
            redo = _ratefit.fit_all([taus_list[i] for i in idx], workers=workers)
            for i, f in zip(idx, redo):
                fits[i] = f
    return fits


def _rate_record(fitted):
    """(kind, parameters, message, k, N) of _ratefit.fit_one -> the dict an edge carries."""
    kind, p, msg, k, N = fitted
    cls = fitExp2 if kind == 'exp2' else fitExp
    return dict(k=k, N=N, fit=cls._from_fit(p, msg))


def Unitstep(x, x0):
    """Heaviside step function Theta(x - x0): 1 if x > x0, 0.5 if x == x0, 0 if x < x0
    (hop/graph.py:2501-2518)."""
    return 0.5 * (1 + numpy.sign(numpy.asarray(x) - x0))


def survivalfunction(waitingtimes, block_w=200, block_t=1000):
    """S(t): the fraction of particles that have not yet left the site after time t, from a list
    of waiting times (hop/graph.py:2433-2499): S(t) = <Theta(t0 - t)>_{t0}.

    The reference sums one step function per waiting time for every t (O(N * T), blocked to save
    memory); here the waiting times are sorted once and S(t) = (#{t0 > t} + 0.5 #{t0 == t}) / N comes
    from two binary searches -- the same numbers (sums of 0, 0.5 and 1 are exact)."""
    w = numpy.asarray(waitingtimes, dtype=float)

    def S(t):
        return _ratefit.survival_values(w, t)

    S.__doc__ = "Survival function S(t).  t can be a array."
    return S
