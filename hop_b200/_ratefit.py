"""Rate constants from waiting times: the arithmetic of hop.graph.HoppingGraph._rate
(hop/graph.py:1677-1748) with fit_func / fitExp / fitExp2 / survivalfunction (hop/graph.py:2354-2518),
kept free of torch / CUDA imports so that it can run in worker processes.

One fit is ~0.5 ms of scipy MINPACK with Python call-backs (up to 800 function evaluations when the
double exponential does not converge -- hop tells the user to ignore that warning); a hopping graph
has 10^4..10^5 edges, so the fits of a graph are spread over a process pool.
"""
from __future__ import annotations

import os
import warnings

import numpy


def survival_values(waitingtimes, t):
    """S(t) = <Theta(t0 - t)> over the waiting times t0, Theta(0) = 1/2 (hop/graph.py:2433-2518), from
    one sort and two binary searches instead of the reference's sum of N step functions: the same
    numbers (sums of 0, 1/2 and 1 are exact)."""
    w = numpy.sort(numpy.asarray(waitingtimes, dtype=float))
    n = len(w)
    _t = numpy.asarray(t, dtype=float)
    right = numpy.searchsorted(w, _t, side='right')   # #{t0 <= t}
    left = numpy.searchsorted(w, _t, side='left')     # #{t0 <  t}
    return ((n - right) + 0.5 * (right - left)) / n


def f_exp(p, x):
    return numpy.exp(-p[0] * x)


def f_exp2(p, x):
    return p[0] * numpy.exp(-p[1] * x) + (1 - p[0]) * numpy.exp(-p[2] * x)


INITIAL = {'exp': [1.0], 'exp2': [0.5, 0.1, 1e-4]}
FUNCTIONS = {'exp': f_exp, 'exp2': f_exp2}


def leastsq_fit(kind, x, y):
    """fit_func.__init__ (hop/graph.py:2354-2397): scipy.optimize.leastsq on f(p, x) - y from the class's
    initial values.  Returns (parameters as a list-like, message)."""
    import scipy.optimize

    f = FUNCTIONS[kind]
    p0 = list(INITIAL[kind])
    p, msg = scipy.optimize.leastsq(lambda p, x, y: f(p, x) - y, p0, args=(numpy.asarray(x), numpy.asarray(y)))
    try:
        p[0]
    except (TypeError, IndexError):
        p = [p]
    return p, msg


def fit_one(taus):
    """HoppingGraph._rate for one edge: (kind, parameters, message, k, N)."""
    taus = numpy.asarray(taus)
    dt = 1.0
    tmax = numpy.max(taus) + 1.0 * dt
    N = len(taus)
    x = numpy.arange(0, tmax + 1, dt)
    Sx = survival_values(taus, x)
    enough_data = N > 10
    is_single_exp = False
    kind = p = msg = k = None
    if enough_data:
        kind = 'exp2'
        p, msg = leastsq_fit(kind, x, Sx)
        a, k1, k2 = p
        is_single_exp = abs(0.5 - a) > 0.49
        all_k = numpy.array([k1, k2])
        if numpy.any(all_k < 0):
            is_single_exp = True
        else:
            k = all_k[all_k > 0].max()
    if not enough_data or is_single_exp:
        kind = 'exp'
        p, msg = leastsq_fit(kind, x, Sx)
        k, = p
    return kind, numpy.asarray(p, dtype=float), msg, float(k), N


def _fit_chunk(chunk):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # "Number of calls to function has reached maxfev = 800"
        return [fit_one(t) for t in chunk]


def default_workers(n_fits):
    if n_fits < 512:
        return 0
    try:
        n_cpu = len(os.sched_getaffinity(0))
    except AttributeError:
        n_cpu = os.cpu_count() or 1
    return min(n_cpu, 32) if n_cpu >= 4 else 0


def fit_all(taus_list, workers=None):
    """fit_one for every entry of taus_list, in order.  workers: None = automatic (worker processes when
    there are at least 512 fits and 4 cores), 0 / 1 = in this process.

    The workers are fresh interpreters (`python -m hop_b200._ratefit`) fed through pipes -- not forks of
    this process, which holds a CUDA context, and not multiprocessing's spawn, which would re-import the
    caller's main script."""
    n = len(taus_list)
    if workers is None:
        workers = default_workers(n)
    workers = min(int(workers), n)
    if workers <= 1:
        return [fit_one(t) for t in taus_list]
    import pickle
    import subprocess
    import sys
    from concurrent.futures import ThreadPoolExecutor

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    env.setdefault("OMP_NUM_THREADS", "1")
    shares = [list(range(w, n, workers)) for w in range(workers)]   # interleaved: long and short fits mix

    def run(share):
        blob = pickle.dumps([numpy.asarray(taus_list[i]) for i in share], pickle.HIGHEST_PROTOCOL)
        proc = subprocess.Popen([sys.executable, "-m", "hop_b200._ratefit"], stdin=subprocess.PIPE,
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env)
        out, _ = proc.communicate(blob)
        if proc.returncode != 0:
            raise RuntimeError("rate-fit worker failed")
        return pickle.loads(out)

    try:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            parts = list(pool.map(run, shares))
    except (OSError, RuntimeError, pickle.UnpicklingError, EOFError):   # same numbers, serially
        return [fit_one(t) for t in taus_list]
    out = [None] * n
    for share, part in zip(shares, parts):
        for i, r in zip(share, part):
            out[i] = r
    return out


def _worker_main():
    import pickle
    import sys

    data = pickle.load(sys.stdin.buffer)
    result = _fit_chunk(data)
    sys.stdout.buffer.write(pickle.dumps(result, pickle.HIGHEST_PROTOCOL))
    sys.stdout.buffer.flush()


if __name__ == "__main__":
    _worker_main()
