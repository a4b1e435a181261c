"""The multi-rank path on CPU: two (and three) processes over gloo run hop_b200.exchange with the
CPU mirrors of the stitching kernels (tests/emul) and must reproduce the single-process oracle --
counts all-reduce, slab-summary all-gather + replay, transition-count matrix all-reduce."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EV = np.dtype([("frame", "i4"), ("iatom", "i4"), ("s0", "i4"), ("s", "i4"), ("tau", "i4"), ("tbar", "i4")])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lab, counts_parts, n_chunks, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from hop_b200 import exchange
    from tests.emul.build import build

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emul = ctypes.CDLL(build())
        emul.emul_slab_scan.restype = ctypes.c_int64
        emul.emul_fixup.restype = ctypes.c_int64
        F, N = lab.shape
        f0, f1 = exchange.slab_bounds(F, world, rank)
        slab_lab = np.ascontiguousarray(lab[f0:f1])
        Fl = f1 - f0
        # stage 1 message: the count grid
        counts = torch.from_numpy(counts_parts[rank].copy())
        exchange.reduce_counts(counts)
        # stage 3/4: chunk summaries of this slab, slab summary, all-gather, replay, fix-up
        chunk_len = -(-Fl // n_chunks)
        nc = -(-Fl // chunk_len)
        summ = np.zeros((nc, 11, N), np.int32)
        cap = Fl * N + 8
        ev = np.zeros(cap, EV)
        vp = ctypes.c_void_p
        n1 = emul.emul_slab_scan(slab_lab.ctypes.data_as(vp), ctypes.c_int64(Fl), ctypes.c_int64(N), ctypes.c_int64(f0),
                                 nc, ctypes.c_int64(chunk_len), summ.ctypes.data_as(vp), ev.ctypes.data_as(vp), ctypes.c_int64(cap))
        slab = np.zeros((11, N), np.int32)
        emul.emul_combine(summ.ctypes.data_as(vp), ctypes.c_int64(N), nc, ctypes.c_int64(chunk_len), ctypes.c_int64(Fl),
                          slab.ctypes.data_as(vp))
        gathered = exchange.gather_slab_summaries(torch.from_numpy(slab)).numpy()
        assert gathered.shape == (world, 11, N)
        state = np.zeros((4, N), np.int32)
        prev = np.ascontiguousarray(gathered[:rank])
        emul.emul_advance(prev.ctypes.data_as(vp), rank, ctypes.c_int64(N), 1, state.ctypes.data_as(vp))
        ev2 = np.zeros(cap, EV)
        n2 = emul.emul_fixup(summ.ctypes.data_as(vp), nc, ctypes.c_int64(N), state.ctypes.data_as(vp),
                             ev2.ctypes.data_as(vp), ctypes.c_int64(cap))
        final = np.zeros((4, N), np.int32)
        allg = np.ascontiguousarray(gathered)
        emul.emul_advance(allg.ctypes.data_as(vp), world, ctypes.c_int64(N), 1, final.ctypes.data_as(vp))
        events = np.concatenate([ev[:n1], ev2[:n2]])
        # stage 4 message: dense transition-count matrix
        W = int(lab.max()) + 2
        M = torch.zeros((W, W), dtype=torch.int32)
        for e in events:
            M[e["s0"] + 1, e["s"] + 1] += 1
        exchange.reduce_count_matrix(M)
        out_q.put((rank, counts.numpy(), events, final, M.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_chunks", [(2, 3), (3, 1)])
def test_frame_sharded_exchange_over_gloo(world, n_chunks):
    import torch.multiprocessing as mp

    from oracle import hop_oracle as O
    from tests.emul.build import build
    from tests.test_machine_emul import random_labels

    build()
    rng = np.random.default_rng(5)
    F, N = 101, 37
    lab = random_labels(rng, F, N, nsites=4, p_stay=0.85, p_out=0.08, p_zero=0.3)
    ev_ref, nodes_ref = O.analyze_hops_vectorised(lab)
    total_counts = rng.integers(0, 50, size=(6, 6, 6)).astype(np.int32)
    parts = [rng.integers(0, 20, size=(6, 6, 6)).astype(np.int32) for _ in range(world - 1)]
    parts.append(total_counts - sum(parts))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lab, parts, n_chunks, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=120)
        results[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    events = np.concatenate([results[r][1] for r in range(world)])
    order = np.lexsort((events["iatom"], events["frame"]))
    got = np.stack([events[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbar")], axis=1).astype(np.int64)
    assert np.array_equal(got, ev_ref)                      # every transition exactly once, same tau/tbarrier
    pairs, cnts = O.events_to_counts(ev_ref)
    for r in range(world):
        counts, _, final, M = results[r]
        assert np.array_equal(counts, total_counts)           # identical reduced grid on every rank
        assert M.sum() == len(ev_ref)
        for (a, b), c in zip(pairs, cnts):
            assert M[a + 1, b + 1] == c
        ia = np.nonzero(final[0] != 0)[0]
        nodes = np.stack([ia, final[0][ia], (F - 1) - final[1][ia]], axis=1)
        assert np.array_equal(nodes, nodes_ref)               # node records from the replayed final state


def test_slab_bounds_cover_all_frames():
    from hop_b200.exchange import slab_bounds

    for F in (1, 7, 100, 10000, 100003):
        for world in (1, 2, 3, 8):
            edges = [slab_bounds(F, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1
