"""Host-side logic of the N>1 path on CPU: world_size-2 gloo.

What can be checked without GPUs: shard geometry (even boundaries, OWL range intersection), the
unique-id exchange, max-over-ranks timing, and the property the sharded solver rests on — every
rank replays the scalar line-search state machine on bit-identical all-reduced scalars and
therefore takes identical decisions."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rust_lbfgs_b200 as R
from rust_lbfgs_b200 import dist as D


def test_shard_range_properties():
    for n in (2, 10, 100, 101, 1 << 20, 10**8, (1 << 31)):
        for w in (1, 2, 3, 4, 8):
            if n // 2 < w:
                continue
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            assert all(lo % 2 == 0 for lo, _ in spans)            # Rosenbrock pairs stay on one rank
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 3
    assert D.shard_range(1 << 31, 7, 8) == (7 << 28, 1 << 31)     # configs[4]: 2^28 per GPU


def test_owl_range_local():
    assert D.owl_range_local(1, 21, 21, 0, 21) == (1, 21)
    assert D.owl_range_local(0, None, 100, 50, 100) == (0, 50)
    assert D.owl_range_local(10, 60, 100, 50, 100) == (0, 10)
    assert D.owl_range_local(10, 40, 100, 50, 100) is None
    assert D.owl_range_local(0, 1000, 100, 0, 50) == (0, 50)    # end clamped to n (orthantwise.rs:63)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        # 1. unique-id exchange
        uid = D.broadcast_unique_id(lambda: bytes(range(128)))
        out["uid_ok"] = uid == bytes(range(128))
        # 2. timing reduction
        out["max"] = D.max_over_ranks(10.0 + rank)
        out["sum"] = D.sum_over_ranks(1.0 + rank)

        # 3. sharded line search: each rank owns half of x; f and dg partials are all-reduced (sum), then
        #    every rank feeds its own copy of the state machine.
        n = 64
        lo, hi = D.shard_range(n, rank, world)
        x0 = np.empty(n); x0[0::2] = -1.2; x0[1::2] = 1.0

        def rosen(x):
            t1 = 1.0 - x[0::2]; t2 = 10.0 * (x[1::2] - x[0::2] ** 2)
            g = np.empty_like(x); g[1::2] = 20.0 * t2; g[0::2] = -2.0 * (x[0::2] * g[1::2] + t1)
            return float(np.sum(t1 * t1 + t2 * t2)), g

        def allsum(*vals):
            t = torch.tensor(vals, dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()
        xs = x0[lo:hi].copy()
        f, g = rosen(xs)
        d = -g
        finit, dginit, dd = allsum(f, float(g @ d), float(d @ d))
        L = R.lib()
        p = R.default_param()
        h = L.lbfgsb200_linesearch_begin(C.byref(p), 0, finit, dginit, 1.0 / np.sqrt(dd))
        steps = []
        stp = C.c_double()
        while L.lbfgsb200_linesearch_next(h, C.byref(stp)):
            steps.append(stp.value)
            ft, gt = rosen(xs + stp.value * d)
            fs, dgs = allsum(ft, float(gt @ d))
            L.lbfgsb200_linesearch_feed(h, 1, fs, dgs)
        ncall, step = C.c_int64(), C.c_double()
        err = L.lbfgsb200_linesearch_result(h, C.byref(ncall), C.byref(step))
        L.lbfgsb200_linesearch_end(h)
        out["ls"] = (steps, ncall.value, step.value, err)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r]["uid_ok"]
        assert res[r]["max"] == 11.0 and res[r]["sum"] == 3.0
    assert res[0]["ls"] == res[1]["ls"]                 # replicated scalar control: identical decisions
    steps, ncall, step, err = res[0]["ls"]
    assert err == 0 and ncall == len(steps) >= 1
