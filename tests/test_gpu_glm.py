"""The fused one-pass GLM objective (csrc/objectives.cu: k_glm_fused, rows staged by TMA bulk copies) against
the oracle's objective (tests/owlqn.rs:22-43 for Poisson; the logistic definition in oracle/lbfgs_oracle.cpp)
and against the two-pass kernels, over the shapes that select every template instantiation and the fallbacks:
odd ncol (8-byte aligned rows: scalar shared-memory reads, even rows per stage), nrow < number of CTAs, ragged last
column pair slots, ncol up to 10 240 per CTA and beyond it with the columns split over a thread-block cluster."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import rust_lbfgs_b200 as R
from gpu_util import ck, dev, host, stream

pytestmark = pytest.mark.gpu

P2, FUSED, ODD, CLUSTER = R._lib.GLM_PATH_TWO_PASS, R._lib.GLM_PATH_FUSED, R._lib.GLM_PATH_FUSED_ODD, R._lib.GLM_PATH_FUSED_CLUSTER
# (nrow, ncol, the kernels that must have run)
SHAPES = [(7, 2, FUSED), (500, 22, FUSED), (1000, 512, FUSED), (333, 1000, FUSED), (4097, 2050, FUSED), (64, 4096, FUSED),
          (200, 6144, FUSED), (300, 8190, FUSED), (150, 10000, FUSED), (150, 10240, FUSED),
          # odd ncol (rows 8-byte aligned): the reference's own fixture shape, short last groups with an odd row count
          (500, 21, ODD), (501, 21, ODD), (7, 1, ODD), (333, 1001, ODD), (129, 4097, ODD), (75, 6143, ODD), (40, 6145, P2),
          # ncol > 10 240: columns split over a thread-block cluster of 2 / 4 / 8 CTAs
          (100, 10242, CLUSTER), (150, 20000, CLUSTER), (77, 20480, CLUSTER), (60, 40000, CLUSTER), (30, 81920, CLUSTER),
          (20, 20001, P2)]


def eval_gpu(obj, w):
    import torch
    L = R.lib()
    wd, gd, fd = dev(w), dev(np.zeros_like(w)), dev(np.zeros(1))
    ck(L.lbfgsb200_objective_eval(obj._user_ptr(0), wd.data_ptr(), gd.data_ptr(), w.size, stream(), fd.data_ptr()))
    torch.cuda.synchronize()
    return float(host(fd)[0]), host(gd)


@pytest.mark.parametrize("nrow,ncol,path", SHAPES)
@pytest.mark.parametrize("kind", ["poisson", "logistic"])
def test_glm_objective_matches_oracle(oracle, kind, nrow, ncol, path):
    rng = np.random.default_rng(nrow * 131 + ncol)
    X = rng.standard_normal((nrow, ncol)) / np.sqrt(ncol)
    X[:, 0] = 1.0
    w = rng.standard_normal(ncol) * 0.5
    z = X @ w
    y = rng.poisson(np.exp(np.clip(z, -3, 3))).astype(np.float64) if kind == "poisson" else (rng.random(nrow) < 1 / (1 + np.exp(-z))).astype(np.float64)
    ob = oracle.Objective.glm(kind, X, y)
    gr = np.zeros(ncol)
    err = C.c_int(0)
    fr = getattr(oracle.lib(), "oracle_eval_" + kind)(ob.user, w.ctypes.data, gr.ctypes.data, ncol, C.byref(err))
    obj = R.Glm(kind, dev(X), dev(y))
    f, g = eval_gpu(obj, w)
    assert R.lib().lbfgsb200_objective_last_path(obj._user_ptr(0)) == path
    scale = max(abs(fr), float(np.sum(np.abs(y)) + nrow))
    assert abs(f - fr) <= 1e-12 * scale, (f, fr)
    assert np.max(np.abs(g - gr)) <= 1e-11 * max(np.max(np.abs(gr)), 1.0)
    # evaluating again gives the same bits (deterministic reductions, stage ring re-armed correctly)
    f2, g2 = eval_gpu(obj, w)
    assert f2 == f and np.array_equal(g2, g)


def test_glm_fused_equals_two_pass_in_a_solve():
    """Same OWL-QN solve with the fused kernel and (LBFGSB200_GLM_FUSED=0, in a child process) the two-pass
    kernels: same iteration/evaluation counts and fx to 1e-10 while the trajectories are comparable (the two
    kernels sum in different orders, and L-BFGS amplifies last-bit differences), the same minimum and the same
    sparsity pattern at the end."""
    code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
import rust_lbfgs_b200 as R
rng = np.random.default_rng(9)
nrow, ncol = 20000, 400
X = rng.standard_normal((nrow, ncol)); X[:, 0] = 1.0
wt = np.zeros(ncol); wt[rng.choice(ncol, 12, replace=False)] = rng.standard_normal(12)
y = (rng.random(nrow) < 1 / (1 + np.exp(-(X @ wt)))).astype(np.float64)
Xd, yd = torch.tensor(X, device="cuda:0"), torch.tensor(y, device="cuda:0")
w = torch.zeros(ncol, dtype=torch.float64, device="cuda:0")
tr = []
rep = R.lbfgs().with_orthantwise(40.0, 1).with_epsilon(1e-5).with_max_iterations(60).minimize(
    w, R.Glm("logistic", Xd, yd), lambda p: tr.append((p.niter, p.neval, p.ncall, p.fx)) and False)
print(json.dumps(dict(status=rep.status_name, trace=tr, nnz=int((w != 0).sum()), w=w.cpu().tolist())))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import json
    outs = []
    for fused in ("1", "0"):
        env = dict(os.environ, LBFGSB200_GLM_FUSED=fused)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    a, b = outs
    assert a["status"] == b["status"]
    k = min(12, len(a["trace"]), len(b["trace"]))
    assert [t[:3] for t in a["trace"][:k]] == [t[:3] for t in b["trace"][:k]]
    for s, t in zip(a["trace"][:k], b["trace"][:k]):
        assert abs(s[3] - t[3]) <= 1e-10 * abs(s[3])
    assert abs(a["trace"][-1][3] - b["trace"][-1][3]) <= 1e-8 * abs(a["trace"][-1][3])
    assert a["nnz"] == b["nnz"] and a["nnz"] < 400
    assert np.array_equal(np.sign(a["w"]), np.sign(b["w"]))
