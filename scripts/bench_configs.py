"""Secondary measurements: BASELINE.json configs[0], [2], [3] (the headline bench.py covers [1] and [4]).

    python scripts/bench_configs.py [--full] [--only cfg1,cfg3,cfg4,small]

Each line is a JSON record {config, n, wall_s, iterations, evaluations, it_per_s, ...}.  Synthetic data per
SURVEY.md §8(d).  --full uses the full sizes (1e6 x 1e4 GLM = 80 GB of X; 1e5 LJ atoms); the default sizes
finish in seconds."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rust_lbfgs_b200 as R

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
DEV = torch.device("cuda", LOCAL)
COMM = None
if WORLD > 1:   # torchrun: one process per GPU; GLM rows / LJ atoms are sharded over the ranks
    import torch.distributed as dist
    from rust_lbfgs_b200 import dist as D
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=DEV)
    COMM = D.Comm(RANK, WORLD, LOCAL)


def sync_time():
    torch.cuda.synchronize()
    return time.perf_counter()


def solve(builder, x, obj, tag, extra=None, max_iter=None):
    if max_iter:
        builder = builder.with_max_iterations(max_iter)
    ncalls = []
    t0 = sync_time()
    try:
        rep = builder.minimize(x, obj, lambda p: ncalls.append(p.ncall) and False)
        status = rep.status_name
    except R.LbfgsError as e:
        rep, status = e.report, e.status_name
    t1 = sync_time()
    rec = dict(config=tag, n=int(x.numel()), wall_s=t1 - t0, status=status, iterations=len(ncalls) - 1,
               evaluations=rep.neval, fx=rep.fx, gnorm=rep.gnorm,
               it_per_s=(len(ncalls) - 1) / (t1 - t0), ms_per_evaluation=1e3 * (t1 - t0) / max(1, rep.neval))
    rec.update(extra or {})
    if "objective_ms_alone" in rec and rec["iterations"] > 0:   # what the solver adds around the objective, per iteration
        rec["solver_ms_per_iteration"] = (1e3 * (t1 - t0) - rec["evaluations"] * rec["objective_ms_alone"]) / rec["iterations"]
    rec["n_gpus"] = WORLD
    if RANK == 0:
        print(json.dumps(rec), flush=True)
    return rec


def cfg1():
    """examples/sample.rs: Rosenbrock N=100, defaults."""
    x = torch.empty(100, dtype=torch.float64, device=DEV)
    for rep in range(2):  # second run: warm
        x[0::2], x[1::2] = -1.2, 1.0
        solve(R.lbfgs(), x, R.Rosenbrock(), f"cfg1 rosenbrock n=100 (run {rep})")


def small_n():
    """launch-latency regime: Rosenbrock at n = 1e4 .. 1e7, 40 iterations."""
    for n in (10_000, 300_000, 1_000_000, 10_000_000):
        x = torch.empty(n, dtype=torch.float64, device=DEV)
        x[0::2], x[1::2] = -1.2, 1.0
        solve(R.lbfgs(), x, R.Rosenbrock(), f"rosenbrock n={n}", max_iter=41)


def make_glm(nrow, ncol, seed=2024):
    """X: column 0 = 1, others N(0,1); w* 1% non-zeros; y ~ Bernoulli(sigmoid(X w*)).  Generated on the device in
    row chunks (80 GB at the full size)."""
    g0 = torch.Generator(device=DEV)
    g0.manual_seed(2024)            # the true model is the same on every rank ...
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)             # ... the rows are this rank's own
    X = torch.empty((nrow, ncol), dtype=torch.float64, device=DEV)
    wstar = torch.zeros(ncol, dtype=torch.float64, device=DEV)
    nz = max(1, ncol // 100)
    idx = torch.randperm(ncol, generator=g0, device=DEV)[:nz]
    wstar[idx] = torch.randn(nz, generator=g0, device=DEV, dtype=torch.float64)
    y = torch.empty(nrow, dtype=torch.float64, device=DEV)
    chunk = max(1, min(nrow, (1 << 28) // ncol))
    for r0 in range(0, nrow, chunk):
        r1 = min(nrow, r0 + chunk)
        X[r0:r1].normal_(generator=g)
        X[r0:r1, 0] = 1.0
        p = torch.sigmoid(X[r0:r1] @ wstar)
        y[r0:r1] = (torch.rand(r1 - r0, generator=g, device=DEV, dtype=torch.float64) < p).to(torch.float64)
    return X, y


def cfg3(full):
    """OWL-QN L1-regularised logistic regression, intercept unpenalised (start = 1), mirrors tests/owlqn.rs:46-49."""
    nrow_all, ncol = (1_000_000, 10_000) if full else (100_000, 2_000)
    nrow = nrow_all // WORLD          # this rank's block of rows (w is replicated, the solver runs unsharded)
    X, y = make_glm(nrow, ncol, seed=3000 + RANK)
    obj = R.Glm("logistic", X, y)
    if COMM is not None:
        obj.shard(COMM)
    w = torch.zeros(ncol, dtype=torch.float64, device=DEV)
    gx = torch.empty_like(w)
    # objective alone
    import ctypes as C
    L = R.lib()
    fx = torch.zeros(1, dtype=torch.float64, device=DEV)
    st = int(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        L.lbfgsb200_objective_eval(obj._user_ptr(LOCAL), w.data_ptr(), gx.data_ptr(), ncol, st, fx.data_ptr())
    t0 = sync_time()
    reps = 5
    for _ in range(reps):
        L.lbfgsb200_objective_eval(obj._user_ptr(LOCAL), w.data_ptr(), gx.data_ptr(), ncol, st, fx.data_ptr())
    t1 = sync_time()
    xbytes = 8.0 * nrow * ncol
    glm_ms = 1e3 * (t1 - t0) / reps
    if RANK == 0:
        print(json.dumps(dict(config=f"cfg3 glm objective alone {nrow_all}x{ncol} over {WORLD} GPU(s)",
                              ms_per_evaluation=1e3 * (t1 - t0) / reps, X_GB_per_gpu=xbytes / 1e9,
                              GBps_per_gpu_one_pass_equivalent=xbytes / 1e9 / ((t1 - t0) / reps))), flush=True)
    c = 1.0 * nrow_all / 500.0   # tests/owlqn.rs uses c = 1 with 500 rows
    # (a two-iteration solve first: the process's first solver pays for the memory pool, pinned scalars, module loading)
    R.lbfgs().with_orthantwise(c, 1).with_max_iterations(3).minimize(w.clone(), obj, None)
    rec = solve(R.lbfgs().with_orthantwise(c, 1).with_epsilon(1e-4), w, obj, f"cfg3 owlqn logistic {nrow_all}x{ncol} c={c}",
                extra=dict(X_GB_per_gpu=xbytes / 1e9, objective_ms_alone=glm_ms), max_iter=60)
    # the objective alone again, at the solution and with a synchronisation after every evaluation as inside the solve
    # (w = 0 makes every exp() argument 0, and back-to-back evaluations hide the kernel's ramp-up): this is the number
    # the solver's share per iteration should be read against
    t0 = sync_time()
    for _ in range(reps):
        L.lbfgsb200_objective_eval(obj._user_ptr(LOCAL), w.data_ptr(), gx.data_ptr(), ncol, st, fx.data_ptr())
        torch.cuda.synchronize()
    t1 = sync_time()
    glm_ms_sol = 1e3 * (t1 - t0) / reps
    if RANK == 0:
        print(json.dumps(dict(config=f"cfg3 glm objective alone at the solution, one synchronisation per evaluation",
                              ms_per_evaluation=glm_ms_sol,
                              solver_ms_per_iteration=(1e3 * rec["wall_s"] - rec["evaluations"] * glm_ms_sol) / max(1, rec["iterations"]))),
              flush=True)
    if RANK == 0:
        print(json.dumps(dict(config="cfg3 sparsity", nonzeros=int((w != 0).sum()), ncol=ncol)), flush=True)


def cfg4(full):
    """Lennard-Jones cluster on a jittered simple-cubic lattice (spacing 1.12, jitter 0.05, seed 7).
    LJ PARITY IS UNPINNED: the reference has no test for examples/lj.rs:20-64 (and its vecdist comes from the
    un-vendored vecfx crate); the oracle restates it and the CUDA kernels are checked against that restatement only."""
    import ctypes as C
    side = 47 if full else 16           # 47^3 = 103823 ~ 1e5 atoms
    rng = np.random.default_rng(7)
    g = np.arange(side, dtype=np.float64) * 1.12
    p = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    p += rng.uniform(-0.05, 0.05, p.shape)
    na = p.shape[0]
    flat = p.ravel()
    offs = None
    if COMM is not None:   # atoms sharded: solver vectors and forces local, positions gathered per evaluation
        offs = [D.shard_range(flat.size, r, WORLD, granule=6)[0] for r in range(WORLD)] + [flat.size]
        x0 = torch.tensor(flat[offs[RANK]:offs[RANK + 1]], dtype=torch.float64, device=DEV)
    else:
        x0 = torch.tensor(flat, dtype=torch.float64, device=DEV)
    # the objective alone, both per-pair arithmetics: ms per evaluation and ordered pairs per second
    L = R.lib()
    st = int(torch.cuda.current_stream().cuda_stream)
    gx, fx = torch.empty_like(x0), torch.zeros(1, dtype=torch.float64, device=DEV)
    lj_ms = {}
    for fast in (False, True):
        lj = R.LennardJones(fast=fast)
        if COMM is not None:
            lj.shard(COMM, offs)
        h = lj._user_ptr(LOCAL)
        for _ in range(2):
            L.lbfgsb200_objective_eval(h, x0.data_ptr(), gx.data_ptr(), x0.numel(), st, fx.data_ptr())
        reps = 5
        t0 = sync_time()
        for _ in range(reps):
            L.lbfgsb200_objective_eval(h, x0.data_ptr(), gx.data_ptr(), x0.numel(), st, fx.data_ptr())
        t1 = sync_time()
        lj_ms[fast] = 1e3 * (t1 - t0) / reps
        if RANK == 0:
            ms = 1e3 * (t1 - t0) / reps
            print(json.dumps(dict(config=f"cfg4 lj objective alone, {na} atoms over {WORLD} GPU(s)",
                                  arithmetic="1/r^2 + FMA (opt-in)" if fast else "reference per-pair arithmetic",
                                  ms_per_evaluation=ms, ordered_pairs_per_s=na * (na - 1) / (ms / 1e3),
                                  parity="UNPINNED (no reference test for examples/lj.rs; oracle restatement only)")), flush=True)
        lj.close()
    # (a short solve first: the process's first solver pays for the memory pool, pinned scalars, module loading — the
    # 0.7 ms per iteration by which the first configuration below exceeded the others in profiles/r02k_configs_8gpu_full.log)
    bw, ljw = R.lbfgs().with_damping(True).with_max_iterations(3), R.LennardJones()
    if COMM is not None:
        ljw.shard(COMM, offs)
        bw = bw.with_shard(COMM, flat.size, offs[RANK])
    bw.minimize(x0.clone(), ljw, None)
    ljw.close()
    for tag, mk in (("gradient-only max_linesearch=2", lambda: R.lbfgs().with_gradient_only().with_max_linesearch(2)),
                    ("damped", lambda: R.lbfgs().with_damping(True))):
        for fast in (False, True):
            b, lj = mk(), R.LennardJones(fast=fast)
            if COMM is not None:
                lj.shard(COMM, offs)
                b = b.with_shard(COMM, flat.size, offs[RANK])
            x = x0.clone()
            solve(b, x, lj, f"cfg4 lj {na} atoms {tag}" + (" [fast arithmetic]" if fast else ""),
                  extra=dict(pairs=na * (na - 1) // 2, parity="UNPINNED", objective_ms_alone=lj_ms[fast]),
                  max_iter=21 if full else 41)
            lj.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="cfg1,small,cfg3,cfg4")
    a = ap.parse_args()
    which = a.only.split(",")
    if "cfg1" in which and WORLD == 1:
        cfg1()
    if "small" in which and WORLD == 1:
        small_n()
    if "cfg4" in which:
        cfg4(a.full)
    if "cfg3" in which:
        cfg3(a.full)
    if COMM is not None:
        COMM.close()
        dist.destroy_process_group()
