#!/usr/bin/env python
"""bench.py — the headline measurement: L-BFGS iterations/s and HBM GB/s at n = 1e8, m = 6.

    python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[1]): Rosenbrock, x0 = (-1.2, 1.0) repeated (examples/sample.rs:10-17),
n = 1e8 f64 per GPU, m = 6, MoreThuente, device-resident evaluate.  One "step" is one L-BFGS iteration
(`propagate`, src/lbfgs.rs:503-560: line search with its evaluations, history update, two-loop).
Multi-GPU is weak scaling: every rank owns a contiguous 1e8-element shard of an N*1e8 vector and the
only exchange is the solver's scalar all-reduce.  `value` counts iterations/s normalised to n = 1e8
(iterations/s * n_global / 1e8), i.e. plain iterations/s on one GPU.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_REF = 100_000_000
METRIC = "lbfgs_iterations_per_sec_at_n1e8"
UNIT = "it/s"
WORKLOAD = ("Rosenbrock n=1e8 f64 per GPU, m=6, MoreThuente, device-resident evaluate "
            "(BASELINE.json configs[1])")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Samples that arrived inside [t_begin, t_end] (the timed region); all samples if that leaves none."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.15)]
        window = "timed region"
        if not rows:
            rows, window = [r for _, r in self.rows], "whole run (no sample fell inside the timed region)"
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


def algorithmic_bytes_survey(n, launches, kbytes, evaluations=None):
    # evaluations = unfused evaluate callbacks + fused trial evaluations (several may share one probe pass)
    """per iteration (6t + 6 + 8b) V solver-only + 2V t for the Rosenbrock evaluate,
    V = 8n bytes: trial step 3V t + post-eval dots 3V t + history update 7V + two-loop (8b - 1) V.
    t and b are the MEASURED evaluation / two-loop trip counts of the timed region."""
    V = 8.0 * n
    t = evaluations if evaluations is not None else launches["evaluate"] + launches.get("trial_eval", 0) + launches.get("probe", 0)
    iters = launches["history"] + launches.get("commit", 0)
    return 8.0 * t * V + 7.0 * iters * V + kbytes["backward"] + kbytes["forward"]


# ---------------------------------------------------------------------------------------------------
def cpu_oracle_rate(n_sample, warmup, steps, m=6, budget_s=None):
    """Times the oracle (CPU port of the reference, single thread) on Rosenbrock n_sample: iterations
    warmup+1 .. warmup+steps of the same workload.  Returns (it/s at n_sample, seconds, iterations timed)."""
    import numpy as np
    from oracle import oracle_lib as O
    x = np.empty(n_sample)
    x[0::2], x[1::2] = -1.2, 1.0
    stamps = []

    def on_progress(rec):
        stamps.append(time.perf_counter())
        if budget_s is not None and len(stamps) > warmup + 2 and stamps[-1] - stamps[warmup] > budget_s:
            return True
        return False
    p = O.default_param(m=m, max_iterations=1 + warmup + steps)
    O.minimize(p, x, O.Objective.builtin("rosenbrock"), progress=on_progress)
    # stamps[i] is the end of propagate #i+1; propagate #1 is the no-op (src/lbfgs.rs:507-510)
    done = len(stamps) - 1 - warmup
    if done < 1:
        return None, 0.0, 0
    dt = stamps[-1] - stamps[warmup]
    return done / dt, dt, done


def isometric_oracle_trace(n_global, m, iters, owl_c=None, record_x=False):
    """The reference algorithm (oracle) on the n = 1e8 workload WITHOUT summation error: with x0 = (-1.2, 1) repeated
    every pair of the vector is identical for ever, so the solve lives in a 2-dimensional subspace; u = sqrt(n/2) x
    is an isometry from that subspace to R^2, and the reference solver run on F(u) = (n/2) f(u / sqrt(n/2)) follows
    the same trajectory (same dot products, same line-search decisions).  Checked against the real oracle at
    n = 100, 1e5 and 2e6 (identical evaluation counts over 51 iterations, x to 3e-12).  Used as the CHECKER of the
    timed run: the CUDA path must take the same number of evaluations in every iteration and see the same f."""
    import math
    import numpy as np
    from oracle import oracle_lib as O
    K = n_global // 2
    rK = math.sqrt(K)

    def f(u, g):
        x0, x1 = u[0] / rK, u[1] / rK
        t1 = 1.0 - x0
        t2 = 10.0 * (x1 - x0 * x0)
        g1 = 20.0 * t2
        g0 = -2.0 * (x0 * g1 + t1)
        g[0], g[1] = rK * g0, rK * g1
        return K * (t1 * t1 + t2 * t2)
    kw = {}
    if owl_c is not None:   # c * sum |x_i| = c sqrt(n/2) (|u_0| + |u_1|): OWL-QN on all of x maps to OWL-QN on u
        kw = dict(orthantwise=1, owl_c=owl_c * rK, owl_start=0, owl_end=-1)
    r = O.minimize(O.default_param(m=m, max_iterations=iters, **kw), np.array([-1.2 * rK, 1.0 * rK]), O.Objective.python(f),
                   record_x=record_x)
    if record_x:
        for t in r["trace"]:
            t["x_pair"] = (t["x"][0] / rK, t["x"][1] / rK)
    return r["trace"]


def host_numa_cpus(local_rank):
    """CPUs of the NUMA node the GPU hangs off (so pinned buffers and the launching thread are local to it)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(path + "/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        return sorted(cpus & allowed) or None
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path, on the QUOTED configuration (n = 1e8,
    m = 6).  The reference is Rust and this image has no rustc/cargo, so it is the oracle port
    (oracle/lbfgs_oracle.cpp), single-threaded like the reference (README.md:24-25: rayon/SIMD are unchecked
    TODOs).  One step = one L-BFGS iteration at n = 1e8 (19 vectors x 0.8 GB = 15.2 GB of host RAM, about 5-8 s per
    iteration); W warm-up and K timed iterations as asked, cut short (and `steps` says so) if the timed part would
    exceed --ref-budget seconds.  Under torchrun only rank 0 runs; its sample is one GPU's shard (n = 1e8)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = int(args.ref_n)
    rate, dt, done = cpu_oracle_rate(n_sample, args.warmup, args.steps, m=args.m, budget_s=args.ref_budget)
    value = rate * n_sample / N_REF
    world = max(1, args.gpus)
    sample = (f"oracle (C++ port of the reference, 1 thread) Rosenbrock n={n_sample} (x0 = (-1.2, 1) repeated, m={args.m}, "
              f"MoreThuente), iterations {args.warmup + 1}..{args.warmup + done}"
              + ("" if n_sample == N_REF else "; it/s scaled by n_sample/1e8"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(N_REF, N_REF * world, args.m),
        "run": {"n_timed": n_sample, "seconds_timed": dt, "steps_requested": args.steps,
                "reference_arm": "CPU oracle port (the reference is Rust; no rustc here), 1 thread as the reference",
                "note": None if world == 1 else "timed at one GPU's shard (n = 1e8): the metric is normalised to n = 1e8 "
                        "and the CPU cost per element does not depend on n"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_local, n_global, m):
    """The `config` block: identical in both arms (everything that varies from run to run lives in `run`)."""
    return {"workload": WORKLOAD if (n_local == N_REF and m == 6) else
            f"Rosenbrock n={n_local} f64 per GPU, m={m}, MoreThuente, device-resident evaluate",
            "n_per_gpu": n_local, "n_global": n_global, "m": m, "linesearch": "MoreThuente",
            "l2_policy": "inputs larger than L2 (19 vectors x 0.8 GB vs 126 MB)"}


# ---------------------------------------------------------------------------------------------------
def timed_iterations(R, D, dev, comm, world, n_local, m, K, W, fused, barrier, sampler=None, dom="backward",
                     profile_pass=0, direction=None):
    """W warm-up + K timed L-BFGS iterations on a device-resident x0 = (-1.2, 1) repeated; CUDA events on the
    solver's stream (torch's current stream), max over ranks.  Returns a dict of raw measurements."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    n_global, goff = n_local * world, rank * n_local
    b = R.lbfgs().with_m(m).with_fused_trial(fused)
    if direction is not None:
        b = b.with_direction(direction)
    if comm is not None:
        b = b.with_shard(comm, n_global, goff)
    x = torch.empty(n_local, dtype=torch.float64, device=dev)
    x[0::2] = -1.2
    x[1::2] = 1.0
    obj = R.Rosenbrock()
    state = b.build(x, obj)
    state.propagate()                       # propagate #1 is the reference's no-op (src/lbfgs.rs:507-510)
    for _ in range(W):
        state.propagate()
    # CUDA events around the dominant kernel only inside the timed region (events around every launch cost
    # ~2 % at n = 1e8); the other kernels are timed in a separate pass after it
    state.profile_enable(True, kinds=[dom])
    state.profile_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    wall_begin = time.time()
    ev0.record()
    seen = []
    for _ in range(K):
        p = state.propagate()
        seen.append((p.niter, p.ncall, p.fx, p.xnorm, p.gnorm, p.step))
    ev1.record()
    torch.cuda.synchronize()
    wall_end = time.time()
    barrier()
    clocks = sampler.stop(wall_begin, wall_end) if sampler is not None else None
    out = {"ms_total": D.max_over_ranks(ev0.elapsed_time(ev1)), "prof": state.profile(), "final": state.report(),
           "seen": seen, "clocks": clocks, "prof_all": None, "profile_iterations": profile_pass,
           "transport": comm.transport if comm is not None else None}
    if profile_pass:   # per-kernel profile pass (NOT part of `value`): more iterations with events around every launch
        state.profile_enable(True)
        state.profile_reset()
        for _ in range(profile_pass):
            state.propagate()
        out["prof_all"] = state.profile()
    state.finish()
    state.close()
    obj.close()
    del x
    torch.cuda.empty_cache()
    return out


def isometric_parity(seen, n_global, m, iters):
    """The timed trajectory against the reference algorithm on the isometric 2-variable image of the workload."""
    try:
        ref = {t["niter"]: t for t in isometric_oracle_trace(n_global, m, iters)}
        worst, same = 0.0, True
        for (it, nc, fx, xn, gn, stp) in seen:
            t = ref.get(it)
            if t is None or t["ncall"] != nc:
                same = False
                break
            for a, b in ((fx, t["fx"]), (xn, t["xnorm"]), (gn, t["gnorm"]), (stp, t["step"])):
                worst = max(worst, abs(a - b) / max(abs(b), 1e-300))
        return {"checker": "oracle on the isometric 2-variable image of the workload (bench.py: isometric_oracle_trace)",
                "iterations_checked": len(seen), "evaluations_per_iteration_identical": same,
                "max_rel_err_fx_xnorm_gnorm_step": worst,
                "bar": "north_star: identical evaluation counts, 1e-10 relative (not widened)",
                "bar_met": bool(same and worst <= 1e-10)}
    except Exception as e:  # the checker must never break the measurement
        return {"error": repr(e)}


def nondegenerate_sharded_parity(R, D, dev, comm, world, rank, n=100_002, iters=40):
    """A small NON-degenerate solve through the same sharded production path (TREE reductions, peer exchange when
    world > 1): n = 100 002, x0 = (-1.2, 1) repeated scaled by linspace(0.9, 1.1) — every element different, uneven
    shards — against the ORACLE (faithful, and compensated for the drift scale) on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist
    lo, hi = D.shard_range(n, rank, world)
    x0 = np.empty(n)
    x0[0::2], x0[1::2] = -1.2, 1.0
    x0 *= np.linspace(0.9, 1.1, n)
    x = torch.tensor(x0[lo:hi], dtype=torch.float64, device=dev)
    b = R.lbfgs().with_max_iterations(iters)
    if comm is not None:
        b = b.with_shard(comm, n, lo)
    spans = [D.shard_range(n, r, world) for r in range(world)]
    width = max(h - l for l, h in spans)
    trace = []

    def on_progress(p):
        xs = p.x
        if world > 1:   # gather the iterate on every rank (same collective order everywhere: replicated control flow)
            pad = torch.zeros(width, dtype=torch.float64, device=dev)
            pad[: hi - lo] = xs
            allx = torch.empty(world * width, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allx, pad)
            if rank == 0:
                xs = torch.cat([allx[r * width: r * width + (h - l)] for r, (l, h) in enumerate(spans)])
        if rank == 0:
            trace.append(dict(niter=p.niter, ncall=p.ncall, fx=p.fx, xnorm=p.xnorm, gnorm=p.gnorm, x=xs.cpu().numpy().copy()))
        return False
    obj = R.Rosenbrock()
    rep = b.minimize(x, obj, on_progress)
    obj.close()
    if rank != 0:
        return None
    from oracle import oracle_lib as O
    ref = O.minimize(O.default_param(max_iterations=iters), x0.copy(), O.Objective.builtin("rosenbrock"), record_x=True)
    alt = O.minimize(O.default_param(max_iterations=iters, reduction_mode=1), x0.copy(), O.Objective.builtin("rosenbrock", 1),
                     record_x=True)
    same = len(trace) == len(ref["trace"]) and [t["ncall"] for t in trace] == [t["ncall"] for t in ref["trace"]]
    ex = ef = drift = 0.0
    strict_through, widened_ok = 0, True
    for a, c, g in zip(ref["trace"], alt["trace"], trace):
        if a["ncall"] != g["ncall"]:
            break
        sx = max(float(np.max(np.abs(a["x"]))), 1e-300)
        ex_i = float(np.max(np.abs(a["x"] - g["x"]))) / sx
        ef_i = abs(a["fx"] - g["fx"]) / max(abs(a["fx"]), a["gnorm"] * a["xnorm"])
        drift = max(drift, float(np.max(np.abs(a["x"] - c["x"]))) / sx,
                    abs(a["fx"] - c["fx"]) / max(abs(a["fx"]), a["gnorm"] * a["xnorm"]))
        ex, ef = max(ex, ex_i), max(ef, ef_i)
        if ex <= 1e-10 and ef <= 1e-10:
            strict_through = a["niter"]
        if max(ex_i, ef_i) > max(1e-10, 100.0 * drift):
            widened_ok = False
    counts_ok = bool(same and rep.status_name == ref["status_name"])
    return {"checker": "oracle (faithful CPU restatement of the reference, sequential sums) on rank 0",
            "n": n, "x0": "(-1.2, 1) repeated * linspace(0.9, 1.1)", "ranks": world, "iterations": len(trace),
            "status": rep.status_name, "oracle_status": ref["status_name"],
            "evaluations_per_iteration_identical": bool(same),
            "max_rel_err_x": ex, "max_rel_err_fx": ef,
            "drift_between_two_cpu_summation_orders": drift,
            # which bar the production (tree-sum) path met: north_star's 1e-10 as written holds while the two CPU
            # summation orders of the reference algorithm (sequential vs compensated) themselves agree to 1e-12;
            # beyond that the bar is 100x their drift (the tolerance rule of tests/gpu_util.py)
            "strict_1e-10_holds_through_iteration": strict_through,
            "bar": "identical status and evaluation counts; x and fx within max(1e-10, 100 x the drift between two CPU "
                   "summation orders of the oracle) in every iteration",
            "bar_met": bool(counts_ok and widened_ok)}


def compact_block(R, D, dev, comm, world, rank, n_local, m, K, W, fused, barrier, peak, check, n_ref, unit):
    """The same workload with the opt-in compact search direction (`with_direction("compact")`, csrc/compact.cu): the
    reference's recursion and element-wise operations, its 2m scalars from inner products of the unmodified ring
    vectors — two passes over the ring per iteration instead of 2m dependent ones.  NOT the headline `value` (that
    is the reference's own arithmetic, trip by trip); reported beside it with its own parity check."""
    r = timed_iterations(R, D, dev, comm, world, n_local, m, K, W, fused, barrier, None, "forward", direction="compact",
                         profile_pass=max(3, min(10, K)))
    launches, kbytes, kms, moved, evals = kernel_tables(r["prof"], n_local, K, r["ms_total"], peak)
    pa = r["prof_all"]
    names = {"backward": "pass A (k_gram)", "forward": "pass B (k_direction)", "probe": "probe", "commit": "commit"}
    out = {
        "what": "opt-in LBFGSB200_DIRECTION_COMPACT: alpha_j / beta_j from S^T Y, Y^T Y kept on the device; pass A (gram) + "
                "scalar recursions + pass B (direction) instead of 2m trips; same element-wise arithmetic",
        "ms_per_step": r["ms_total"] / K, "value": K / (r["ms_total"] / 1e3) * (n_local * world) / n_ref, "unit": unit,
        "algorithmic_bytes_per_iteration_per_gpu": moved / max(1, K),
        "algorithmic_GBps_per_gpu": moved / 1e9 / (r["ms_total"] / 1e3),
        "frac_of_peak_per_gpu": moved / 1e9 / (r["ms_total"] / 1e3) / peak,
        "evaluations_per_iteration": sum(t[1] for t in r["seen"]) / max(1, K),
        "line_search_passes_per_iteration": evals / max(1, K),
        "direction_kernel_GBps": (kbytes["forward"] / 1e9) / (kms["forward"] / 1e3) if kms["forward"] > 0 else None,
        "profile_pass_kernel_GBps": {names.get(k, k): round(pa["bytes"][k] / 1e9 / (pa["ms"][k] / 1e3), 1) for k in pa["ms"]
                                     if pa["ms"][k] > 0 and pa["bytes"][k] > 0},
        "profile_pass_kernel_ms_per_iteration": {names.get(k, k): round(pa["ms"][k] / max(3, min(10, K)), 3) for k in pa["ms"]
                                                 if pa["ms"][k] > 0},
        "launches_per_iteration": sum(launches.values()) / max(1, K),
        "allreduces": r["prof"]["allreduces"],
        "parity": isometric_parity(r["seen"], n_local * world, m, 1 + W + K) if (check and rank == 0) else None,
    }
    return out


def kernel_tables(prof, n_local, K, ms_total, peak):
    launches, kbytes, kms = dict(prof["launches"]), dict(prof["bytes"]), dict(prof["ms"])
    kbytes["evaluate"] = 2.0 * 8.0 * n_local * launches["evaluate"]     # Rosenbrock: 1R 1W per evaluation
    moved = sum(kbytes.values())
    evals = launches["evaluate"] + launches.get("trial_eval", 0) + launches.get("probe", 0)
    return launches, kbytes, kms, moved, evals


def run_ours(args):
    import torch
    import rust_lbfgs_b200 as R
    from rust_lbfgs_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rust_lbfgs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpus = host_numa_cpus(local_rank)      # launching thread + pinned buffers on the GPU's own NUMA node
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except Exception:
            cpus = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = D.Comm(rank, world, local_rank)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    n_local = int(args.n)
    n_global = n_local * world
    K, W, m = args.steps, args.warmup, args.m
    fused = False if args.unfused_trial else ("trial" if args.fused_trial_only else "probe")
    peak, peak_src = load_peaks()

    # ---- device-resident run: `value`, roofline ------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()                     # nvidia-smi needs ~0.5 s to produce its first sample: start it early
    DOM = "backward"
    P = max(3, min(10, K))
    r = timed_iterations(R, D, dev, comm, world, n_local, m, K, W, fused, barrier, sampler, DOM, profile_pass=P)
    ms_total, clocks, seen, final = r["ms_total"], r["clocks"], r["seen"], r["final"]
    launches, kbytes, kms, moved, evals = kernel_tables(r["prof"], n_local, K, ms_total, peak)
    gpu_launches = int(sum(launches.values()))
    value = K / (ms_total / 1e3) * n_global / N_REF

    dom_gbs = (kbytes[DOM] / 1e9) / (kms[DOM] / 1e3) if kms[DOM] > 0 else None
    traffic = load_traffic()
    roofline = {
        "bound": "hbm", "kernel": f"two_loop_{DOM}_step (k_{DOM})", "achieved": dom_gbs, "peak": peak,
        "unit": "GB/s", "frac": (dom_gbs / peak) if dom_gbs else None,
        "traffic": (traffic or {}).get("dram_bytes_per_launch"),
        "peak_source": peak_src,
        "launches": int(launches[DOM]), "avg_launch_ms": kms[DOM] / max(1, launches[DOM]),
        "algorithmic_bytes_per_launch": kbytes[DOM] / max(1, launches[DOM]),
    }
    surv_bytes = algorithmic_bytes_survey(n_local, launches, kbytes, evaluations=sum(t[1] for t in seen))
    pa = r["prof_all"]
    iteration = {
        # bytes the launched kernels must move (DESIGN.md §3 per-kernel passes x 8n) / wall time of the K steps
        "algorithmic_GBps": moved * world / 1e9 / (ms_total / 1e3),
        "algorithmic_GBps_per_gpu": moved / 1e9 / (ms_total / 1e3),
        "frac_of_peak_per_gpu": moved / 1e9 / (ms_total / 1e3) / peak,
        "algorithmic_bytes_per_iteration_per_gpu": moved / max(1, K),
        # SURVEY.md §8(d)'s formula prices a trial at 8V (K1 + evaluate + K2); probes move 2V, so this
        # "unfused-equivalent" rate can exceed what the HBM actually carried — reported for comparison only
        "survey_formula_equivalent_GBps_per_gpu": surv_bytes / 1e9 / (ms_total / 1e3),
        "line_search_trials": ("probe + commit" if launches.get("probe", 0) > 0 else
                               "fused trial" if launches.get("trial_eval", 0) > 0 else "unfused (K1 + evaluate + K2)"),
        "evaluations_per_iteration": sum(t[1] for t in seen) / max(1, K),
        # several trial points share one pass over xp and d when the search extrapolates (lbfgsb200_probe_multi_fn)
        "line_search_passes_per_iteration": evals / max(1, K),
        "kernel_ms_timed_region": {k: round(v, 3) for k, v in kms.items() if v > 0},
        "profile_pass": {
            "note": f"{P} extra iterations after the timed region with CUDA events around every launch",
            "kernel_ms": {k: round(v, 3) for k, v in pa["ms"].items() if v > 0},
            "kernel_GBps": {k: round(pa["bytes"][k] / 1e9 / (pa["ms"][k] / 1e3), 1) for k in pa["ms"]
                            if pa["ms"][k] > 0 and pa["bytes"][k] > 0},
        },
        "host_syncs": r["prof"]["host_syncs"], "allreduces": r["prof"]["allreduces"],
        "allreduce_transport": r["transport"],
    }

    # ---- end to end through the public API with HOST buffers: `e2e` ---------------------------------
    # The reference-shaped call: x is a HOST slice (src/lbfgs.rs:399), passed to the C ABI's host-buffer entry.
    # Timed: the whole call — pinned host -> device copy of x0, solver creation, a complete minimize() of W+K
    # iterations (build + line searches + two-loops), device -> host copy of the result, teardown.  The copies
    # happen once per solve, so bytes/step are 8n/(W+K) each way (+ the few scalars read back per iteration).
    iters_e2e = W + K
    xh = torch.empty(n_local, dtype=torch.float64, pin_memory=True)
    xh[0::2] = -1.2
    xh[1::2] = 1.0
    builder = R.lbfgs().with_m(m).with_fused_trial(fused).with_max_iterations(iters_e2e + 1)
    if comm is not None:
        builder = builder.with_shard(comm, n_global, rank * n_local)
    obj = R.Rosenbrock()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rep = builder.minimize_host(xh, obj, None, device=local_rank)   # ONE C-ABI call: lbfgsb200_minimize_host_ex
    t1 = time.perf_counter()
    barrier()
    e2e_s = D.max_over_ranks(t1 - t0)
    e2e_iters = rep.niter - 1
    e2e = {
        "value": (e2e_iters / e2e_s) * n_global / N_REF, "unit": UNIT,
        "h2d_bytes_per_step": 8.0 * n_local * world / max(1, e2e_iters),
        "d2h_bytes_per_step": 8.0 * n_local * world / max(1, e2e_iters) + 64.0 * 3,
        "iterations": e2e_iters, "seconds": e2e_s, "evaluations": rep.neval,
        "host_buffer": "pinned, allocated on the GPU's NUMA node" if cpus else "pinned",
        "note": "one lbfgsb200_minimize_host_ex() call on a pinned HOST buffer: H2D of x0, solver creation, build, "
                "W+K iterations, D2H of x, teardown",
    }
    # the same call with the opt-in compact search direction (reported inside the `compact_direction` block)
    e2e_compact = None
    if not args.no_compact and m <= 32:
        try:
            xh[0::2] = -1.2
            xh[1::2] = 1.0
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            repc = builder.with_direction("compact").minimize_host(xh, obj, None, device=local_rank)
            t1 = time.perf_counter()
            barrier()
            sc = D.max_over_ranks(t1 - t0)
            e2e_compact = {"value": ((repc.niter - 1) / sc) * n_global / N_REF, "unit": UNIT, "iterations": repc.niter - 1,
                           "seconds": sc, "evaluations": repc.neval,
                           "note": "one lbfgsb200_minimize_host_ex() call on the pinned host buffer, as `e2e` above"}
        except Exception as e:
            e2e_compact = {"error": repr(e)}
    # what the two bulk copies of that call cost on their own, all ranks copying at the same time (the limiter of
    # e2e on N GPUs of one host: they share the host's memory and PCIe root complexes)
    xd = torch.empty(n_local, dtype=torch.float64, device=dev)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    xd.copy_(xh, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    xh.copy_(xd, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    h2d_s, d2h_s = D.max_over_ranks(t1 - t0), D.max_over_ranks(t2 - t1)
    e2e["bulk_copies_alone"] = {"h2d_seconds": h2d_s, "d2h_seconds": d2h_s,
                                "h2d_GBps_per_gpu": 8.0 * n_local / 1e9 / h2d_s, "d2h_GBps_per_gpu": 8.0 * n_local / 1e9 / d2h_s,
                                "share_of_e2e_seconds": (h2d_s + d2h_s) / e2e_s}
    del xh, xd
    obj.close()

    # ---- parity of what was just timed (checker only; rank 0) -----------------------------------------------
    parity = None
    if not args.no_cpu_baseline:
        parity = {}
        if rank == 0 and n_global % 2 == 0:
            parity["timed_trajectory"] = isometric_parity(seen, n_global, m, 1 + W + K)
        try:
            nd = nondegenerate_sharded_parity(R, D, dev, comm, world, rank)
        except Exception as e:
            nd = {"error": repr(e)}
        if rank == 0:
            parity["nondegenerate_sharded_solve"] = nd

    # ---- the opt-in compact search direction on the headline workload ------------------------------------------------
    compact = None
    if not args.no_compact and m <= 32:
        try:
            compact = compact_block(R, D, dev, comm, world, rank, n_local, m, K, W, fused, barrier, peak,
                                    not args.no_cpu_baseline, N_REF, UNIT)
        except Exception as e:   # never break the headline measurement
            compact = {"error": repr(e)}
        compact["e2e"] = e2e_compact

    # ---- BASELINE configs[4]: Rosenbrock n = 2^31, m = 20 sharded over 8 GPUs = 2^28 elements per GPU -----------
    # Run at every N (weak scaling at 2^28 per GPU), so the driver's own N = 1, 2, 4, 8 set yields north_star's
    # "sharded n = 2^31, m = 20 scales >= 6x from 1 to 8 GPUs" from its per-N lines.
    config5 = None
    if not args.no_config5:
        n5, m5, K5, W5 = 1 << 28, 20, max(10, min(K, 20)), 3
        R.lib().lbfgsb200_trim_pool(local_rank)    # hand the n = 1e8 arenas back before asking for 92 GiB
        free_b, _ = torch.cuda.mem_get_info()
        if free_b > (2 * m5 + 7) * 8 * n5 * 1.03:
            r5 = timed_iterations(R, D, dev, comm, world, n5, m5, K5, W5, fused, barrier, None, DOM)
            l5, kb5, km5, moved5, ev5 = kernel_tables(r5["prof"], n5, K5, r5["ms_total"], peak)
            config5 = {
                "workload": "Rosenbrock n=2^28 f64 per GPU (n=2^31 on 8 GPUs), m=20, MoreThuente (BASELINE.json configs[4])",
                "n_per_gpu": n5, "n_global": n5 * world, "m": m5, "steps": K5, "warmup": W5, "scaling": "weak",
                "ms_per_step": r5["ms_total"] / K5,
                # element-iterations/s in units of 2^28 elements: plain iterations/s on one GPU; the 1 -> N factor is
                # this value at N over this value at 1
                "value": K5 / (r5["ms_total"] / 1e3) * world, "unit": "it/s x n_global/2^28",
                "algorithmic_GBps": moved5 * world / 1e9 / (r5["ms_total"] / 1e3),
                "algorithmic_GBps_per_gpu": moved5 / 1e9 / (r5["ms_total"] / 1e3),
                "frac_of_peak_per_gpu": moved5 / 1e9 / (r5["ms_total"] / 1e3) / peak,
                "evaluations_per_iteration": sum(t[1] for t in r5["seen"]) / K5,
                "line_search_passes_per_iteration": ev5 / K5,
                "k_backward_GBps": (kb5[DOM] / 1e9) / (km5[DOM] / 1e3) if km5[DOM] > 0 else None,
                "parity": isometric_parity(r5["seen"], n5 * world, m5, 1 + W5 + K5) if (rank == 0 and not args.no_cpu_baseline) else None,
            }
            if not args.no_compact:
                try:
                    config5["compact_direction"] = compact_block(R, D, dev, comm, world, rank, n5, m5, K5, W5, fused, barrier,
                                                                 peak, not args.no_cpu_baseline, n5, "it/s x n_global/2^28")
                except Exception as e:
                    config5["compact_direction"] = {"error": repr(e)}
        else:
            config5 = {"skipped": f"needs {(2 * m5 + 7) * 8 * n5 / 2**30:.0f} GiB of free HBM, {free_b / 2**30:.0f} GiB free"}

    # ---- the reference's CPU path on this host (rank 0, N=1 only), on the quoted size ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sample = int(args.cpu_n)
        rate, dt, done = cpu_oracle_rate(n_sample, 1, args.cpu_steps, m=m, budget_s=args.cpu_budget)
        if rate:
            cpu = {"value": rate * n_sample / N_REF, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle (C++ port of the reference, 1 thread) Rosenbrock n={n_sample}, m={m}: iterations 2.."
                             f"{1 + done} in {dt:.1f} s" + ("" if n_sample == N_REF else "; it/s scaled by n_sample/1e8"),
                   "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_local, n_global, m),
            "run": {"ncall_per_iteration": [s[1] for s in seen], "final_fx": final.fx, "final_gnorm": final.gnorm,
                    "host_numa_cpus": len(cpus) if cpus else None},
            "roofline": roofline, "iteration": iteration, "cpu_baseline": cpu, "parity": parity, "e2e": e2e,
            "compact_direction": compact, "config5": config5, "gpu_launches": gpu_launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        comm.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # (--elems / --history: torchrun's own parser chokes on "--n" / "--m" as ambiguous abbreviations)
    ap.add_argument("--n", "--elems", dest="n", type=float, default=1e8, help="elements per GPU")
    ap.add_argument("--m", "--history", dest="m", type=int, default=6)
    ap.add_argument("--ref-n", type=float, default=1e8, help="--impl reference: elements (the quoted configuration: 1e8)")
    ap.add_argument("--ref-budget", type=float, default=420.0, help="--impl reference: seconds of timed iterations at most")
    ap.add_argument("--cpu-n", type=float, default=1e8, help="cpu_baseline: elements")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="cpu_baseline: seconds of timed iterations at most")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline and the oracle-side parity checks")
    ap.add_argument("--no-config5", action="store_true", help="skip the 2^28-per-GPU, m=20 block")
    ap.add_argument("--no-compact", action="store_true", help="skip the opt-in compact-direction blocks")
    ap.add_argument("--unfused-trial", action="store_true",
                    help="line-search trials as K1 + evaluate + K2 (three passes)")
    ap.add_argument("--fused-trial-only", action="store_true",
                    help="the one-pass trial that writes x and g (round 1's path) instead of probe + commit")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
