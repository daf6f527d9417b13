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


def algorithmic_bytes_survey(n, launches, kbytes):
    # evaluations = unfused evaluate callbacks + fused trial evaluations
    """per iteration (6t + 6 + 8b) V solver-only + 2V t for the Rosenbrock evaluate,
    V = 8n bytes: trial step 3V t + post-eval dots 3V t + history update 7V + two-loop (8b - 1) V.
    t and b are the MEASURED evaluation / two-loop trip counts of the timed region."""
    V = 8.0 * n
    t = launches["evaluate"] + launches.get("trial_eval", 0) + launches.get("probe", 0)
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


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path.  The reference is Rust and this
    image has no rustc/cargo, so it is the oracle port (oracle/lbfgs_oracle.cpp), single-threaded like the
    reference (README.md:24-25: rayon/SIMD are unchecked TODOs).  Each step is one L-BFGS iteration on a
    bounded sample n_sample of the n = 1e8 workload; the rate is normalised to n = 1e8."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = int(args.ref_n)
    rate, dt, done = cpu_oracle_rate(n_sample, args.warmup, args.steps, m=args.m)
    value = rate * n_sample / N_REF
    sample = (f"Rosenbrock n={n_sample} (same x0 pattern, m={args.m}, MoreThuente), iterations "
              f"{args.warmup + 1}..{args.warmup + done}; it/s scaled by n_sample/1e8")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done * (N_REF / n_sample),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_per_gpu": N_REF, "n_global": N_REF, "m": args.m, "linesearch": "MoreThuente",
                   "reference_arm": "CPU oracle port (the reference is Rust; no rustc here), 1 thread as the reference",
                   "n_sample": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
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
    goff = rank * n_local
    K, W, m = args.steps, args.warmup, args.m

    def make_builder():
        b = R.lbfgs().with_m(m).with_fused_trial(not args.unfused_trial)
        if comm is not None:
            b = b.with_shard(comm, n_global, goff)
        return b

    def fill_x0(t):
        t[0::2] = -1.2
        t[1::2] = 1.0

    obj = R.Rosenbrock()

    # ---- device-resident run: `value`, roofline ------------------------------------------------
    x = torch.empty(n_local, dtype=torch.float64, device=dev)
    fill_x0(x)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                     # nvidia-smi needs ~0.5 s to produce its first sample: start it early
    state = make_builder().build(x, obj)
    state.propagate()                       # propagate #1 is the reference's no-op (src/lbfgs.rs:507-510)
    for _ in range(W):
        state.propagate()
    # CUDA events around the dominant kernel only inside the timed region (events around every launch cost
    # ~2 % at n = 1e8); the other kernels are timed in a separate pass after it
    DOM = "backward"
    state.profile_enable(True, kinds=[DOM])
    state.profile_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    wall_begin = time.time()
    ev0.record()
    ncalls, seen = [], []
    for _ in range(K):
        p = state.propagate()
        ncalls.append(p.ncall)
        seen.append((p.niter, p.ncall, p.fx, p.xnorm, p.gnorm, p.step))
    ev1.record()
    torch.cuda.synchronize()
    wall_end = time.time()
    barrier()
    clocks = sampler.stop(wall_begin, wall_end) if rank == 0 else None
    ms_total = D.max_over_ranks(ev0.elapsed_time(ev1))
    prof = state.profile()
    final = state.report()
    # per-kernel profile pass (NOT part of `value`): a few more iterations with events around every launch
    state.profile_enable(True)
    state.profile_reset()
    P = max(3, min(10, K))
    for _ in range(P):
        state.propagate()
    prof_all = state.profile()
    state.finish()
    state.close()

    launches, kbytes, kms = prof["launches"], prof["bytes"], prof["ms"]
    gpu_launches = int(sum(launches.values()))
    it_per_s = K / (ms_total / 1e3)
    value = it_per_s * n_global / N_REF

    peak, peak_src = load_peaks()
    dom = DOM
    dom_gbs = (kbytes[dom] / 1e9) / (kms[dom] / 1e3) if kms[dom] > 0 else None
    traffic = load_traffic()
    roofline = {
        "bound": "hbm", "kernel": f"two_loop_{dom}_step (k_{dom})", "achieved": dom_gbs, "peak": peak,
        "unit": "GB/s", "frac": (dom_gbs / peak) if dom_gbs else None,
        "traffic": (traffic or {}).get("dram_bytes_per_launch"),
        "peak_source": peak_src,
        "launches": int(launches[dom]), "avg_launch_ms": kms[dom] / max(1, launches[dom]),
        "algorithmic_bytes_per_launch": kbytes[dom] / max(1, launches[dom]),
    }
    kbytes["evaluate"] = 2.0 * 8.0 * n_local * launches["evaluate"]     # Rosenbrock: 1R 1W per evaluation
    surv_bytes = algorithmic_bytes_survey(n_local, launches, kbytes)
    moved = sum(kbytes.values())
    iteration = {
        # bytes the launched kernels must move (DESIGN.md §3 per-kernel passes x 8n) / wall time of the K steps
        "algorithmic_GBps": moved / 1e9 / (ms_total / 1e3),
        "frac_of_peak": moved / 1e9 / (ms_total / 1e3) / peak,
        "algorithmic_bytes_per_iteration": moved / max(1, K),
        # SURVEY.md §8(d)'s formula prices a trial at 8V (K1 + evaluate + K2); the fused trial moves 4V, so this
        # "unfused-equivalent" rate can exceed what the HBM actually carried — reported for comparison only
        "survey_formula_equivalent_GBps": surv_bytes / 1e9 / (ms_total / 1e3),
        "line_search_trials": ("probe + commit" if launches.get("probe", 0) > 0 else
                               "fused trial" if launches.get("trial_eval", 0) > 0 else "unfused (K1 + evaluate + K2)"),
        "evaluations_per_iteration": (launches["evaluate"] + launches.get("trial_eval", 0) + launches.get("probe", 0)) / max(1, K),
        "kernel_ms_timed_region": {k: round(v, 3) for k, v in kms.items() if v > 0},
        "profile_pass": {
            "note": f"{P} extra iterations after the timed region with CUDA events around every launch",
            "kernel_ms": {k: round(v, 3) for k, v in prof_all["ms"].items() if v > 0},
            "kernel_GBps": {k: round(prof_all["bytes"][k] / 1e9 / (prof_all["ms"][k] / 1e3), 1) for k in prof_all["ms"]
                            if prof_all["ms"][k] > 0 and prof_all["bytes"][k] > 0},
        },
        "host_syncs": prof["host_syncs"], "allreduces": prof["allreduces"],
        "allreduce_transport": comm.transport if comm is not None else None,
    }
    del x
    torch.cuda.empty_cache()

    # ---- end to end through the public API with HOST buffers: `e2e` ---------------------------------
    # The reference-shaped call: x is a HOST slice (src/lbfgs.rs:399), passed to the C ABI's host-buffer entry.
    # Timed: the whole call — pinned host -> device copy of x0, solver creation, a complete minimize() of W+K
    # iterations (build + line searches + two-loops), device -> host copy of the result, teardown.  The copies
    # happen once per solve, so bytes/step are 8n/(W+K) each way (+ the few scalars read back per iteration).
    iters_e2e = W + K
    xh = torch.empty(n_local, dtype=torch.float64, pin_memory=True)
    fill_x0(xh)
    builder = make_builder().with_max_iterations(iters_e2e + 1)
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
        "note": "one lbfgsb200_minimize_host_ex() call on a pinned HOST buffer: H2D of x0, solver creation, build, "
                "W+K iterations, D2H of x, teardown",
    }
    del xh

    # ---- the timed trajectory against the reference algorithm (checker only; rank 0) -------------------------
    parity = None
    if rank == 0 and not args.no_cpu_baseline and n_global % 2 == 0:
        try:
            ref = {t["niter"]: t for t in isometric_oracle_trace(n_global, m, 1 + W + K)}
            worst = 0.0
            same = True
            for (it, nc, fx, xn, gn, stp) in seen:
                t = ref.get(it)
                if t is None or t["ncall"] != nc:
                    same = False
                    break
                for a, b in ((fx, t["fx"]), (xn, t["xnorm"]), (gn, t["gnorm"]), (stp, t["step"])):
                    worst = max(worst, abs(a - b) / max(abs(b), 1e-300))
            parity = {"checker": "oracle on the isometric 2-variable image of the workload (bench.py: isometric_oracle_trace)",
                      "iterations_checked": len(seen), "evaluations_per_iteration_identical": same,
                      "max_rel_err_fx_xnorm_gnorm_step": worst}
        except Exception as e:  # the checker must never break the measurement
            parity = {"error": repr(e)}

    # ---- the reference's CPU path on this host (rank 0, N=1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sample = int(args.cpu_n)
        rate, dt, done = cpu_oracle_rate(n_sample, min(W, 2), args.cpu_steps, m=m, budget_s=25.0)
        if rate:
            cpu = {"value": rate * n_sample / N_REF, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle (C++ port of the reference, 1 thread) Rosenbrock n={n_sample}, m={m}, "
                             f"{done} iterations after {min(W, 2)} warm-up in {dt:.1f} s; it/s scaled by n_sample/1e8",
                   "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if (n_local == N_REF and m == 6) else
                       f"Rosenbrock n={n_local} f64 per GPU, m={m}, MoreThuente, device-resident evaluate",
                       "n_per_gpu": n_local, "n_global": n_global, "m": m, "linesearch": "MoreThuente",
                       "l2_policy": "inputs larger than L2 (19 vectors x 0.8 GB vs 126 MB)",
                       "ncall_per_iteration": ncalls, "final_fx": final.fx, "final_gnorm": final.gnorm},
            "roofline": roofline, "iteration": iteration, "cpu_baseline": cpu, "parity": parity, "e2e": e2e,
            "gpu_launches": gpu_launches, "clocks": clocks,
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
    ap.add_argument("--ref-n", type=float, default=5e6, help="--impl reference: sample size")
    ap.add_argument("--cpu-n", type=float, default=2e7, help="cpu_baseline sample size")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused-trial", action="store_true",
                    help="line-search trials as K1 + evaluate + K2 (three passes) instead of the fused one-pass trial")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
