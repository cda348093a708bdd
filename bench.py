#!/usr/bin/env python
"""Benchmark of the GP-surrogate hot path (BASELINE.json metric: GP posterior+grad candidate evals/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one pass of the hot path over one batch of synthetic candidates: for every candidate and all n_obj
objectives the posterior mean, std, d mean/dx, d std/dx and the fused UCB acquisition value + gradient.
Workload: N=2000 training points, d=20, n_obj=3, Matern-5/2, 2^20 candidates per GPU.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is produced.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gp_posterior_grad_candidate_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = dict(workload="c4: synthetic GP N=2000 d=20 n_obj=3, Matern-5/2, UCB value+gradient over 2^20 candidates per GPU",
                n_train=2000, n_var=20, n_obj=3, nu=5, acquisition="ucb", candidates_per_gpu=1 << 20,
                l2_policy="inputs larger than L2: every chunk of 16384 candidates streams 3.2 GB of K*, W, V, g/r slabs (126 MB L2), outputs 528 MB per step")


WORKLOADS = {
    "c4": WORKLOAD,
    # BASELINE.json configs[4]: the sharding sweep shape (8 M candidates over 8 GPUs = 2^20 per GPU, weak scaling)
    "c5": dict(workload="c5: synthetic GP N=4000 d=32 n_obj=3, Matern-5/2, UCB value+gradient over 2^20 candidates per GPU",
               n_train=4000, n_var=32, n_obj=3, nu=5, acquisition="ucb", candidates_per_gpu=1 << 20,
               l2_policy="inputs larger than L2: every chunk of 16384 candidates streams 6.4 GB of K*, W, V, g/r slabs (126 MB L2)"),
}


def make_problem(n_train, d, m, seed=0):
    """Synthetic data of SURVEY.md §8(d): X ~ U[0,1]^{N x d}, Y = sin(XW) + 0.1 X^2 |W|; fixed hyper-parameters."""
    rng = np.random.default_rng(seed)
    X = rng.random((n_train, d))
    W = rng.standard_normal((d, m))
    Y = np.sin(X @ W) + 0.1 * (X ** 2) @ np.abs(W)
    rng1 = np.random.default_rng(seed + 1)
    thetas = np.stack([np.concatenate([[0.0], rng1.uniform(-0.5, 1.0, d), [math.log(0.01)]]) for _ in range(m)])
    return X, Y, thetas


def candidates(n, d, seed):
    return np.random.default_rng(seed).random((n, d))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def blas_threads():
    """Lets the CPU baseline use every host core even under torchrun (which exports OMP_NUM_THREADS=1): returns the
    (context manager, thread count) to run the oracle under."""
    from threadpoolctl import threadpool_info, threadpool_limits
    n = os.cpu_count() or 1
    ctx = threadpool_limits(limits=n)
    used = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    return ctx, used


def cpu_port_rate(n_train, d, m, nu, budget_s, chunk=1024):
    """The CPU oracle (vectorised numpy restatement of gp.py:141-253 + ucb.py) on a bounded sample of the workload."""
    ctx, _ = blas_threads()
    with ctx:
        return _cpu_port_rate(n_train, d, m, nu, budget_s, chunk)


def _cpu_port_rate(n_train, d, m, nu, budget_s, chunk=1024):
    from oracle import gp_oracle as go
    X, Y, thetas = make_problem(n_train, d, m)
    orc = go.GPOracle(d, m, nu=nu).fit(X, Y, thetas=thetas)
    for st in orc.states:
        st.Kinv  # K^-1 formed once outside the timed region, as the reference caches gp._K_inv (gp.py:154-157)
    Xs = candidates(chunk, d, 2)
    val = orc.evaluate(Xs[:64], std=True, gradient=True, var_form="kinv", chunk=64)  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        # the reference's own formulation: one dense GEMM against the explicit inverse (gp.py:160-161, 201-205)
        val = orc.evaluate(Xs, std=True, gradient=True, var_form="kinv", chunk=512)
        go.acquisition("ucb", val, True, False, n_sample=n_train)
        done += chunk
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return done / el, done, el


def reference_literal(w, n_train=None, rows=512, loop_rows=24):
    """SURVEY.md §8(d) baselines B1 / B3, informational: the UNMODIFIED reference classes (`autooed/mobo/surrogate_model/gp.py`,
    `acquisition/ucb.py`, staged under the git-ignored baseline/_ref/ by `__graft_entry__.build()` where the reference tree is
    mounted) with sklearn's own `GaussianProcessRegressor` underneath, on the workload's training set and fixed theta
    (`optimizer=None`, so `fit` is one factorisation).  The reference can only produce d sigma / dx for ONE row per call
    (SURVEY B-1), so "std + gradient" is a row loop; the batched calls are its `std=True` (no gradient) and its
    `std=False, gradient=True` forms.  Not the `value` of the reference arm — that stays the vectorised port, the stronger
    comparator."""
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline", "_ref")
    if not os.path.isfile(os.path.join(root, "autooed", "mobo", "surrogate_model", "gp.py")):
        return {"unavailable": "baseline/_ref holds no staged reference modules (run __graft_entry__.build() where /root/reference is mounted)"}
    import warnings
    os.environ["AOED_REFERENCE_ROOT"] = root
    try:
        from oracle import ref_loader as rl
        ns = rl.load()
    except Exception as e:  # a dependency of the reference modules is missing on this box
        return {"unavailable": f"reference modules do not import here: {type(e).__name__}: {e}"[:300]}
    n_train = n_train or w["n_train"]
    d, m, nu = w["n_var"], w["n_obj"], w["nu"]
    X, Y, thetas = make_problem(n_train, d, m)
    ctx, cores = blas_threads()
    out = {"source": "unmodified autooed/mobo/surrogate_model/gp.py + acquisition/ucb.py on sklearn GaussianProcessRegressor",
           "n_train": n_train, "cores": cores}
    with ctx, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ns.GaussianProcess(rl.FakeProblem(d, m), nu=nu)
        for g, th in zip(model.gps, thetas):
            g.kernel = g.kernel.clone_with_theta(th)
            g.optimizer = None
        t0 = time.perf_counter()
        model.fit(X, Y)
        out["fit_fixed_theta_s"] = time.perf_counter() - t0
        for g in model.gps:
            g._K_inv = None  # sklearn >= 0.24 no longer defines the attribute gp.py:154 reads
        acq = ns.UpperConfidenceBound(model)
        acq.fit(X, Y)
        Xs = candidates(rows, d, 2)
        model.evaluate(Xs[:8], std=True)  # forms gp._K_inv once, as the reference caches it (gp.py:154-157)

        def rate(fn, n):
            t0 = time.perf_counter()
            fn()
            return n / (time.perf_counter() - t0)

        out["evaluate_std_batched_per_s"] = rate(lambda: model.evaluate(Xs, std=True), rows)
        out["evaluate_mean_gradient_batched_per_s"] = rate(lambda: model.evaluate(Xs, std=False, gradient=True), rows)
        out["ucb_value_gradient_row_loop_per_s"] = rate(lambda: [acq.evaluate(Xs[i:i + 1], gradient=True) for i in range(loop_rows)], loop_rows)
        g0 = model.gps[0]
        t0 = time.perf_counter()
        g0.log_marginal_likelihood(g0.kernel_.theta, eval_gradient=True)
        out["sklearn_lml_gradient_call_s"] = time.perf_counter() - t0
    out["sample"] = f"{rows} rows batched, {loop_rows} rows in the one-row loop, one LML + gradient call of one objective"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    cores = blas_threads()[1]
    rates = []
    for _ in range(args.warmup):
        cpu_port_rate(w["n_train"], w["n_var"], w["n_obj"], w["nu"], 1.0)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        rate, done, el = cpu_port_rate(w["n_train"], w["n_var"], w["n_obj"], w["nu"], args.cpu_seconds)
        rates.append(rate); t_tot += el; n_tot += done
    value = n_tot / t_tot
    sample = f"{n_tot // max(args.steps, 1)} candidates per step of the c4 workload (time-bounded {args.cpu_seconds}s per step)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(w),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    try:
        line["reference_literal"] = reference_literal(w)
    except Exception as e:  # informational: never take the arm down
        line["reference_literal"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(line), flush=True)


def dgemm_peak(torch, dev):
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / best * 1e-9


def run_ours(args):
    # exactly ONE line goes to stdout: library chatter (e.g. NCCL's version banner) is diverted to stderr until the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from autooed_b200 import GaussianProcess, UpperConfidenceBound, _lib
    from autooed_b200.problem import SimpleProblem
    import __graft_entry__ as ge
    ge.build()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    w = WORKLOAD
    N, d, m, C = w["n_train"], w["n_var"], w["n_obj"], args.candidates or w["candidates_per_gpu"]

    X, Y, thetas = make_problem(N, d, m)
    gp = GaussianProcess(SimpleProblem(d, m), nu=w["nu"], device=local)
    gp.fit_fixed(X, Y, thetas)
    acq = UpperConfidenceBound(gp)
    acq.fit(X, Y)
    Xs = candidates(C, d, 2 + rank)  # every rank owns its own shard of candidates (weak scaling, no data-path collective)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (`value`) --------------------------------------------------------------------
    dX = torch.from_numpy(Xs).to(dev)
    dA = torch.empty((C, m), dtype=torch.float64, device=dev)
    ddA = torch.empty((C, m, d), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        gp.evaluate_device(dX, C, std=True, gradient=True, acq_kind=_lib.ACQ_UCB, acq_param=[float(N)], d_A=dA, d_dA=ddA,
                           stream=stream)

    for _ in range(args.warmup):
        step_device()
    barrier()
    gp.reset_counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = gp.counters()["launches"]
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg (`e2e`): public API, host numpy in / host numpy out -----------------------------------
    # warm-up keeps results alive exactly like the timed loop does (previous result still referenced while the next call
    # runs), so that the pinned result-buffer pool reaches its steady state of two sets before timing starts
    F = dF = None
    for _ in range(max(2, args.warmup // 2 + 1)):
        F, dF, _ = acq.evaluate(Xs, dtype='normalized', gradient=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        F, dF, _ = acq.evaluate(Xs, dtype='normalized', gradient=True)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = t.tolist()

    # numerics guard: device leg and host leg agree (different chunk sizes => different partial-sum splits, not bitwise)
    same = bool(np.allclose(dA.cpu().numpy(), F, rtol=1e-11, atol=1e-12))
    del F, dF

    extras = {}
    if args.extras != "none":
        ctxt = dict(torch=torch, dist=dist, world=world, rank=rank, local=local, dev=dev, barrier=barrier, args=args)

        def guarded(name, fn, *a, **k):  # an extra leg must never cost the contract line: its failure is reported in its place
            try:
                extras.update(fn(*a, **k))
            except Exception as e:  # noqa: BLE001
                extras[name + "_error"] = f"{type(e).__name__}: {e}"[:300]

        guarded("fit_c4", extra_fit_c4, gp_model=(X, Y), **ctxt)
        if world > 1:
            guarded("sharded_eval", extra_sharded_eval, gp, dX, **ctxt)
        if world > 1:
            guarded("sharded_selection", extra_sharded_selection, **ctxt)
        guarded("c5", extra_c5, **ctxt)
        if rank == 0:
            guarded("mobo", extra_mobo)
        barrier()

    if rank == 0:
        # ---- roofline of the dominant kernel (triangular FP64 DMMA product), timed with events per launch -------
        gp.set_profiling(True)
        gp.reset_counters()
        step_device()
        torch.cuda.synchronize()
        cnt = gp.counters()
        gp.set_profiling(False)
        achieved = cnt["trmm_flops"] / (cnt["trmm_ms"] * 1e-3) * 1e-12
        peak = dgemm_peak(torch, dev)
        step_ms_profiled = cnt["trmm_ms"]
        variant = os.environ.get("AOED_TRMM", "tma")
        kname = {"tma": "trmm_tma_kernel (persistent TMA + mbarrier ring + FP64 DMMA m8n8k4 triangular product)",
                 "tile64": "trmm_kernel<.,64> (cp.async 128x64 tiles, FP64 DMMA m8n8k4 triangular product)",
                 "tile128": "trmm_kernel<.,128> (cp.async 128x128 tiles, FP64 DMMA m8n8k4 triangular product)"}.get(variant, variant)
        traffic, traffic_src = None, None
        prof = os.path.join(ROOT, "profiles", "ncu_trmm_summary.json")
        if variant == "tma" and os.path.exists(prof):  # dram bytes per launch from the committed `ncu --set full` capture
            ks = [k for k in json.load(open(prof))["kernels"] if "trmm_tma" in k["kernel"] and "dram_traffic_bytes" in k]
            if ks:
                traffic = sum(k["dram_traffic_bytes"] for k in ks) / len(ks)
                traffic_src = "profiles/ncu_trmm_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, mean of the LOWER and UPPER launch)"
        issue_peak = None
        probe = os.path.join(ROOT, "profiles", "fp64_pipe_probe_r01.json")
        if os.path.exists(probe):  # DMMA issue-rate peak measured by tools/fp64_peaks.cu on this pool's B200 (no memory traffic)
            issue_peak = json.load(open(probe)).get("dmma884_w32_tflops")
        roofline = {"bound": "tensor", "kernel": kname,
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_dmma_issue": issue_peak, "frac_of_dmma_issue_peak": (achieved / issue_peak) if issue_peak else None,
                    "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": 8.0 * m * (0.5 * N * N + 2.0 * N * min(C, 16384)),
                    "peak_source": "cuBLAS DGEMM fp64 8192^3 measured live in this run (MEASURED_PEAKS.json holds HBM and bf16 "
                                   "only; FP64 DMMA issue peak measured at 37.2 TFLOP/s, profiles/fp64_pipe_probe_r01.json)",
                    "flops_per_launch": cnt["trmm_flops"] / max(cnt["trmm_launches"], 1),
                    "avg_launch_ms": cnt["trmm_ms"] / max(cnt["trmm_launches"], 1),
                    "kernel_share_of_step": step_ms_profiled / (ms_dev / args.steps)}
        cpu_rate, cpu_done, cpu_el = cpu_port_rate(N, d, m, w["nu"], args.cpu_seconds)
        total = C * world * args.steps
        value = total / (ms_dev * 1e-3)
        e2e_value = total / (ms_e2e * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": dict(w, candidates_per_gpu=C),
                "parallelism": f"candidate-shard x{world}",
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": C * d * 8,
                        "d2h_bytes_per_step": C * m * (1 + d) * 8},
                "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": blas_threads()[1], "kind": "port",
                                 "sample": f"{cpu_done} candidates of the same workload in {cpu_el:.1f}s (oracle/gp_oracle.py, numpy BLAS threads unrestricted)"},
                "clocks": clocks, "device_host_legs_agree": same}
        if world == 1 and args.extras == "all" and w is WORKLOADS["c4"]:
            # SURVEY.md §8(d) B1 / B3 beside the port: the unmodified reference classes on sklearn (informational)
            try:
                line["cpu_baseline"]["reference_literal"] = reference_literal(w)
            except Exception as e:
                line["cpu_baseline"]["reference_literal"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        line.update(extras)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================================
# extra keys of the contract line (outside its timed region): the configurations BASELINE.json names beside c4's evaluate
# =====================================================================================================================
def _max_over_ranks(torch, dist, world, dev, vals):
    if world == 1:
        return list(vals)
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def extra_fit_c4(gp_model, torch, dist, world, rank, local, dev, barrier, args):
    """BASELINE config c4, the fit half: N=2000, d=20, m=3, 16 L-BFGS-B starts per objective (start 0 = theta0 of
    gp.py:126-132, 15 log-uniform restarts, sklearn `n_restarts_optimizer` semantics), all 48 runs in lock-step on the
    blocked factorisation.  Checked against the CPU oracle AT the optimum: same LML value, and the optimum is one of the
    oracle's function too (projected gradient of -LML within L-BFGS-B's stopping tolerances).  With WORLD_SIZE > 1 the
    same problems are then fitted with `fit_sharded` (6 problems per GPU at 8) and must give the same thetas."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.sharding import fit_sharded
    from oracle import gp_oracle as go
    X, Y = gp_model
    N, d = X.shape
    m = Y.shape[1]
    out = {}
    thetas_single = None
    if rank == 0:
        gp = GaussianProcess(SimpleProblem(d, m), nu=5, n_restarts=15, seed=3, device=local)
        Yn = (Y - Y.mean(0)) / Y.std(0)
        gp._set_train(X, Yn)
        th0 = np.tile(go.initial_theta(d), (48, 1))
        gp.log_marginal_likelihood_batch(np.arange(48) % m, th0, True)  # allocates the scratch of the blocked path
        t0 = time.perf_counter()
        gp.fit(X, Y)
        wall = time.perf_counter() - t0
        thetas_single = np.stack([g.kernel_.theta for g in gp.gps])
        lml_star = [float(g.log_marginal_likelihood_value_) for g in gp.gps]
        info = dict(gp.fit_info)
        ctx, threads = blas_threads()
        lo, hi = go.theta_bounds(d).T
        rel, pg = [], []
        with ctx:
            t0 = time.perf_counter()
            for o in range(m):
                v, g = go.lml_and_grad(thetas_single[o], gp._Xn, gp._Yn[:, o], 5)
                rel.append(abs(v - lml_star[o]) / abs(v))
                g = -g  # gradient of the minimised -LML; projected onto the box (scipy's `pg` of L-BFGS-B)
                proj = np.where((thetas_single[o] <= lo) & (g > 0), 0.0, np.where((thetas_single[o] >= hi) & (g < 0), 0.0, g))
                pg.append(float(np.abs(proj).max()))
            cpu_s = (time.perf_counter() - t0) / m
        out["fit_c4"] = {"what": "full fit, N=2000 d=20 m=3, 16 starts per objective, 48 L-BFGS-B runs in lock-step",
                         "wall_s": wall, "n_problems": info["n_problems"], "n_batches": info["n_batches"],
                         "n_lml_evals": info["n_lml_evals"], "ms_per_lml_eval": 1e3 * wall / max(info["n_lml_evals"], 1),
                         "lml_star": lml_star, "oracle_lml_rel_err_at_theta_star": rel, "oracle_lml_tol": 1e-6,
                         "oracle_projected_gradient_inf_norm": pg,
                         "relative_projected_gradient": [p / max(1.0, abs(v)) for p, v in zip(pg, lml_star)],
                         "parity_ok": bool(max(rel) <= 1e-6),
                         "cpu_port_s_per_lml_eval": cpu_s, "cpu_threads": threads,
                         "cpu_port_s_extrapolated_same_evals": cpu_s * info["n_lml_evals"]}
        del gp
    if world > 1:
        barrier()
        gp2 = GaussianProcess(SimpleProblem(d, m), nu=5, n_restarts=15, seed=3, device=local)
        t0 = time.perf_counter()
        fit_sharded(gp2, X, Y)
        torch.cuda.synchronize()
        (wall_sh,) = _max_over_ranks(torch, dist, world, dev, [time.perf_counter() - t0])
        if rank == 0:
            th2 = np.stack([g.kernel_.theta for g in gp2.gps])
            out["fit_sharded_s"] = wall_sh
            out["fit_sharded"] = {"problems_per_rank": gp2.fit_info["n_problems_local"], "world": world,
                                  "thetas_equal_single_rank_bitwise": bool(np.array_equal(th2, thetas_single)),
                                  "max_abs_theta_diff": float(np.abs(th2 - thetas_single).max()),
                                  "collectives": "broadcast_object_list(starts) + all_reduce(MIN) of 48 x (p+2) records"}
        del gp2
    return out


def extra_sharded_eval(gp, dX_local, torch, dist, world, rank, local, dev, barrier, args):
    """SURVEY.md §8(e) row 1 with the collective in the timed region: ONE fixed set of 2^20 candidates (strong scaling),
    each rank evaluates its contiguous shard on the device and the results are all-gathered over NVLink
    (`all_gather_into_tensor`, 25 MB of values + 503 MB of gradients).  Rank 0 re-evaluates one 16384-row chunk of the LAST
    rank's shard alone and compares it with the gathered rows bit for bit."""
    from autooed_b200 import _lib
    from autooed_b200.sharding import evaluate_sharded_device, shard_bounds
    N, d, m = gp._Xn.shape[0], gp.n_var, gp.n_obj
    C = 1 << 20
    Xs = candidates(C, d, 99)  # the same set on every rank
    dX = torch.from_numpy(Xs).to(dev)
    outbuf = None
    A = dA = None
    best = None
    for rep in range(3):  # first repetition warms NCCL's communicator up
        barrier()
        t0 = time.perf_counter()
        A, dA, tm = evaluate_sharded_device(gp, dX, C, _lib.ACQ_UCB, [float(N)], gradient=True, out=outbuf)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        outbuf = (A, dA) if A.shape[0] == C else None
        vals = _max_over_ranks(torch, dist, world, dev, [tm["compute_ms"], tm["gather_ms"], tm["compute_ms"] + tm["gather_ms"], 1e3 * wall])
        if rep > 0 and (best is None or vals[2] < best[2]):
            best = vals
    res = {}
    if rank == 0:
        lo, hi = shard_bounds(C, world, world - 1)
        rows = min(16384, hi - lo)
        a1 = torch.empty((rows, m), dtype=torch.float64, device=dev)
        da1 = torch.empty((rows, m, d), dtype=torch.float64, device=dev)
        gp.evaluate_device(dX[lo:lo + rows], rows, std=True, gradient=True, acq_kind=_lib.ACQ_UCB, acq_param=[float(N)], d_A=a1,
                           d_dA=da1, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        eq = bool(torch.equal(a1, A[lo:lo + rows]) and torch.equal(da1, dA[lo:lo + rows]))
        res = {"sharded_equals_single": eq, "gather_ms": best[1],
               "sharded_eval": {"what": "strong scaling: 2^20 candidates in total, shards evaluated on the device, results all-gathered",
                                "candidates_total": C, "compute_ms": best[0], "gather_ms": best[1], "total_ms": best[2],
                                "wall_ms": best[3], "gathered_bytes": C * m * (1 + d) * 8,
                                "value_evals_per_s": C / (best[2] * 1e-3), "rows_compared_bitwise": rows,
                                "collective": "2 x ncclAllGather (in place) per call"}}
    del dX, A, dA
    return res


def extra_sharded_selection(torch, dist, world, rank, local, dev, barrier, args):
    """SURVEY.md §8(e) rows 3-4 on real devices: greedy HVI batch selection with the candidates sharded (one
    (contribution, index) pair per rank and round over NCCL) and the non-dominated mask with the rows sharded; rank 0 repeats
    both on its own GPU and compares."""
    from autooed_b200.selection.hvi import hvi_select_indices
    from autooed_b200.sharding import hvi_select_sharded, pareto_mask_sharded
    from autooed_b200.utils.pareto import check_pareto
    rng = np.random.default_rng(7)  # the same data on every rank
    m, C, N, batch = 3, 1 << 17, 2000, 8
    Y = rng.random((N, m)) * 0.5 + 0.5
    Yc = rng.random((C, m)) ** 0.5
    out = {}
    for rep in range(2):
        np.random.seed(11)
        barrier()
        t0 = time.perf_counter()
        idx = hvi_select_sharded(Yc, Y, batch)
        (t_hvi,) = _max_over_ranks(torch, dist, world, dev, [time.perf_counter() - t0])
        barrier()
        t0 = time.perf_counter()
        mask = pareto_mask_sharded(Yc)
        (t_par,) = _max_over_ranks(torch, dist, world, dev, [time.perf_counter() - t0])
    if rank == 0:
        np.random.seed(11)
        t0 = time.perf_counter()
        idx1 = hvi_select_indices(Yc, Y, batch, device=local)
        t1 = time.perf_counter() - t0
        mask1 = check_pareto(Yc, device=local)
        out["sharded_selection"] = {"candidates": C, "n_obj": m, "batch": batch, "world": world,
                                    "hvi_sharded_ms": 1e3 * t_hvi, "hvi_single_gpu_ms": 1e3 * t1,
                                    "hvi_batch_equals_single": bool(list(idx) == list(idx1)),
                                    "pareto_mask_sharded_ms": 1e3 * t_par, "pareto_mask_equals_single": bool(np.array_equal(mask, mask1)),
                                    "collectives": "all_gather of one (contribution, index) pair per rank and round; all_gather of mask shards",
                                    "note": "the sharded HVI path scores shards with one host-synchronised call per round; the single-GPU "
                                            "path keeps all rounds on the device, so it is the faster one at this size"}
    return out


def extra_c5(torch, dist, world, rank, local, dev, barrier, args):
    """BASELINE config c5 per GPU: N=4000, d=32, m=3, UCB value + gradient over 2^20 candidates per GPU (8 M at 8 GPUs),
    device-resident, one timed step after a short warm-up, max over ranks."""
    from autooed_b200 import GaussianProcess, _lib
    from autooed_b200.problem import SimpleProblem
    w = WORKLOADS["c5"]
    N, d, m, C = w["n_train"], w["n_var"], w["n_obj"], w["candidates_per_gpu"]
    X, Y, thetas = make_problem(N, d, m)
    gp = GaussianProcess(SimpleProblem(d, m), nu=w["nu"], device=local)
    gp.fit_fixed(X, Y, thetas)
    dX = torch.from_numpy(candidates(C, d, 2 + rank)).to(dev)
    dA = torch.empty((C, m), dtype=torch.float64, device=dev)
    ddA = torch.empty((C, m, d), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step(rows):
        gp.evaluate_device(dX, rows, std=True, gradient=True, acq_kind=_lib.ACQ_UCB, acq_param=[float(N)], d_A=dA, d_dA=ddA,
                           stream=stream)

    step(1 << 16)
    step(1 << 16)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step(C)
    e1.record()
    barrier()
    (ms,) = _max_over_ranks(torch, dist, world, dev, [e0.elapsed_time(e1)])
    del gp, dX, dA, ddA
    torch.cuda.empty_cache()
    return {"c5_value": C * world / (ms * 1e-3),
            "c5": {"workload": w["workload"], "n_gpus": world, "ms_per_step": ms, "steps": 1, "candidates_total": C * world,
                   "unit": UNIT}}


def extra_mobo():
    """MOBO iteration wall time (the second half of BASELINE.json's metric) on c1 (ZDT1 d=6 m=2, identity, NSGA-II pop 200
    x 200 generations, batch 10) and c2 (DTLZ2 d=10 m=3, EI, pop 1000, batch 20): fit + solve + HVI select per iteration
    on the B200 path, averaged over the iterations after the first (which pays allocations)."""
    import random
    from autooed_b200 import ExpectedImprovement, GaussianProcess, HypervolumeImprovement, Identity
    from autooed_b200.mobo import MOBO
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.solver import NSGA2
    from autooed_b200.solver.nsga2 import lhs
    out = {}
    detail = {}
    for name, f, d, m, acq_cls, pop, batch, n_init, iters in (("c1", _zdt1, 6, 2, Identity, 200, 10, 20, 8),
                                                             ("c2", lambda X: _dtlz2(X, 3), 10, 3, ExpectedImprovement, 1000, 20, 30, 5)):
        np.random.seed(0)
        random.seed(0)
        prob = SimpleProblem(d, m)
        X = lhs(d, n_init)
        Y = f(X)
        sur = GaussianProcess(prob, nu=1)
        opt = MOBO(prob, sur, acq_cls(sur), NSGA2(prob, n_gen=200, pop_size=pop), HypervolumeImprovement(sur))
        times = []
        for _ in range(iters):
            t0 = time.perf_counter()
            Xn = opt.optimize(X, Y, None, batch)
            times.append(time.perf_counter() - t0)
            X, Y = np.vstack([X, Xn]), np.vstack([Y, f(Xn)])
        out[f"mobo_{name}_s"] = float(np.mean(times[2:]))  # the first two iterations load kernels / build the CUDA graph
        detail[name] = {"iterations_timed": iters - 2, "s_each": [round(t, 5) for t in times], "n_train_last": int(len(X) - batch)}
    out["mobo"] = dict(detail, what="s per MOBO iteration (fit + 200 NSGA-II generations + HVI batch selection), B200 path; the "
                                    "CPU arm with identical iterations is `bench.py --mode mobo`")
    return out


# =====================================================================================================================
# secondary measurements (not the contract line): `--mode fit` and `--mode mobo` print one JSON line per measurement
# =====================================================================================================================
def _fit_data(N, d, m, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.1 * (X ** 2) @ np.abs(rng.standard_normal((d, m))) \
        + 0.02 * rng.standard_normal((N, m))
    return X, Y


def run_fit_mode(args):
    """Hyper-parameter fit path: batched LML + gradient per device batch and full L-BFGS-B fits,
    next to the CPU oracle (sklearn's formulation restated in numpy) with all host threads."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    from oracle import gp_oracle as go
    ctx, threads = blas_threads()

    def lml_batches(N, d, m, nu, batches, reps=5, cpu_calls=2):
        X, Y = _fit_data(N, d, m)
        Yn = (Y - Y.mean(0)) / Y.std(0)
        gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
        gp._set_train(X, Yn)
        lo, hi = go.theta_bounds(d).T
        rng = np.random.default_rng(1)
        with ctx:
            t0 = time.perf_counter()
            for i in range(cpu_calls):
                go.lml_and_grad(rng.uniform(lo, hi) * 0.3, X, Yn[:, i % m], nu)
            cpu_s = (time.perf_counter() - t0) / cpu_calls
        for B in batches:
            obj = np.arange(B) % m
            th = rng.uniform(lo, hi, (B, d + 2)) * 0.3
            t_warm = time.perf_counter()  # the CPU baseline above left the GPU idle: bring its clocks back up before timing
            while time.perf_counter() - t_warm < 0.3:
                gp.log_marginal_likelihood_batch(obj, th, True)
            t0 = time.perf_counter()
            for _ in range(reps):
                gp.log_marginal_likelihood_batch(obj, th, True)
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"mode": "fit", "what": "lml+grad batch", "N": N, "d": d, "nu": nu, "problems": B,
                              "ms_per_batch": 1e3 * dt, "problems_per_s": B / dt, "cpu_port_ms_per_problem": 1e3 * cpu_s,
                              "cpu_threads": threads, "speedup_vs_cpu_serial": cpu_s * B / dt,
                              "path": "one-CTA shared-memory kernel" if N <= 232 else "batched blocked factorisation (tgemm / potrf_diag)"}), flush=True)

    def full_fit(N, d, m, nu, n_restarts):
        X, Y = _fit_data(N, d, m)
        gp = GaussianProcess(SimpleProblem(d, m), nu=nu, n_restarts=n_restarts, seed=0)
        gp.fit(X, Y)
        t0 = time.perf_counter()
        gp.fit(X, Y)
        gpu_s = time.perf_counter() - t0
        with ctx:
            t0 = time.perf_counter()
            orc = go.GPOracle(d, m, nu=nu).fit(X, Y)
            cpu_s = time.perf_counter() - t0
        print(json.dumps({"mode": "fit", "what": "full fit", "N": N, "d": d, "m": m, "nu": nu,
                          "gpu_starts_per_objective": 1 + n_restarts, "gpu_s": gpu_s,
                          "gpu_fit_info": {k: v for k, v in gp.fit_info.items() if k != "nfev"},
                          "cpu_port_s_single_start": cpu_s, "cpu_threads": threads, "cpu_nfev": orc.nfev}), flush=True)

    def hessian_calls(N, d, m, nu, C, acq_name):
        """c3-like: the ParetoDiscovery access pattern — value + gradient + Hessian of the acquisition for a small batch."""
        from autooed_b200 import ExpectedImprovement, Identity
        X, Y = _fit_data(N, d, m)
        thetas = np.stack([np.concatenate([[0.0], np.full(d, 0.3), [math.log(0.01)]]) for _ in range(m)])
        gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
        gp.fit_fixed(X, Y, thetas)
        acq = (Identity if acq_name == "identity" else ExpectedImprovement)(gp)
        acq.fit(X, Y)
        Xs = np.random.default_rng(3).random((C, d))
        acq.evaluate(Xs, gradient=True, hessian=True)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            acq.evaluate(Xs, gradient=True, hessian=True)
        gpu_s = (time.perf_counter() - t0) / reps
        orc = go.GPOracle(d, m, nu=nu).fit(X, Y, thetas=thetas)
        n_cpu = min(C, 16)
        with ctx:
            t0 = time.perf_counter()
            val = orc.evaluate(Xs[:n_cpu], std=acq_name != "identity", gradient=True, hessian=True, var_form="kinv", chunk=1)
            go.acquisition(acq_name, val, True, True, n_sample=N, y_min=Y.min(0))
            cpu_s = (time.perf_counter() - t0) / n_cpu
        print(json.dumps({"mode": "fit", "what": "acquisition value+gradient+hessian call", "N": N, "d": d, "m": m, "nu": nu,
                          "acquisition": acq_name, "rows_per_call": C, "gpu_ms_per_call": 1e3 * gpu_s,
                          "gpu_rows_per_s": C / gpu_s, "cpu_port_ms_per_row": 1e3 * cpu_s, "cpu_threads": threads,
                          "speedup": cpu_s * C / gpu_s}), flush=True)

    hessian_calls(200, 30, 2, 5, 1, "identity")
    hessian_calls(200, 30, 2, 5, 100, "identity")
    hessian_calls(200, 30, 2, 5, 100, "ei")
    hessian_calls(2000, 20, 3, 5, 100, "ei")
    lml_batches(200, 6, 2, 5, [1, 2, 16, 148, 592])
    lml_batches(500, 10, 3, 5, [3, 48])
    lml_batches(2000, 20, 3, 5, [3, 48], reps=2, cpu_calls=1)
    full_fit(200, 6, 2, 5, 0)
    full_fit(200, 6, 2, 5, 15)
    full_fit(500, 10, 3, 5, 0)


class _OracleSurrogate:
    """CPU stand-in with the plugin surface the solver / selection need (test infrastructure: oracle/gp_oracle.py)."""

    def __init__(self, problem, nu):
        from autooed_b200.problem import ContinuousTransformation
        from autooed_b200.utils.normalization import StandardNormalization
        self.problem, self.nu = problem, nu
        self.n_var, self.n_obj = problem.n_var, problem.n_obj
        self.transformation = ContinuousTransformation()
        self.normalization = StandardNormalization(np.array([problem.xl, problem.xu]))
        self.fitted = False

    def fit(self, X, Y):
        from oracle import gp_oracle as go
        self.orc = go.GPOracle(self.n_var, self.n_obj, self.problem.xl, self.problem.xu, self.nu).fit(X, Y)
        self.fitted = True


class _OracleAcquisition:
    def __init__(self, surrogate, kind):
        self.s, self.kind, self.fitted = surrogate, kind, False

    def fit(self, X, Y):
        self.n_sample, self.y_min, self.fitted = len(X), np.min(Y, axis=0), True

    def evaluate(self, X, dtype='continuous', gradient=False, hessian=False):
        from oracle import gp_oracle as go
        need_std = self.kind != "identity"
        val = self.s.orc.evaluate(np.atleast_2d(X), std=need_std, gradient=gradient, hessian=hessian, var_form="kinv")
        return go.acquisition(self.kind, val, gradient, hessian, n_sample=self.n_sample, y_min=self.y_min)


class _OracleTS:
    """Thompson-sampling acquisition on the CPU (oracle/ts_oracle.py restates acquisition/ts.py)."""

    def __init__(self, surrogate):
        self.s, self.fitted = surrogate, False

    def fit(self, X, Y):
        from oracle import ts_oracle as to
        orc = self.s.orc
        Xn, Yn = orc.norm.do_x(np.asarray(X, float)), orc.norm.do_y(np.asarray(Y, float))
        self.sample = to.ts_fit(Xn, Yn, [st.theta for st in orc.states], self.s.nu)
        self.fitted = True

    def evaluate(self, X, dtype='continuous', gradient=False, hessian=False):
        from oracle import ts_oracle as to
        orc = self.s.orc
        return to.ts_evaluate(orc.norm.do_x(np.atleast_2d(X)), *self.sample, orc.norm.y_mean, orc.norm.y_scale, gradient, hessian)


class _OracleHVI:
    def select(self, X_candidate, Y_candidate, X, Y, batch_size):
        from oracle import moo_oracle as mo
        return X_candidate[mo.hvi_select(Y_candidate, Y, batch_size)]


def _zdt1(X):
    f1 = X[:, 0]
    g = 1 + 9.0 / (X.shape[1] - 1) * X[:, 1:].sum(1)
    return np.stack([f1, g * (1 - np.sqrt(f1 / g))], 1)


def _dtlz2(X, m=3):
    g = ((X[:, m - 1:] - 0.5) ** 2).sum(1)
    F = []
    for i in range(m):
        f = 1 + g
        f = f * np.prod(np.cos(X[:, :m - 1 - i] * np.pi / 2.0), axis=1)
        if i > 0:
            f = f * np.sin(X[:, m - 1 - i] * np.pi / 2.0)
        F.append(f)
    return np.stack(F, 1)


def run_mobo_mode(args):
    """MOBO iteration wall time: fit + acquisition.fit + NSGA-II solve + HVI select per iteration on the
    BASELINE configs c1 (ZDT1 d=6 m=2, identity acquisition, pop 200 x 200 generations, batch 10) and c2 (DTLZ2 d=10 m=3,
    EI, pop 1000, batch 20), plus TSEMO (the reference's CLI default: Thompson sampling + NSGA-II + HVI) on the c1 problem:
    B200 path vs the CPU port (host NSGA-II with the same operators, same seeds)."""
    import random
    from autooed_b200 import (ExpectedImprovement, GaussianProcess, HypervolumeImprovement, Identity, ThompsonSampling)
    from autooed_b200.mobo import MOBO
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.solver import NSGA2
    from autooed_b200.solver.nsga2 import lhs
    from autooed_b200.utils.pareto import calc_hypervolume
    from oracle import moo_oracle as mo
    ctx, threads = blas_threads()
    # both arms run the SAME number of iterations from the same seed (the training set grows by `batch` per iteration, so
    # iteration counts must match for the per-iteration times to be comparable); --mobo-iters bounds the CPU arm's minutes
    it1, it2 = min(args.mobo_iters, 6), max(2, min(args.mobo_iters // 3, 3))
    configs = [dict(name="tsemo-default", f=_zdt1, d=6, m=2, nu=1, acq="ts", pop=200, gens=200, batch=10, n_init=20,
                    iters=it1, cpu_iters=it1),
               dict(name="c1", f=_zdt1, d=6, m=2, nu=1, acq="identity", pop=200, gens=200, batch=10, n_init=20,
                    iters=it1, cpu_iters=it1),
               dict(name="c2", f=lambda X: _dtlz2(X, 3), d=10, m=3, nu=1, acq="ei", pop=1000, gens=200, batch=20, n_init=30,
                    iters=it2, cpu_iters=it2)]
    for c in configs:
        prob = SimpleProblem(c["d"], c["m"])
        for arm in ("b200", "cpu_port"):
            np.random.seed(0)
            random.seed(0)
            X = lhs(c["d"], c["n_init"])
            Y = c["f"](X)
            if arm == "b200":
                sur = GaussianProcess(prob, nu=c["nu"])
                acq = {"identity": Identity, "ei": ExpectedImprovement, "ts": ThompsonSampling}[c["acq"]](sur)
                opt = MOBO(prob, sur, acq, NSGA2(prob, n_gen=c["gens"], pop_size=c["pop"]), HypervolumeImprovement(sur))
                n_it = c["iters"]
            else:
                sur = _OracleSurrogate(prob, c["nu"])
                acq = _OracleTS(sur) if c["acq"] == "ts" else _OracleAcquisition(sur, c["acq"])
                opt = MOBO(prob, sur, acq, NSGA2(prob, n_gen=c["gens"], pop_size=c["pop"], rank_fn=mo.pareto_rank), _OracleHVI())
                n_it = c["cpu_iters"]
            # where the iteration goes: surrogate fit, acquisition calls + non-dominated sorting inside the solver, selection
            split = {"fit": 0.0, "solve": 0.0, "acq_evaluate": 0.0, "rank": 0.0, "select": 0.0}

            def timed(obj, name, key):
                fn = getattr(obj, name)

                def wrapper(*a, **k):
                    t0 = time.perf_counter()
                    try:
                        return fn(*a, **k)
                    finally:
                        split[key] += time.perf_counter() - t0
                setattr(obj, name, wrapper)

            timed(sur, "fit", "fit")
            timed(opt.solver, "solve", "solve")  # contains acq_evaluate and rank when the generation loop runs on the host
            timed(acq, "evaluate", "acq_evaluate")
            timed(opt.solver.algo, "rank_fn", "rank")
            timed(opt.selection, "select", "select")
            times = []
            for it in range(n_it):
                t0 = time.perf_counter()
                if arm == "cpu_port":
                    with ctx:
                        Xn = opt.optimize(X, Y, None, c["batch"])
                else:
                    Xn = opt.optimize(X, Y, None, c["batch"])
                times.append(time.perf_counter() - t0)
                X, Y = np.vstack([X, Xn]), np.vstack([Y, c["f"](Xn)])
            ref = np.max(c["f"](lhs(c["d"], 200)), axis=0) * 1.1
            hv = mo.hypervolume(Y, ref) if arm == "cpu_port" else calc_hypervolume(Y, ref)
            print(json.dumps({"mode": "mobo", "config": c["name"], "arm": arm, "acquisition": c["acq"], "pop": c["pop"],
                              "generations": c["gens"], "batch": c["batch"], "iterations": n_it, "n_final": len(X),
                              "solver": "device-resident NSGA-II" if (arm == "b200" and opt.solver._device_eligible(len(X))) else "host NSGA-II",
                              "s_per_iteration": float(np.mean(times)), "s_first": times[0], "s_last": times[-1],
                              # the first iteration pays context / module loading / graph instantiation on the device arm
                              "s_per_iteration_steady": float(np.mean(times[1:])) if len(times) > 1 else times[0],
                              "s_per_iteration_split": {k: v / n_it for k, v in split.items()},
                              # non-dominated sorting is 75-90 % of the numpy arm (O(n^2 m) per generation on the host; the
                              # device arm does it inside its generation loop): the iteration without it, for a ratio that
                              # says something about the surrogate path rather than about that one function
                              "s_per_iteration_excluding_rank": float(np.mean(times)) - split["rank"] / n_it,
                              "hypervolume_of_evaluated_set": hv, "cpu_threads": threads if arm == "cpu_port" else None}),
                  flush=True)


def run_discovery_mode(args):
    """BASELINE config c3 shape: the ParetoDiscovery inner loop on 100 samples — local optimisation,
    KKT directions from the Hessians, first-order expansion — batched B200 path vs the per-sample CPU port
    (oracle/discovery_oracle.py on the GP oracle), same model, same seed."""
    from autooed_b200 import GaussianProcess, Identity
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.solver.pareto_discovery import pareto_discover
    from autooed_b200.surrogate_problem import SurrogateProblem
    from oracle import discovery_oracle as do
    from oracle import gp_oracle as go
    ctx, threads = blas_threads()
    d, m, N, pop, n_grid = 10, 2, 100, 100, 10
    rng = np.random.default_rng(0)
    X = rng.random((N, d))
    f1 = X[:, 0]
    gz = 1 + 9.0 / (d - 1) * X[:, 1:].sum(1)
    Y = np.stack([f1, gz * (1 - np.sqrt(f1 / gz) - f1 / gz * np.sin(10 * np.pi * f1))], 1)  # ZDT3
    prob = SimpleProblem(d, m)
    gp = GaussianProcess(prob, nu=5)
    gp.fit(X, Y)
    acq = Identity(gp)
    acq.fit(X, Y)
    sp = SurrogateProblem(prob, acq)
    calls = {"n": 0, "rows": 0}

    def eval_func(x, return_values_of):
        calls["n"] += 1
        calls["rows"] += len(np.atleast_2d(x))
        return sp.evaluate(x, return_values_of=return_values_of)

    # ---- BASELINE config c3 end to end: the ParetoDiscovery solver (pop 100 x 10 generations) proposing a batch of 10 ----
    from autooed_b200 import init_solver
    solver = init_solver('discovery', {'n_gen': 10, 'pop_size': pop}, prob)
    np.random.seed(1)
    t0 = time.perf_counter()
    Xc, Yc = solver.solve(X, Y, 10, acq)
    dt = time.perf_counter() - t0
    algo = solver.algo
    print(json.dumps({"mode": "discovery", "arm": "b200", "what": "c3 end to end: ParetoDiscovery solver, ZDT3 d=10, GP nu=5/2 + identity, "
                      "pop 100 x 10 generations, batch 10 (buffer + stochastic sampling + batched inner loop + family proposal)",
                      "s_total": dt, "s_per_generation": dt / 10, "surrogate_rows_evaluated": int(solver.n_eval),
                      "buffer_samples": int(algo.buffer.sample_count), "families": int(len(set(int(l) for l in algo.fam_lbls))),
                      "approx_front_size": int(len(algo.approx_front)), "proposed": [int(Xc.shape[0]), int(Xc.shape[1])],
                      "sparse_approximation": "aoed_graph_cut (own alpha-expansion; pygco absent)"}), flush=True)

    xs = rng.random((pop, d))
    bounds = np.stack([np.zeros(d), np.ones(d)])
    origin = Y.min(axis=0)
    times = []
    for rep in range(4):
        np.random.seed(5)
        calls["n"] = calls["rows"] = 0
        t0 = time.perf_counter()
        out = pareto_discover(xs, eval_func, bounds, None, 0.3, origin, 1e-2, n_grid)
        times.append(time.perf_counter() - t0)
    line = {"mode": "discovery", "arm": "b200", "what": "ParetoDiscovery inner loop, one generation", "samples": pop, "N": N,
            "d": d, "m": m, "n_grid_sample": n_grid, "s_per_generation": min(times[1:]), "s_first": times[0],
            "device_calls": calls["n"], "rows_per_call": calls["rows"] / max(calls["n"], 1), "patch_rows": int(len(out[0]))}
    print(json.dumps(line), flush=True)
    thetas = np.stack([g.kernel_.theta for g in gp.gps])
    orc = go.GPOracle(d, m, nu=5).fit(X, Y, thetas=thetas)
    n_cpu = 20
    with ctx:
        np.random.seed(5)
        t0 = time.perf_counter()
        ref = do.pareto_discover(xs[:n_cpu], do.oracle_eval_func(orc), bounds, 0.3, origin, 1e-2, n_grid)
        dt = time.perf_counter() - t0
    line = {"mode": "discovery", "arm": "cpu_port", "what": "the reference's per-sample loop on the GP oracle", "samples": n_cpu,
            "s_measured": dt, "s_per_generation_extrapolated": dt * pop / n_cpu, "cpu_threads": threads}
    # d > m: the minimiser of |F(x) - z| is a (d - m)-dimensional set, so the two arms' stopping POINTS may differ while
    # the distance they reach agrees; compare that (each patch starts with its locally optimised sample)
    from autooed_b200.solver.pareto_discovery import reference_point
    np.random.seed(5)
    small = pareto_discover(xs[:n_cpu], eval_func, bounds, None, 0.3, origin, 1e-2, n_grid)  # the same 20 samples on the GPU
    x_gpu = small[0][np.r_[0, np.flatnonzero(np.diff(small[1])) + 1]]
    ys = sp.evaluate(xs[:n_cpu], return_values_of=['F'])
    zs = np.array([reference_point(y, y - small[3], 0.3) for y in ys])
    reach = lambda xo: np.linalg.norm(sp.evaluate(xo, return_values_of=['F']) - zs, axis=1)
    r_gpu, r_cpu = reach(x_gpu), reach(ref["x_opt"])
    line["max_diff_of_reached_distance"] = float(np.abs(r_gpu - r_cpu).max())
    line["mean_reached_distance"] = {"b200": float(r_gpu.mean()), "cpu_port": float(r_cpu.mean())}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU per step (default 2^20)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sample budget per step")
    ap.add_argument("--mode", default="evals", choices=["evals", "fit", "mobo", "discovery"],
                    help="evals = the contract line (default); fit / mobo = secondary measurements, one JSON line each")
    ap.add_argument("--mobo-iters", type=int, default=20)
    ap.add_argument("--extras", default="all", choices=["all", "none"],
                    help="all = also run the full c4 fit, the c5 shape, the sharded (gathered) legs at N > 1 and the MOBO "
                         "iterations, reported as extra keys of the one JSON line; none = the contract legs only")
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS),
                    help="c4 = the configuration the metric is quoted on (default); c5 = the N=4000, d=32 sharding-sweep shape")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.workload]
    if args.mode == "fit":
        return run_fit_mode(args)
    if args.mode == "mobo":
        return run_mobo_mode(args)
    if args.mode == "discovery":
        return run_discovery_mode(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
