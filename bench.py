#!/usr/bin/env python
"""bench.py -- agent-days/s of the daily agent-stepping path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is ONE full simulation: the C2 workload of BASELINE.json/SURVEY.md section 8 -- n = 10 M agents
(synthetic Lombardy-like population), T = 60 days, Italy contact matrix and severity tables,
5 initial cases, lockdown on day 46 (/10), tracing never switching on -- i.e. n*(T-1) = 5.9e8
agent-days.  With N > 1 (launched under torchrun, one rank per GPU) every rank runs its own seed of
the same workload (seed ensembles need no communication): weak scaling.

value      device time of the whole run (state init + day loop + history), CUDA events on the launching stream,
           population and tables already resident in HBM; max over ranks.
e2e        the same through the C ABI with HOST buffers: per-run inputs (Home_real, initial cases,
           seed) copied from pinned host memory, the [T,9,101] tallies and counters read back, wall clock.
roofline   the algorithmic bytes of SURVEY.md section 8d (sum_t 2n + 16 A(t-1) + 64 I(t)) against the measured HBM
           peak over the WHOLE device step (init + day-loop kernels + seir_materialize_kernel): the 2n term is
           paid by the materialise kernel, the 16 A + 64 I terms by the day loop, so neither kernel alone owns the
           figure; the per-kernel split is in roofline.kernels.
config     further legs, each with its own numbers: c1, c4_like, c5_like (replica ensembles), c3_single_gpu, the
           drop-in seir_individual.run_model call, the stress variant, tracing as the reference runs as compiled;
           with N > 1 also sharded_c3 (ONE population over the N GPUs, peer-memory exchange) with a bit-exact
           sharded-vs-single-GPU assertion.
cpu_baseline  the CPU oracle (oracle/oracle_seir.c, provider "mt": the reference's algorithm and
           Mersenne-Twister stream, bit-identical to the reference) on one host core, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_italy_default.npz")
METRIC = "agent_days_per_sec"
UNIT = "agent-days/s"


def workload_params(n, T=60):
    """C2 parameters: run_simulation.py:52-170 (Italy), as frozen in the golden fixture."""
    fx = np.load(GOLDEN)
    p = dict(zip(fx["param_keys"].tolist(), [float(v) for v in fx["param_vals"]]))
    p["n"] = float(n)
    p["T"] = float(T)
    p["initial_infected_fraction"] = 5.0 / n          # run_simulation.py:66
    p["t_tracing_start"] = 1.0e6                      # tier-1 semantics: tracing never switches on
    tabs = {k: fx[k] for k in ("contact_matrix", "p_mild_severe", "p_severe_critical", "p_critical_death",
                               "mean_time_to_isolate_factor", "lockdown_factor_age")}
    tabs["p_infect_household"] = float(fx["p_infect_household_scalar"])
    return p, tabs


def algorithmic_bytes(tal, n):
    """B_alg = sum_{t=1}^{T-1} 2 n + 16 A(t-1) + 64 I(t), from the run's own tallies."""
    active = tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2))          # E + Mild + Severe + Critical per day
    S = tal[:, 0, :].sum(axis=1)
    new_inf = S[:-1] - S[1:]
    T = tal.shape[0]
    return float(2.0 * n * (T - 1) + 16.0 * active[:-1].sum() + 64.0 * new_inf.sum()), int(active[:-1].sum()), int(new_inf.sum())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  nvidia-smi needs ~0.1 s to
    deliver its first row, longer than a whole timed region here, so the sampler is started before the warm-up
    steps (every 20 ms) and finish() keeps the rows stamped inside [start, end] of the timed region (and, when the
    region is shorter than the sampling period, the rows right before and after it)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.idx = gpu_index
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def finish(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.mark_end()
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [r for (ts, r) in self.rows if t0 <= ts <= self.t1]
        if len(inside) < 3:  # a region shorter than a few sampling periods: add the neighbours
            before = [r for (ts, r) in self.rows if ts < t0][-2:]
            after = [r for (ts, r) in self.rows if ts > self.t1][:2]
            inside = before + inside + after
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_all": len(self.rows), "reasons": sorted(reasons)}


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
        return float(d[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


TAL_ROWS_DEVICE = 11  # uint32 rows per day and age the device keeps (csrc/seir_kernels.cuh TAL_ROWS)


def make_population(n, country="Italy", seed=1):
    """the repo's own census-driven samplers (household types, World_Age_2019 ages,
    comorbidity prevalences of sample_comorbidities.py), vectorised; seeded explicitly."""
    from covid19_demography_b200 import population
    return population.census_population(n, seed=seed, country=country)


# --------------------------------------------------------------------------- CPU legs (oracle)
_W = {}


def _cpu_worker_init(n_sample, T):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle  # test/bench infrastructure: the reference's algorithm restated in C
    p, tabs = workload_params(n_sample, T)
    hh, age, diab, hyp, groups = make_population(n_sample)
    _W.update(oracle=oracle, p=p, tabs=tabs, pop=(hh, age, groups, diab, hyp), n=n_sample, T=T)
    oracle.lib()


def _cpu_worker_run(seed):
    o, p, tabs = _W["oracle"], _W["p"], _W["tabs"]
    hh, age, groups, diab, hyp = _W["pop"]
    n = _W["n"]
    t0 = time.perf_counter()
    o.run_model(seed, hh, age, groups, diab, hyp, tabs["contact_matrix"], tabs["p_mild_severe"],
                tabs["p_severe_critical"], tabs["p_critical_death"], tabs["mean_time_to_isolate_factor"],
                tabs["lockdown_factor_age"], np.full(n, tabs["p_infect_household"]), np.zeros(101), p,
                provider="mt", tracing_enabled=1)
    return time.perf_counter() - t0


def _mem_available():
    try:
        return int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        return 16 << 30


def _bytes_per_run(n, T):
    return int(n * (9 * T + 400 + 200 + 6 * 8 + 60) * 1.15)


def cpu_baseline_single(n_sample, T):
    if n_sample <= 0:  # full C2 size when the host has the memory (about 14 GB), else a 2 M-agent sample
        n_sample = 10_000_000 if _mem_available() > 2 * _bytes_per_run(10_000_000, T) else 2_000_000
    _cpu_worker_init(n_sample, T)
    dt = _cpu_worker_run(1)
    return {"value": n_sample * (T - 1) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/oracle_seir.c provider=mt (reference algorithm + MT19937 stream), one run of "
                      "n=%d agents, T=%d, C2 parameters, %.1f s" % (n_sample, T, dt)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU algorithm on all the host cores the memory allows (independent
    seeds in separate processes: the reference's SLURM-array mode, job.parameter_sweep.sh) on the SAME workload as
    the GPU arm: the C2 population of 10 M agents, one full 60-day run per worker and step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    T = 60
    n_sample = args.cpu_sample if args.cpu_sample > 0 else args.n
    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, int(_mem_available() * 0.7 / _bytes_per_run(n_sample, T)), args.cpu_workers or 64))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_worker_init, initargs=(n_sample, T)) as pool:
        # (the pool initializer already touched the library and built the population; one untimed round pages the
        #  output buffers in -- a CPU run has no further warm-up state, so more untimed rounds would only cost minutes)
        for w in range(min(args.warmup, 1)):
            pool.map(_cpu_worker_run, range(workers))
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_worker_run, [1000 + k * workers + j for j in range(workers)])
        dt = time.perf_counter() - t0
    value = workers * n_sample * (T - 1) * args.steps / dt
    sample = ("oracle port of the reference (MT19937 stream, bit-identical to the reference's run_model); each step = "
              "%d concurrent single-threaded runs of n=%d agents, T=%d, C2 parameters (%d host cores, memory allows "
              "%d concurrent runs)" % (workers, n_sample, T, cores, workers))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": C2_WORKLOAD % (n_sample, T, workers), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


C2_WORKLOAD = ("C2 Lombardy 10 M with lockdown: %d agents (census-driven household sampler), T=%d, lockdown day 46 (/10), "
               "5 initial cases, %d concurrent seed(s), tracing never switches on")


# --------------------------------------------------------------------------- GPU arm
def _engine(engine, tables, p, tabs, pop, R, device, mode="native", launch="persistent", history="eager"):
    hh, age, diab, hyp, groups = pop
    n = len(age)
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=(mode == "replay"))
    return engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb,
                         n_replicas=R, device=device, mode=mode, launch=launch, history=history)


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from covid19_demography_b200 import engine, tables

    sampler = ClockSampler(local)  # (nvidia-smi takes a few hundred ms to deliver rows: start it well before the timed region)
    sampler.start()
    n, T = args.n, 60
    p, tabs = workload_params(n, T)
    pop = make_population(n)
    t0 = time.perf_counter()
    eng = _engine(engine, tables, p, tabs, pop, args.replicas, local, args.mode, args.launch)
    create_s = time.perf_counter() - t0
    R = args.replicas
    n_init = int(p["initial_infected_fraction"] * n)
    home = engine.pin(np.zeros(n, dtype=np.uint8))  # fraction_stay_home = 0 in C2 (run_simulation.py:53-54)

    # Weak scaling means the SAME work per GPU at every N.  A C2 run costs 1.1-2.2 ms depending on how large its epidemic
    # grows, so ranks that drew their own seeds would be measured by their luck (the line is the maximum over ranks).  The
    # K timed steps therefore use the same K (seed, initial cases) pairs on every rank, rank r starting r pairs further on:
    # every GPU runs a different trajectory at any moment and the same K trajectories over the timed region.
    def inputs(step):
        slot = 100000 - step if step < 0 else (step + rank) % max(1, args.steps)  # (negative: warm-up steps)
        rng = np.random.default_rng(1000 + slot)
        seeds = [slot * 64 + r for r in range(R)]
        init = [rng.choice(n, size=n_init, replace=False).astype(np.int64) for _ in range(R)]
        return seeds, [home] * R, init

    def sync_all():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    flush = None
    if args.flush_l2:
        import torch
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local)

    for w in range(args.warmup):
        eng.run(*inputs(-1 - w))
    # the K timed steps: the SAME seeds and initial cases for the device-timed pass and for the end-to-end pass (a
    # run costs 1.1-2.0 ms depending on how large its epidemic grows, so different draws would not be comparable)
    ins = [inputs(k) for k in range(args.steps)]
    for _, _, init in ins:
        for a in init:
            engine.pin(a)
    # ---- device-timed region
    sync_all()
    sampler.mark_start()
    dev_ms = loop_ms = mat_sum = init_sum = 0.0
    launches = 0
    balg = a_sum = i_sum = 0.0
    wall0 = time.perf_counter()
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()
            import torch
            torch.cuda.synchronize()
        eng.run(*ins[k])
        i_ms, l_ms, m_ms = eng.last_run_ms3()
        loop_ms += l_ms
        mat_sum += m_ms
        init_sum += i_ms
        dev_ms += i_ms + l_ms + m_ms
        launches += eng.counters()["launches"]
    sync_all()
    sampler.mark_end()
    wall = time.perf_counter() - wall0
    clocks = sampler.finish()
    # ---- end to end through the C ABI with host buffers (wall clock): per-run inputs from pinned host memory, the
    # tallies and counters of every run back on the host; the algorithmic bytes of every step come from those tallies
    tal_out = [[np.zeros((T, 9, 101), dtype=np.int64) for _ in range(R)] for _ in range(args.steps)]  # host result buffers of the K steps
    sync_all()
    e0 = time.perf_counter()
    tals = []
    for k in range(args.steps):
        eng.run(*ins[k])
        tals.append([eng.tallies(r, out=tal_out[k][r]) for r in range(R)])
    sync_all()
    e2e_s = time.perf_counter() - e0
    for step in tals:
        for tal in step:
            b_, a_, i_ = algorithmic_bytes(tal, n)
            balg += b_; a_sum += a_; i_sum += i_
    home_used = 1 <= p["t_stayhome_start"] < T  # (C2: stay-home never starts, Home_real is never read nor copied)
    h2d = R * ((n if home_used else 0) + 8 * n_init + 8)
    d2h = R * (T * 9 * 101 * 8 + 32)  # the run's [T, 9, 101] int64 counts (accumulated on the device) + the run counters, to pinned host memory

    agent_days_step = float(n) * (T - 1) * R
    local_ms = {"dev": dev_ms, "loop": loop_ms, "mat": mat_sum, "init": init_sum}  # the roofline line describes rank 0's own kernels
    if dist is not None:
        import torch
        tt = torch.tensor([dev_ms, e2e_s, loop_ms], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, loop_ms = [float(x) for x in tt.tolist()]
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda:%d" % local)
        dist.all_reduce(ll)
        launches = int(ll.item())
    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        K = args.steps
        step_s = local_ms["dev"] / K / 1e3
        achieved = balg / K / step_s / 1e9
        day_bytes = (16.0 * a_sum + 64.0 * i_sum) / K
        row_bytes = 2.0 * n * (T - 1) * R
        line = {
            "metric": METRIC, "value": world * agent_days_step * K / (dev_ms / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": C2_WORKLOAD % (n, T, R * world),
                       "mode": args.mode, "launch": args.launch, "replicas_per_gpu": R,
                       "seeds": "the same K trajectories on every rank, rank r starting r steps further on (equal work per GPU)",
                       "l2": ("explicit 256 MiB flush between steps" if args.flush_l2 else
                              "per-run working set %.2f GB (history %d x %d B + state) > 126 MB L2; state re-initialised every step"
                              % ((T * n + 110 * n) / 1e9, T, n)),
                       "day_loop_ms_per_step": loop_ms / K, "engine_create_s": create_s,
                       "sim_runs_per_hour_device": 3600.0 * world * R / (dev_ms / K / 1e3),
                       "sim_runs_per_hour_e2e": 3600.0 * world * R / (e2e_s / K),
                       "wall_s_timed_region": wall},
            "clocks": clocks,
            "e2e": {"value": world * agent_days_step * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / K * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": args.traffic if args.traffic is not None else measured_traffic("step"),
                         "kernel": "device step = seir_init_* + day loop (seir_sparse_kernel + seir_dense_solo_kernel in turn) + seir_materialize_kernel"
                                   if args.launch == "persistent" else "seir_day_kernel x T + seir_materialize_kernel",
                         "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": balg / K, "active_agent_days": a_sum / K, "infections": i_sum / K,
                         "note": "SURVEY.md 8d bytes (2n + 16 A + 64 I per day) over the WHOLE device step: the 2n row term is "
                                 "paid once by seir_materialize_kernel as a streaming T x n byte write, the 16 A + 64 I terms "
                                 "by the event-driven day loop, which is latency-bound (L2-resident working set)",
                         "kernels": {
                             "day_loop": {"kernels": "seir_sparse_kernel + seir_dense_solo_kernel in turn (single replica); seir_persistent_kernel for replica ensembles, tracing and sharded runs", "ms": local_ms["loop"] / K, "algorithmic_bytes": day_bytes,
                                                        "GBps": day_bytes / (local_ms["loop"] / K / 1e3) / 1e9,
                                                        "bound": "latency (one warp's dependent chain per sparse day, five staged CTA phases per dense day, one grid barrier per day)",
                                                        "frac_if_charged_all_bytes": balg / K / (local_ms["loop"] / K / 1e3) / 1e9 / peak,
                                                        "dram_bytes_per_launch_ncu": measured_traffic("seir_persistent_kernel")},
                             "seir_materialize_kernel": {"ms": local_ms["mat"] / K, "algorithmic_bytes": row_bytes,
                                                         "bytes_moved": R * (T * n + n / 4.0) + 32.0 * i_sum / K,
                                                         "GBps": (R * (T * n + n / 4.0) + 32.0 * i_sum / K) / (local_ms["mat"] / K / 1e3) / 1e9,
                                                         "frac": (R * (T * n + n / 4.0) + 32.0 * i_sum / K) / (local_ms["mat"] / K / 1e3) / 1e9 / peak,
                                                         "bound": "hbm write"},
                             "seir_init_state_kernel + seir_init_infected_kernel": {"ms": local_ms["init"] / K}}},
        }
    eng.close()
    del eng
    extra = {}
    if world == 1 and rank == 0 and not args.quick:
        legs = [("as_reference_tracing", tracing_leg), ("stress_variant", stress_leg), ("c1", c1_leg),
                ("c4_like", c4_leg), ("c5_like", c5_leg), ("dropin_run_model", dropin_leg), ("c3_single_gpu", c3_single_leg)]
        for name, fn in legs:
            if name in args.skip:
                continue
            t0 = time.perf_counter()
            try:
                extra[name] = fn(args, pop if name in ("as_reference_tracing", "stress_variant", "dropin_run_model") else None, local)
            except Exception as e:  # a leg must not cost the headline line
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            extra[name]["leg_wall_s"] = time.perf_counter() - t0
    if world > 1 and not args.quick and "sharded_c3" not in args.skip:
        try:
            sh = sharded_c3_leg(args, dist, rank, world, local)
        except Exception as e:
            sh = {"error": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            extra["sharded_c3"] = sh
    if rank == 0:
        line["config"].update(extra)
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample, T)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def c3_params(n, n_initial=5):
    """C3 parameters: the China block of run_simulation.py (:35-39, :62, :72, :94) as frozen in the golden fixture:
    T = 127, lockdown on day 69 (/50), China contact matrix; tracing never switches on."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_china_default.npz"))
    p = dict(zip(fx["param_keys"].tolist(), [float(v) for v in fx["param_vals"]]))
    p["n"] = float(n)
    p["initial_infected_fraction"] = n_initial / n
    p["t_tracing_start"] = 1.0e6
    tabs = {k: fx[k] for k in ("contact_matrix", "p_mild_severe", "p_severe_critical", "p_critical_death",
                               "mean_time_to_isolate_factor", "lockdown_factor_age")}
    tabs["p_infect_household"] = float(fx["p_infect_household_scalar"])
    return p, tabs


def _timed_runs(eng, runs, peak, n, T, R=1):
    """Warm-up run + timed runs of one engine: device ms per run, end-to-end (wall, tallies read back) ms per run, and
    the roofline figure of the leg from its own tallies.  `runs` = list of (seeds, home, init)."""
    go = (lambda r: eng.run(**r)) if isinstance(runs[0], dict) else (lambda r: eng.run(*r))
    go(runs[0])
    dev, balg, a_sum, i_sum = [], 0.0, 0, 0
    e0 = time.perf_counter()
    for r in runs[1:]:
        go(r)
        dev.append(sum(eng.last_run_ms3()))
        for k in range(R):
            b_, a_, i_ = algorithmic_bytes(eng.tallies(k), n)
            balg += b_; a_sum += a_; i_sum += i_
    e2e = (time.perf_counter() - e0) / len(dev)
    ms = float(np.mean(dev))
    K = len(dev)
    return {"ms_per_launch": ms, "ms_per_run": ms / R, "e2e_ms_per_run": e2e * 1e3 / R,
            "agent_days_per_sec": float(n) * (T - 1) * R / (ms / 1e3), "sim_runs_per_hour_device": 3600.0 * R / (ms / 1e3),
            "sim_runs_per_hour_e2e": 3600.0 * R / e2e, "algorithmic_bytes_per_launch": balg / K,
            "active_agent_days_per_run": a_sum / K / R, "infections_per_run": i_sum / K / R,
            "roofline_frac": balg / K / (ms / 1e3) / 1e9 / peak, "day_loop_ms": eng.last_run_ms3()[1]}


def tracing_leg(args, pop, device):
    """The same workload with the reference's own t_tracing_start = 37 (run_simulation.py:115-116): as
    compiled the reference traces contacts from that day.  One warm-up + one timed run."""
    from covid19_demography_b200 import engine, tables
    n, T = args.n, 60
    p, tabs = workload_params(n, T)
    p["t_tracing_start"] = 37.0
    eng = _engine(engine, tables, p, tabs, pop, 1, device, args.mode, args.launch)
    rng = np.random.default_rng(7)
    ms = []
    for k in range(2):
        init = [rng.choice(n, size=int(p["initial_infected_fraction"] * n), replace=False).astype(np.int64)]
        eng.run([900 + k], None, init)
        ms.append(eng.last_run_ms()[1])
    tal = eng.tallies(0)
    eng.close()
    return {"t_tracing_start": 37, "ms_per_run": ms[-1], "agent_days_per_sec": float(n) * (T - 1) / (ms[-1] / 1e3),
            "documented_at_end": int(tal[-1, 3].sum()), "deaths_at_end": int(tal[-1, 7].sum())}


def stress_leg(args, pop, device):
    """with 5 initial cases the active fraction stays tiny for 30 days, so the same workload is
    also run as a stress variant -- 0.1 % of the agents infected at t = 0, no lockdown -- in which up to half of the
    population is in E/Mild/Severe/Critical at once.  One warm-up + two timed runs."""
    from covid19_demography_b200 import engine, tables
    n, T = args.n, 60
    peak, _ = measured_peak()
    p, tabs = workload_params(n, T)
    p["initial_infected_fraction"] = 0.001
    p["t_lockdown"] = 1.0e6
    eng = _engine(engine, tables, p, tabs, pop, 1, device, args.mode, args.launch)
    rng = np.random.default_rng(11)
    runs = [([700 + k], None, [rng.choice(n, size=int(0.001 * n), replace=False).astype(np.int64)]) for k in range(3)]
    out = _timed_runs(eng, runs, peak, n, T)
    tal = eng.tallies(0)
    out.update({"initial_infected_fraction": 0.001, "lockdown": "never", "visits_per_run": eng.counters()["visits"],
                "peak_active_fraction": float(tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2)).max()) / n,
                "dram_bytes_per_launch_ncu": measured_traffic("seir_persistent_kernel.stress")})
    eng.close()
    return out


def c1_leg(args, pop, device):
    """C1: run_simulation.py's scenario at 1 M agents, one seed (the reference's own CPU-runnable case)."""
    from covid19_demography_b200 import engine, tables
    n, T = max(20000, args.n // 10), 60
    peak, _ = measured_peak()
    p, tabs = workload_params(n, T)
    eng = _engine(engine, tables, p, tabs, make_population(n), 1, device, args.mode, args.launch)
    rng = np.random.default_rng(21)
    runs = [([300 + k], None, [rng.choice(n, size=5, replace=False).astype(np.int64)]) for k in range(6)]
    out = _timed_runs(eng, runs, peak, n, T)
    eng.close()
    out["workload"] = "C1 run_simulation.py Italy scenario: %d agents, T=%d, 5 initial cases, one seed per launch" % (n, T)
    return out


def policy_params(n, T=120):
    """C4: simulate_agepolicies.py's constants on top of the Italy block: T = 120 (:205-206, :247), asymptomatic
    isolation never (:269), no documentation of mild cases (:306), the 60+ sheltered from the lockdown day on."""
    p, tabs = workload_params(n, T)
    p["mean_time_to_isolate_asympt"] = float("inf")
    p["p_documented_in_mild"] = 0.0
    p["t_stayhome_start"] = p["t_lockdown"]
    fsh = np.zeros(101)
    fsh[60:] = 0.5
    return p, tabs, fsh


def _ensemble_leg(args, device, n, T, R, p, tabs, fsh, label, launches=3, grid=None):
    """Replica ensembles: R runs per launch share the device-resident population; the prologue's picks are made on
    the device (keyed prologue), no T x n history is written (tallies are what ensembles and sweeps consume), and with
    `grid` the replicas of one launch belong to DIFFERENT grid points of a sweep (per-replica overrides)."""
    from covid19_demography_b200 import engine, tables
    peak, _ = measured_peak()
    pop = make_population(n)
    eng = _engine(engine, tables, p, tabs, pop, R, device, args.mode, args.launch, history="on_demand")
    overrides = None
    if grid:
        gs = np.array([len(g) for g in pop[4]])
        sets, pcs, points = [], [], []
        for k, (pc, mult, shift) in enumerate(grid):
            q = dict(p)
            q["p_infect_given_contact"] = pc
            sev = [np.clip(tabs[x] * mult ** (1.0 / 3.0), 0.0, 1.0) for x in ("p_mild_severe", "p_severe_critical", "p_critical_death")]
            sets.append(tables.build_tables(tabs["contact_matrix"], gs, tables.params_to_dict(q), tabs["mean_time_to_isolate_factor"], *sev,
                                            with_replay=(args.mode == "replay")))
            pcs.append(pc)
            points.append({"T": T - shift, "t_lockdown": int(p["t_lockdown"]) - shift, "table_set": k})
        eng.set_table_bank(sets, pcs)
        overrides = [points[r % len(points)] for r in range(R)]
    runs = [dict(seeds=[5000 + k * R + r for r in range(R)], keyed=(fsh, p["initial_infected_fraction"]), overrides=overrides)
            for k in range(launches + 1)]
    out = _timed_runs(eng, runs, peak, n, T, R)
    # no T x n history is written in this mode, so the 2 n bytes per day of SURVEY.md 8d's formula are not owed: the leg's
    # own figure charges only the bytes of the agents and infections it touches; the survey's formula is given beside it
    out["roofline_frac_survey_8d_formula"] = out["roofline_frac"]
    touched = 16.0 * out["active_agent_days_per_run"] * R + 64.0 * out["infections_per_run"] * R
    out["roofline_frac"] = touched / (out["ms_per_launch"] / 1e3) / 1e9 / peak
    out["algorithmic_bytes_per_launch_without_row_term"] = touched
    if grid:  # (the replicas' horizons differ: count their own agent-days)
        ad = float(n) * sum(o["T"] - 1 for o in overrides)
        out["agent_days_per_sec"] = ad / (out["ms_per_launch"] / 1e3)
    eng.close()
    out.update({"workload": label, "replicas_per_launch": R, "prologue": "device (keyed permutation picks)",
                "history": "on demand (not written)"})
    return out


def c4_leg(args, pop, device):
    n, T, R = args.n, 120, 8
    p, tabs, fsh = policy_params(n, T)
    return _ensemble_leg(args, device, n, T, R, p, tabs, fsh,
                         "C4-like: simulate_agepolicies.py parameters, %d agents, T=%d, %d seeds per launch on one GPU" % (n, T, R))


def c5_leg(args, pop, device):
    n, T, R = max(20000, args.n // 10), 73, 64
    p, tabs = workload_params(n, T)
    p["initial_infected_fraction"] = 5.0 / n
    # eight points of parameter_sweep.py's grid (:54-56): transmissibility x mortality multiplier x start date
    grid = [(pc, mult, shift) for pc in (0.021, 0.025, 0.029, 0.033) for mult, shift in ((2.0, 0), (4.0, 9))]
    return _ensemble_leg(args, device, n, T, R, p, tabs, np.zeros(101),
                         "C5-like: parameter_sweep.py grid, %d agents, T in {64, 73}, %d runs of 8 grid points per launch on one GPU" % (n, R),
                         grid=grid)


def c3_single_leg(args, pop, device):
    """C3 on ONE GPU: 59 M agents, T = 127 (the sharded leg of the N > 1 runs measures the same workload split over N)."""
    from covid19_demography_b200 import engine, tables
    n = args.c3_n
    peak, _ = measured_peak()
    p, tabs = c3_params(n, 5)
    T = int(p["T"])
    eng = _engine(engine, tables, p, tabs, make_population(n, "China"), 1, device)
    rng = np.random.default_rng(3)
    runs = [([50 + k], None, [rng.choice(n, size=5, replace=False).astype(np.int64)]) for k in range(4)]
    out = _timed_runs(eng, runs, peak, n, T)
    eng.close()
    out["workload"] = "C3 Hubei: %d agents, T=%d, lockdown day 69 (/50), 5 initial cases, one GPU" % (n, T)
    return out


def dropin_leg(args, pop, device):
    """The reference's own entry point: seir_individual.run_model(seed, households, age, ...) once per seed as
    run_simulation.py:437-444 calls it -- tables + the reference's MT19937 prologue picks + the device run + the ten
    per-agent result vectors back on the host.  The first call creates the engine (population upload), the later ones
    find it in the cache."""
    sys.path.insert(0, os.path.join(ROOT, "covid19-demography_b200"))
    import seir_individual as drop
    n, T = args.n, 60
    p, tabs = workload_params(n, T)
    hh, age, diab, hyp, groups = pop
    a = (hh, age, groups, diab, hyp, tabs["contact_matrix"], tabs["p_mild_severe"], tabs["p_severe_critical"],
         tabs["p_critical_death"], tabs["mean_time_to_isolate_factor"], tabs["lockdown_factor_age"],
         np.full(n, tabs["p_infect_household"]), np.zeros(101), p)
    import contextlib, io
    ms, ms_reads, ms_eager, ms_numpy = [], [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for seed in range(4):
            t0 = time.perf_counter()
            out = drop.run_model(seed, *a, tracing="off")
            deaths = int(out[7].sum(axis=1)[-1])
            ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms = sum(drop.last_engine.last_run_ms3())
        for seed in range(4, 7):   # + what run_simulation.py:448-460 reads of a run: two vectors, three final rows
            t0 = time.perf_counter()
            out = drop.run_model(seed, *a, tracing="off")
            te, nib = out[15], out[9]
            early = np.logical_and(te <= 20, te > 0)
            r0 = float(nib[early].mean()) if early.any() else 0.0
            final = int(out[7][-1].sum()) + int(out[6][-1].sum()) + int(out[0][-1].sum())
            ms_reads.append((time.perf_counter() - t0) * 1e3)
        for seed in range(7, 9):   # all ten vectors as host arrays at once (the reference's contract taken literally)
            t0 = time.perf_counter()
            out = drop.run_model(seed, *a, tracing="off", vectors="eager")
            ms_eager.append((time.perf_counter() - t0) * 1e3)
        for seed in range(9, 11):  # round 2's first version of this call: NumPy replay of the picks + eager vectors
            t0 = time.perf_counter()
            out = drop.run_model(seed, *a, tracing="off", vectors="eager", prologue="numpy")
            ms_numpy.append((time.perf_counter() - t0) * 1e3)
        del out, te, nib
        drop.clear_engine_cache()
    return {"first_call_ms": ms[0], "warm_call_ms": float(np.mean(ms[1:])),
            "warm_call_plus_driver_reads_ms": float(np.mean(ms_reads)),
            "warm_call_eager_vectors_ms": float(np.mean(ms_eager)),
            "warm_call_eager_vectors_numpy_prologue_ms": float(np.mean(ms_numpy)),
            "device_ms_last_call": dev_ms, "deaths_last_call": deaths, "r0_early_last_read": r0, "final_rows_sum": final,
            "what": "seir_individual.run_model per seed: build_tables + the reference's MT19937 prologue picks "
                    "(seir_prologue_mt, host) + seir_run + tallies; the nine histories and the ten float64[n] vectors "
                    "come back lazy and are copied when first used (plus_driver_reads: time_exposed, num_infected_by, "
                    "D[-1], R[-1], S[-1]); eager_vectors: all ten vectors D2H at once; engine cached after call 1"}


def sharded_c3_leg(args, dist, rank, world, local):
    """ONE population (C3: 59 M agents, T = 127) split by whole households over the N GPUs of the job, cross-shard
    contacts resolved in peer memory over NVLink inside the persistent kernel, one device-side barrier per day.
    Rank 0 first times the same runs unsharded (speed-up denominator); every rank then checks, on two fixtures,
    that the sharded run is bit-identical to its own unsharded run (tallies summed over the ranks, packed history
    and per-agent vectors of the owned agents)."""
    import torch
    from covid19_demography_b200 import engine, sharding, tables

    class _D:  # sharding.attach only needs rank / size / object gather / barrier
        get_rank, get_world_size = staticmethod(dist.get_rank), staticmethod(dist.get_world_size)
        all_gather_object, barrier = staticmethod(dist.all_gather_object), staticmethod(dist.barrier)

    # ---- parity on two fixtures
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    parity = []
    for name in ("italy_seeded", "china_default"):
        fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name))
        p = dict(zip(fx["param_keys"].tolist(), [float(v) for v in fx["param_vals"]]))
        p["t_tracing_start"] = 1.0e6
        age = fx["age"].astype(np.int64)
        n = len(age)
        hh = fx["households"].astype(np.int64)
        groups = tuple(np.where(age == a)[0] for a in range(101))
        tabs = {k: fx[k] for k in ("contact_matrix", "p_mild_severe", "p_severe_critical", "p_critical_death",
                                   "mean_time_to_isolate_factor")}
        tabs["p_infect_household"] = float(fx["p_infect_household_scalar"])
        popf = (hh, age, fx["diabetes"].astype(np.int64), fx["hypertension"].astype(np.int64), groups)
        home, init = tables.prologue(int(fx["seed"]), groups, n, fx["fraction_stay_home"], p["initial_infected_fraction"])
        single = _engine(engine, tables, p, tabs, popf, 1, local)
        single.run([11], [home], [init])
        ref_tal, ref_hist = single.tallies(0), single.history("packed")
        ref_vec = {k: single.agent_vector(k) for k in ("num_infected_by", "time_exposed", "time_to_recovery", "time_infected")}
        single.close()
        eng = _engine(engine, tables, p, tabs, popf, 1, local)
        bounds = sharding.attach(eng, hh, _D)
        eng.run([11], [home], [init])
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        ok = np.array_equal(eng.history("packed")[:, lo:hi], ref_hist[:, lo:hi])
        for k, v in ref_vec.items():
            ok = ok and np.array_equal(eng.agent_vector(k)[lo:hi], v[lo:hi], equal_nan=True)
        tal = torch.from_numpy(eng.tallies(0)).to("cuda:%d" % local)
        dist.all_reduce(tal)
        ok = ok and np.array_equal(tal.cpu().numpy(), ref_tal)
        flag = torch.tensor([0 if ok else 1], device="cuda:%d" % local)
        dist.all_reduce(flag)
        parity.append((name, int(flag.item()) == 0))
        eng.close()
    # ---- C3, 5 initial cases and the heavy variant
    n = args.c3_n
    pop = make_population(n, "China")
    out = {"parity": "bit-exact" if all(ok for _, ok in parity) else "MISMATCH", "parity_cases": dict(parity), "n_gpus": world}
    for label, n_initial in (("reference_5_cases", 5), ("heavy_30000_cases", 30000)):
        p, tabs = c3_params(n, n_initial)
        T = int(p["T"])
        rng = np.random.default_rng(3)  # the same initial cases and seeds on every rank
        runs = [([50 + k], None, [rng.choice(n, size=n_initial, replace=False).astype(np.int64)]) for k in range(4)]
        eng = _engine(engine, tables, p, tabs, pop, 1, local)
        one_ms = None
        if rank == 0:  # the same runs on one GPU
            eng.run(*runs[0])
            t_ = []
            for r in runs[1:]:
                eng.run(*r)
                t_.append(sum(eng.last_run_ms3()))
            one_ms = float(np.mean(t_))
        dist.barrier()
        bounds = sharding.attach(eng, pop[0], _D)
        eng.run(*runs[0])
        dev, infected = [], 0
        e0 = time.perf_counter()
        for r in runs[1:]:
            eng.run(*r)
            dev.append(sum(eng.last_run_ms3()))
            tal = eng.tallies(0)
            infected = int(tal[0, 0].sum() - tal[-1, 0].sum())
        dist.barrier()
        e2e = (time.perf_counter() - e0) / len(dev)
        tt = torch.tensor([float(np.mean(dev)), e2e, float(infected)], dtype=torch.float64, device="cuda:%d" % local)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt)
        eng.close()
        if rank == 0:
            ms = float(mx[0].item())
            inf = int(tt[2].item())
            # payload that crosses NVLink, from the run's own infection count (upper bound: as if every day were a sharded
            # one): per infection (world-1)/world of the hitters are remote -- an 8 B atomicMax and its 8 B answer, a 4 B
            # list entry -- and every other GPU reads the 4 B list entry next morning (its copy of the infection days)
            nv = inf * ((world - 1.0) / world * 20.0 + 4.0 * (world - 1))
            out[label] = {"ms_per_run": ms, "agent_days_per_sec": float(n) * (T - 1) / (ms / 1e3),
                          "e2e_ms_per_run": float(mx[1].item()) * 1e3, "one_gpu_ms_per_run": one_ms,
                          "speedup_vs_1gpu": one_ms / ms, "infections": inf,
                          "nvlink_payload_bytes_per_run_upper_bound": nv, "nvlink_payload_bytes_per_day": nv / (T - 1),
                          "shard_bounds": [int(b) for b in bounds]}
    out["workload"] = "C3 Hubei: %d agents, T=127, ONE population sharded by whole households over %d GPUs" % (n, world)
    out["exchange"] = ("inside seir_persistent_kernel, no NCCL call in the day loop: replicated (no exchange at all) while a day's lists hold "
                       "fewer than 393216 entries; then per first hit one system-scope atomicMax in the owner's HBM over NVLink, the new "
                       "infectees staged per CTA and appended to the owner's list with one system-scope atomic per CTA and owner; every "
                       "morning each GPU completes its own copy of the infection days from all owners' lists (bulk reads), so the "
                       "susceptibility test of a candidate contact is local; one device-side barrier across the GPUs per day")
    return out


def run_sharded_arm(args):
    """--sharded: only the sharded C3 leg (strong scaling), as its own line."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world < 2:
        raise SystemExit("--sharded needs torchrun with 2, 4 or 8 ranks")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = sharded_c3_leg(args, dist, rank, world, local)
    if rank == 0:
        print(json.dumps({"metric": METRIC, "n_gpus": world, "scaling": "strong", "config": {"sharded_c3": out}}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--replicas", type=int, default=1)
    ap.add_argument("--mode", default="native", choices=["native", "replay"])
    ap.add_argument("--launch", default="persistent", choices=["persistent", "per_day"])
    ap.add_argument("--flush-l2", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="agents per CPU run (0 = the GPU arm's population size)")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="only the headline C2 leg (no further config legs)")
    ap.add_argument("--skip", default="", help="comma-separated legs to leave out (c1,c4_like,c5_like,c3_single_gpu,...)")
    ap.add_argument("--c3-n", type=int, default=59_000_000, help="agents of the C3 legs")
    ap.add_argument("--sharded", action="store_true", help="only the sharded C3 leg (needs torchrun with N >= 2 ranks)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from ncu (profiles/), if known")
    args = ap.parse_args()
    args.skip = set(x for x in args.skip.split(",") if x)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3) if os.environ.get("SEIR_BENCH_STRICT_WARMUP", "1") == "1" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.sharded:
        run_sharded_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
