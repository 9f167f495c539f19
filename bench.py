#!/usr/bin/env python
"""bench.py -- agent-days/s of the daily agent-stepping path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is ONE full simulation: the C2 workload of BASELINE.json/SURVEY.md section 8 -- n = 10 M agents
(synthetic Lombardy-like population), T = 60 days, Italy contact matrix and severity tables,
5 initial cases, lockdown on day 46 (/10), tracing never switching on -- i.e. n*(T-1) = 5.9e8
agent-days.  With N > 1 (launched under torchrun, one rank per GPU) every rank runs its own seed of
the same workload (seed ensembles need no communication): weak scaling.

value      device time of the whole run (state init + day loop), CUDA events on the launching stream,
           population and tables already resident in HBM; max over ranks.
e2e        the same through the C ABI with HOST buffers: per-run inputs (Home_real, initial cases,
           seed) copied from pinned host memory, the [T,9,101] tallies and counters read back, wall clock.
roofline   seir_persistent_kernel (one launch = the whole day loop) against the measured HBM peak,
           with the ALGORITHMIC bytes of SURVEY.md section 8d: sum_t 2n + 16 A(t-1) + 64 I(t).
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


# --------------------------------------------------------------------------- CPU legs (oracle)
_W = {}


def _cpu_worker_init(n_sample, T):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle  # test/bench infrastructure: the reference's algorithm restated in C
    from covid19_demography_b200 import population
    p, tabs = workload_params(n_sample, T)
    hh, age, diab, hyp, groups = population.synthetic_population(n_sample, seed=1)
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
    if n_sample <= 0:  # full C2 size when the host has the memory (about 12 GB), else a 2 M-agent sample
        n_sample = 10_000_000 if _mem_available() > 2 * _bytes_per_run(10_000_000, T) else 2_000_000
    _cpu_worker_init(n_sample, T)
    dt = _cpu_worker_run(1)
    return {"value": n_sample * (T - 1) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/oracle_seir.c provider=mt (reference algorithm + MT19937 stream), one run of "
                      "n=%d agents, T=%d, C2 parameters, %.1f s" % (n_sample, T, dt)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU algorithm on all host cores (independent seeds in
    separate processes: the reference's SLURM-array mode, job.parameter_sweep.sh)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    T = 60
    # a bounded sample of the workload per step: every worker runs ONE simulation of 4 M agents (C2 parameters, the same
    # 5 initial cases); --cpu-sample 10000000 times the full C2 population (about 22 s per step on 16 cores)
    n_sample = args.cpu_sample if args.cpu_sample > 0 else 4_000_000
    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, int(_mem_available() * 0.6 / _bytes_per_run(n_sample, T)), args.cpu_workers or 32))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_worker_init, initargs=(n_sample, T)) as pool:
        # (the pool initializer already touched the library and built the population; the warm-up
        #  steps page the output buffers in once and are not timed)
        for w in range(min(args.warmup, 1)):
            pool.map(_cpu_worker_run, range(workers))
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_worker_run, [1000 + k * workers + j for j in range(workers)])
        dt = time.perf_counter() - t0
    value = workers * n_sample * (T - 1) * args.steps / dt
    sample = ("oracle port of the reference (MT19937 stream, bit-identical to the reference's run_model); each step = "
              "%d concurrent single-threaded runs of n=%d agents, T=%d, C2 parameters" % (workers, n_sample, T))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": "C2 Lombardy-like 10000000 agents, T=60, lockdown day 46 (/10), 1 seed(s) per GPU, "
                                   "tracing never switches on", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
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
    from covid19_demography_b200 import engine, population, tables

    sampler = ClockSampler(local)  # (nvidia-smi takes a few hundred ms to deliver rows: start it well before the timed region)
    sampler.start()
    n, T = args.n, 60
    p, tabs = workload_params(n, T)
    hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
    gs = np.array([len(g) for g in groups])
    tb = tables.build_tables(tabs["contact_matrix"], gs, tables.params_to_dict(p), tabs["mean_time_to_isolate_factor"],
                             tabs["p_mild_severe"], tabs["p_severe_critical"], tabs["p_critical_death"],
                             with_replay=(args.mode == "replay"))
    t0 = time.perf_counter()
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb,
                        n_replicas=args.replicas, device=local, mode=args.mode, launch=args.launch)
    create_s = time.perf_counter() - t0
    del hh
    R = args.replicas
    n_init = int(p["initial_infected_fraction"] * n)
    rng = np.random.default_rng(100 + rank)
    home = engine.pin(np.zeros(n, dtype=np.uint8))  # fraction_stay_home = 0 in C2 (run_simulation.py:53-54)

    def inputs(step):
        seeds = [(rank * 100003 + step) * 64 + r for r in range(R)]
        init = [rng.choice(n, size=n_init, replace=False).astype(np.int64) for _ in range(R)]
        return seeds, [home] * R, init

    def sync_all():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    flush = None
    if args.flush_l2:
        import ctypes as C
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
    dev_ms = loop_ms = 0.0
    launches = 0
    wall0 = time.perf_counter()
    last_tal = None
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()
            import torch
            torch.cuda.synchronize()
        eng.run(*ins[k])
        a, b = eng.last_run_ms()
        loop_ms += a
        dev_ms += b
        launches += eng.counters()["launches"]
    sync_all()
    sampler.mark_end()
    wall = time.perf_counter() - wall0
    clocks = sampler.finish()
    last_tal = eng.tallies(0)
    # ---- end to end through the C ABI with host buffers (wall clock)
    sync_all()
    e0 = time.perf_counter()
    for k in range(args.steps):
        eng.run(*ins[k])
        for r in range(R):
            eng.tallies(r)
    sync_all()
    e2e_s = time.perf_counter() - e0
    home_used = 1 <= p["t_stayhome_start"] < T  # (C2: stay-home never starts, Home_real is never read nor copied)
    h2d = R * ((n if home_used else 0) + 8 * n_init + 8)
    d2h = R * (T * 10 * 101 * 4 + 32)  # the device's ten tally rows per day and age (uint32) + the run counters, to pinned host memory

    agent_days_step = float(n) * (T - 1) * R
    loop_ms_local = loop_ms  # the roofline line describes rank 0's own kernel
    if dist is not None:
        import torch
        tt = torch.tensor([dev_ms, e2e_s, loop_ms], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, loop_ms = [float(x) for x in tt.tolist()]
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda:%d" % local)
        dist.all_reduce(ll)
        launches = int(ll.item())
    if rank == 0:
        peak, peak_src = measured_peak()
        balg, a_sum, i_sum = algorithmic_bytes(last_tal, n)
        init_ms, _, mat_ms = eng.last_run_ms3()  # (last step)
        kern_s = loop_ms_local / args.steps / 1e3
        achieved = balg * R / kern_s / 1e9
        line = {
            "metric": METRIC, "value": world * agent_days_step * args.steps / (dev_ms / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": "C2 Lombardy-like %d agents, T=%d, lockdown day 46 (/10), %d seed(s) per GPU, "
                                   "tracing never switches on" % (n, T, R),
                       "mode": args.mode, "launch": args.launch, "replicas_per_gpu": R,
                       "l2": ("explicit 256 MiB flush between steps" if args.flush_l2 else
                              "per-run working set %.2f GB (history %d x %d B + state) > 126 MB L2; state re-initialised every step"
                              % ((T * n + 46 * n) / 1e9, T, n)),
                       "day_loop_ms_per_step": loop_ms / args.steps, "engine_create_s": create_s,
                       "sim_runs_per_hour_device": 3600.0 * world * R / (dev_ms / args.steps / 1e3),
                       "sim_runs_per_hour_e2e": 3600.0 * world * R / (e2e_s / args.steps),
                       "wall_s_timed_region": wall},
            "clocks": clocks,
            "e2e": {"value": world * agent_days_step * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": args.traffic if args.traffic is not None else measured_traffic("seir_persistent_kernel"),
                         "kernel": "seir_persistent_kernel" if args.launch == "persistent" else "seir_day_kernel x T",
                         "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0, "algorithmic_bytes_per_launch": balg * R,
                         "active_agent_days": a_sum, "infections": i_sum,
                         "note": "the day loop visits only E/Mild/Severe/Critical agents (L2-resident); the 2n bytes per "
                                 "agent-day of the algorithmic figure are paid once by seir_materialize_kernel as a streaming "
                                 "T x n byte write",
                         "other_kernels": {"seir_materialize_kernel": {"ms": mat_ms, "GBps": R * (T * n + n / 4.0 + 32.0 * i_sum) / (mat_ms / 1e3) / 1e9 if mat_ms > 0 else None,
                                                                       "frac": R * (T * n + n / 4.0 + 32.0 * i_sum) / (mat_ms / 1e3) / 1e9 / peak if mat_ms > 0 else None},
                                           "seir_init_state_kernel + seir_init_infected_kernel": {"ms": init_ms}}},
        }
        if world == 1 and not args.no_tracing_leg:
            line["config"]["as_reference_tracing"] = tracing_leg(engine, population, tables, args, n, T, local)
            line["config"]["stress_variant"] = stress_leg(engine, population, tables, args, n, T, local, peak)
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample, T)
        print(json.dumps(line))
    eng.close()
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


def run_sharded_arm(args):
    """--sharded: ONE population (default C3: 59 M agents, T = 127) split by whole households over the N GPUs
    (strong scaling).  Cross-shard contacts are resolved in peer memory over NVLink inside the persistent
    kernel; the ranks meet at a device-side barrier once per day.  N = 1 runs the same workload unsharded."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from covid19_demography_b200 import engine, population, sharding, tables
    sampler = ClockSampler(local)
    sampler.start()
    n = args.n if args.n != 10_000_000 else 59_000_000
    p, tabs = c3_params(n, args.n_initial)
    T = int(p["T"])
    hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1, country="China")
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=False)
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1,
                        device=local, mode="native", launch="persistent")
    bounds = None
    if world > 1:
        class _D:  # sharding.attach only needs rank / size / object gather / barrier
            get_rank, get_world_size = staticmethod(dist.get_rank), staticmethod(dist.get_world_size)
            all_gather_object, barrier = staticmethod(dist.all_gather_object), staticmethod(dist.barrier)
        bounds = sharding.attach(eng, hh, _D)
    del hh
    rng = np.random.default_rng(3)  # the same initial cases and seeds on every rank
    runs = [([50 + k], None, [rng.choice(n, size=args.n_initial, replace=False).astype(np.int64)])
            for k in range(args.warmup + args.steps)]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        eng.run(*runs[k])
    sync_all()
    sampler.mark_start()
    dev_ms = loop_ms = 0.0
    launches = 0
    e0 = time.perf_counter()
    infected = 0
    for k in range(args.warmup, args.warmup + args.steps):
        eng.run(*runs[k])
        a, b = eng.last_run_ms()
        loop_ms += a
        dev_ms += b
        launches += eng.counters()["launches"]
        tal = eng.tallies(0)  # d2h of the step's result
        infected = int(tal[0, 0].sum() - tal[-1, 0].sum())
    sync_all()
    sampler.mark_end()
    e2e_s = time.perf_counter() - e0
    clocks = sampler.finish()
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s, loop_ms], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, loop_ms = [float(x) for x in tt.tolist()]
        ll = torch.tensor([launches, infected], dtype=torch.int64, device="cuda:%d" % local)
        dist.all_reduce(ll)
        launches, infected = int(ll[0].item()), int(ll[1].item())
    if rank == 0:
        agent_days = float(n) * (T - 1) * args.steps
        line = {"metric": METRIC, "value": agent_days / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
                "config": {"workload": "C3 Hubei-like %d agents, T=%d, lockdown day 69 (/50), %d initial cases, ONE population "
                                       "sharded by whole households over %d GPU(s), tracing never switches on" % (n, T, args.n_initial, world),
                           "mode": "native", "launch": "persistent", "exchange": "peer-memory atomics over NVLink + device-side day barrier"
                           if world > 1 else "none (single GPU)", "shard_bounds": None if bounds is None else [int(b) for b in bounds],
                           "day_loop_ms_per_step": loop_ms / args.steps, "infections_last_step": infected,
                           "l2": "per-run working set %.1f GB > 126 MB L2; state re-initialised every step" % ((T * n + 46 * n) / 1e9)},
                "clocks": clocks,
                "e2e": {"value": agent_days / e2e_s, "unit": UNIT, "h2d_bytes_per_step": world * (8 * args.n_initial + 8),
                        "d2h_bytes_per_step": world * T * 9 * 101 * 8, "ms_per_step": e2e_s / args.steps * 1e3},
                "gpu_launches": launches}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def tracing_leg(engine, population, tables, args, n, T, device):
    """The same workload with the reference's own t_tracing_start = 37 (run_simulation.py:115-116): as
    compiled the reference traces contacts from that day.  One warm-up + one timed run."""
    p, tabs = workload_params(n, T)
    p["t_tracing_start"] = 37.0
    hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=(args.mode == "replay"))
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1,
                        device=device, mode=args.mode, launch=args.launch)
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


def stress_leg(engine, population, tables, args, n, T, device, peak):
    """with 5 initial cases the active fraction stays tiny for 30 days, so the same workload is
    also run as a stress variant -- 0.1 % of the agents infected at t = 0, no lockdown -- in which up to half of the
    population is in E/Mild/Severe/Critical at once.  One warm-up + two timed runs; the roofline figure uses the same
    algorithmic-byte formula on this run's own tallies."""
    p, tabs = workload_params(n, T)
    p["initial_infected_fraction"] = 0.001
    p["t_lockdown"] = 1.0e6
    hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=(args.mode == "replay"))
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1,
                        device=device, mode=args.mode, launch=args.launch)
    rng = np.random.default_rng(11)
    loop, total = [], []
    for k in range(3):
        init = [rng.choice(n, size=int(p["initial_infected_fraction"] * n), replace=False).astype(np.int64)]
        eng.run([700 + k], None, init)
        a, b = eng.last_run_ms()
        loop.append(a)
        total.append(b)
    tal = eng.tallies(0)
    eng.close()
    balg, a_sum, i_sum = algorithmic_bytes(tal, n)
    act = tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2))
    return {"initial_infected_fraction": 0.001, "lockdown": "never", "ms_per_run": float(np.mean(total[1:])),
            "day_loop_ms": float(np.mean(loop[1:])), "agent_days_per_sec": float(n) * (T - 1) / (np.mean(total[1:]) / 1e3),
            "active_agent_days": a_sum, "peak_active_fraction": float(act.max()) / n, "infections": i_sum,
            "algorithmic_bytes_per_launch": balg, "roofline_frac": balg / (loop[-1] / 1e3) / 1e9 / peak}


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
    ap.add_argument("--cpu-sample", type=int, default=0, help="agents per CPU run (0 = full C2 size if memory allows)")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tracing-leg", action="store_true")
    ap.add_argument("--sharded", action="store_true", help="one C3 population sharded over the GPUs (strong scaling)")
    ap.add_argument("--n-initial", type=int, default=5, help="initial cases of the --sharded workload")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from ncu (profiles/), if known")
    args = ap.parse_args()
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
