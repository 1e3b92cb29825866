#!/usr/bin/env python
"""bench.py - fixed-level coupled path-samples/s at L=6, M=4 (BASELINE.json metric, config #5).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's numpy algorithm on host cores

A "step" is one pass of the hot path over one batch: `Option.looper(batch, l=6, M=4)` for
Euro_GBM(X0=125,K=105,r=0.02,sig=0.5) - one kernel launch that simulates `batch` coupled
fine/coarse paths of 4096/1024 Euler-Maruyama steps and reduces the seven sums.  Per-GPU batch is
fixed (weak scaling); ranks take disjoint global path-id ranges; the only message is one all-reduce
of 7 doubles per step.

value  : samples/s, K launches timed on the device with CUDA events, outputs left in HBM
e2e    : the same through the public host call (Option.looper: launch + 56-byte D2H + all-reduce
         when N>1), wall clock around K calls
The kernel reads no global memory (its only traffic is the 56-byte result), so there is no L2
state to flush between iterations.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_BENCH, M_BENCH = 6, 4
OPT_KW = dict(X0=125, K=105, r=0.02, sig=0.5)
METRIC = "coupled path-samples/sec at L=6,M=4"
UNIT = "samples/s"
# algorithmic FP64 issue slots per coupled fine step with the canonical fp64
# Box-Muller normal (3.5 for the EM update at M=4 + 41 for the normal).
W_ALG_FP64_SLOTS = 44.5


# ------------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    seed, n = args
    from oracle import np_oracle as O
    np.random.seed(seed)
    return O.looper(O.Params(**OPT_KW), O.GBM, O.EURO, n, L_BENCH, M_BENCH)


def cpu_reference_rate(paths_per_core, repeats=1, cores=None):
    """The reference's algorithm (oracle/np_oracle.py: the numpy restatement that is bit-identical
    to MLMC_RSCAM.looper) on all host cores: one forked worker per core, distinct numpy seeds,
    7-vectors added.  Returns (samples/s best of `repeats`, cores, all rates)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    rates = []
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(i, 8) for i in range(cores)])          # warm the workers
        for r in range(repeats):
            t0 = time.perf_counter()
            rows = pool.map(_ref_worker, [(1000 * r + i, paths_per_core) for i in range(cores)])
            dt = time.perf_counter() - t0
            assert np.isfinite(np.sum(rows))
            rates.append(cores * paths_per_core / dt)
    return max(rates), cores, rates


def c_port_rate(n, threads):
    from oracle import c_oracle as co, np_oracle as O
    co.build()
    t0 = time.perf_counter()
    co.level_sums(O.Params(**OPT_KW), O.GBM, O.EURO, L_BENCH, M_BENCH, n, seed=1, nthreads=threads)
    return n / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_core = int(os.environ.get("MLMCB_BENCH_REF_PATHS", 10000))   # one step = one reference-sized chunk (Npl=1e4, :919) per core, ~1 s
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_ref_worker, [(w * cores + i, 1000) for i in range(cores)])
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_ref_worker, [(7919 * (k + 1) + i, per_core) for i in range(cores)])
        dt = time.perf_counter() - t0
    value = args.steps * cores * per_core / dt
    sample = f"{args.steps} steps x {cores} workers x looper({per_core}, l=6, M=4) of the numpy port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Euro_GBM(X0=125,K=105,r=0.02,sig=0.5) looper l=6 M=4 (config #5)",
                   "paths_per_step": cores * per_core},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, period=0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._halt = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", str(int(self.period * 1000))], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if self._halt.is_set():
                break
            parts = [p.strip() for p in line.split(",")]
            try:
                self.sm.append(float(parts[0]))
                self.sm_max = float(parts[1])
            except Exception:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)

    def stop(self):
        self._halt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------- this repo's arm
def run_b200(args):
    import torch
    import multilevel_mc_sdes_b200 as mm
    from multilevel_mc_sdes_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    dev = torch.device("cuda", local)
    sampler = args.sampler
    opt = mm.Euro_GBM(seed=args.seed, device=local, sampler=sampler, distributed=world > 1, verbose=False, **OPT_KW)
    batch = int(args.batch)                      # paths per GPU per step
    p = opt._params()
    flags = opt._flags()
    out = torch.zeros(7, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    nsteps_path = M_BENCH ** L_BENCH

    def launch(step):
        # global ids: step-major, then rank (disjoint Philox subsequences by construction)
        off = (step * world + rank) * batch
        _lib.check(lib.mlmcb_level_sums(C.byref(p), _lib.GBM, _lib.EURO, L_BENCH, M_BENCH, batch, args.seed, off,
                                        flags, local, stream, out.data_ptr(), None, None))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-timed value -------------------------------------------------------------------
    for w in range(args.warmup):
        launch(w)
    barrier()
    sampler_thread = ClockSampler(local)
    sampler_thread.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        launch(args.warmup + k)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler_thread.stop()
    sums_last = out.cpu().numpy()
    assert np.isfinite(sums_last).all() and sums_last[3] > 0
    total = float(batch) * world * args.steps
    value = total / (ms * 1e-3)

    # ---- end to end through the public host call ------------------------------------------------
    base = (args.warmup + args.steps) * world * batch
    opt._next_path[L_BENCH] = base
    for w in range(max(1, args.warmup // 2)):
        opt.looper(batch * world, L_BENCH, M_BENCH, False)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        s_host = opt.looper(batch * world, L_BENCH, M_BENCH, False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dt)
    e2e = total / dt
    price_f = s_host[2] / (batch * world)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline denominators measured live (rank 0) -------------------------------------------
    ips, pms = C.c_double(), C.c_float()
    _lib.check(lib.mlmcb_pipe_probe(0, local, C.byref(ips), C.byref(pms)))
    dfma_lane_per_s = ips.value * 32.0
    steps_per_s_gpu = value * nsteps_path / world
    achieved = steps_per_s_gpu * W_ALG_FP64_SLOTS
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # the real bound of this kernel: warp-instruction dispatch
    issue = None
    try:
        mix = json.load(open(os.path.join(ROOT, "multilevel_mc_sdes_b200", "csrc", "sass_mix.json")))[sampler]
        f_sm = (clocks.get("sm_mhz") or 1965.0) * 1e6
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        slots_per_s = n_sm * 4 * f_sm                             # dispatch cycles/s over all SMSPs
        warp_steps_per_s = steps_per_s_gpu / 32.0
        cyc = slots_per_s / warp_steps_per_s
        issue = {
            "warp_inst_per_fine_step_static": mix["inst_per_4_steps"] / 4.0,
            "dispatch_cycles_per_fine_step_model": mix["dispatch_cycles_per_4_steps"] / 4.0,
            "measured_smsp_cycles_per_warp_step": cyc,
            "issue_frac": mix["inst_per_4_steps"] / 4.0 / cyc,
            "dispatch_frac": mix["dispatch_cycles_per_4_steps"] / 4.0 / cyc,
            "note": "ALU-class and IMAD instructions occupy the dispatch port 2 cycles on B200 "
                    "(profiles/r01_pipe_mix.txt); dispatch_frac is the fraction of that bound reached",
        }
    except Exception:
        pass
    roofline = {
        "bound": "fp64_pipe", "achieved": achieved / 1e9, "peak": dfma_lane_per_s / 1e9, "issue": issue,
        "unit": "G FP64 issue-slots/s (algorithmic 44.5 per coupled fine step, SURVEY 8(d))",
        "frac": achieved / dfma_lane_per_s,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
        # kernel (profiles/r01_ncu_gbm_euro_f32bm.txt: 71.7 KB read - kernel image and constants - 0 written)
        "traffic": 71680 if sampler == "f32bm" else None,
        "peak_source": "DFMA chain micro-benchmark (mlmcb_pipe_probe) measured in this run; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "kernel_ms_per_launch": ms / args.steps,
        "path_steps_per_s_per_gpu": steps_per_s_gpu,
        "hbm_gbs_measured_peak": peaks.get("hbm_gbs"),
        "algorithmic_hbm_bytes_per_launch": 56,
    }

    # the second half of BASELINE.json's metric: mlmc(eps) wall time (N=1 only; ~0.3 s of GPU time)
    mlmc_wall = None
    if world == 1:
        mlmc_wall = {}
        for name, o, eps, kw in (("asian_gbm_eps1e-4_M4", mm.Asian_GBM(verbose=False, device=local), 1e-4, dict(M=4)),
                                 ("readme_euro_gbm_eps0.01", mm.Euro_GBM(verbose=False, device=local, **OPT_KW), 0.01, {})):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sums_m, N_m = o.mlmc(eps, **kw)
            mlmc_wall[name] = {"seconds": time.perf_counter() - t0, "price": float(np.sum(sums_m[0] / N_m)),
                               "levels": len(N_m), "samples": float(N_m.sum())}
        mlmc_wall["readme_euro_gbm_eps0.01"]["BS"] = 35.17419437739385

    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        per_core = 20000                     # ~3 s per repeat per core at ~6e3 samples/s/core
        best, cores, rates = cpu_reference_rate(per_core, repeats=3, cores=cores)
        cpu = {"value": best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"best of 3 x ({cores} workers x looper({per_core}, l=6, M=4)) of oracle/np_oracle.py "
                         f"(numpy restatement, bit-identical to the reference); all runs: "
                         + ", ".join(f"{r:.3g}" for r in rates),
               "c_port_all_threads": c_port_rate(4000 * cores, cores)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Euro_GBM(X0=125,K=105,r=0.02,sig=0.5) looper l=6 M=4 (BASELINE config #5)",
                   "paths_per_gpu_per_step": batch, "fine_steps_per_path": nsteps_path, "sampler": sampler,
                   "parallelism": f"paths sharded over {world} GPU(s), one 56-byte all-reduce per step",
                   "l2": "no inputs: the kernel reads no global memory, nothing to flush"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(_lib.Params),
                "d2h_bytes_per_step": 56, "ms_per_step": 1e3 * dt / args.steps},
        "gpu_launches": args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "mlmc_wall": mlmc_wall,
        "check": {"mean_Pf_last_step": price_f, "path_steps_per_s": value * nsteps_path},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=float, default=float(1 << 26), help="paths per GPU per step")
    ap.add_argument("--sampler", default="f32bm", choices=["f32bm", "f64bm"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
