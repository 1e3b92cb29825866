#!/usr/bin/env python
"""bench.py - fixed-level coupled path-samples/s at L=6, M=4 (BASELINE.json metric, config #5).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's numpy algorithm on host cores

A "step" is one pass of the hot path over one batch: `Option.looper(batch, l=6, M=4)` for
Euro_GBM(X0=125,K=105,r=0.02,sig=0.5) - one kernel launch that simulates `batch` coupled
fine/coarse paths of 4096/1024 Euler-Maruyama steps and reduces the seven sums; ranks take disjoint
global path-id ranges and the only message is one all-reduce of the sums per step (inside the timed
region for N > 1).

value  : samples/s with the DEFAULT (all-fp64) sampler, K launches timed on the device with CUDA events
e2e    : the same through the public host call (Option.looper: launch + all-reduce when N>1 + D2H of
         the sums), wall clock around K calls
value_f32bm / e2e_f32bm : the opt-in fp32-sampler kernel on the same workload (a few steps)
sampler_check : level moments of both samplers from the timed launches, against each other and against the
         pinned C port's 10^7-sample moments (tests/golden/cport_moments.json) as z-scores
roofline : what bounds the kernel - warp-instruction issue/dispatch - with the FP64-pipe utilisation and
         SURVEY 8(d)'s algorithmic FP64-slot figure as sub-fields; pipe rates re-measured in this run
mlmc_wall : Option.mlmc(eps) wall times (second half of BASELINE's metric) on all N GPUs
strong : a fixed total of 2^28 coupled samples split over the N GPUs (BASELINE config #5 is a fixed total)
The kernel reads no global memory (its only traffic is the 56-byte result), so there is no L2
state to flush between iterations.
"""
import argparse
import ctypes as C
import json
import math
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
WORKLOAD = "Euro_GBM(X0=125,K=105,r=0.02,sig=0.5) looper l=6 M=4 (BASELINE config #5)"
# algorithmic FP64 issue slots per coupled fine step with a canonical (libdevice) fp64
# Box-Muller normal: 3.5 for the EM update at M=4 + 41 for the normal.  Reported as `alg_fp64_equiv`.
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
        "config": {"workload": WORKLOAD, "paths_per_step": cores * per_core, "fine_steps_per_path": M_BENCH ** L_BENCH,
                   "sampler": "numpy legacy MT19937 + polar normals (fp64)"},
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
PROBES = {0: "DFMA", 3: "IMAD.WIDE", 4: "LOP3", 2: "FFMA", 9: "DFMA+LOP3", 10: "DFMA+IMAD.WIDE", 11: "IMAD.WIDE+LOP3",
          12: "FFMA+LOP3", 13: "DFMA+FFMA", 14: "DFMA+IMAD"}


def _moments(sums, n):
    m_d, m_f = sums[0] / n, sums[2] / n
    return m_d, sums[1] / n - m_d ** 2, m_f, sums[3] / n - m_f ** 2


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
    strong = args.scaling == "strong"
    batch = int(args.total) // world if strong else int(args.batch)          # paths per GPU per step
    stream = torch.cuda.current_stream(dev).cuda_stream
    nsteps_path = M_BENCH ** L_BENCH

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

    def measure(sampler, steps, warmup, with_clocks):
        """-> dict(value, ms, e2e, e2e_ms, sums (7,) over the timed launches of all ranks, clocks)."""
        opt = mm.Euro_GBM(seed=args.seed, device=local, sampler=sampler, distributed=world > 1, verbose=False, **OPT_KW)
        p, flags = opt._params(), opt._flags()
        outs = torch.zeros((steps + warmup, 8), dtype=torch.float64, device=dev)

        def launch(step):
            # global ids: step-major, then rank (disjoint Philox subsequences by construction)
            off = (step * world + rank) * batch
            row = outs[step]
            _lib.check(lib.mlmcb_level_sums(C.byref(p), _lib.GBM, _lib.EURO, L_BENCH, M_BENCH, batch, args.seed, off,
                                            flags, local, stream, row.data_ptr(), None, None))
            if dist is not None:
                dist.all_reduce(row, op=dist.ReduceOp.SUM)         # the step's only message, on the device

        for w in range(warmup):
            launch(w)
        barrier()
        sampler_thread = ClockSampler(local) if with_clocks else None
        if sampler_thread:
            sampler_thread.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for k in range(steps):
            launch(warmup + k)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler_thread.stop() if sampler_thread else None
        sums = outs[warmup:, :7].sum(dim=0).cpu().numpy()
        assert np.isfinite(sums).all() and sums[3] > 0
        total = float(batch) * world * steps
        # end to end through the public host call
        opt._next_path[L_BENCH] = (warmup + steps) * world * batch
        for w in range(max(1, warmup // 2)):
            opt.looper(batch * world, L_BENCH, M_BENCH, False)
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            opt.looper(batch * world, L_BENCH, M_BENCH, False)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return dict(value=total / (ms * 1e-3), ms=ms / steps, e2e=total / dt, e2e_ms=1e3 * dt / steps, sums=sums,
                    n=total, clocks=clocks)

    main_s = args.sampler
    other_s = "f32bm" if main_s == "f64bm" else "f64bm"
    head = measure(main_s, args.steps, args.warmup, True)
    alt = measure(other_s, min(args.steps, 3), 3, False)
    value, ms, clocks = head["value"], head["ms"], head["clocks"]

    # ---- the second half of BASELINE.json's metric: mlmc(eps) wall time, on all N GPUs ----------------
    mlmc_wall = {}
    for name, ctor, eps, kw in (("asian_gbm_eps1e-4_M4", lambda: mm.Asian_GBM(verbose=False, device=local, distributed=world > 1), 1e-4, dict(M=4)),
                                ("readme_euro_gbm_eps0.01", lambda: mm.Euro_GBM(verbose=False, device=local, distributed=world > 1, **OPT_KW), 0.01, {}),
                                # BASELINE config #4 (Merton factor walk + the l = 0 loop): one run, about 5 s on one GPU
                                ("digital_merton_eps1e-3_M2", lambda: mm.Digital_Merton(beta_0=0.5, verbose=False, device=local, distributed=world > 1), 1e-3, dict(M=2))):
        best = None
        for rep in range(1 if name.startswith("digital_merton") else 2):   # second run: warm library state (pools, pinned buffers, NCCL)
            o = ctor()
            barrier()
            t0 = time.perf_counter()
            sums_m, N_m = o.mlmc(eps, **kw)
            dt_m = max_over_ranks(time.perf_counter() - t0)
            if best is None or dt_m < best["seconds"]:
                best = {"seconds": dt_m, "price": float(np.sum(sums_m[0] / N_m)), "levels": len(N_m), "samples": float(N_m.sum())}
        mlmc_wall[name] = best
    mlmc_wall["readme_euro_gbm_eps0.01"]["BS"] = 35.17419437739385
    mlmc_wall["digital_merton_eps1e-3_M2"]["exact"] = 44.77499       # Poisson mixture of Black-Scholes digitals (tests/test_gpu_configs.py)
    mlmc_wall["sampler"] = "f64bm (default)"

    # ---- fixed total work split over the GPUs (what can bend as N grows) ------------------------------
    strong_total = 1 << 28
    o = mm.Euro_GBM(seed=args.seed + 1, device=local, distributed=world > 1, verbose=False, **OPT_KW)
    o.looper(strong_total // 8, L_BENCH, M_BENCH, False)
    barrier()
    t0 = time.perf_counter()
    s_strong = o.looper(strong_total, L_BENCH, M_BENCH, False)
    dt_s = max_over_ranks(time.perf_counter() - t0)
    strong_leg = {"total_paths": strong_total, "seconds": dt_s, "value": strong_total / dt_s, "unit": UNIT,
                  "mean_Pf": float(s_strong[2] / strong_total),
                  "note": "one Option.looper call (sampler f64bm) over a fixed total, sharded by global path id"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- sampler check: level moments of both samplers from the timed launches -------------------------
    a, b = _moments(head["sums"], head["n"]), _moments(alt["sums"], alt["n"])
    check = {main_s: {"n": head["n"], "E_dP": a[0], "E_Pf": a[2]}, other_s: {"n": alt["n"], "E_dP": b[0], "E_Pf": b[2]},
             "z_between_samplers": {"E_dP": (a[0] - b[0]) / math.sqrt(a[1] / head["n"] + b[1] / alt["n"]),
                                    "E_Pf": (a[2] - b[2]) / math.sqrt(a[3] / head["n"] + b[3] / alt["n"])}}
    try:
        ref = json.load(open(os.path.join(ROOT, "tests", "golden", "cport_moments.json")))["euro_gbm_bench_l6M4"]
        r = _moments(np.array(ref["sums"]), ref["n"])
        for nm, m, n in ((main_s, a, head["n"]), (other_s, b, alt["n"])):
            check[f"z_{nm}_vs_c_port_1e7"] = {"E_dP": (m[0] - r[0]) / math.sqrt(m[1] / n + r[1] / ref["n"]),
                                              "E_Pf": (m[2] - r[2]) / math.sqrt(m[3] / n + r[3] / ref["n"])}
    except Exception as e:  # noqa: BLE001
        check["c_port"] = f"unavailable: {e}"

    # ---- roofline: pipe rates measured live, instruction mix of the built kernel ------------------------
    f_sm = (clocks.get("sm_mhz") or 1965.0) * 1e6
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    slots_per_s = n_sm * 4 * f_sm                                 # issue/dispatch cycles per second over all SMSPs
    probe = {}
    for which, name in PROBES.items():
        ips, pms = C.c_double(), C.c_float()
        _lib.check(lib.mlmcb_pipe_probe(which, local, C.byref(ips), C.byref(pms)))
        per = 2.0 if which >= 9 else 1.0
        probe[name] = {"warp_inst_per_s": ips.value, "smsp_cycles_per_" + ("pair" if per == 2 else "inst"): per * slots_per_s / ips.value}
    dfma_lane_per_s = probe["DFMA"]["warp_inst_per_s"] * 32.0
    steps_per_s_gpu = value * nsteps_path / world
    warp_steps_per_s = steps_per_s_gpu / 32.0
    cyc = slots_per_s / warp_steps_per_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = {"bound": "issue/dispatch", "unit": "G warp-instructions/s", "peak": slots_per_s / 1e9,
                "kernel_ms_per_launch": ms, "path_steps_per_s_per_gpu": steps_per_s_gpu,
                "measured_smsp_cycles_per_warp_step": cyc,
                "peak_source": "one warp instruction per SMSP per clock: SMs x 4 x SM clock sampled during the run; pipe "
                               "rates from mlmcb_pipe_probe in this run (MEASURED_PEAKS.json has no issue or FP64 entry)",
                "pipe_probe": probe, "hbm_gbs_measured_peak": peaks.get("hbm_gbs"), "algorithmic_hbm_bytes_per_launch": 56,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the ncu --set full capture of this
                # kernel (profiles/r02_ncu_gbm_euro_f64bm.txt, r01_ncu_gbm_euro_f32bm.txt): kernel image, constants and
                # the sampler tables; the kernel reads no data and its 56-byte result stays in L2
                "traffic": {"f64bm": 154112, "f32bm": 71680}.get(main_s)}
    try:
        mix = json.load(open(os.path.join(ROOT, "multilevel_mc_sdes_b200", "csrc", "sass_mix.json")))[main_s]
        inst = mix["inst_per_4_steps"] / 4.0
        fp64 = mix["classes_per_iteration"].get("fp64", 0) / mix["steps_per_loop_iteration"]
        dfma_cyc = probe["DFMA"]["smsp_cycles_per_inst"]
        roofline.update({
            "achieved": inst * warp_steps_per_s / 1e9, "frac": inst / cyc,
            "warp_inst_per_fine_step_static": inst,
            "dispatch_cycles_per_fine_step_model": mix["dispatch_cycles_per_4_steps"] / 4.0,
            "dispatch_frac": mix["dispatch_cycles_per_4_steps"] / 4.0 / cyc,
            "fp64_inst_per_fine_step": fp64, "fp64_pipe_frac": fp64 * dfma_cyc / cyc,
            "note": "frac = executed warp instructions / issue slots; INT32-class instructions (LOP3, IADD, SHF, IMAD) hold "
                    "the dispatch port two cycles on B200 (pipe_probe: LOP3 alone, DFMA+LOP3, IMAD.WIDE+LOP3), so "
                    "dispatch_frac - the 2-cycle-weighted count over the measured cycles - is the fraction of the bound "
                    "this instruction mix allows; fp64_pipe_frac = FP64 instructions x measured DFMA issue interval"})
    except Exception as e:  # noqa: BLE001
        roofline["mix_error"] = str(e)
    roofline["alg_fp64_equiv"] = {
        "slots_per_fine_step": W_ALG_FP64_SLOTS, "achieved_G_slots_per_s": steps_per_s_gpu * W_ALG_FP64_SLOTS / 1e9,
        "dfma_peak_G_lane_per_s": dfma_lane_per_s / 1e9, "ratio": steps_per_s_gpu * W_ALG_FP64_SLOTS / dfma_lane_per_s,
        "note": "SURVEY 8(d): 44.5 FP64 slots per step with a libdevice fp64 Box-Muller; the shipped samplers spend fewer, "
                "so this ratio is a throughput equivalence, not a pipe utilisation"}

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

    h2d = C.sizeof(_lib.Params)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "paths_per_step": batch * world, "paths_per_gpu_per_step": batch,
                   "fine_steps_per_path": nsteps_path,
                   "sampler": main_s + (" (default: 52-bit uniforms, fp64 Box-Muller)" if main_s == "f64bm" else " (opt-in: fp32 Box-Muller normals)"),
                   "parallelism": f"paths sharded over {world} GPU(s) by global path id, one all-reduce of the sums per step",
                   "l2": "no inputs: the kernel reads no global memory, nothing to flush"},
        "e2e": {"value": head["e2e"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 56,
                "ms_per_step": head["e2e_ms"]},
        f"value_{other_s}": alt["value"], f"e2e_{other_s}": alt["e2e"], f"ms_per_step_{other_s}": alt["ms"],
        "gpu_launches": args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "mlmc_wall": mlmc_wall, "strong": strong_leg,
        "sampler_check": check,
        "check": {"mean_Pf": a[2], "path_steps_per_s": value * nsteps_path},
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
    ap.add_argument("--batch", type=float, default=float(1 << 26), help="paths per GPU per step (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --total paths per step split over the GPUs (BASELINE config #5 is a fixed total)")
    ap.add_argument("--total", type=float, default=float(1 << 29), help="paths per step over all GPUs with --scaling strong")
    ap.add_argument("--sampler", default="f64bm", choices=["f32bm", "f64bm"])
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
