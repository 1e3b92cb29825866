#!/usr/bin/env python
"""BASELINE config #5 as stated: fixed level l=6, M=4, Euro_GBM(125,105,.02,.5), 1e10 coupled samples
sharded over the ranks (global path ids [0,1e10) split contiguously), one all-reduce of 7 doubles.

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/config5.py [--total 1e10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import multilevel_mc_sdes_b200 as mm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=float, default=1e10)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
opt = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5, seed=0, device=local, distributed=world > 1, verbose=False)
opt.looper(1_000_000, 6, 4, False)                      # warm-up: context, module load, NCCL
opt._next_path[6] = 0
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
s = opt.looper(int(a.total), 6, 4, False)
dt = time.perf_counter() - t0
if rank == 0:
    n = a.total
    print(json.dumps({"config": "Euro_GBM(125,105,.02,.5) l=6 M=4", "total_samples": n, "n_gpus": world, "seconds": dt,
                      "samples_per_s": n / dt, "path_steps_per_s": n * 4096 / dt,
                      "mean_dP": s[0] / n, "var_dP": s[1] / n - (s[0] / n) ** 2, "mean_Pf": s[2] / n,
                      "se_mean_Pf": float(np.sqrt((s[3] / n - (s[2] / n) ** 2) / n)), "sums": list(map(float, s))}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
