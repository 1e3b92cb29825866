#!/usr/bin/env python
"""Multi-GPU invariance: level sums and an mlmc run sharded over G ranks with
`distributed=True` equal the single-GPU result for the same seed up to fp64 summation order.

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/dist_check.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import multilevel_mc_sdes_b200 as mm  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rep = {"world": world}
worst = 0.0
for cls, l, M, n in ((mm.Euro_GBM, 6, 4, 1_000_003), (mm.Asian_GBM, 3, 2, 5_000_001), (mm.Lookback_Merton, 4, 4, 2_000_000),
                     (mm.Digital_Merton, 2, 2, 3_000_000), (mm.Euro_Merton, 0, 2, 7_000_001), (mm.Asian_Merton, 4, 4, 500_000),
                     (mm.Amer_GBM, 5, 2, 1_000_000)):
    d = cls(seed=11, device=local, distributed=True, verbose=False).looper(n, l, M, False)
    s = cls(seed=11, device=local, distributed=False, verbose=False).looper(n, l, M, False)
    rel = float(np.max(np.abs(d - s) / np.maximum(np.abs(s), 1e-300)[np.abs(s) > 0].max()))
    worst = max(worst, rel)
    rep[cls.__name__] = {"max_rel_diff": rel}
    assert np.allclose(d, s, rtol=1e-12, atol=0), (cls.__name__, d, s)
# a full driver run: identical decisions (N per level) and prices on every rank
o = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5, seed=5, device=local, distributed=True, verbose=False)
sums, N = o.mlmc(0.02)
ref = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5, seed=5, device=local, distributed=False, verbose=False)
sums1, N1 = ref.mlmc(0.02)
rep["mlmc"] = {"price_dist": float(np.sum(sums[0] / N)), "price_single": float(np.sum(sums1[0] / N1)),
               "same_N": bool(np.array_equal(N, N1))}
assert np.array_equal(N, N1) and abs(rep["mlmc"]["price_dist"] - rep["mlmc"]["price_single"]) < 1e-9
gathered = [None] * world
dist.all_gather_object(gathered, rep["mlmc"]["price_dist"])
assert len(set(gathered)) == 1            # every rank holds the same all-reduced result
rep["worst_rel_diff"] = worst
if rank == 0:
    print(json.dumps(rep))
dist.barrier()
dist.destroy_process_group()
