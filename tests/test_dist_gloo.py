"""CPU, world_size 2 over gloo: the N>1 host path (contiguous id slices + one all-reduce)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_row(level, n, off):
    """A 'level function' that is a pure function of the path-id range, so that sharding is checkable:
    row k = sum over ids of (id+1)^(k+1) mod 1e6-ish weights."""
    ids = np.arange(off, off + n, dtype=np.float64)
    return np.array([np.sum(np.cos(ids * (k + 1) * 0.001 + level)) for k in range(7)])


class _FakeLib:
    def mlmcb_level_sweep_host(self, p, model, payoff, M, k, levels, n, off, seed, flags, dev, st, out):
        import ctypes
        arr = np.ctypeslib.as_array(ctypes.cast(out, ctypes.POINTER(ctypes.c_double)), shape=(k * 7,))
        for i in range(k):
            arr[7 * i:7 * i + 7] = _fake_row(levels[i], n[i], off[i])
        return 0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import multilevel_mc_sdes_b200 as mm
    from multilevel_mc_sdes_b200 import _lib
    _lib.load = lambda: _FakeLib()
    o = mm.Euro_GBM(distributed=True)
    r1 = o.looper(1001, 3, 4, False)
    r2 = o.looper(500, 3, 4, False)
    sw = o.level_sweep(Lmax=2, Nsamples=77, M=2)
    n_loc, start = o._shard(10, 100)
    q.put((rank, r1, r2, sw, n_loc, start))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank holds the same all-reduced rows, equal to the single-process answer
    for rank, r1, r2, sw, n_loc, start in res:
        np.testing.assert_allclose(r1, _fake_row(3, 1001, 0), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(r2, _fake_row(3, 500, 1001), rtol=1e-12, atol=1e-9)
        for l in range(3):
            np.testing.assert_allclose(sw[:, l], _fake_row(l, 77, 0), rtol=1e-12, atol=1e-9)
    assert (res[0][4], res[0][5]) == (5, 100) and (res[1][4], res[1][5]) == (5, 105)
