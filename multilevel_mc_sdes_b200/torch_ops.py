"""`torch.ops.mlmcb.level_sums`: the level function as a PyTorch custom op (device tensor in the graph
of a torch program, current stream, no host synchronisation).  A thin registration over the C ABI
(include/mlmcb.h) - no separate C++ build; importing this module needs torch.

    import multilevel_mc_sdes_b200.torch_ops        # registers the op
    out = torch.ops.mlmcb.level_sums(params, model, payoff, l, M, n_paths, seed, path_offset, flags, device)
    # params: the 13 doubles of struct mlmcb_params in order (a tabulated American boundary is not supported
    # through this op); out: float64[7] on cuda:device
"""
import ctypes as C

import torch

from . import _lib


@torch.library.custom_op("mlmcb::level_sums", mutates_args=())
def level_sums(params: list[float], model: int, payoff: int, l: int, M: int, n_paths: int, seed: int,
               path_offset: int, flags: int, device: int) -> torch.Tensor:
    p = _lib.Params(*params)
    dev = torch.device("cuda", device)
    out = torch.empty(7, dtype=torch.float64, device=dev)
    _lib.check(_lib.load().mlmcb_level_sums(
        C.byref(p), model, payoff, l, M, n_paths, seed, path_offset, flags, device,
        int(torch.cuda.current_stream(dev).cuda_stream), out.data_ptr(), None, None))
    return out


@level_sums.register_fake
def _(params, model, payoff, l, M, n_paths, seed, path_offset, flags, device):
    return torch.empty(7, dtype=torch.float64, device=torch.device("cuda", device))


def option_params(opt):
    """The `params` list of an Option from this package."""
    p = opt._params()
    return [getattr(p, name) for name, ctype in _lib.Params._fields_ if ctype is C.c_double]
