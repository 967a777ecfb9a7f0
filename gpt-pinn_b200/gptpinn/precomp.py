"""Closed-form evaluation of a trained full-PINN and its derivative channels (no autograd).

Replaces the autograd calls in the reference's inputs() (Klein-Gordon/KG_precomp.py:7-21,34-52;
Burgers/B_precomp.py:6-16,22-55; Allen-Cahn/AC_precompute.py:7-18,26-58).  Channel names:
  v = u,  x = u_x,  t = u_t,  q = "P_xx" = u_xx + u_xt,  r = "P_tt" = u_tt + u_xt
(the mixed terms are what the reference's second autograd call actually returns).
"""
import ctypes as C

import torch

from . import _lib

_CH = ("v", "x", "t", "q", "r")
ACT = {"cos": 0, "tanh": 1}


def _linears(pinn):
    return [(l.weight, l.bias) for l in pinn.linears]


def mlp_channels(pinn_or_layers, act, xt, **outs):
    """Write the requested channels at points xt[N,2] into 1-D (possibly strided) destination views.

    `outs`: v=, x=, t=, q=, r= -> tensors of N elements (e.g. `buffer[:, i]`), written in place.
    """
    layers = _linears(pinn_or_layers) if hasattr(pinn_or_layers, "linears") else list(pinn_or_layers)
    if not xt.is_cuda:
        raise _lib.GptPinnError("mlp_channels: xt must be a CUDA tensor (no CPU fallback)")
    xt = xt.detach()
    if xt.dtype != torch.float32 or xt.dim() != 2 or xt.shape[1] != 2:
        raise ValueError("xt must be float32 [N, 2]")
    xt = xt.contiguous()
    N = xt.shape[0]
    m = _lib.Mlp()
    m.nlayers = len(layers) + 1
    if m.nlayers > 8:
        raise ValueError("at most 7 linear layers")
    keep = []
    for i, (W, b) in enumerate(layers):
        W = W.detach().to(torch.float32).contiguous()
        b = b.detach().to(torch.float32).contiguous().reshape(-1)
        if not W.is_cuda:
            raise _lib.GptPinnError("mlp_channels: weights must live on the GPU")
        keep += [W, b]
        m.layers[i] = W.shape[1]
        m.layers[i + 1] = W.shape[0]
        m.W[i] = W.data_ptr()
        m.b[i] = b.data_ptr()
    m.act = ACT[act]
    ptrs = (C.c_void_p * 5)()
    lds = (C.c_longlong * 5)()
    for k, name in enumerate(_CH):
        t = outs.get(name)
        if t is None:
            ptrs[k] = None
            lds[k] = 0
            continue
        if not t.is_cuda or t.dtype != torch.float32:
            raise ValueError(f"output {name} must be a float32 CUDA tensor")
        if t.dim() == 2 and t.shape[1] == 1:
            t = t[:, 0]
        if t.dim() != 1 or t.shape[0] != N:
            raise ValueError(f"output {name} must have {N} elements, got {tuple(t.shape)}")
        ptrs[k] = t.data_ptr()
        lds[k] = t.stride(0) if N > 1 else 1
    with torch.cuda.device(xt.device):
        stream = torch.cuda.current_stream(xt.device).cuda_stream
        rc = _lib.lib().gptpinn_mlp_channels(C.byref(m), xt.data_ptr(), N, C.byref(ptrs), C.byref(lds),
                                             C.c_void_p(stream))
    _lib.check(rc, "gptpinn_mlp_channels")
    return keep
