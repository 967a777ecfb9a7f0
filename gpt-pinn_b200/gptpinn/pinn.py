"""Full-PINN training (reference KG_train.py:48-59, B_train.py:76-90): Adam on the PINN loss.

`train_fused` runs every epoch inside one persistent cooperative CUDA kernel (gptpinn_pinn_train):
forward-mode derivative channels, PDE residual, hand-derived reverse pass, gradient reduction and
torch.optim.Adam's update, with the weights resident on the device for the whole run.
`train_adam` is the host loop over torch autograd; it is kept only for callers that hand over an
arbitrary loss closure (it is not used by the drop-in pinn_train for Klein-Gordon / Burgers).
"""
import ctypes as C

import torch

from . import _lib

PDE = {"kg": 0, "burgers": 1}
ACT = {"cos": 0, "tanh": 1}


def _vec(t, rows, name, keep):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.GptPinnError(f"{name} must be a CUDA tensor (no CPU fallback)")
    t = t.detach().to(torch.float32).reshape(-1).contiguous()
    if t.numel() != rows:
        raise ValueError(f"{name} has {t.numel()} entries, expected {rows}")
    keep.append(t)
    return t.data_ptr()


def train_fused(model, pde, pde_params, xt_resid, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, epochs, lr,
                xcos=None, tol=None, trace_every=0, want_grad=False, betas=(0.9, 0.999), eps=1e-8):
    """Train `model` (an nn.Module with `.linears`, updated in place).  Returns dict(loss, epochs[, trace, grad]).

    loss = the last evaluated loss (the value the reference prints), as a 0-dim CUDA tensor (no sync here)."""
    dev = xt_resid.device
    if dev.type != "cuda":
        raise _lib.GptPinnError("train_fused: tensors must live on the GPU (no CPU fallback)")
    keep = []
    m = _lib.Mlp()
    lin = list(model.linears)
    m.nlayers = len(lin) + 1
    n_params = 0
    for i, layer in enumerate(lin):
        for p in (layer.weight, layer.bias):
            if not (p.is_cuda and p.dtype == torch.float32 and p.data.is_contiguous()):
                raise _lib.GptPinnError("model parameters must be contiguous float32 CUDA tensors")
        m.layers[i] = layer.weight.shape[1]
        m.layers[i + 1] = layer.weight.shape[0]
        m.W[i] = layer.weight.data.data_ptr()
        m.b[i] = layer.bias.data.data_ptr()
        n_params += layer.weight.numel() + layer.bias.numel()
    m.act = ACT[getattr(model, "activation_name", "tanh")]
    NR, NI, NB = xt_resid.shape[0], IC_xt.shape[0], BC_xt.shape[0]
    q = _lib.PinnProblem()
    q.pde = PDE[pde]
    pp = list(pde_params) + [0.0, 0.0, 0.0]
    q.p0, q.p1, q.p2 = float(pp[0]), float(pp[1]), float(pp[2])
    q.NR, q.NI, q.NB = NR, NI, NB
    q.xt_resid = _vec(xt_resid, 2 * NR, "xt_resid", keep)
    q.f_hat = _vec(f_hat, NR, "f_hat", keep)
    q.xcos = _vec(xcos, NR, "xcos", keep)
    q.IC_xt = _vec(IC_xt, 2 * NI, "IC_xt", keep)
    q.IC_u1 = _vec(IC_u1, NI, "IC_u", keep)
    q.IC_u2 = _vec(IC_u2, NI, "IC_u2", keep)
    q.BC_xt = _vec(BC_xt, 2 * NB, "BC_xt", keep)
    q.BC_u = _vec(BC_u, NB, "BC_u", keep)
    a = _lib.TrainArgs()
    a.epochs, a.lr, a.beta1, a.beta2, a.eps = int(epochs), float(lr), float(betas[0]), float(betas[1]), float(eps)
    a.tol = float(tol) if tol is not None else -1.0
    trace = grad = None
    if trace_every and trace_every > 0 and epochs > 0:
        trace = torch.zeros((epochs - 1) // trace_every + 1, dtype=torch.float32, device=dev)
        a.trace_every, a.loss_trace_dev = int(trace_every), trace.data_ptr()
    if want_grad:
        grad = torch.zeros(n_params, dtype=torch.float32, device=dev)
        a.grad_out_dev = grad.data_ptr()
    status = torch.zeros(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        h = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.lib().gptpinn_pinn_train(h, C.byref(q), C.byref(m), C.byref(a), status.data_ptr(), C.c_void_p(stream))
    _lib.check(rc, "gptpinn_pinn_train")
    out = dict(loss=status[0], status=status, trace=trace, grad=grad, n_params=n_params, _keep=keep)
    return out


def epochs_done(result):
    """Epochs entered (syncs)."""
    return int(result["status"][1:2].view(torch.int32).item())


def train_adam(model, loss_fn, epochs, lr, tol=None, record_every=None):
    """Host loop: `epochs` Adam steps on an arbitrary loss closure (torch autograd).  With `tol`, stops
    *before* the step once loss < tol (B_train.py:84).  Returns the final loss, or with `record_every`
    a dict {epoch: loss evaluated after that epoch's step}."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    loss = None
    rec = {}
    for i in range(1, epochs + 1):
        loss = loss_fn()
        if tol is not None and loss < tol:
            break
        opt.zero_grad()
        loss.backward()
        opt.step()
        if record_every and (i % record_every == 0 or i == epochs):
            rec[i] = loss_fn().item()
    if record_every:
        return rec
    return float(loss.item()) if loss is not None else float("nan")
