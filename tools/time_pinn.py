"""Time the fused full-PINN training kernel against the torch autograd + Adam loop on the same GPU."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
import B_data, B_models, KG_data, KG_models, KG_precomp  # noqa: E402
from gptpinn import pinn  # noqa: E402



def _torch_adam(model, loss_fn, epochs, lr):
    """the reference's loop (torch autograd + torch.optim.Adam), for the timing comparison only"""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    for _ in range(epochs):
        loss = loss_fn()
        opt.zero_grad()
        loss.backward()
        opt.step()
    return float(loss)


def kg(epochs_fused=3000, epochs_torch=60):
    xt, f_hat, _ = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, 100, 40)
    IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, 512, 512)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = (x.cuda() for x in (xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u))
    xcos = KG_precomp.xcos_term(xt[:, 0:1], xt[:, 1:2])
    net = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, xcos).cuda()
    pinn.train_fused(net, "kg", (-1.5, 0.5, 0.5), xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 10, 5e-4, xcos=xcos)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = pinn.train_fused(net, "kg", (-1.5, 0.5, 0.5), xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, epochs_fused, 5e-4, xcos=xcos)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"KG fused : {1e6 * dt / epochs_fused:9.1f} us/epoch  ({epochs_fused} epochs, loss {float(r['loss']):.4e})  -> 75000 epochs = {75000 * dt / epochs_fused:.1f} s")
    for nb in (2, 3, 4, 6, 8):                 # several test parameters per launch (pinn_test): throughput per training
        mus = [(-1.5 + 0.05 * k, 0.5, 0.5) for k in range(nb)]
        nets = [KG_models.NN(np.array([2, 40, 40, 1]), *mu, xcos).cuda() for mu in mus]
        pinn.train_fused_batch(nets, "kg", mus, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 10, 5e-4, xcos=xcos)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pinn.train_fused_batch(nets, "kg", mus, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, epochs_fused // 2, 5e-4, xcos=xcos)
        torch.cuda.synchronize()
        dtb = time.perf_counter() - t0
        print(f"KG fused, batch of {nb}: {1e6 * dtb / (epochs_fused // 2):9.1f} us/epoch per launch = {1e6 * dtb / (epochs_fused // 2) / nb:7.1f} us/epoch per training")
    net2 = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, xcos).cuda()
    xr, ir = xt.clone().requires_grad_(), IC_xt.clone().requires_grad_()
    fn = lambda: net2.loss(xr, f_hat, ir, IC_u1, IC_u2, BC_xt, BC_u)
    _torch_adam(net2, fn, 5, 5e-4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _torch_adam(net2, fn, epochs_torch, 5e-4)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    print(f"KG torch : {1e6 * dt2 / epochs_torch:9.1f} us/epoch  (autograd + Adam on the same GPU)   speed-up {dt2 / epochs_torch / (dt / epochs_fused):.1f}x")


def burgers(epochs_fused=3000, epochs_torch=60):
    xt, f_hat, _ = B_data.residual_data(-1.0, 1.0, 0.0, 1.0, 100, 100)
    IC_xt, IC_u, BC_xt, BC_u = B_data.ICBC_data(-1.0, 1.0, 0.0, 1.0, 200, 200)
    xt, f_hat, IC_xt, IC_u, BC_xt, BC_u = (x.cuda() for x in (xt, f_hat, IC_xt, IC_u, BC_xt, BC_u))
    net = B_models.NN(np.array([2, 20, 20, 20, 20, 1]), 0.5025).cuda()
    pinn.train_fused(net, "burgers", (0.5025,), xt, f_hat, IC_xt, IC_u, None, BC_xt, BC_u, 10, 5e-3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = pinn.train_fused(net, "burgers", (0.5025,), xt, f_hat, IC_xt, IC_u, None, BC_xt, BC_u, epochs_fused, 5e-3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"B  fused : {1e6 * dt / epochs_fused:9.1f} us/epoch  ({epochs_fused} epochs, loss {float(r['loss']):.4e})")
    net2 = B_models.NN(np.array([2, 20, 20, 20, 20, 1]), 0.5025).cuda()
    xr = xt.clone().requires_grad_()
    fn = lambda: net2.loss(xr, IC_xt, IC_u, BC_xt, BC_u, f_hat)
    _torch_adam(net2, fn, 5, 5e-3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _torch_adam(net2, fn, epochs_torch, 5e-3)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    print(f"B  torch : {1e6 * dt2 / epochs_torch:9.1f} us/epoch   speed-up {dt2 / epochs_torch / (dt / epochs_fused):.1f}x")


if __name__ == "__main__":
    kg()
    burgers()
