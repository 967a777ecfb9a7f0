"""Full-PINN training loop (reference KG_train.py:48-59, B_train.py:76-90): Adam on the PINN loss.

Host loop over torch autograd + torch.optim.Adam, i.e. the same algorithm as the reference on the
same device; the fused forward/derivative/backward/Adam CUDA step is the next row of the scope
table (K9/B5) and is not part of this module yet.
"""
import torch


def train_adam(model, loss_fn, epochs, lr, tol=None, record_every=None):
    """Runs `epochs` Adam steps.  With `tol`, stops *before* the step once loss < tol (B_train.py:84).

    Returns the final loss as float, or with `record_every` a dict {epoch: loss after that epoch}
    (evaluated like the reference: an extra loss evaluation after the step)."""
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
            rec[i if i % record_every == 0 else i] = loss_fn().item()
    if record_every:
        return rec
    return float(loss.item()) if loss is not None else float("nan")
