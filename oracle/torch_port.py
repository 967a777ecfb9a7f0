"""Restatement of the reference's PyTorch code path for the Klein-Gordon sweep, op for op.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs, tests).
It exists because the reference itself (a Python tree under /root/reference) cannot travel to the
GPU box; this file issues the same ATen calls in the same order, so its CPU timing is the
"reference's torch CPU path" the north star asks to be shown next to the GPU number.

Follows: KG_precomp.py:23-29 (T, G2), KG_optimizer.py:34-68 (update), KG_models.py:28-50 (loss),
KG_train.py:12-43 (offline_generation incl. early exit), KG_test.py:10-44 (gpt_test loop).
Pinned by tests/test_oracle_golden.py::test_torch_port_matches_reference against tests/golden.
"""
import torch
import torch.nn.functional as F


class KGUpdate:
    """One parameter's grad_descent + GPT pair (the objects offline_generation builds per parameter)."""

    def __init__(self, alpha, beta, gamma, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, xcos, f_hat,
                 IC_u1, IC_u2, BC_u, lr):
        self.gamma, self.lr = gamma, lr
        self.P, self.PIC, self.PICt, self.PBC = out, out_IC, out_IC_t, out_BC
        self.T = torch.add(torch.add(out_tt, torch.mul(alpha, out_xx)), torch.mul(beta, out))
        self.G2 = torch.mul(2 * gamma, out)
        self.xcos, self.f_hat, self.IC_u1, self.IC_u2, self.BC_u = xcos, f_hat, IC_u1, IC_u2, BC_u
        self.fac1, self.fac2, self.fac3 = 2 / out.shape[0], 2 / out_BC.shape[0], 2 / out_IC.shape[0]

    def update(self, c, cr):
        u = torch.matmul(self.P, cr)
        term1 = torch.matmul(self.T, cr)
        term2 = torch.add(term1, torch.mul(self.gamma, torch.square(u)))
        first = torch.add(term2, self.xcos)
        second = torch.add(self.T, torch.mul(self.G2, u))
        g = torch.mul(self.fac1, torch.sum(torch.mul(first, second), axis=0))
        bc = torch.sub(torch.matmul(self.PBC, cr), self.BC_u)
        ic1 = torch.sub(torch.matmul(self.PIC, cr), self.IC_u1)
        g += torch.mul(self.fac2, torch.sum(torch.mul(bc, self.PBC), axis=0))
        g += torch.mul(self.fac3, torch.sum(torch.mul(ic1, self.PIC), axis=0))
        ic2 = torch.matmul(self.PICt, cr)
        g += torch.mul(self.fac3, torch.sum(torch.mul(ic2, self.PICt), axis=0))
        return torch.sub(c, torch.mul(self.lr, g))

    def loss(self, cr):
        u = torch.matmul(self.P, cr)
        f = torch.add(torch.add(torch.matmul(self.T, cr), torch.mul(self.gamma, torch.square(u))), self.xcos)
        lR = F.mse_loss(f, self.f_hat)
        lI1 = F.mse_loss(torch.matmul(self.PIC, cr), self.IC_u1)
        lI2 = F.mse_loss(torch.matmul(self.PICt, cr), self.IC_u2)
        lB = F.mse_loss(torch.matmul(self.PBC, cr), self.BC_u)
        return lR + lI1 + lI2 + lB


def kg_offline_generation(grid, c_initial, mats, epochs, lr, early_exit=True):
    """KG_train.py:12-43.  Returns (largest_loss, index, param_epochs_executed)."""
    largest_loss, largest_idx, done = 0, -1, 0
    for p, (alpha, beta, gamma) in enumerate(grid):
        st = KGUpdate(alpha, beta, gamma, *mats, lr)
        c = c_initial
        cr = c.unsqueeze(1)
        loss = st.loss(cr)
        for _ in range(1, epochs + 1):
            if early_exit and loss < largest_loss:
                break
            c = st.update(c, cr)
            cr = c.unsqueeze(1)
            loss = st.loss(cr)
            done += 1
        if loss > largest_loss:
            largest_idx, largest_loss = p, loss
    return largest_loss, largest_idx, done


def kg_trajectory(mu, c_initial, mats, epochs, lr):
    """Loss and c after `epochs` updates for one parameter (no early exit)."""
    st = KGUpdate(*mu, *mats, lr)
    c = c_initial
    cr = c.unsqueeze(1)
    for _ in range(epochs):
        c = st.update(c, cr)
        cr = c.unsqueeze(1)
    return c, st.loss(cr)
