"""WeightOffsets — same module surface as the reference (e4t/weightoffsets.py:5-23): parameters `v`, `linear1`,
`linear2`, `linear_column`, `linear_row`; `forward()` -> Δ of shape (column_dim, row_dim).

The reference materialises Δ with two dense square GEMMs per call; algebraically Δ = b·aᵀ + s·b_cᵀ + b_r·1ᵀ with
a = W_c(w1 v+β1), b = W_r(w2 v+β2), s = W_r·1 (rank ≤ 3).  `forward()` returns that closed
form (API accessor, differentiable torch ops); the attention hot path never calls it — it goes through
e4t_b200.functional.WOEffectiveFn, which fuses the factor mat-vecs, the W ⊙ (1+Δ) modulation and the whole
backward on the sm_100a kernels."""
import torch
from torch import nn


class WeightOffsets(nn.Module):
    def __init__(self, row_dim, column_dim):
        super().__init__()
        self.v = nn.Parameter(torch.ones(1))
        self.linear1 = nn.Linear(1, row_dim)
        self.linear2 = nn.Linear(1, column_dim)
        self.linear_column = nn.Linear(row_dim, row_dim)
        self.linear_row = nn.Linear(column_dim, column_dim)

    def factors(self):
        vx = self.linear1.weight[:, 0] * self.v + self.linear1.bias
        vy = self.linear2.weight[:, 0] * self.v + self.linear2.bias
        a = self.linear_column.weight @ vx
        b = self.linear_row.weight @ vy
        s = self.linear_row.weight.sum(dim=1)
        return a, b, s

    def forward(self):
        a, b, s = self.factors()
        return b[:, None] * a[None, :] + s[:, None] * self.linear_column.bias[None, :] + self.linear_row.bias[:, None]

    def kernel_params(self):
        """The nine tensors in the order the C-ABI expects (include/e4t_b200.h: e4t_wo_factors_fwd / e4t_wo_bwd)."""
        return (self.v, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                self.linear_column.weight, self.linear_column.bias, self.linear_row.weight, self.linear_row.bias)
