"""Reference: avssl/module/projections.py:6-29 (MLPLayers: Linear -> nonlinearity -> Dropout stack, the last pair dropped).
Inactive in every shipped config (the ``*_projection`` keys of kwClip.py:1149-1187 are absent from the YAMLs); kept on the module
surface and run on the same kernels as the trainable head: TF32 tensor-core GEMMs on fp32 rows, ReLU and the
counter-based dropout of ``scb_dropout_rows``."""
import torch
from torch import nn

from speechclip_b200.functional import DropoutState, DropoutFn, LinearFn, ReluFn

__all__ = ["MLPLayers"]


class MLPLayers(nn.Module):
    def __init__(self, units=[512, 512, 512], nonlin=nn.ReLU(), dropout=0.1):
        super().__init__()
        if not isinstance(nonlin, nn.ReLU):
            raise NotImplementedError("MLPLayers on B200: ReLU (the reference's default) is the only nonlinearity with a kernel")
        self.nonlin = nonlin
        self.dropout = dropout
        # same module layout as the reference (nn.Sequential of Linear / ReLU / Dropout, last two removed): same state-dict keys
        sequence = []
        for u0, u1 in zip(units[:-1], units[1:]):
            sequence += [nn.Linear(u0, u1), self.nonlin, nn.Dropout(self.dropout)]
        self.sequential = nn.Sequential(*sequence[:-2])
        self._scb_dropout = DropoutState()

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if not X.is_cuda:
            raise RuntimeError("MLPLayers: CUDA tensors required (no CPU path)")
        site = 16
        state = self._scb_dropout.advance(X.device) if self.training and self.dropout > 0 else None
        for m in self.sequential:
            if isinstance(m, nn.Linear):
                X = LinearFn.apply(X, m.weight, m.bias)
            elif isinstance(m, nn.ReLU):
                X = ReluFn.apply(X)
            elif isinstance(m, nn.Dropout) and state is not None:
                X = DropoutFn.apply(X, (float(self.dropout), state, site))
                site += 1
        return X
