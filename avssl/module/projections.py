"""Reference: avssl/module/projections.py:6-29.  Inactive in every shipped config (kwClip.py:1149-1187 keys are absent from
the YAMLs); the class is kept so configs naming it fail with a clear message instead of an AttributeError."""
from torch import nn

__all__ = ["MLPLayers"]


class MLPLayers(nn.Module):
    def __init__(self, units=(512, 512, 512), nonlin=None, dropout=0.1):
        super().__init__()
        raise NotImplementedError("MLPLayers projections are outside the B200 hot path (no shipped config enables them)")
