"""Reference: avssl/model/base_model.py:11-26 (BaseLightningModel).  Uses pytorch_lightning when it is importable; otherwise
a minimal stand-in with the LightningModule services the model code relies on: ``device``, ``log`` /
``log_dict``, ``save_hyperparameters``, ``load_from_checkpoint``, ``current_epoch`` / ``global_step``."""
import torch
from torch import nn

from ..base import OrderedNamespace

try:  # pragma: no cover - pytorch_lightning is absent from this image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class _Base(nn.Module):
        def __init__(self):
            super().__init__()
            self.logged = {}
            self.hparams = {}
            self.current_epoch = 0
            self.global_step = 0
            self.logger = None

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        def save_hyperparameters(self, **kw):
            self.hparams = {"config": getattr(self, "config", None)}

        def log(self, name, value, **kw):
            self.logged[name] = value

        def log_dict(self, d, **kw):
            self.logged.update(d)

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kw):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hp = ckpt.get("hyper_parameters", {})
            config = hp.get("config", hp)
            model = cls(config if isinstance(config, OrderedNamespace) else OrderedNamespace(config))
            model.load_state_dict(ckpt["state_dict"], strict=kw.get("strict", True))
            return model


class BaseLightningModel(_Base):
    def __init__(self, config: OrderedNamespace):
        super().__init__()
        self.config = config
        self.save_hyperparameters()

    def forward(self, batch):
        raise NotImplementedError

    def training_step(self, batch, batch_idx=None):
        raise NotImplementedError

    def configure_optimizers(self):
        raise NotImplementedError
