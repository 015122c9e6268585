"""Reference: avssl/model/base_model.py:11-26 (BaseLightningModel).  Uses pytorch_lightning when it is importable; otherwise
a minimal stand-in with the LightningModule services the model code relies on: ``device``, ``log`` /
``log_dict``, ``save_hyperparameters``, ``load_from_checkpoint``, ``current_epoch`` / ``global_step``."""
import logging

import torch
from torch import nn

from speechclip_b200.params import restoring_from_checkpoint

from ..base import OrderedNamespace

logger = logging.getLogger(__name__)

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


class BaseLightningModel(_Base):
    def __init__(self, config: OrderedNamespace):
        super().__init__()
        self.config = config
        self.save_hyperparameters()

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, hparams_file=None, strict=True, **kwargs):
        """``example.py:10`` / ``base_task.py:64``: rebuild the model from the ``OrderedNamespace`` pickled under
        ``hyper_parameters["config"]`` (``base_model.py:15`` -> ``save_hyperparameters``) and restore ``state_dict``.

        Same contract as Lightning's classmethod, with one addition: while the model is constructed the towers do not look for
        their separate pretrained files (``audio_encoder.pretrained: true`` in a released checkpoint's config would otherwise
        demand ``hubert_base_ls960.pt`` although the .ckpt holds every ``audio_encoder.encoder.*`` / ``clip.model.*`` tensor).
        Optimizer / scheduler state (``optimizer_states``, ``lr_schedulers``) is left to the Trainer's resume path
        (``FusedAdam.load_state_dict`` accepts torch-Adam's layout)."""
        from ..module.speech_encoder_plus import load_checkpoint_lenient
        ckpt = load_checkpoint_lenient(checkpoint_path) if map_location is None else torch.load(
            checkpoint_path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        config = hp.get("config", hp)
        if not isinstance(config, OrderedNamespace):
            config = OrderedNamespace(config)
        with restoring_from_checkpoint():
            model = cls(config)
        missing, unexpected = model.load_state_dict(ckpt["state_dict"], strict=False)
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_from_checkpoint({checkpoint_path}): missing keys {missing[:8]}, unexpected keys {unexpected[:8]}")
        if missing or unexpected:
            logger.warning("load_from_checkpoint: %d missing / %d unexpected keys", len(missing), len(unexpected))
        if hasattr(model, "on_load_checkpoint"):
            try:
                model.on_load_checkpoint(ckpt)
            except TypeError:
                pass
        return model

    def forward(self, batch):
        raise NotImplementedError

    def training_step(self, batch, batch_idx=None):
        raise NotImplementedError

    def configure_optimizers(self):
        raise NotImplementedError
