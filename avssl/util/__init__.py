"""Reference: avssl/util/__init__.py:1-6.  The helpers on the hot path (``get_keypadding_mask``, ``freeze_model`` /
``unfreeze_model``) live here.  The rest of the reference's ``avssl.util`` (argument parsing, logging set-up, weight init,
penalty schedule: ``args.py``, ``log.py``, ``init_model.py``, ``penalty_scheduler.py``) is control plane and is NOT rebuilt:
when the reference tree is also on ``sys.path`` ("overlay"), this package extends its search path over the
reference's ``avssl/util`` and resolves those names from the reference's own files, so ``run_task.py`` /
``avssl/task/base_task.py:14`` import unchanged."""
import importlib
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

from .data_utils import get_keypadding_mask  # noqa: E402
from .model_utils import freeze_model, unfreeze_model  # noqa: E402

_REFERENCE_ONLY = {"add_general_arguments": "args", "set_logging": "log", "set_pl_logger": "log", "init_weights": "init_model",
                   "PenaltyScheduler": "penalty_scheduler"}


def __getattr__(name):
    mod = _REFERENCE_ONLY.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    try:
        return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
    except ModuleNotFoundError as e:
        raise ImportError(f"avssl.util.{name} is part of the reference's control plane (avssl/util/{mod}.py); put the reference "
                          f"tree on sys.path BEHIND this repo to use it ({e})") from e
