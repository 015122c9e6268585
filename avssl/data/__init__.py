"""Reference: avssl/data/__init__.py:1-5.  Only the step that hands batches to the hot path is rebuilt here — ``collate_general``
(SURVEY.md §8 row f3).  The datasets, audio / image transforms and their file formats are the reference's data plane and are NOT
rebuilt: when the reference tree is also on ``sys.path`` ("overlay") this package extends its search path over the
reference's ``avssl/data`` and resolves ``FlickrDataset`` / ``CoCoDataset`` / ``random_crop_max_length`` /
``get_simple_image_transform`` from the reference's own files, so ``avssl/task/base_task.py:13`` imports unchanged."""
import importlib
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)

from .collate_function import collate_general, collate_packed, unpack_on_device  # noqa: E402

_REFERENCE_ONLY = {"random_crop_max_length": "audio_transforms", "CoCoDataset": "coco_dataset", "FlickrDataset": "flickr_dataset",
                   "get_simple_image_transform": "image_transforms"}


def __getattr__(name):
    mod = _REFERENCE_ONLY.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    try:
        return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
    except ModuleNotFoundError as e:
        raise ImportError(f"avssl.data.{name} is part of the reference's data plane (avssl/data/{mod}.py); put the reference tree on "
                          f"sys.path BEHIND this repo to use it ({e})") from e
