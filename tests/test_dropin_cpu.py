"""The drop-in boundary at package level, on CPU:

* the reference's own YAML configs + vocabulary tables build the model (needs /root/reference; skipped on the GPU box);
* with this repo in front of the reference on ``sys.path`` the reference's task runner (``run_task.py:6`` ->
  ``avssl/task/base_task.py``) imports, resolving ``avssl.model`` / ``avssl.module`` / ``avssl.base`` here and its control
  plane (``avssl.task``, the datasets of ``avssl.data``, ``avssl.util.{args,log}``) in the reference;
* a reference-layout Lightning checkpoint (pickled ``OrderedNamespace`` hparams with ``pretrained: true``, duplicate
  ``cascaded_branch.clip.*`` keys, ``criterion.*`` buffers, torch-Adam ``optimizer_states``) loads through
  ``load_from_checkpoint`` and ``FusedAdam.load_state_dict``.
"""
import copy
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "config")), reason="the reference tree is not on this machine")

YAMLS = ["model_base/spchclp_p.yaml", "model_base/spchclp_c.yaml", "model_large/flickr/spchclp_p.yaml",
         "model_large/flickr/spchclp_c.yaml", "model_large/coco/spchclp_p.yaml", "model_large/coco/spchclp_c.yaml"]
# trainable parameters of the shipped configurations (weighted sum + branch [+ temperature])
EXPECTED_TRAINABLE = {"model_base/spchclp_p.yaml": 7_483_917, "model_large/flickr/spchclp_p.yaml": 13_386_522,
                      "model_large/coco/spchclp_p.yaml": 13_386_522}


def _load_yaml(rel):
    import yaml
    cfg = yaml.load(open(os.path.join(REF, "config", "speechCLIP", rel)), Loader=yaml.FullLoader)
    cfg["audio_encoder"]["pretrained"] = False  # hubert_*.pt / ViT-*.pt are not reachable offline: seeded weights, same shapes
    npy = cfg["clip"].get("reduce_subword_embbedding")
    if npy:
        cfg["clip"]["reduce_subword_embbedding"] = os.path.normpath(os.path.join(REF, npy))
    return cfg


@needs_reference
@pytest.mark.parametrize("rel", YAMLS)
def test_reference_yaml_builds_the_model(rel):
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    cfg = _load_yaml(rel)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg))
    n_train = sum(p.numel() for p in model.getTrainableParams() if p.requires_grad)
    if rel in EXPECTED_TRAINABLE:
        assert n_train == EXPECTED_TRAINABLE[rel], n_train
    cascaded = cfg["model_settings"]["cascaded_objective_weight"] > 0
    assert (model.cascaded_branch is not None) == cascaded and (model.parallel_branch is not None) == (not cascaded)
    sd = model.state_dict()
    large = "model_large" in rel
    d, L = (1024, 24) if large else (768, 12)
    assert sd["audio_encoder.encoder.post_extract_proj.weight"].shape == (d, 512)
    assert f"audio_encoder.encoder.encoder.layers.{L - 1}.fc2.weight" in sd and f"audio_encoder.encoder.encoder.layers.{L}.fc2.weight" not in sd
    assert sd["audio_encoder.weightedsum_layer.weights"].shape == (L + 1,)
    assert ("criterion.temperature" in sd) == bool(cfg["cl_loss"]["args"]["temperature_trainable"])
    if cascaded:
        import numpy as np
        table = np.load(cfg["clip"]["reduce_subword_embbedding"])
        assert sd["clip.model.token_embedding.weight"].shape[0] == len(table)          # reduced vocabulary (clip_official.py:75-86)
        assert sd["clip.original_text_emb_weight"].shape[0] == 49408
        assert "cascaded_branch.clip.model.token_embedding.weight" in sd                 # the shared ClipModel is registered twice
        assert sd["cascaded_branch.bn_layer.bn_layer.weight"].shape == (8 * sd["clip.model.token_embedding.weight"].shape[1],)
    opts, scheds = model.configure_optimizers()
    assert len(opts) == 1 and scheds[0]["interval"] == "step"


# --------------------------------------------------------------------------------------------------------------- overlay
_STUBS = {
    "pytorch_lightning/__init__.py": """
        import torch
        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, *a, **k): pass
            def log(self, *a, **k): pass
            def log_dict(self, *a, **k): pass
        class Callback: pass
        class Trainer:
            def __init__(self, *a, **k): pass
        def seed_everything(seed): return seed
        """,
    "pytorch_lightning/callbacks.py": "class ModelCheckpoint:\n    def __init__(self, *a, **k): pass\nclass TQDMProgressBar: pass\n",
    "pytorch_lightning/loggers.py": "class LightningLoggerBase: pass\nclass WandbLogger: pass\nclass CSVLogger: pass\n",
    "clip/__init__.py": "def load(*a, **k): raise RuntimeError('stub')\ndef tokenize(*a, **k): raise RuntimeError('stub')\n",
    "clip/simple_tokenizer.py": "class SimpleTokenizer: pass\n",
    "librosa/__init__.py": "",
}


@needs_reference
def test_reference_task_runner_imports_with_this_repo_in_front(tmp_path):
    """run_task.py:6 ``from avssl import task`` with PYTHONPATH = <this repo>:<reference>.  pytorch_lightning / clip / librosa
    are absent from this image: empty stand-ins (created here, test-only) let the reference's own files import so that the
    package wiring — not those libraries — is what is exercised."""
    for rel, src in _STUBS.items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(textwrap.dedent(src))
    code = textwrap.dedent(f"""
        import sys
        from avssl import task                                   # run_task.py:6
        import avssl.model, avssl.module, avssl.base, avssl.util, avssl.data, avssl.optim
        runner = task.TrainKWClip_GeneralTransformer()           # run_task.py:16
        here, ref = {ROOT!r}, {REF!r}
        assert task.__file__.startswith(ref)
        for m in (avssl.model, avssl.module, avssl.base, avssl.optim, avssl.util, avssl.data):
            assert m.__file__.startswith(here), m.__file__
        from avssl.data import FlickrDataset, CoCoDataset, collate_general as cg   # datasets: the reference's own files
        assert sys.modules[FlickrDataset.__module__].__file__.startswith(ref) and cg.__module__ == "avssl.data.collate_function"
        assert sys.modules[cg.__module__].__file__.startswith(here)
        from avssl.task import train_KWClip
        assert train_KWClip.KWClip_GeneralTransformer is avssl.model.KWClip_GeneralTransformer
        assert train_KWClip.KWClip_GeneralTransformer.__module__ == "avssl.model.kwClip"
        from avssl.task.base_task import OrderedNamespace, collate_general, add_general_arguments, set_logging
        assert OrderedNamespace is avssl.base.OrderedNamespace and add_general_arguments.__module__ == "avssl.util.args"
        import argparse
        args = runner.add_args(argparse.ArgumentParser()).parse_args(["--config", "x.yaml"])
        assert args.seed == 7122
        # names the reference's avssl/module/__init__.py:1-7 exports
        from avssl.module import (AttentivePoolingLayer, ClipModel, FairseqSpeechEncoder_Hubert, MaskedContrastiveLoss, MeanPoolingLayer,
                                  MLPLayers, S3prlSpeechEncoderPlus, SupConLoss, WeightedSumLayer, mutualRetrieval)
        print("OK")
        """)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, REF, str(tmp_path)]))
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-3000:]


def test_package_is_a_namespace_portion():
    """No top-level ``avssl/__init__.py``: the repo's ``avssl`` must stay a namespace portion so the reference's own portion
    (``avssl/task``, ``avssl/data``) remains importable beside it."""
    assert not os.path.exists(os.path.join(ROOT, "avssl", "__init__.py"))
    from avssl.util import get_keypadding_mask  # noqa: F401
    if not os.path.isdir(REF) or REF not in sys.path:
        with pytest.raises(ImportError):
            from avssl.util import add_general_arguments  # noqa: F401


# --------------------------------------------------------------------------------------------------------------- checkpoints
def _reference_layout_ckpt(model, cfg, path, with_optimizer=True):
    """What Lightning 1.5.10 writes for the reference model (base_task.py:176-193): state_dict + pickled OrderedNamespace
    hparams (+ optimizer / scheduler state).  ``pretrained: true`` as in every shipped YAML."""
    from avssl.base import OrderedNamespace
    cfg = copy.deepcopy(cfg)
    cfg["audio_encoder"]["pretrained"] = True
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ckpt = {"epoch": 7, "global_step": 1234, "pytorch-lightning_version": "1.5.10", "state_dict": sd,
            "hyper_parameters": {"config": OrderedNamespace(cfg)}, "callbacks": {}, "lr_schedulers": [{"last_epoch": 1234}]}
    if with_optimizer:
        params = [p for p in model.getTrainableParams() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-6)
        g = torch.Generator().manual_seed(11)
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
        opt.step()
        for p in params:
            p.grad = None
        ckpt["optimizer_states"] = [opt.state_dict()]
        ckpt["state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
    torch.save(ckpt, path)
    return ckpt


@pytest.mark.parametrize("kind", ["cascaded", "parallel_large"])
def test_reference_layout_lightning_checkpoint_loads(tmp_path, kind):
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import cascaded_config, parallel_config, write_synthetic_vocab_usage
    if kind == "cascaded":
        cfg = cascaded_config("tiny", write_synthetic_vocab_usage(str(tmp_path / "vocab.npy"), n=40, vocab=96))
    else:
        cfg = parallel_config("tiny_large")
    torch.manual_seed(5)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg))
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.01 * torch.randn_like(p))
        if kind == "cascaded":
            model.cascaded_branch.bn_layer.bn_layer.running_mean.normal_()
    path = str(tmp_path / "last.ckpt")
    ckpt = _reference_layout_ckpt(model, cfg, path)
    sd = ckpt["state_dict"]
    # the layout the reference writes
    assert {"criterion.eye_mat", "criterion.neg_eye_mat", "criterion.eye_mat_fl"} <= set(sd)
    assert any(k.startswith("audio_encoder.encoder.encoder.layers.0.self_attn.q_proj.") for k in sd)
    if kind == "cascaded":
        dup = [k for k in sd if k.startswith("cascaded_branch.clip.model.")]
        assert dup and all(k[len("cascaded_branch."):] in sd for k in dup)
        assert "cascaded_branch.bn_layer.bn_layer.running_mean" in sd and "clip.original_text_emb_weight" in sd
    else:
        assert "criterion.temperature" in sd

    # pretrained: true is pickled in the config, no fairseq / openai file exists here: the .ckpt alone must be enough
    again = KWClip_GeneralTransformer.load_from_checkpoint(path)
    assert again.config.audio_encoder.pretrained is True
    sd2 = again.state_dict()
    assert list(sd2) == list(sd)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k
    # building the same config outside load_from_checkpoint still refuses to invent pretrained weights
    with pytest.raises(FileNotFoundError):
        KWClip_GeneralTransformer(again.config)

    # optimizer state written by torch.optim.Adam (what the reference's checkpoints hold) -> FusedAdam and back
    opt = again.configure_optimizers()[0][0]
    opt.load_state_dict(ckpt["optimizer_states"][0])
    out = opt.state_dict()
    ref_state = ckpt["optimizer_states"][0]["state"]
    assert set(out["state"]) == set(ref_state) and len(ref_state) == len([p for p in again.getTrainableParams() if p.requires_grad])
    for i, st in ref_state.items():
        assert int(out["state"][i]["step"]) == 2
        assert torch.equal(out["state"][i]["exp_avg"], st["exp_avg"]) and torch.equal(out["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
    assert out["param_groups"][0]["lr"] == ckpt["optimizer_states"][0]["param_groups"][0]["lr"]


def test_collate_general_matches_reference_fixture(golden):
    """avssl.data.collate_general (host code) on the rows the reference's own collate_general was run on
    (tests/golden/make_golden.py): same keys in the same order, wav_len appended, zero padding, stacking, LongTensor ids."""
    import numpy as np
    from avssl.data import collate_general, collate_packed
    z = golden("ref_collate.npz")
    rows = []
    for i in range(3):
        rows.append({"wav": torch.from_numpy(z[f"row{i}_wav"]), "image": torch.from_numpy(z[f"row{i}_image"]), "id": i * 3,
                     "text": torch.from_numpy(z[f"row{i}_text"])})
    out = collate_general(rows)
    assert list(out.keys()) == [str(k) for k in z["out_keys"]]
    for k, v in out.items():
        ref = torch.from_numpy(z["out_" + k])
        assert v.dtype == ref.dtype and torch.equal(v, ref), k
    # the packed variant carries the same information in fewer bytes
    pk = collate_packed(rows)
    assert pk["wav"].numel() == 11 + 7 + 15 and pk["wav_len"].tolist() == [11, 7, 15] and pk["wav_offset"].tolist() == [0, 11, 18]
    for i in range(3):
        assert torch.equal(pk["wav"][pk["wav_offset"][i]:pk["wav_offset"][i] + pk["wav_len"][i]], rows[i]["wav"])
    assert torch.equal(pk["image"], out["image"]) and torch.equal(pk["id"], out["id"])
