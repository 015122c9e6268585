"""Host-side logic that runs without a GPU: config container, schedules, model construction / state-dict layout,
loud failure on CPU tensors, and the world_size-2 (gloo) feature gather + gradient routing."""
import os
import pickle
from argparse import Namespace
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avssl.base import OrderedNamespace
from avssl.optim import get_scheduler
from avssl.util import get_keypadding_mask
from speechclip_b200.configs import parallel_config


def test_ordered_namespace_semantics():
    # mirrors the reference's own test (test/test_dict.py:7-67)
    d_1 = {"a": 1, "b": [2, {"c": 3}], "d": {"e": 4, "f": "g"}, "h": SimpleNamespace(i=5, j={"k": 6})}
    ons_1, ons_2 = OrderedNamespace(d_1), OrderedNamespace(**d_1)
    assert ons_1.a == ons_1["a"] and ons_1.b == ons_1["b"] and ons_1.b[0] == 2 and ons_1.b[1].c == 3
    assert ons_1.d.e == ons_1["d"]["e"] == ons_1.d["e"] == 4
    assert ons_1.h.i == 5 and ons_1.h.j.k == 6 and ons_1 == ons_2 and len(ons_1) == 4 and len(ons_1.keys()) == 4 and "a" in ons_1
    d_2, od_1 = ons_1.pydict, ons_1.odict
    assert ons_1.keys() == d_2.keys() == od_1.keys() and d_2 == ons_1.to_dict() and od_1 == ons_1.to_odict()
    assert isinstance(d_2, dict) and isinstance(od_1, OrderedDict)
    assert OrderedNamespace({"a": 1, "b": 2}) == OrderedNamespace(SimpleNamespace(a=1, b=2)) == OrderedNamespace(Namespace(a=1, b=2))
    ons_6 = OrderedNamespace([{"a": 1, "b": 2, "c": {"d": 3}}, Namespace(e=4, f=SimpleNamespace(g=5))])
    assert (ons_6.a, ons_6.c.d, ons_6.e, ons_6.f.g, ons_6.f["g"]) == (1, 3, 4, 5, 5)
    # checkpoint format: the pickle round-trips through __getstate__/__setstate__ (base_model.py:15)
    back = pickle.loads(pickle.dumps(ons_1))
    assert back == ons_1 and back.h.j.k == 6
    assert ons_1.get("zz", 7) == 7 and not hasattr(ons_1, "zz")
    assert dict(**ons_1.d) == {"e": 4, "f": "g"}


def test_scheduler_matches_reference_fixture(golden):
    lrs = golden("ref_scheduler.npz")["lrs"]
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
    s = get_scheduler("linear_warmup_decay", opt, warmup=5, max_step=20, final_lr=1e-8)
    mine = []
    for _ in range(20):
        mine.append(opt.param_groups[0]["lr"])
        opt.step()
        s.step()
    assert np.allclose(mine, lrs, rtol=1e-12, atol=0)
    with pytest.raises(NotImplementedError):
        get_scheduler("cosine", opt)


def test_keypadding_mask_matches_reference_fixture(golden):
    z = golden("ref_keypad.npz")
    assert torch.equal(get_keypadding_mask(12, torch.from_numpy(z["lens"])), torch.from_numpy(z["mask"]))


def test_model_builds_with_reference_state_dict_layout():
    from avssl.model import KWClip_GeneralTransformer
    m = KWClip_GeneralTransformer(OrderedNamespace(parallel_config("base")))
    keys = set(m.state_dict().keys())
    for k in ("audio_encoder.encoder.feature_extractor.conv_layers.0.0.weight", "audio_encoder.encoder.feature_extractor.conv_layers.0.2.weight",
              "audio_encoder.encoder.post_extract_proj.weight", "audio_encoder.encoder.encoder.pos_conv.0.weight_g",
              "audio_encoder.encoder.encoder.layers.11.self_attn.q_proj.weight", "audio_encoder.encoder.encoder.layers.0.final_layer_norm.bias",
              "audio_encoder.encoder.mask_emb", "audio_encoder.encoder.label_embs_concat", "audio_encoder.weightedsum_layer.weights",
              "clip.model.visual.conv1.weight", "clip.model.visual.transformer.resblocks.11.attn.in_proj_weight", "clip.model.visual.proj",
              "clip.model.token_embedding.weight", "clip.model.text_projection", "clip.model.logit_scale",
              "criterion.eye_mat", "criterion.neg_eye_mat", "criterion.eye_mat_fl", "parallel_branch.cls",
              "parallel_branch.self_att.model.layers.0.self_attn.in_proj_weight", "parallel_branch.self_att.model.layers.0.linear2.bias",
              "parallel_branch.self_att.model.norm.weight", "parallel_branch.linear_proj.weight"):
        assert k in keys, k
    trainable = sum(p.numel() for p in m.getTrainableParams())
    assert trainable == 7089408 + 768 + 768 * 512 + 512 + 13  # encoder layer + norm, cls, linear_proj, layer weights
    assert all(not p.requires_grad for p in m.audio_encoder.encoder.parameters())
    assert all(not p.requires_grad for p in m.clip.parameters())
    assert m.audio_encoder.out_dim == 768 and m.audio_encoder.downsample_rate == 320 and m.audio_encoder.upstream_model_hiddenstates_len == 13
    assert m.subword_embd_dim == 512 and m.recall_at == [1, 5, 10]
    large = KWClip_GeneralTransformer(OrderedNamespace(parallel_config("tiny_large")))
    assert "criterion.temperature" in large.state_dict() and large.criterion.temperature.requires_grad


def test_unsupported_configs_fail_loudly():
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import cascaded_config
    cfg = cascaded_config("tiny")
    cfg["model_settings"]["cascaded_branch"]["transformer_type"] = "TransformerEncoder"
    with pytest.raises(NotImplementedError):
        KWClip_GeneralTransformer(OrderedNamespace(cfg))
    cfg = cascaded_config("tiny")
    cfg["model_settings"]["cascaded_branch"]["vq"]["args"]["use_gumbel"] = True
    with pytest.raises(NotImplementedError):
        KWClip_GeneralTransformer(OrderedNamespace(cfg))
    cfg = cascaded_config("tiny")
    cfg["model_settings"]["cascaded_branch"]["keyword"]["batchnorms"]["parallel"] = False
    with pytest.raises(NotImplementedError):
        KWClip_GeneralTransformer(OrderedNamespace(cfg))
    cfg = parallel_config("tiny")
    cfg["audio_encoder"]["trainable"] = True
    with pytest.raises(NotImplementedError):
        KWClip_GeneralTransformer(OrderedNamespace(cfg))
    cfg = parallel_config("tiny")
    cfg["audio_encoder"]["pretrained"] = True
    with pytest.raises(FileNotFoundError):
        KWClip_GeneralTransformer(OrderedNamespace(cfg))


def test_cascaded_model_mirrors_reference_structure(tmp_path):
    """State-dict keys / parameter counts of the cascaded configuration (kwClip.py:697-826, spchclp_c.yaml)."""
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import cascaded_config, write_synthetic_vocab_usage
    npy = write_synthetic_vocab_usage(str(tmp_path / "usage.npy"))
    m = KWClip_GeneralTransformer(OrderedNamespace(cascaded_config("base", npy)))
    assert m.parallel_branch is None and m.cascaded_branch is not None
    keys = set(m.state_dict().keys())
    for k in ("cascaded_branch.cls", "cascaded_branch.self_att.multihead_attn_layer.in_proj_weight",
              "cascaded_branch.self_att.multihead_attn_layer.out_proj.bias", "cascaded_branch.self_att.attentionBlock_Norm.weight",
              "cascaded_branch.linear_proj.weight", "cascaded_branch.bn_layer.bn_layer.weight", "cascaded_branch.bn_layer.bn_layer.running_var",
              "cascaded_branch.bn_layer.bn_layer.num_batches_tracked", "cascaded_branch.vector_quantizer.curr_temp",
              "cascaded_branch.clip.model.token_embedding.weight", "clip.model.token_embedding.weight"):
        assert k in keys, k
    cb = m.cascaded_branch
    assert cb.cls.shape == (1, 8, 768) and m.clip.model.token_embedding.weight.shape == (8112, 512)
    assert cb.bn_layer.bn_layer.weight.shape == (8 * 512,)
    emb = m.clip.model.token_embedding.weight
    torch.testing.assert_close(cb.bn_layer.bn_layer.bias.data, emb.mean(0).repeat(8))
    torch.testing.assert_close(cb.bn_layer.bn_layer.weight.data, emb.std(0).repeat(8))
    assert m.clip.special_tokens() == (2, 3) and cb.vector_quantizer.temperature() == pytest.approx(0.1)
    # trainable: weighted-sum (13) + cls + MHA (4 d^2 + 4 d) + LN (2 d) + linear_proj + BatchNorm affine
    d = 768
    want = 13 + 8 * d + 4 * d * d + 4 * d + 2 * d + d * 512 + 512 + 2 * 8 * 512
    assert sum(p.numel() for p in m.getTrainableParams() if p.requires_grad) == want
    assert all(not p.requires_grad for p in m.clip.parameters())


def test_product_path_has_no_cpu_fallback():
    from avssl.model import KWClip_GeneralTransformer
    from avssl.module import MaskedContrastiveLoss, mutualRetrieval
    m = KWClip_GeneralTransformer(OrderedNamespace(parallel_config("tiny")))
    b = {"wav": torch.randn(2, 4000), "wav_len": torch.tensor([4000, 3000]), "image": torch.randn(2, 3, 32, 32), "id": torch.arange(2)}
    with pytest.raises(RuntimeError, match="CUDA"):
        m.training_step(b)
    from speechclip_b200.configs import cascaded_config
    mc = KWClip_GeneralTransformer(OrderedNamespace(cascaded_config("tiny")))
    with pytest.raises(RuntimeError, match="CUDA"):
        mc.training_step(b)
    with pytest.raises(RuntimeError, match="CUDA"):
        mc.clip.encode_text(torch.zeros(2, 8, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        mc.cascaded_branch.vector_quantizer(torch.randn(2, 8, 96))
    with pytest.raises(RuntimeError, match="CUDA"):
        MaskedContrastiveLoss()(torch.randn(4, 8), torch.randn(4, 8), torch.arange(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        mutualRetrieval(torch.randn(4, 2), torch.randn(2, 4), torch.arange(4) // 2, torch.arange(2), [1])


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import avssl.model, speechclip_b200.engine, speechclip_b200.head, speechclip_b200.optim; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported by the product'")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root)


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avssl.model.kwClip import gather_features
    torch.manual_seed(rank)
    a = torch.randn(3, 4, requires_grad=True)
    feats = gather_features({"id": torch.arange(3) + 10 * rank, "image_feat": torch.randn(3, 4), "parallel_audio_feat": a, "tag": 1.0})
    assert feats["parallel_audio_feat"].shape == (3 * world, 4) and feats["id"].tolist() == [0, 1, 2, 10, 11, 12]
    w = torch.arange(3 * world * 4, dtype=torch.float32).view(3 * world, 4)
    (feats["parallel_audio_feat"] * w).sum().backward()  # every rank evaluates the same global objective
    assert torch.equal(a.grad, w[rank * 3:(rank + 1) * 3])  # ... and receives exactly the gradient rows of its own slice
    q.put((rank, feats["parallel_audio_feat"].detach().clone()))
    dist.destroy_process_group()


def test_gather_features_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(got[0], got[1])  # identical global feature matrix on both ranks


def _allreduce_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avssl.model import KWClip_GeneralTransformer
    torch.manual_seed(0)
    model = KWClip_GeneralTransformer(OrderedNamespace(parallel_config("tiny_large")))  # learnable temperature (model_large YAMLs)
    params = [p for p in model.getTrainableParams() if p.requires_grad]
    for p in params:
        p.grad = torch.full_like(p, float(rank + 1))
    model.on_after_backward()
    t = model.criterion.temperature
    others = [p for p in params if p is not t]
    q.put((rank, float(t.grad), sorted({float(p.grad.flatten()[0]) for p in others}), all(bool((p.grad == p.grad.flatten()[0]).all()) for p in others)))
    dist.destroy_process_group()


def test_gradient_allreduce_counts_the_temperature_once_world_size_2_gloo():
    """every rank differentiates the same GLOBAL loss, so each holds the FULL gradient of the criterion's
    learnable temperature while the branch gradients are per-rank partial sums: after ``on_after_backward`` the branch gradients
    are the sum over ranks and the temperature gradient is counted once (rank 0's), as in the reference's single-process DP."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, t_grad, other_vals, uniform in got:
        assert t_grad == 1.0, (rank, t_grad)          # rank 0 contributed 1, rank 1 contributed 0
        assert other_vals == [3.0] and uniform, (rank, other_vals)   # 1 + 2


def test_gelu_h16_fit_constants_in_the_cuda_header():
    """The sigmoid-polynomial erf-GELU of the 16-bit GEMM / conv0 epilogues (csrc/common.cuh: gelu_h16): the constants compiled
    into the kernels keep |error| <= 3e-5 on the GELU value over the whole real line, including beyond the clamp."""
    import math
    import os
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speechclip_b200", "csrc", "common.cuh")).read()
    body = src[src.index("__device__ __forceinline__ float gelu_h16(float x)"):]
    body = body[:body.index("}")]
    c2, c1, c0 = (float(v) for v in re.findall(r"(-?\d+\.\d+)f \* -1\.4426950408889634f", body))
    clamp = float(re.search(r"fminf\(x \* x, (\d+\.\d+)f\)", body).group(1))
    x = np.linspace(-40.0, 40.0, 400001)
    s = np.minimum(x * x, clamp)
    approx = x / (1.0 + np.exp(-x * (c0 + c1 * s + c2 * s * s)))
    ref = np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in x])
    assert np.abs(approx - ref).max() < 3e-5
    assert c0 + c1 * clamp + c2 * clamp * clamp > 3.0  # beyond the clamp the argument keeps growing with |x|: Phi saturates at 0 / 1


def test_conv0_layernorm_statistics_from_the_gram_matrix():
    """The identity behind the HuBERT-large conv0 kernel (csrc/frontend.cu: conv0_ln_apply_kernel): the LayerNorm mean and
    variance of a frame's 512 conv outputs follow from its 10 samples and 77 channel sums of the weights."""
    import numpy as np
    rng = np.random.default_rng(0)
    C, K = 512, 10
    w, b = rng.normal(0, 0.45, (C, K)), rng.normal(0, 0.3, C)
    x = rng.normal(0, 1.0, (200, K))
    o = x @ w.T + b
    gram, wbar, wb = w.T @ w / C, w.mean(0), w.T @ b / C
    mean = x @ wbar + b.mean()
    e2 = np.einsum("fk,kj,fj->f", x, gram, x) + 2.0 * x @ wb + (b * b).mean()
    assert np.abs(mean - o.mean(1)).max() < 1e-12
    assert np.abs((e2 - mean * mean) - o.var(1)).max() < 1e-10


def _holder_from(state):
    """nn.Module tree with exactly these (dotted) parameter names — what a TorchScript archive of a model carries."""
    from torch import nn

    class Holder(nn.Module):
        pass
    root = Holder()
    for name, v in state.items():
        m = root
        parts = name.split(".")
        for p in parts[:-1]:
            if not hasattr(m, p):
                setattr(m, p, Holder())
            m = getattr(m, p)
        setattr(m, parts[-1], nn.Parameter(v.clone(), requires_grad=False))
    return root


def test_clip_checkpoint_formats(tmp_path):
    """`ClipModel(ckpt_path=...)` takes openai's released format (a TorchScript archive, fp16 weights, extra scalars) and a plain
    state-dict pickle; every parameter arrives (strict load over the module's own openai-named keys)."""
    from avssl.module import ClipModel
    from oracle import clip as oc
    from speechclip_b200.init import seeded_init_
    src = seeded_init_(oc.CLIP(oc.ClipCfg.named("tiny")), 99).state_dict()  # openai key names
    fp16 = {k: v.half() if v.is_floating_point() and v.dim() > 0 else v for k, v in src.items()}
    jit_path, sd_path = str(tmp_path / "clip_jit.pt"), str(tmp_path / "clip_sd.pt")
    torch.jit.script(_holder_from(fp16)).save(jit_path)
    torch.save({"state_dict": dict(fp16, input_resolution=torch.tensor(32), context_length=torch.tensor(77))}, sd_path)
    for path in (jit_path, sd_path):
        m = ClipModel("tiny", ckpt_path=path)
        got = m.model.state_dict()
        assert set(got) <= set(src)
        for k, v in got.items():
            assert v.dtype == torch.float32 and torch.equal(v, fp16[k].float()), k


def test_hubert_checkpoint_fairseq_layout(tmp_path):
    """`FairseqSpeechEncoder_Hubert(pretrained=True, ckpt_path=...)` reads fairseq's checkpoint layout ({"model": state_dict, "cfg": ...},
    fairseq key names — the layout torchaudio's fairseq importer pins in tests/test_oracle_crosscheck.py) and refuses a missing file."""
    from avssl.module import FairseqSpeechEncoder_Hubert
    from oracle import hubert as oh
    from speechclip_b200.init import seeded_init_
    src = seeded_init_(oh.HubertModel(oh.HubertCfg.named("tiny")), 5).state_dict()
    path = str(tmp_path / "hubert_tiny.pt")
    torch.save({"model": src, "cfg": {"model": {"_name": "hubert"}}, "args": None}, path)
    enc = FairseqSpeechEncoder_Hubert("tiny", pretrained=True, ckpt_path=path, feat_select_idx="weighted_sum")
    got = enc.encoder.state_dict()
    assert set(got) <= set(src)
    for k, v in got.items():
        assert torch.equal(v, src[k].float()), k
    with pytest.raises(FileNotFoundError):
        FairseqSpeechEncoder_Hubert("tiny", pretrained=True, ckpt_path=str(tmp_path / "absent.pt"))
    # a real fairseq file pickles omegaconf / fairseq objects in "cfg": the model weights must load without those packages
    import sys
    import types
    fake = types.ModuleType("scb_fake_omegaconf")

    class DictConfig(dict):
        pass
    DictConfig.__module__, DictConfig.__qualname__ = "scb_fake_omegaconf", "DictConfig"
    fake.DictConfig = DictConfig
    sys.modules["scb_fake_omegaconf"] = fake
    path2 = str(tmp_path / "hubert_tiny_cfg.pt")
    try:
        torch.save({"model": src, "cfg": DictConfig(model="hubert")}, path2)
    finally:
        del sys.modules["scb_fake_omegaconf"]
    enc2 = FairseqSpeechEncoder_Hubert("tiny", pretrained=True, ckpt_path=path2)
    for k, v in enc2.encoder.state_dict().items():
        assert torch.equal(v, src[k].float()), k


def test_lightning_checkpoint_roundtrip(tmp_path):
    """`KWClip_GeneralTransformer.load_from_checkpoint` (example.py:10; base_task.py:60-77): a Lightning-style file
    ({"state_dict", "hyper_parameters": {"config": OrderedNamespace}}) rebuilds the model from the pickled config and restores
    every tensor; the tower keys carry the fairseq / openai names under `audio_encoder.encoder.` / `clip.model.`."""
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import parallel_config
    torch.manual_seed(3)
    model = KWClip_GeneralTransformer(OrderedNamespace(parallel_config("tiny")))
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.01 * torch.randn_like(p))
    sd = model.state_dict()
    assert any(k.startswith("audio_encoder.encoder.feature_extractor.conv_layers.0.0.weight") for k in sd)
    assert any(k.startswith("clip.model.visual.transformer.resblocks.0.attn.in_proj_weight") for k in sd)
    path = str(tmp_path / "model.ckpt")
    torch.save({"state_dict": sd, "hyper_parameters": {"config": model.config}, "epoch": 3, "global_step": 120}, path)
    again = KWClip_GeneralTransformer.load_from_checkpoint(path)
    sd2 = again.state_dict()
    assert list(sd2) == list(sd)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k


def test_tower_pipeline_order_and_slots_without_a_gpu():
    """runtime.TowerPipeline: the towers of batch i + 1 are launched BEFORE batch i is handed to the loop body, tower buffer slots
    alternate 1 / 2, every batch carries its own handle, and a prefetcher without a look-ahead slot is refused."""
    from speechclip_b200.runtime import TowerPipeline

    log = []

    class FakeModel:
        def precompute_towers(self, batch, slot=0):
            log.append(("launch", batch["i"], slot))
            return {"for": batch["i"], "slot": slot}

    pipe = TowerPipeline(FakeModel(), head_priority=False)
    seen = []
    for b in pipe.iterate({"i": i} for i in range(5)):
        log.append(("step", b["i"]))
        assert b["_scb_towers"]["for"] == b["i"]
        seen.append((b["i"], b["_scb_towers"]["slot"]))
    assert seen == [(0, 1), (1, 2), (2, 1), (3, 2), (4, 1)]
    order = [e[:2] for e in log]
    for i in range(4):   # launch(i + 1) precedes step(i)
        assert order.index(("launch", i + 1)) < order.index(("step", i))
    assert order.index(("launch", 0)) == 0 and order[-1] == ("step", 4)
    assert list(pipe.iterate(iter(()))) == []   # empty epoch

    class NoLookahead:
        lookahead = 0

        def __iter__(self):
            return iter(())

    import pytest
    with pytest.raises(ValueError, match="lookahead=1"):
        list(TowerPipeline(FakeModel(), head_priority=False).iterate(NoLookahead()))
