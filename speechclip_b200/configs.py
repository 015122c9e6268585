"""Programmatic equivalents of the reference's YAML configs (config/speechCLIP/model_{base,large}/**/spchclp_{p,c}.yaml) for the
Parallel and Cascaded SpeechCLIP paths — bench.py and the tests run where /root/reference does not exist.  Field names and values follow
the YAMLs; ``pretrained`` is off and the vocabulary reduction is disabled because neither the checkpoints nor the reference's
``*_stat/*.npy`` tables travel with this repo."""
from __future__ import annotations

import copy

_BASE = {
    "data": {"batch_size": 256, "dev_batch_size": 8},
    "model_settings": {
        "cascaded_objective_weight": 0.0,
        "parallel_objective_weight": 1.0,
        "parallel_branch": {
            "transformer_type": "TransformerEncoder",
            "transformer_args": {"n_layers": 1, "d_model": 768, "nhead": 8, "dim_feedforward": 3072, "dropout": 0.1,
                                 "activation": "gelu", "layer_norm_eps": 1.0e-5, "batch_first": True, "norm_first": False},
            "need_projection": True,
        },
        "cascaded_branch": {
            "type": "KW_CascadedBranch", "transformer_type": "MultiheadAttentionAndNorm",
            "transformer_args": {"n_layers": 1, "d_model": 768, "nhead": 1, "dim_feedforward": 3072, "dropout": 0.1,
                                 "activation": "gelu", "layer_norm_eps": 1.0e-5, "batch_first": True, "norm_first": False},
            "keyword": {"number": 8, "detokenized_K_neighbors": 5, "retrieve_method": "cosine",
                        "batchnorms": {"type": "eachKw", "std_scale": 1.0, "learnable": True, "parallel": True}},
            "vq": {"bn_before_vq": True, "activation": "gelu", "type": "SimpleVectorQuantizer",
                   "args": {"temp": "fixed=0.1", "time_first": True, "use_gumbel": False, "hard": True}},
        },
    },
    "cl_loss": {"type": "MaskedContrastiveLoss",
                "args": {"temperature": 0.07, "temperature_trainable": False, "margin": 0.0, "dcl": False, "a2b": True, "b2a": True}},
    "retrieval": {"audio_feat_src": "parallel", "recall_at": [1, 5, 10]},
    "clip": {"name": "ViT-B/32", "image_encoder_trainable": False, "text_encoder_trainable": False, "reduce_subword_embbedding": None},
    "audio_encoder": {
        "type": "FairseqHubert", "name": "hubert", "pretrained": False, "trainable": False, "feat_select_idx": "weighted_sum",
        "layer_drop": 0.0, "max_audio_len": 102400, "normalize_hiddenstates": False,
        "optim": {"name": "Adam", "args": {"lr": 1.0e-4, "weight_decay": 1.0e-6}},
        "scheduler": {"name": "linear_warmup_decay", "warmup": 5000, "max_step": 50000, "final_lr": 1.0e-8},
    },
    "trainer": {"max_steps": 50000, "gradient_clip_val": 4, "accumulate_grad_batches": 1, "precision": 16, "strategy": "dp"},
    "log_setting": {"log_detokenize_results": True},
}


def parallel_config(size: str = "base") -> dict:
    """'base' = HuBERT-base + ViT-B/32 (spchclp_p.yaml, model_base); 'large' = HuBERT-large-ll60k + ViT-L/14 with a learnable
    temperature and normalised hidden states (model_large/flickr/spchclp_p.yaml); 'tiny' / 'tiny_large' = structure-preserving
    miniatures for the parity tests."""
    c = copy.deepcopy(_BASE)
    ta = c["model_settings"]["parallel_branch"]["transformer_args"]
    if size == "base":
        return c
    if size == "large":
        c["clip"]["name"] = "ViT-L/14"
        c["audio_encoder"].update(name="hubert_large_ll60k", normalize_hiddenstates=True, normalize_type="s3prl")
        ta.update(d_model=1024, dim_feedforward=4096)
        c["cl_loss"]["args"]["temperature_trainable"] = True
        return c
    if size in ("tiny", "tiny_large"):
        c["clip"]["name"] = "tiny"
        c["audio_encoder"]["name"] = size
        ta.update(d_model=64, nhead=4, dim_feedforward=128)
        if size == "tiny_large":
            c["audio_encoder"].update(normalize_hiddenstates=True, normalize_type="s3prl")
            c["cl_loss"]["args"]["temperature_trainable"] = True
        return c
    raise KeyError(size)


def cascaded_config(size: str = "base", vocab_usage_npy: str = None) -> dict:
    """Cascaded SpeechCLIP (spchclp_c.yaml): the parallel config of the same size with the objective weights swapped and
    ``retrieval.audio_feat_src: cascaded``.  ``vocab_usage_npy``: path of a ``text_clip_vocab_usage_byfreq.npy``-style table
    (``write_synthetic_vocab_usage`` makes a stand-in of the reference's 8112-entry Flickr table)."""
    c = parallel_config(size)
    ms = c["model_settings"]
    ms["cascaded_objective_weight"], ms["parallel_objective_weight"] = 1.0, 0.0
    c["retrieval"]["audio_feat_src"] = "cascaded"
    d = ms["parallel_branch"]["transformer_args"]["d_model"]
    ms["cascaded_branch"]["transformer_args"].update(d_model=d, dim_feedforward=4 * d if size in ("base", "large") else 2 * d)
    if size in ("tiny", "tiny_large"):
        c["clip"]["name"] = "tiny_c"
    c["clip"]["reduce_subword_embbedding"] = vocab_usage_npy
    return c


def write_synthetic_vocab_usage(path: str, n: int = 8112, vocab: int = 49408) -> str:
    """A stand-in for avssl/data/flickr_stat/text_clip_vocab_usage_byfreq.npy (int64 [n, 2] = (token id, count), sorted by
    count): same length, [SOT] / [EOT] at rows 2 / 3 like the real table, so ids 0, 2, 3 are the ones the quantiser masks."""
    import numpy as np
    sot, eot = vocab - 2, vocab - 1
    ids = [0, 320 % (vocab - 2) or 1, sot, eot]
    rest = [i for i in range(1, vocab - 2) if i != ids[1]][: n - 4]
    ids = np.asarray(ids + rest, dtype=np.int64)
    counts = np.maximum(1, (2_000_000 / (1 + np.arange(len(ids))) ** 1.1)).astype(np.int64)
    np.save(path, np.stack([ids, counts], 1))
    return path
