"""Build ``transformers`` Hubert / CLIP models carrying the SAME weights as an oracle module.

Used to cross-check the oracle's restatement of the absent third-party towers (fairseq HuBERT,
openai CLIP) against an independent implementation that ships in this image.
"""
import torch
import transformers


def hf_hubert_from_oracle(om):
    c = om.cfg
    large = c.extractor_mode == "layer_norm"
    cfg = transformers.HubertConfig(
        hidden_size=c.embed_dim, num_hidden_layers=c.layers, num_attention_heads=c.heads,
        intermediate_size=c.ffn_dim, num_conv_pos_embeddings=c.pos_kernel,
        num_conv_pos_embedding_groups=c.pos_groups, feat_extract_norm="layer" if large else "group",
        do_stable_layer_norm=c.layer_norm_first, conv_bias=c.conv_bias, hidden_act="gelu",
        hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
        layerdrop=0.0, feat_proj_layer_norm=True, apply_spec_augment=False, attn_implementation="eager")
    hf = transformers.HubertModel(cfg).eval()
    src = om.state_dict()
    dst = {}
    for i in range(7):
        dst[f"feature_extractor.conv_layers.{i}.conv.weight"] = src[f"feature_extractor.conv_layers.{i}.0.weight"]
        if c.conv_bias:
            dst[f"feature_extractor.conv_layers.{i}.conv.bias"] = src[f"feature_extractor.conv_layers.{i}.0.bias"]
        if large:
            for wb in ("weight", "bias"):
                dst[f"feature_extractor.conv_layers.{i}.layer_norm.{wb}"] = src[f"feature_extractor.conv_layers.{i}.2.1.{wb}"]
    if not large:
        for wb in ("weight", "bias"):
            dst[f"feature_extractor.conv_layers.0.layer_norm.{wb}"] = src[f"feature_extractor.conv_layers.0.2.{wb}"]
    for wb in ("weight", "bias"):
        dst[f"feature_projection.layer_norm.{wb}"] = src[f"layer_norm.{wb}"]
        dst[f"feature_projection.projection.{wb}"] = src[f"post_extract_proj.{wb}"]
        dst[f"encoder.layer_norm.{wb}"] = src[f"encoder.layer_norm.{wb}"]
    dst["encoder.pos_conv_embed.conv.bias"] = src["encoder.pos_conv.0.bias"]
    dst["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = src["encoder.pos_conv.0.weight_g"]
    dst["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = src["encoder.pos_conv.0.weight_v"]
    for l in range(c.layers):
        for wb in ("weight", "bias"):
            for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
                dst[f"encoder.layers.{l}.attention.{p}.{wb}"] = src[f"encoder.layers.{l}.self_attn.{p}.{wb}"]
            dst[f"encoder.layers.{l}.layer_norm.{wb}"] = src[f"encoder.layers.{l}.self_attn_layer_norm.{wb}"]
            dst[f"encoder.layers.{l}.feed_forward.intermediate_dense.{wb}"] = src[f"encoder.layers.{l}.fc1.{wb}"]
            dst[f"encoder.layers.{l}.feed_forward.output_dense.{wb}"] = src[f"encoder.layers.{l}.fc2.{wb}"]
            dst[f"encoder.layers.{l}.final_layer_norm.{wb}"] = src[f"encoder.layers.{l}.final_layer_norm.{wb}"]
    missing, unexpected = hf.load_state_dict(dst, strict=False)
    assert not unexpected, unexpected
    assert set(missing) <= {"masked_spec_embed"}, missing
    return hf


def _clip_tower(dst, src, hf_prefix, o_prefix, layers, width):
    for l in range(layers):
        hp, op = f"{hf_prefix}.encoder.layers.{l}", f"{o_prefix}.resblocks.{l}"
        w, b = src[f"{op}.attn.in_proj_weight"], src[f"{op}.attn.in_proj_bias"]
        for j, p in enumerate(("q_proj", "k_proj", "v_proj")):
            dst[f"{hp}.self_attn.{p}.weight"] = w[j * width:(j + 1) * width]
            dst[f"{hp}.self_attn.{p}.bias"] = b[j * width:(j + 1) * width]
        for wb in ("weight", "bias"):
            dst[f"{hp}.self_attn.out_proj.{wb}"] = src[f"{op}.attn.out_proj.{wb}"]
            dst[f"{hp}.layer_norm1.{wb}"] = src[f"{op}.ln_1.{wb}"]
            dst[f"{hp}.layer_norm2.{wb}"] = src[f"{op}.ln_2.{wb}"]
            dst[f"{hp}.mlp.fc1.{wb}"] = src[f"{op}.mlp.c_fc.{wb}"]
            dst[f"{hp}.mlp.fc2.{wb}"] = src[f"{op}.mlp.c_proj.{wb}"]


def hf_clip_from_oracle(om):
    c = om.cfg
    src = om.state_dict()
    vcfg = transformers.CLIPVisionConfig(hidden_size=c.v_width, intermediate_size=4 * c.v_width,
                                         num_hidden_layers=c.v_layers, num_attention_heads=c.v_heads,
                                         image_size=c.image_size, patch_size=c.patch, projection_dim=c.embed_dim,
                                         hidden_act="quick_gelu", attention_dropout=0.0, attn_implementation="eager")
    hv = transformers.CLIPVisionModelWithProjection(vcfg).eval()
    d = {"vision_model.embeddings.class_embedding": src["visual.class_embedding"],
         "vision_model.embeddings.patch_embedding.weight": src["visual.conv1.weight"],
         "vision_model.embeddings.position_embedding.weight": src["visual.positional_embedding"],
         "visual_projection.weight": src["visual.proj"].t().contiguous()}
    for wb in ("weight", "bias"):
        d[f"vision_model.pre_layrnorm.{wb}"] = src[f"visual.ln_pre.{wb}"]
        d[f"vision_model.post_layernorm.{wb}"] = src[f"visual.ln_post.{wb}"]
    _clip_tower(d, src, "vision_model", "visual.transformer", c.v_layers, c.v_width)
    missing, unexpected = hv.load_state_dict(d, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)

    tcfg = transformers.CLIPTextConfig(vocab_size=c.vocab, hidden_size=c.t_width, intermediate_size=4 * c.t_width,
                                       num_hidden_layers=c.t_layers, num_attention_heads=c.t_heads,
                                       max_position_embeddings=c.context, projection_dim=c.embed_dim,
                                       hidden_act="quick_gelu", attention_dropout=0.0, eos_token_id=2,
                                       attn_implementation="eager")
    ht = transformers.CLIPTextModelWithProjection(tcfg).eval()
    d = {"text_model.embeddings.token_embedding.weight": src["token_embedding.weight"],
         "text_model.embeddings.position_embedding.weight": src["positional_embedding"],
         "text_projection.weight": src["text_projection"].t().contiguous()}
    for wb in ("weight", "bias"):
        d[f"text_model.final_layer_norm.{wb}"] = src[f"ln_final.{wb}"]
    _clip_tower(d, src, "text_model", "transformer", c.t_layers, c.t_width)
    missing, unexpected = ht.load_state_dict(d, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    return hv, ht
