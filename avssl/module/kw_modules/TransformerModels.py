"""Branch encoders (reference: avssl/module/kw_modules/TransformerModels.py:48-135).

``TransformerEncoder`` holds the parameters of ``n_layers`` torch ``nn.TransformerEncoderLayer`` + a final LayerNorm under
the reference's state-dict keys (``model.layers.N.*``, ``model.norm.*``) and evaluates them with the sm_100a kernels.
``forward`` / ``extract_hidden_states`` compute every row (inference surface); training goes through
``KW_ParallelBranch.forward`` which only needs the [CLS] row (speechclip_b200/head.py).
"""
import logging
import math

import torch
from torch import nn

from speechclip_b200 import ops
from speechclip_b200.functional import workspace
from speechclip_b200.head import L0, PARAM_ORDER, ParallelHead
from speechclip_b200.params import ParamTree

logger = logging.getLogger(__name__)

__all__ = ["TransformerEncoder", "MultiheadAttentionAndNorm"]


def _layer_shapes(d_model, dim_feedforward, prefix):
    return {
        prefix + "self_attn.in_proj_weight": (3 * d_model, d_model), prefix + "self_attn.in_proj_bias": (3 * d_model,),
        prefix + "self_attn.out_proj.weight": (d_model, d_model), prefix + "self_attn.out_proj.bias": (d_model,),
        prefix + "linear1.weight": (dim_feedforward, d_model), prefix + "linear1.bias": (dim_feedforward,),
        prefix + "linear2.weight": (d_model, dim_feedforward), prefix + "linear2.bias": (d_model,),
        prefix + "norm1.weight": (d_model,), prefix + "norm1.bias": (d_model,),
        prefix + "norm2.weight": (d_model,), prefix + "norm2.bias": (d_model,),
    }


class TransformerEncoder(nn.Module):
    def __init__(self, n_layers: int = 1, d_model: int = 768, nhead: int = 8, dim_feedforward: int = 3072, dropout: float = 0.1,
                 activation: str = "gelu", layer_norm_eps: float = 1e-5, batch_first: bool = True, norm_first: bool = False) -> None:
        super().__init__()
        if n_layers != 1 or norm_first or activation != "gelu" or not batch_first:
            raise NotImplementedError("B200 branch encoder: 1 post-LN GELU layer, batch_first (what every shipped config uses)")
        logger.info(f"Using {n_layers} layer transformer encoder")
        shapes = _layer_shapes(d_model, dim_feedforward, "layers.0.")
        shapes["norm.weight"], shapes["norm.bias"] = (d_model,), (d_model,)
        self.model = ParamTree.from_shapes(shapes)
        self.d_model, self.nhead, self.layer_norm_eps = d_model, nhead, layer_norm_eps
        self.dropout = dropout  # applied in train mode by the branch's kernel sequence (speechclip_b200/head.py: four sites of the layer)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """torch defaults: xavier-uniform in_proj, zero attention biases, kaiming-uniform(a=sqrt 5) linears, unit LayerNorms."""
        sd = dict(self.model.named_parameters())
        nn.init.xavier_uniform_(sd["layers.0.self_attn.in_proj_weight"])
        for k in ("layers.0.self_attn.in_proj_bias", "layers.0.self_attn.out_proj.bias"):
            sd[k].zero_()
        for w, b in (("layers.0.self_attn.out_proj.weight", None), ("layers.0.linear1.weight", "layers.0.linear1.bias"),
                     ("layers.0.linear2.weight", "layers.0.linear2.bias")):
            nn.init.kaiming_uniform_(sd[w], a=math.sqrt(5))
            if b is not None:
                bound = 1 / math.sqrt(sd[w].shape[1])
                nn.init.uniform_(sd[b], -bound, bound)
        for k in ("layers.0.norm1", "layers.0.norm2", "norm"):
            sd[k + ".weight"].fill_(1.0)
            sd[k + ".bias"].zero_()

    def head_params(self, prefix: str = "self_att.") -> dict:
        """Parameters keyed the way speechclip_b200.head expects (names relative to KW_ParallelBranch)."""
        return {prefix + "model." + k: v for k, v in self.model.named_parameters()}

    def _run(self, src: torch.Tensor, key_padding_mask: torch.Tensor):
        if not src.is_cuda:
            raise RuntimeError("TransformerEncoder: CUDA tensors required (no CPU path)")
        B, L, d = src.shape
        kv_len = torch.empty(B, device=src.device, dtype=torch.int32)
        # valid keys are a prefix (get_keypadding_mask): count them
        lens = (~key_padding_mask).sum(dim=1).to(torch.int64).contiguous()
        ops.lengths_to_i32(lens, 0, L, kv_len)
        p = {k: v for k, v in self.head_params().items()}
        p["cls"] = src.new_zeros(1, 1, d)
        head = ParallelHead(d, self.nhead, self.layer_norm_eps)
        return head, p, kv_len

    @torch.no_grad()
    def forward(self, src: torch.Tensor, key_padding_mask: torch.Tensor) -> torch.Tensor:
        return self._full(src, key_padding_mask)[0]

    @torch.no_grad()
    def extract_hidden_states(self, src: torch.Tensor, key_padding_mask: torch.Tensor):
        return tuple(self._full(src, key_padding_mask)[1])

    def _full(self, src, key_padding_mask):
        head, p, kv_len = self._run(src, key_padding_mask)
        return head.full_forward_src(workspace(src.device), p, src.float().contiguous(), kv_len)


class MultiheadAttentionAndNorm(nn.Module):
    """``LayerNorm(MultiheadAttention(src, src, src) + src)`` (reference: TransformerModels.py:99-135); parameters under the
    reference's keys ``multihead_attn_layer.{in_proj_weight,in_proj_bias,out_proj.weight,out_proj.bias}`` and
    ``attentionBlock_Norm.{weight,bias}``.  Training goes through ``KW_CascadedBranch.forward`` (keyword rows only,
    speechclip_b200/cascaded.py); ``forward`` / ``extract_hidden_states`` here evaluate every row (inference surface)."""

    def __init__(self, d_model: int = 768, nhead: int = 1, dim_feedforward: int = 3072, dropout: float = 0.1, activation: str = "gelu",
                 layer_norm_eps: float = 1e-5, batch_first: bool = True, norm_first: bool = False, n_layers: int = 1, **kwargs) -> None:
        super().__init__()
        if not batch_first:
            raise NotImplementedError("MultiheadAttentionAndNorm on B200: batch_first only (every shipped config)")
        self.model_shapes = {
            "multihead_attn_layer.in_proj_weight": (3 * d_model, d_model), "multihead_attn_layer.in_proj_bias": (3 * d_model,),
            "multihead_attn_layer.out_proj.weight": (d_model, d_model), "multihead_attn_layer.out_proj.bias": (d_model,),
            "attentionBlock_Norm.weight": (d_model,), "attentionBlock_Norm.bias": (d_model,),
        }
        tree = ParamTree.from_shapes(self.model_shapes)
        self.multihead_attn_layer = tree.multihead_attn_layer
        self.attentionBlock_Norm = tree.attentionBlock_Norm
        self.d_model, self.nhead, self.layer_norm_eps = d_model, nhead, layer_norm_eps
        self.dropout = dropout  # attention dropout, applied in train mode by the keyword attention kernel (speechclip_b200/cascaded.py)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """torch defaults of nn.MultiheadAttention / nn.LayerNorm."""
        m = self.multihead_attn_layer
        nn.init.xavier_uniform_(m.in_proj_weight)
        m.in_proj_bias.zero_()
        nn.init.kaiming_uniform_(m.out_proj.weight, a=math.sqrt(5))
        m.out_proj.bias.zero_()
        self.attentionBlock_Norm.weight.fill_(1.0)
        self.attentionBlock_Norm.bias.zero_()

    def head_params(self, prefix: str = "self_att.") -> dict:
        """Parameters keyed the way speechclip_b200.cascaded expects (names relative to KW_CascadedBranch)."""
        return {prefix + k: v for k, v in self.named_parameters()}

    def _all_rows(self, src: torch.Tensor, key_padding_mask: torch.Tensor):
        from speechclip_b200.cascaded import CascadedHead
        if not src.is_cuda:
            raise RuntimeError("MultiheadAttentionAndNorm: CUDA tensors required (no CPU path)")
        B, L, d = src.shape
        kv_len = torch.empty(B, device=src.device, dtype=torch.int32)
        lens = (~key_padding_mask).sum(dim=1).to(torch.int64).contiguous()
        ops.lengths_to_i32(lens, 0, L, kv_len)
        head = CascadedHead(d, self.nhead, 0, 0, self.layer_norm_eps)
        return head.all_rows(workspace(src.device), self.head_params(), src.float().contiguous(), kv_len)

    @torch.no_grad()
    def forward(self, src: torch.Tensor, key_padding_mask: torch.Tensor) -> torch.Tensor:
        return self._all_rows(src, key_padding_mask)[1]

    @torch.no_grad()
    def extract_hidden_states(self, src: torch.Tensor, key_padding_mask: torch.Tensor):
        return tuple(self._all_rows(src, key_padding_mask))

    def extract_attention_map(self, src: torch.Tensor, key_padding_mask: torch.Tensor):
        raise NotImplementedError("attention-map visualisation (TransformerModels.py:130-135) is outside the B200 hot path")
