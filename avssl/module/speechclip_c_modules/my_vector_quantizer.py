"""SimpleVectorQuantizer (reference: avssl/module/speechclip_c_modules/my_vector_quantizer.py:12-164).

Hard straight-through quantiser over the vocabulary scores: special ids masked, argmax one-hot in the forward value,
softmax(x / temp) in the backward.  ``KW_CascadedBranch.forward`` uses the fused kernels directly and never materialises the
dense [B, K, V] ``subword_prob``; this module is the standalone surface with the reference's result dictionary."""
import logging

import torch
import torch.nn as nn

from speechclip_b200 import ops
from speechclip_b200.functional import OneHotSTFn, VectorQuantizeFn

logger = logging.getLogger(__name__)

__all__ = ["SimpleVectorQuantizer"]


def vq_statistics(cos: torch.Tensor, idx: torch.Tensor, stats: torch.Tensor, bsz: int, tsz: int) -> dict:
    """code / probability perplexity, per-keyword entropy and diversity loss (my_vector_quantizer.py:84-118,146-158)."""
    R, V = cos.shape
    dev = cos.device
    hist, avg, ent = torch.zeros(V, device=dev), torch.zeros(V, device=dev), torch.empty(R, device=dev)
    ops.vq_diagnostics(cos, stats, idx, hist, avg, ent)
    hard_probs, avg_probs = hist / R, avg / R
    out = {"num_vars": V}
    out["code_perplexity"] = torch.exp(-torch.sum(hard_probs * torch.log(hard_probs + 1e-7), dim=-1)).sum()
    out["ent_per_t"] = ent.view(bsz, tsz).mean(dim=0)
    out["prob_perplexity"] = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-7), dim=-1)).sum()
    return out


class SimpleVectorQuantizer(nn.Module):
    def __init__(self, temp, groundTruthPerplexity=None, time_first=True, use_gumbel=False, hard=True):
        super().__init__()
        self.time_first, self.use_gumbel, self.hard = time_first, use_gumbel, hard
        if use_gumbel or not hard:
            raise NotImplementedError("SimpleVectorQuantizer on B200: hard softmax straight-through only (use_gumbel=false, hard=true)")
        if isinstance(temp, str):
            import ast
            if temp.startswith("learnable="):
                raise NotImplementedError("learnable VQ temperature: no shipped config uses it")
            elif temp.startswith("fixed="):
                self.temp_type = "fixed"
                temp = ast.literal_eval(temp.replace("fixed=", ""))
                self.register_buffer("curr_temp", torch.FloatTensor([temp]))
                self._temp_value = float(temp)
                logger.info("Setting vq temp fixed={}".format(temp))
            else:
                self.temp_type = "scheduled"
                temp = ast.literal_eval(temp)
                assert len(temp) == 3, f"{temp}, {len(temp)}"
                self.max_temp, self.min_temp, self.temp_decay = temp
                logger.info("Setting vq temp scheduled = ({},{},{})".format(*temp))
                self.curr_temp = self.max_temp
        self.codebook_indices = None
        self.groundTruthPerplexity = groundTruthPerplexity

    def temperature(self) -> float:
        """Host value of the temperature (no device sync: the fixed value is cached at construction / load)."""
        if self.temp_type == "fixed":
            return self._temp_value
        return float(self.curr_temp)

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        super()._load_from_state_dict(state_dict, prefix, *a, **k)
        if getattr(self, "temp_type", None) == "fixed" and prefix + "curr_temp" in state_dict:
            self._temp_value = float(state_dict[prefix + "curr_temp"].reshape(-1)[0])

    def set_num_updates(self, num_updates):
        if self.temp_type == "scheduled":
            self.curr_temp = max(self.max_temp * self.temp_decay ** num_updates, self.min_temp)

    def results(self, cos, idx, stats, bsz, tsz, subword_prob=None, produce_targets=True) -> dict:
        """The reference's result dictionary from the fused kernels' outputs."""
        result = vq_statistics(cos, idx, stats, bsz, tsz)
        result["temp"] = self.temperature()
        if subword_prob is not None:
            result["subword_prob"] = subword_prob
        if self.groundTruthPerplexity is not None:
            gt = torch.tensor(float(self.groundTruthPerplexity), device=cos.device)
            result["diversity_loss"] = (result["prob_perplexity"] - gt) ** 2 / (result["num_vars"] - self.groundTruthPerplexity) ** 2
        else:
            result["diversity_loss"] = (result["num_vars"] - result["prob_perplexity"]) / result["num_vars"]
        if produce_targets:
            result["targets"] = idx.view(bsz, tsz, 1).detach()
        return result

    def forward(self, x, prob_msk=[0, 2, 3], produce_targets=True):
        if not x.is_cuda:
            raise RuntimeError("SimpleVectorQuantizer: CUDA tensors required (no CPU path)")
        if not self.time_first:
            x = x.transpose(1, 2)
        bsz, tsz, fsz = x.shape
        x = x.float()
        mask = torch.tensor(list(prob_msk), device=x.device, dtype=torch.int32) if len(prob_msk) else None
        temp = self.temperature()
        cos, idx, stats = VectorQuantizeFn.apply(x.detach(), mask, temp)
        if self.training:
            prob = OneHotSTFn.apply(x, cos, idx, stats, temp)
        else:
            prob = torch.zeros(bsz * tsz, fsz, device=x.device).scatter_(1, idx.view(-1, 1), 1.0).view(bsz, tsz, fsz)
        return self.results(cos, idx, stats, bsz, tsz, subword_prob=prob, produce_targets=produce_targets)
