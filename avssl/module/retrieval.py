"""Reference: avssl/module/retrieval.py:6-121 — recall@k both directions.  The reference argsorts every score row on the
device, moves the permutation to the host and walks it row by row in python; here one kernel (``scb_retrieval_rank``)
finds, per query, how many candidates beat the best-scoring correct answer, and recall@k = mean(rank < k)."""
from typing import Tuple

import torch

from speechclip_b200 import ops


def _recalls(score: torch.Tensor, cand_ids: torch.Tensor, answers: torch.Tensor, recall_at, title: str) -> dict:
    rows, cols = score.shape
    rank = torch.empty(rows, device=score.device, dtype=torch.int32)
    ops.retrieval_rank(score, cand_ids, answers, rank, None)
    rank = rank.cpu()
    out = {}
    for k in recall_at:
        if k > cols:
            print("recall@{} is not eligible for #{} {} samples".format(k, cols, title))
        out["recall@{}".format(k)] = 100.0 * float((rank < min(k, cols)).sum().item()) / rows
    return out


def mutualRetrieval(score_per_A: torch.Tensor, score_per_B: torch.Tensor, AB_answers: torch.Tensor, BA_answers: torch.Tensor,
                    recall_at: list, modality_A_title: str = "audio", modality_B_title: str = "image") -> Tuple[dict, dict, dict]:
    assert score_per_A.dim() == 2 and score_per_B.dim() == 2 and AB_answers.dim() == 1 and BA_answers.dim() == 1
    assert score_per_A.shape == (len(AB_answers), len(BA_answers)), "{} , {}".format(score_per_A.shape, (len(AB_answers), len(BA_answers)))
    assert score_per_B.shape == (len(BA_answers), len(AB_answers)), "{} , {}".format(score_per_B.shape, (len(BA_answers), len(AB_answers)))
    if not score_per_A.is_cuda:
        raise RuntimeError("mutualRetrieval: CUDA score matrices required (no CPU path)")
    dev = score_per_A.device
    ab = AB_answers.to(device=dev, dtype=torch.int64).contiguous()
    ba = BA_answers.to(device=dev, dtype=torch.int64).contiguous()
    res_ab = _recalls(score_per_A.float().contiguous(), ba, ab, recall_at, modality_B_title)
    res_ba = _recalls(score_per_B.float().contiguous(), ab, ba, recall_at, modality_A_title)
    mean = {k: (res_ab[k] + res_ba[k]) / 2.0 for k in res_ab}
    return res_ab, res_ba, mean
