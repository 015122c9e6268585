"""Execution plans for the SpeechCLIP hot path on one B200: weight preparation + kernel sequencing.

Nothing here does arithmetic on activations with torch: torch provides device memory, the current stream
and (at weight-load time only) layout conversion of parameters.  Every activation op is a call into
libspeechclip_b200.so through ``speechclip_b200.ops``.

Towers (all frozen in every shipped config):

* ``HubertPlan``   fairseq HuBERT as driven by ``avssl/module/speech_encoder_plus.py:29-107``
                   conv0(+GroupNorm | +LayerNorm)+GELU -> conv1..6 as tap-walk GEMMs -> LN -> post_extract_proj
                   -> masked-zero + grouped positional conv GEMM (+GELU +residual) -> [LN] -> L encoder layers;
                   returns the L+1 hidden states.
* ``VitPlan``      openai CLIP VisionTransformer as driven by ``avssl/module/clip_official.py:200-209``.
* ``EncoderLayerPlan``  one post-/pre-LN transformer layer on [B, T, d] (fairseq layer, CLIP block, torch
                   nn.TransformerEncoderLayer) — QKV GEMM -> attention -> out-proj(+residual) -> LN -> MLP.

Activations are IEEE fp16 (like the reference's ``precision: 16`` autocast) with fp32 accumulation, fp32
residual streams / LayerNorm statistics / softmax.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops

CONV_SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2  # fairseq conv_feature_layers
H = torch.float16
SLACK = 2048  # elements of tail padding on conv activations: the last pair-row of an odd-length sequence is over-read


class Workspace:
    """Persistent scratch buffers keyed by name (re-allocated only when the requested size grows)."""

    def __init__(self, device):
        self.device = device
        self._buf: Dict[str, torch.Tensor] = {}

    def get(self, name: str, numel: int, dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(numel, device=self.device, dtype=dtype)
            self._buf[name] = t
        return t[:numel]

    def view(self, name: str, shape, dtype, zero: bool = False) -> torch.Tensor:
        return self.get(name, int(math.prod(shape)), dtype, zero).view(*shape)


HIDDEN16 = os.environ.get("SCB_HIDDEN_FP32", "0") != "1"   # post-LN towers: fp16 hidden states / residual stream (see HubertPlan._forward)
LAYER_CHUNKS = int(os.environ.get("SCB_LAYER_CHUNKS", "1"))   # batch slices for the transformer layer stack (L2 residency)
GRAPHS = os.environ.get("SCB_CUDA_GRAPHS", "1") != "0"   # replay the frozen towers as CUDA graphs (launch-bound at small batch)
STREAM_K = os.environ.get("SCB_GEMM_STREAMK", "0") == "1"   # pass the stream-K workspace to the layer GEMMs (see EncoderLayerPlan.forward)
MAX_GRAPHS = 4                                            # per plan; further input shapes run eagerly


class GraphCache:
    """Capture a frozen tower's kernel sequence once per input signature and replay it with one launch.

    The towers are ~115 (HuBERT) / ~100 (ViT) dependent kernels whose launch parameters (TMA tensor maps included) only
    depend on the input shape and on buffer addresses; python + ctypes spend ~30 us on each, which bounds a 32-pair step on
    8 GPUs by the host.  ``run(key, fn)`` executes ``fn(workspace)`` eagerly the first time a signature is seen (this also
    performs the one-off cudaFuncSetAttribute calls), captures it the second time into a graph with a PRIVATE workspace
    (all scratch comes from the graph's memory pool, so addresses stay valid for the graph's lifetime) and replays
    afterwards.  Outputs live in the graph's pool and are overwritten by the next replay.
    """

    def __init__(self, device):
        self.device = device
        self.entries: Dict[tuple, object] = {}
        self.kernels_replayed = 0
        self.replays = 0

    def run(self, key: tuple, eager_ws: "Workspace", fn):
        from . import lib as _l
        if not GRAPHS or ops.PROFILE is not None:  # profiling wants one event pair per kernel: run eagerly
            return fn(eager_ws)
        ent = self.entries.get(key)
        if ent is None:
            if sum(1 for v in self.entries.values() if v != "warm") >= MAX_GRAPHS:
                return fn(eager_ws)
            self.entries[key] = "warm"
            return fn(eager_ws)
        if ent == "warm":
            graph = torch.cuda.CUDAGraph()
            ws = Workspace(self.device)
            n0 = _l.launch_count()
            prof, ops.PROFILE = ops.PROFILE, None  # per-call events cannot be recorded into a capture
            try:
                with torch.cuda.graph(graph):
                    out = fn(ws)
            finally:
                ops.PROFILE = prof
            ent = self.entries[key] = (graph, out, ws, _l.launch_count() - n0)
        graph, out, _, n_kernels = ent
        graph.replay()
        self.kernels_replayed += n_kernels
        self.replays += 1
        for t in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(t, torch.Tensor):
                t._scb_generation = self.replays  # consumers that keep the slab for a backward pass check it (functional.WeightedSumFn)
        return out


def graph_replayed_kernels() -> int:
    """Kernels executed through graph replays (they do not pass through the library's launch counter)."""
    return sum(c.kernels_replayed for c in _ALL_CACHES)


_ALL_CACHES: List[GraphCache] = []


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _h(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).to(H).contiguous()


def conv_out_len(n: int, spec=CONV_SPEC) -> int:
    for _, k, s in spec:
        n = (n - k) // s + 1
    return n


# ====================================================================================================== transformer layer
class EncoderLayerPlan:
    """Weights of one transformer layer in GEMM layout + the kernel sequence for the full [B, T, d] forward."""

    def __init__(self, dev, *, wqkv, bqkv, wo, bo, ln1, w1, b1, w2, b2, ln2, heads: int, pre_ln: bool, act: int, eps: float = 1e-5):
        self.wqkv, self.bqkv = _h(wqkv, dev), _f32(bqkv, dev)
        self.wo, self.bo = _h(wo, dev), _f32(bo, dev)
        self.w1, self.b1 = _h(w1, dev), _f32(b1, dev)
        self.w2, self.b2 = _h(w2, dev), _f32(b2, dev)
        self.ln1 = (_f32(ln1[0], dev), _f32(ln1[1], dev))
        self.ln2 = (_f32(ln2[0], dev), _f32(ln2[1], dev))
        self.heads, self.pre_ln, self.act, self.eps = heads, pre_ln, act, eps
        self.d = self.wo.shape[0]
        self.ffn = self.w1.shape[0]

    def forward(self, ws: Workspace, x32: Optional[torch.Tensor], x16: Optional[torch.Tensor], out32: Optional[torch.Tensor], B: int,
                T: int, kv_len: Optional[torch.Tensor], causal: bool = False, want_x16: bool = True, tag: str = "",
                out16: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """x32: fp32 [B*T, d] residual stream in; x16: its fp16 copy (post-LN only; None for pre-LN).
        out32: fp32 [B*T, d] receives the layer output.  Returns the fp16 copy of the output (post-LN) or None.
        Post-LN layers also run with x32 = out32 = None: the LayerNorm outputs (bounded by gamma / beta, safe in fp16) then exist
        only as fp16 — GEMM operand, residual and hidden state in one tensor — while the pre-LayerNorm sums stay fp32."""
        d, M = self.d, B * T
        hd = d // self.heads
        qkv = ws.view(tag + "qkv", (B, T, 3 * d), H)
        ctx = ws.view(tag + "ctx", (B, T, d), H)
        ffn = ws.view(tag + "ffn", (M, self.ffn), H)
        a16 = ws.view(tag + "a16", (M, d), H)
        # Stream-K scratch of the four GEMMs (zeroed at allocation, then owned by the library; one per launching stream, because
        # the two towers share a workspace and run concurrently).  OFF by default: measured on B200 (tools/streamk_bench.py,
        # profiles/r2_streamk_bench.txt) the split tiles' extra dump + reduce pass costs more than the idle tail of the last wave
        # on every shape of this model (HuBERT fc2 at 32 utterances: 48 -> 61 us; no change at 256).
        sk = None
        if STREAM_K:
            sk = ws.get(f"gemm_sk_{torch.cuda.current_stream().cuda_stream}", ops.gemm_workspace_bytes(), torch.uint8, zero=True)
        if self.pre_ln:
            ops.layernorm(x32, *self.ln1, y16=a16, rows=M, d=d, eps=self.eps)
            src16 = a16
        else:
            src16 = x16
        ops.gemm(src16, self.wqkv, bias=self.bqkv, out=qkv.view(M, 3 * d), scratch=sk)
        ops.attention(qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:3 * d], ctx, self.heads, hd ** -0.5, kv_len, causal)
        if self.pre_ln:
            xa = ws.view(tag + "xa", (M, d), torch.float32)
            ops.gemm(ctx.view(M, d), self.wo, bias=self.bo, residual=x32, out=xa, scratch=sk)
            ops.layernorm(xa, *self.ln2, y16=a16, rows=M, d=d, eps=self.eps)
            ops.gemm(a16, self.w1, bias=self.b1, act=self.act, out=ffn, scratch=sk)
            ops.gemm(ffn, self.w2, bias=self.b2, residual=xa, out=out32, scratch=sk)
            return None
        y = ws.view(tag + "y", (M, d), torch.float32)
        x1 = ws.view(tag + "x1", (M, d), torch.float32) if x32 is not None else None
        ops.gemm(ctx.view(M, d), self.wo, bias=self.bo, residual=x32 if x32 is not None else x16, out=y, scratch=sk)
        ops.layernorm(y, *self.ln1, y32=x1, y16=a16, rows=M, d=d, eps=self.eps)
        ops.gemm(a16, self.w1, bias=self.b1, act=self.act, out=ffn, scratch=sk)
        ops.gemm(ffn, self.w2, bias=self.b2, residual=x1 if x1 is not None else a16, out=y, scratch=sk)
        o16 = out16 if out16 is not None else (ws.view(tag + "o16", (M, d), H) if want_x16 else None)
        ops.layernorm(y, *self.ln2, y32=out32, y16=o16, rows=M, d=d, eps=self.eps)
        return o16


# ====================================================================================================== HuBERT
class HubertPlan:
    """Frozen fairseq HuBERT forward returning all L+1 hidden states (speech_encoder_plus.py:67-107, :29-64)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dev, *, heads: int, layer_norm_first: bool, extractor_layer_norm: bool,
                 pos_groups: int = 16):
        self.dev = dev
        self.pre_ln = layer_norm_first
        self.ext_ln = extractor_layer_norm
        g = lambda k: sd[k]
        has = lambda k: k in sd
        # ---- conv feature extractor
        self.c0_w = _f32(g("feature_extractor.conv_layers.0.0.weight").reshape(512, 10), dev)
        self.c0_b = _f32(g("feature_extractor.conv_layers.0.0.bias"), dev) if has("feature_extractor.conv_layers.0.0.bias") else None
        nk = "feature_extractor.conv_layers.0.2.1" if self.ext_ln else "feature_extractor.conv_layers.0.2"
        self.c0_gamma, self.c0_beta = _f32(g(nk + ".weight"), dev), _f32(g(nk + ".bias"), dev)
        self.convs = []
        for i in range(1, 7):
            w = g(f"feature_extractor.conv_layers.{i}.0.weight")  # [co, ci, k]
            co, ci, k = w.shape
            wk = _h(w.permute(0, 2, 1).reshape(co, k * ci), dev)  # k index = tap*ci + channel: matches the pair-row walk
            b = _f32(g(f"feature_extractor.conv_layers.{i}.0.bias"), dev) if has(f"feature_extractor.conv_layers.{i}.0.bias") else None
            ln = None
            if self.ext_ln:
                ln = (_f32(g(f"feature_extractor.conv_layers.{i}.2.1.weight"), dev), _f32(g(f"feature_extractor.conv_layers.{i}.2.1.bias"), dev))
            self.convs.append((wk, b, ln, k))
        self.ln_feat = (_f32(g("layer_norm.weight"), dev), _f32(g("layer_norm.bias"), dev))
        self.proj_w, self.proj_b = _h(g("post_extract_proj.weight"), dev), _f32(g("post_extract_proj.bias"), dev)
        self.d = self.proj_w.shape[0]
        # ---- positional conv: weight_norm(dim=2) folded, regrouped to [G][cpg][tap*64 + ci] with ci zero-padded to 64
        wg, wv = g("encoder.pos_conv.0.weight_g").float(), g("encoder.pos_conv.0.weight_v").float()
        w = (wg * wv / wv.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()).to(dev)  # [d, cpg, K]
        d, cpg, K = w.shape
        G = d // cpg
        assert G == pos_groups and cpg <= 64 and cpg % 8 == 0, (d, cpg, K)
        wp = torch.zeros(G, cpg, K, 64, device=dev, dtype=torch.float32)
        wp[:, :, :, :cpg] = w.view(G, cpg, cpg, K).permute(0, 1, 3, 2)  # [g, co, tap, ci]
        self.pos_w = wp.reshape(G, cpg, K * 64).to(H).contiguous()
        self.pos_b = _f32(g("encoder.pos_conv.0.bias"), dev)
        self.pos_k, self.pos_g, self.pos_cpg = K, G, cpg
        self.enc_ln = (_f32(g("encoder.layer_norm.weight"), dev), _f32(g("encoder.layer_norm.bias"), dev))
        # ---- encoder layers
        self.layers: List[EncoderLayerPlan] = []
        n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layers."))
        for l in range(n_layers):
            p = f"encoder.layers.{l}."
            self.layers.append(EncoderLayerPlan(
                dev,
                wqkv=torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"), g(p + "self_attn.v_proj.weight")], 0),
                bqkv=torch.cat([g(p + "self_attn.q_proj.bias"), g(p + "self_attn.k_proj.bias"), g(p + "self_attn.v_proj.bias")], 0),
                wo=g(p + "self_attn.out_proj.weight"), bo=g(p + "self_attn.out_proj.bias"),
                ln1=(g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias")),
                w1=g(p + "fc1.weight"), b1=g(p + "fc1.bias"), w2=g(p + "fc2.weight"), b2=g(p + "fc2.bias"),
                ln2=(g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias")),
                heads=heads, pre_ln=self.pre_ln, act=ops.ACT_GELU))
        self.n_hidden = n_layers + 1
        self.graphs = GraphCache(dev)
        _ALL_CACHES.append(self.graphs)

    # -------------------------------------------------------------------------------------------------
    def conv_stack(self, ws: Workspace, wav: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """wav fp32 [B, Tw] (zero padded) -> channel-last fp16 [B, T, 512] features (before layer_norm); returns (buf, T)."""
        B, Tw = wav.shape
        T = (Tw - 10) // 5 + 1
        cur = ws.get("conv_a", B * T * 512 + SLACK, H)
        scratch = ws.get("conv0_scratch", ops.conv0_scratch_bytes(B), torch.uint8)
        if self.ext_ln:
            ops.conv0_layernorm_gelu(wav, Tw, self.c0_w, self.c0_b, self.c0_gamma, self.c0_beta, 1e-5, cur, T * 512, scratch)
        else:
            ops.conv0_groupnorm_gelu(wav, Tw, self.c0_w, self.c0_b, self.c0_gamma, self.c0_beta, 1e-5, cur, T * 512, scratch)
        names = ("conv_b", "conv_a")
        for i, (wk, b, ln, k) in enumerate(self.convs):
            T_out = (T - k) // 2 + 1
            nxt = ws.get(names[i % 2], B * T_out * 512 + SLACK, H)
            ops.gemm_raw(a=cur, a_inner=1024, a_rows=(T + 1) // 2, a_row_stride=1024, a_batch_stride=T * 512, batch=B,
                         m_per_batch=T_out, w=wk, n=512, k=k * 512, kb_per_tap=16, tap_row_shift=1, out=nxt, ldc=512,
                         out_batch_stride=T_out * 512, bias=b, act=ops.ACT_NONE if ln is not None else ops.ACT_GELU)
            if ln is not None:  # large: LayerNorm over channels, then GELU (in place on the fp16 rows)
                ops.layernorm(nxt, ln[0], ln[1], y16=nxt, rows=B * T_out, d=512, eps=1e-5, act=ops.ACT_GELU)
            cur, T = nxt, T_out
        return cur, T

    def forward(self, ws: Workspace, wav: torch.Tensor, valid_frames: Optional[torch.Tensor]) -> Tuple[torch.Tensor, int]:
        """-> (hidden fp32 [L+1, B*T, d], T).  ``wav`` / ``valid_frames`` must be persistent buffers (their addresses are part
        of the graph signature); under graph replay ``hidden`` is overwritten by the next call with the same signature."""
        key = (wav.data_ptr(), tuple(wav.shape), valid_frames.data_ptr() if valid_frames is not None else 0)
        return self.graphs.run(key, ws, lambda w: self._forward(w, wav, valid_frames))

    def _forward(self, ws: Workspace, wav: torch.Tensor, valid_frames: Optional[torch.Tensor]) -> Tuple[torch.Tensor, int]:
        B = wav.shape[0]
        d = self.d
        feats, T = self.conv_stack(ws, wav)
        M = B * T
        f16 = ws.view("feat_ln", (M, 512), H)
        ops.layernorm(feats, *self.ln_feat, y16=f16, rows=M, d=512, eps=1e-5)
        x = ws.view("post_proj", (M, d), torch.float32)
        ops.gemm(f16, self.proj_w, bias=self.proj_b, out=x)
        # masked zero + grouped positional conv (k taps, SamePad drops the last output frame) + GELU + residual
        K, G, cpg = self.pos_k, self.pos_g, self.pos_cpg
        rows_pad = T + K
        xpad = ws.view("xpad", (B, rows_pad, G * 64), H)  # ONE buffer for every (B, T): the pack kernel writes the zero rows / channels too
        ops.posconv_pack(x, valid_frames, xpad, B, T, d, G, K // 2, rows_pad)
        # Post-LN towers (HuBERT-base) keep the hidden states — LayerNorm outputs — in fp16 only: one tensor is the next GEMM's
        # operand, the residual and the state the weighted sum reads (HIDDEN16; SCB_HIDDEN_FP32=1 restores fp32 copies).
        h16 = HIDDEN16 and not self.pre_ln
        hidden = torch.empty(self.n_hidden, M, d, device=self.dev, dtype=H if h16 else torch.float32)
        tgt = hidden[0] if self.pre_ln else x
        ops.gemm_raw(a=xpad, a_inner=G * 64, a_rows=rows_pad, a_row_stride=G * 64, a_batch_stride=rows_pad * G * 64, batch=B,
                     m_per_batch=T, w=self.pos_w, n=cpg, k=K * 64, groups=G, b_group_stride=cpg * K * 64, kb_per_tap=1,
                     tap_row_shift=1, a_group_cols=64, out=tgt, ldc=d, out_batch_stride=T * d, out_group_cols=cpg,
                     bias=self.pos_b, act=ops.ACT_GELU, residual=x, algo_k=K * cpg)
        x16s = None
        if h16:
            ops.layernorm(x, *self.enc_ln, y16=hidden[0], rows=M, d=d, eps=1e-5)
        elif not self.pre_ln:
            x16s = [ws.view("hub_x16_a", (M, d), H), ws.view("hub_x16_b", (M, d), H)]
            ops.layernorm(x, *self.enc_ln, y32=hidden[0], y16=x16s[0], rows=M, d=d, eps=1e-5)
        # The layer stack runs over slices of the batch: with all 256 utterances a layer's intermediates (QKV 376 MB, MLP 502 MB,
        # pre-LN sums 250 MB) stream through HBM between producer and consumer kernels; per slice they stay in the 126 MB L2.
        chunks = max(1, min(LAYER_CHUNKS, B))
        while B % chunks:
            chunks -= 1
        Bc = B // chunks
        Mc = Bc * T
        cur = 0
        for l, layer in enumerate(self.layers):
            for c in range(chunks):
                rows = slice(c * Mc, (c + 1) * Mc)
                vf = valid_frames[c * Bc:(c + 1) * Bc] if valid_frames is not None else None
                if h16:
                    layer.forward(ws, None, hidden[l][rows], None, Bc, T, vf, tag="hub_", out16=hidden[l + 1][rows])
                else:
                    layer.forward(ws, hidden[l][rows], None if self.pre_ln else x16s[cur][rows], hidden[l + 1][rows], Bc, T, vf, tag="hub_",
                                  out16=None if self.pre_ln else x16s[1 - cur][rows])
            cur = 1 - cur  # fp16 copies ping-pong: later slices of this layer still read the previous layer's copy
        return hidden, T


# ====================================================================================================== CLIP ViT
class VitPlan:
    """Frozen CLIP VisionTransformer forward (clip_official.py:200-209 -> openai clip/model.py)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dev, *, heads: int, prefix: str = "visual."):
        g = lambda k: sd[prefix + k]
        self.dev = dev
        w = g("conv1.weight")  # [W, 3, P, P]
        self.width, _, self.patch, _ = w.shape
        kk = 3 * self.patch * self.patch
        self.ldk = (kk + 7) // 8 * 8
        wp = torch.zeros(self.width, self.ldk, dtype=torch.float32, device=dev)
        wp[:, :kk] = w.reshape(self.width, kk).to(dev)
        self.kk = kk
        self.conv_w = wp.to(H).contiguous()
        self.cls = _f32(g("class_embedding"), dev)
        self.pos = _f32(g("positional_embedding"), dev)
        self.tokens = self.pos.shape[0]
        self.ln_pre = (_f32(g("ln_pre.weight"), dev), _f32(g("ln_pre.bias"), dev))
        self.ln_post = (_f32(g("ln_post.weight"), dev), _f32(g("ln_post.bias"), dev))
        self.proj_t = _h(g("proj").t(), dev)  # [E, W]
        self.out_dim = self.proj_t.shape[0]
        n_layers = 1 + max(int(k[len(prefix):].split(".")[2]) for k in sd if k.startswith(prefix + "transformer.resblocks."))
        self.layers = []
        for l in range(n_layers):
            p = f"transformer.resblocks.{l}."
            self.layers.append(EncoderLayerPlan(
                dev, wqkv=g(p + "attn.in_proj_weight"), bqkv=g(p + "attn.in_proj_bias"), wo=g(p + "attn.out_proj.weight"),
                bo=g(p + "attn.out_proj.bias"), ln1=(g(p + "ln_1.weight"), g(p + "ln_1.bias")), w1=g(p + "mlp.c_fc.weight"),
                b1=g(p + "mlp.c_fc.bias"), w2=g(p + "mlp.c_proj.weight"), b2=g(p + "mlp.c_proj.bias"),
                ln2=(g(p + "ln_2.weight"), g(p + "ln_2.bias")), heads=heads, pre_ln=True, act=ops.ACT_QUICK_GELU))
        self.graphs = GraphCache(dev)
        _ALL_CACHES.append(self.graphs)

    def forward(self, ws: Workspace, image: torch.Tensor) -> torch.Tensor:
        """image fp32 [B, 3, S, S] -> un-normalised embedding fp32 [B, E] (ln_post(x[:,0]) @ proj)."""
        B = image.shape[0]
        G2 = self.tokens - 1
        patches = ws.view("vit_patches", (B * G2, self.ldk), H)  # persistent: the graph starts after the im2col of the caller's tensor
        ops.patchify(image, patches, self.patch, self.ldk)
        out = self.graphs.run((patches.data_ptr(), B), ws, lambda w: self._forward(w, patches, B))
        res = torch.empty_like(out)  # the graph's output slot is overwritten by the next replay: hand out a copy
        ops.cast_rows(out, res)
        return res

    def _forward(self, ws: Workspace, patches: torch.Tensor, B: int) -> torch.Tensor:
        Wd, L, G2 = self.width, self.tokens, self.tokens - 1
        xa = ws.view("vit_x0", (B * L, Wd), torch.float32)
        xb = ws.view("vit_x1", (B * L, Wd), torch.float32)
        ops.gemm_raw(a=patches, a_inner=self.ldk, a_rows=G2, a_row_stride=self.ldk, a_batch_stride=G2 * self.ldk, batch=B,
                     m_per_batch=G2, w=self.conv_w, n=Wd, k=self.ldk, out=xa, out_offset=Wd, ldc=Wd, out_batch_stride=L * Wd,
                     residual=self.pos, residual_offset=Wd, residual_ld=Wd, residual_batch_stride=0, algo_k=self.kk)
        ops.broadcast_row(self.cls, self.pos, xa, L * Wd, B, Wd)
        ops.layernorm(xa, *self.ln_pre, y32=xa, rows=B * L, d=Wd, eps=1e-5)
        cur, nxt = xa, xb
        for layer in self.layers:
            layer.forward(ws, cur, None, nxt, B, L, None, tag="vit_")
            cur, nxt = nxt, cur
        c16 = ws.view("vit_cls16", (B, Wd), H)
        ops.layernorm(cur, *self.ln_post, y16=c16, rows=B, d=Wd, x_ld=L * Wd, y_ld=Wd, eps=1e-5)
        out = torch.empty(B, self.out_dim, device=self.dev, dtype=torch.float32)
        ops.gemm(c16, self.proj_t, out=out)
        return out
