"""CrossAttention with E4T WeightOffsets — mirror of e4t/models/cross_attention.py:22-282 (module surface, attribute
names, processor protocol) running on the sm_100a kernels.

Per module the reference does, for every forward (cross_attention.py:506-538):
    W_eff = to_X.weight * (1 + wo_X())  (X = q,k,v; two dense square GEMMs + an elementwise pass each)
    q,k,v = F.linear(.., W_eff) ; SDPA ; to_out
Here: the closed-form WeightOffsets factors and W ⊙ (1+Δ) are produced ONCE per optimisation step for the q/k/v
group (shared by the two UNet passes of pretrain_e4t.py:624,636), the group is one fused tcgen05 GEMM, the attention
core is the fused flash kernel, and to_out carries bias + residual in its epilogue.
"""
from typing import Optional

import torch
from torch import nn

from e4t.weightoffsets import WeightOffsets
from e4t_b200 import functional as FN
from e4t_b200._lib import E4TError


class CrossAttention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias=False, upcast_attention: bool = False, upcast_softmax: bool = False,
                 cross_attention_norm: bool = False, added_kv_proj_dim: Optional[int] = None,
                 norm_num_groups: Optional[int] = None, processor=None):
        super().__init__()
        if cross_attention_norm or added_kv_proj_dim is not None or norm_num_groups is not None:
            raise NotImplementedError("cross_attention_norm / added_kv_proj_dim / norm_num_groups are not on the "
                                      "SD-v1.x E4T path (SURVEY.md §2 #2)")
        inner_dim = dim_head * heads
        self.is_cross = cross_attention_dim is not None
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.cross_attention_norm = cross_attention_norm
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.group_norm = None
        self.to_q = nn.Linear(query_dim, inner_dim, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else B200AttnProcessor())
        # weight offsets (cross_attention.py:97-99)
        self.wo_q = WeightOffsets(query_dim, inner_dim)
        self.wo_k = WeightOffsets(cross_attention_dim, inner_dim)
        self.wo_v = WeightOffsets(cross_attention_dim, inner_dim)
        self._weff_cache = {}
        self._wo_bank = None      # set by e4t_b200.wobank (batched WeightOffsets kernels across the whole UNet)

    def _project(self, x, group):
        """x @ W_eff(group)ᵀ with the WeightOffsets-modulated, row-concatenated projection weights."""
        bank = self._wo_bank
        if bank is not None:
            w_eff, dweff, token = bank.get(self, group)
            if token is not None:
                return FN.WOLinearBankFn.apply(x, w_eff, dweff, token)
            return FN.LinearFn.apply(x, w_eff, None, None, None)
        w_eff, carrier = self.effective_weights(group)
        return FN.WOLinearFn.apply(x, w_eff, carrier)

    # ---- API kept for the training scripts (pretrain_e4t.py:264-272) -------------------------
    def set_use_memory_efficient_attention_xformers(self, use, attention_op=None):
        return None  # the fused sm_100a kernel is always memory-efficient; xFormers is not used

    def set_attention_slice(self, slice_size):
        if slice_size is not None and slice_size > self.sliceable_head_dim:
            raise ValueError(f"slice_size {slice_size} has to be smaller or equal to {self.sliceable_head_dim}.")
        return None  # nothing to slice: scores never leave the SM

    def set_processor(self, processor):
        self.processor = processor

    def head_to_batch_dim(self, tensor):
        b, n, c = tensor.shape
        h = self.heads
        return tensor.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, tensor):
        bh, n, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        """softmax(scale * Q Kᵀ (+ mask)) on head-batched (B*H, N, dh) tensors — the materialising helper of the
        non-fused processors (cross_attention.py:222-251).  The fused processor never calls it (scores do not leave
        the SM there); it is part of the module surface other processors / visualisation code rely on."""
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        scores = torch.matmul(query, key.transpose(-1, -2)) * self.scale
        if attention_mask is not None:
            scores = scores + attention_mask
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size=None):
        """Pad the mask by `target_length` zeros on the key axis and repeat it per head
        (cross_attention.py:253-282; batch_size=None is the deprecated 1-sample form)."""
        if batch_size is None:
            batch_size = 1
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            attention_mask = torch.nn.functional.pad(attention_mask, (0, target_length), value=0.0)
        if attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    # ---- fused effective weights ---------------------------------------------------------------
    def effective_weights(self, group):
        """group ∈ {'qkv','q','kv'} -> (W_eff bf16 (ΣC,R), fp32 autograd carrier).  Cached until a parameter changes
        or the cached autograd node has been consumed by a backward pass."""
        lins = {"qkv": (self.to_q, self.to_k, self.to_v), "q": (self.to_q,), "kv": (self.to_k, self.to_v)}[group]
        wos = {"qkv": (self.wo_q, self.wo_k, self.wo_v), "q": (self.wo_q,), "kv": (self.wo_k, self.wo_v)}[group]
        flat = [l.weight for l in lins]
        for w in wos:
            flat += list(w.kernel_params())
        grad_on = torch.is_grad_enabled() and any(p.requires_grad for p in flat[len(lins):])
        key = (tuple((p._version, p.data_ptr()) for p in flat), grad_on, FN.WO_EPOCH if grad_on else -1,
               FN.PARAM_EPOCH if any(getattr(p, "_e4t_arena", False) for p in flat) else 0)
        ent = self._weff_cache.get(group)
        if ent is None or ent[0] != key:
            for l in lins:
                if l.bias is not None:
                    raise NotImplementedError("q/k/v projections with bias are not on the SD-v1.x path")
            w_eff, carrier = FN.WOEffectiveFn.apply(len(lins), *flat)
            ent = (key, (w_eff, carrier))
            self._weff_cache[group] = ent
        return ent[1]


def _weight_bf16(lin):
    return FN.prepared(lin.weight, "bf16", lambda w: w.to(torch.bfloat16).contiguous())


class B200AttnProcessor:
    """Processor protocol of the reference (cross_attention.py:285-322 / 490-538):
    processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None) -> (B, N, query_dim).
    Extra optional kwarg `residual` (B,N,query_dim) is added in the to_out GEMM epilogue."""

    def __call__(self, attn: CrossAttention, hidden_states, encoder_hidden_states=None, attention_mask=None,
                 residual=None):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is None on the SD-v1.x E4T path (SURVEY.md §8 a-3)")
        if not hidden_states.is_cuda:
            raise E4TError("e4t CrossAttention runs on the sm_100a kernels only (no CPU fallback)")
        x = FN.as_bf16(hidden_states)
        if encoder_hidden_states is None:
            qkv = attn._project(x, "qkv")
            o = FN.AttentionFn.apply(qkv, None, attn.heads, attn.scale)
        else:
            ctx = FN.as_bf16(encoder_hidden_states)
            q = attn._project(x, "q")
            kv = attn._project(ctx, "kv")
            o = FN.AttentionFn.apply(q, kv, attn.heads, attn.scale)
        out = attn.to_out[0]
        y = FN.LinearFn.apply(o, _weight_bf16(out), out.bias, residual, out.weight)
        if attn.to_out[1].p > 0.0 and attn.training:
            y = attn.to_out[1](y)
        return y


# names the reference exports; all resolve to the fused kernel path
CrossAttnProcessor = B200AttnProcessor
AttnProcessor2_0 = B200AttnProcessor
XFormersCrossAttnProcessor = B200AttnProcessor
AttnProcessor = B200AttnProcessor
