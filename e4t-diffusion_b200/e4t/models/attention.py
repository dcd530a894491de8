"""BasicTransformerBlock / FeedForward / GEGLU — mirror of e4t/models/attention.py:181-430 on the sm_100a kernels.
LayerNorm -> attn1 (self) -> LayerNorm -> attn2 (cross) -> LayerNorm -> GEGLU feed-forward, residual adds fused into
the producing GEMM epilogues."""
from typing import Optional

import torch
from torch import nn

from e4t.models.cross_attention import CrossAttention, _weight_bf16
from e4t_b200 import functional as FN


class GEGLU(nn.Module):
    """attention.py:409-430: proj: dim_in -> 2*dim_out, out = h * gelu(gate) (exact erf GELU)."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        h = FN.LinearFn.apply(hidden_states, _weight_bf16(self.proj), self.proj.bias, None, self.proj.weight)
        return FN.GEGLUFn.apply(h)


class FeedForward(nn.Module):
    """attention.py:335-384 (activation_fn='geglu', the only variant SD-v1.x builds)."""

    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu", final_dropout: bool = False):
        super().__init__()
        if activation_fn != "geglu":
            raise NotImplementedError("only GEGLU feed-forward is on the SD-v1.x path")
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])

    def forward(self, hidden_states, residual=None):
        h = self.net[0](hidden_states)
        return FN.LinearFn.apply(h, _weight_bf16(self.net[2]), self.net[2].bias, residual, self.net[2].weight)


class BasicTransformerBlock(nn.Module):
    """attention.py:181-332."""

    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, dropout=0.0,
                 cross_attention_dim: Optional[int] = None, activation_fn: str = "geglu",
                 num_embeds_ada_norm: Optional[int] = None, attention_bias: bool = False,
                 only_cross_attention: bool = False, upcast_attention: bool = False,
                 norm_elementwise_affine: bool = True, norm_type: str = "layer_norm", final_dropout: bool = False):
        super().__init__()
        if norm_type != "layer_norm" or num_embeds_ada_norm is not None or only_cross_attention:
            raise NotImplementedError("AdaLayerNorm / only_cross_attention are not on the SD-v1.x path")
        self.only_cross_attention = only_cross_attention
        self.attn1 = CrossAttention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
                                    dropout=dropout, bias=attention_bias, upcast_attention=upcast_attention)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)
        self.attn2 = CrossAttention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                                    dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                    upcast_attention=upcast_attention) if cross_attention_dim is not None else None
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine) if self.attn2 is not None else None
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)

    @staticmethod
    def _ln(norm, x):
        return FN.LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, attention_mask=None,
                cross_attention_kwargs=None, class_labels=None):
        kw = cross_attention_kwargs if cross_attention_kwargs is not None else {}
        x = FN.as_bf16(hidden_states)
        x = self.attn1(self._ln(self.norm1, x), encoder_hidden_states=None, attention_mask=attention_mask,
                       residual=x, **kw)                                                   # attention.py:291-302
        if self.attn2 is not None:
            x = self.attn2(self._ln(self.norm2, x), encoder_hidden_states=encoder_hidden_states,
                           attention_mask=attention_mask, residual=x, **kw)                # :304-316
        return self.ff(self._ln(self.norm3, x), residual=x)                                # :318-330
