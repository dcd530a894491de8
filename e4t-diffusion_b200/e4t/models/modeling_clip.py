"""CLIPTextModel with `inputs_embeds` — mirror of e4t/models/modeling_clip.py:9-82 (HF CLIP text tower naming:
text_model.embeddings.{token,position}_embedding, encoder.layers.{i}.{layer_norm1,self_attn.{q,k,v,out}_proj,
layer_norm2,mlp.fc1,mlp.fc2}, final_layer_norm; causal mask; quick_gelu).

Frozen weights, but dX must flow to the injected placeholder row (pretrain_e4t.py:630-634).  SURVEY.md §8 a-14 ranks
this tower "next": it runs on stock torch ops (cuBLAS + SDPA) in the module's dtype, not on the e4t_b200 kernels."""
import json
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from e4t._mixins import BaseOutput


@dataclass
class BaseModelOutputWithPooling(BaseOutput):
    last_hidden_state: torch.Tensor = None
    pooler_output: torch.Tensor = None


class CLIPTextConfig:
    def __init__(self, vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                 num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                 **kwargs):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.max_position_embeddings = max_position_embeddings
        self.hidden_act = hidden_act
        self.layer_norm_eps = layer_norm_eps
        self.use_return_dict = True


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.position_embedding = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)

    def forward(self, input_ids=None, position_ids=None, inputs_embeds=None):
        if inputs_embeds is None:
            inputs_embeds = self.token_embedding(input_ids)
        n = inputs_embeds.shape[1]
        pos = self.position_embedding.weight[:n] if position_ids is None else self.position_embedding(position_ids)
        return inputs_embeds + pos


class _Attention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.hidden_size
        self.heads = cfg.num_attention_heads
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))

    def forward(self, x):
        B, N, D = x.shape
        h = self.heads
        q, k, v = (p(x).view(B, N, h, D // h).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)              # causal mask (modeling_clip.py:45-47)
        return self.out_proj(o.transpose(1, 2).reshape(B, N, D))


class _MLP(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.fc1 = nn.Linear(cfg.hidden_size, cfg.intermediate_size)
        self.fc2 = nn.Linear(cfg.intermediate_size, cfg.hidden_size)
        self.act = cfg.hidden_act

    def forward(self, x):
        h = self.fc1(x)
        h = h * torch.sigmoid(1.702 * h) if self.act == "quick_gelu" else F.gelu(h)
        return self.fc2(h)


class _Layer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.self_attn = _Attention(cfg)
        self.layer_norm1 = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.mlp = _MLP(cfg)
        self.layer_norm2 = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(cfg) for _ in range(cfg.num_hidden_layers)])

    def forward(self, x):
        for l in self.layers:
            x = l(x)
        return x


class _TextTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.final_layer_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class CLIPTextModel(nn.Module):
    def __init__(self, config: Optional[CLIPTextConfig] = None, **kwargs):
        super().__init__()
        self.config = config if config is not None else CLIPTextConfig(**kwargs)
        self.text_model = _TextTransformer(self.config)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kwargs):
        d = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(d, "config.json")) as f:
            cfg = CLIPTextConfig(**json.load(f))
        m = cls(cfg)
        for name in ("pytorch_model.bin",):
            if os.path.exists(os.path.join(d, name)):
                m.load_state_dict(torch.load(os.path.join(d, name), map_location="cpu"), strict=False)
        return m

    @property
    def dtype(self):
        return self.text_model.final_layer_norm.weight.dtype

    @property
    def device(self):
        return self.text_model.final_layer_norm.weight.device

    def get_input_embeddings(self):
        return self.text_model.embeddings.token_embedding

    def resize_token_embeddings(self, new_num_tokens):
        old = self.text_model.embeddings.token_embedding
        if new_num_tokens == old.num_embeddings:
            return old
        new = nn.Embedding(new_num_tokens, old.embedding_dim).to(old.weight.device, old.weight.dtype)
        n = min(new_num_tokens, old.num_embeddings)
        with torch.no_grad():
            new.weight[:n] = old.weight[:n]
        new.weight.requires_grad_(old.weight.requires_grad)
        self.text_model.embeddings.token_embedding = new
        self.config.vocab_size = new_num_tokens
        return new

    def forward(self, input_ids=None, inputs_embeds=None, attention_mask=None, position_ids=None,
                output_attentions=None, output_hidden_states=None, return_dict=None):
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify input_ids or inputs_embeds")
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is never passed on the E4T path")
        if input_ids is not None:
            input_ids = input_ids.view(-1, input_ids.shape[-1])
        x = self.text_model.embeddings(input_ids=input_ids, inputs_embeds=inputs_embeds, position_ids=position_ids)
        x = self.text_model.final_layer_norm(self.text_model.encoder(x))
        pooled = x[:, 0]                                                         # modeling_clip.py:72
        if return_dict is False:
            return (x, pooled)
        return BaseModelOutputWithPooling(last_hidden_state=x, pooler_output=pooled)
