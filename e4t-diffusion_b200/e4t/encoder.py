"""E4TEncoder — mirror of e4t/encoder.py:78-168: CLIP ViT-H/14 vision tower (open_clip parameter names, proj=None,
output_tokens=True) + pooled UNet features -> placeholder-token embedding.

  u      = MLP( cat_k mean_{HW}(unet map_k) )                       (encoder.py:147-149)  — sm_100a mean-pool kernel
  hs     = [ln_post(x_cls), x_patch[1::2]]  (129 "layers")           (encoder.py:154-157)
  out_i  = first_linears[i]( feature_linear([hs_i, u]) )              (encoder.py:159-162)
  result = final_linear( LeakyReLU( mean_i out_i ) )                  (encoder.py:163-168)

The reference loops over the 129 layers in Python (258 tiny GEMMs + 129 cats).  Algebraically
feature_linear([hs_i,u]) = hs_i·Wf[:, :W]ᵀ + u·Wf[:, W:]ᵀ + bf, so the head is one (B·129,W)x(W,W) GEMM, one
broadcast add and ONE batched GEMM over the stacked first_linears.  The 129 first_linears parameters are views of
one stacked (129,W,W) storage (state-dict keys unchanged) so no per-step stacking copy is needed.

The ViT tower is frozen by default (freeze_clip_vision=True) and, being ranked below the UNet kernels in SURVEY.md
§8, runs on stock torch ops (cuBLAS + SDPA) in bf16; random-init construction never touches the network."""
import torch
import torch.nn.functional as F
from torch import nn

from e4t._mixins import ConfigMixin, ModelMixin, register_to_config
from e4t_b200 import functional as FN

_VIT_ARCHS = {
    # arch: (width, layers, heads, mlp, patch, image)
    "ViT-H-14": (1280, 32, 16, 5120, 14, 224),
    "ViT-L-14": (1024, 24, 16, 4096, 14, 224),
    "ViT-B-32": (768, 12, 12, 3072, 32, 224),
    "ViT-tiny-test": (64, 2, 4, 128, 14, 224),
}


class _MHA(nn.Module):
    """nn.MultiheadAttention-compatible parameter names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        nn.init.xavier_uniform_(self.in_proj_weight)

    def forward(self, x, cast):
        B, N, W = x.shape
        q, k, v = F.linear(x, cast(self.in_proj_weight), cast(self.in_proj_bias)).chunk(3, dim=-1)
        h = self.heads
        q, k, v = (t.view(B, N, h, W // h).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v)
        return F.linear(o.transpose(1, 2).reshape(B, N, W), cast(self.out_proj.weight), cast(self.out_proj.bias))


class _MLP(nn.Module):
    def __init__(self, width, mlp):
        super().__init__()
        self.c_fc = nn.Linear(width, mlp)
        self.gelu = nn.GELU()
        self.c_proj = nn.Linear(mlp, width)


class _ResBlock(nn.Module):
    def __init__(self, width, heads, mlp):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = _MHA(width, heads)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = _MLP(width, mlp)

    def forward(self, x, cast):
        h = F.layer_norm(x, (x.shape[-1],), cast(self.ln_1.weight), cast(self.ln_1.bias), 1e-5)
        x = x + self.attn(h, cast)
        h = F.layer_norm(x, (x.shape[-1],), cast(self.ln_2.weight), cast(self.ln_2.bias), 1e-5)
        h = F.gelu(F.linear(h, cast(self.mlp.c_fc.weight), cast(self.mlp.c_fc.bias)))
        return x + F.linear(h, cast(self.mlp.c_proj.weight), cast(self.mlp.c_proj.bias))


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(width, heads, mlp) for _ in range(layers)])


class VisionTransformer(nn.Module):
    """open_clip VisionTransformer with proj=None, output_tokens=True: returns (ln_post(cls), patch tokens)."""

    def __init__(self, width, layers, heads, mlp, patch, image):
        super().__init__()
        self.output_tokens = True
        self.proj = None
        self.patch = patch
        grid = image // patch
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(grid * grid + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads, mlp)
        self.ln_post = nn.LayerNorm(width)

    def forward(self, x, compute_dtype=torch.bfloat16):
        trainable = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if trainable:   # --unfreeze_clip_vision: differentiable casts of the fp32 masters
            cast = lambda p: p.to(compute_dtype)
        else:           # frozen: cached bf16 operand copies
            cast = lambda p: FN.prepared(p, ("cast", compute_dtype), lambda t: t.to(compute_dtype).contiguous())
        W = self.conv1.out_channels
        with torch.set_grad_enabled(trainable):
            x = F.conv2d(x.to(compute_dtype), cast(self.conv1.weight), stride=self.patch)
            x = x.reshape(x.shape[0], W, -1).permute(0, 2, 1)
            cls = cast(self.class_embedding).expand(x.shape[0], 1, W)
            x = torch.cat([cls, x], dim=1) + cast(self.positional_embedding)
            x = F.layer_norm(x, (W,), cast(self.ln_pre.weight), cast(self.ln_pre.bias), 1e-5)
            for blk in self.transformer.resblocks:
                x = blk(x, cast)
            pooled = F.layer_norm(x[:, 0], (W,), cast(self.ln_post.weight), cast(self.ln_post.bias), 1e-5)
            tokens = x[:, 1:]
        return pooled, tokens


class E4TEncoder(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, word_embedding_dim=768, block_out_channels=(320, 640, 1280, 1280), arch="ViT-H-14",
                 version="laion2b_s32b_b79k", antialias=False, freeze_clip_vision=True, **kwargs):
        super().__init__()
        if arch not in _VIT_ARCHS:
            raise ValueError(f"unknown CLIP vision arch {arch}; known: {sorted(_VIT_ARCHS)}")
        width, layers, heads, mlp, patch, image = _VIT_ARCHS[arch]
        # random-init tower; pretrained weights arrive through load_state_dict / e4t.utils.load_e4t_encoder
        self.clip_vision = VisionTransformer(width, layers, heads, mlp, patch, image)
        if freeze_clip_vision:
            self.clip_vision.requires_grad_(False)
        unet_feature_dim = int(kwargs.get("unet_feature_dim", 10880))       # hard-coded 10880 at encoder.py:102
        self.unet_feature_embedder = nn.Sequential(nn.Linear(unet_feature_dim, width), nn.LeakyReLU(),
                                                   nn.Linear(width, width))
        self.feature_linear = nn.Linear(2 * width, width)
        if arch == "ViT-H-14":
            n_odd_layers = 128 + 1                                            # encoder.py:111
        else:
            n_odd_layers = kwargs.get("n_odd_layers", None)
            assert n_odd_layers is not None, "You must specify `n_odd_layers`!"
            n_odd_layers = int(n_odd_layers)
        self.first_linears = nn.ModuleList([nn.Linear(width, width) for _ in range(n_odd_layers)])
        self._restack_first_linears()
        self.act = nn.LeakyReLU()
        self.final_linear = nn.Linear(width, word_embedding_dim)
        self.image_size = image
        self.antialias = antialias
        self.register_buffer("mean", torch.tensor([0.48145466, 0.4578275, 0.40821073]), persistent=False)
        self.register_buffer("std", torch.tensor([0.26862954, 0.26130258, 0.27577711]), persistent=False)

    # ---- stacked storage for the 129 first_linears ----------------------------------------------------
    def _restack_first_linears(self):
        """Re-home the per-layer weights/biases as views of one (n,W,W) / (n,W) storage (values preserved)."""
        n = len(self.first_linears)
        w0 = self.first_linears[0].weight
        W = w0.shape[0]
        wst = torch.empty((n, W, W), device=w0.device, dtype=w0.dtype)
        bst = torch.empty((n, W), device=w0.device, dtype=w0.dtype)
        with torch.no_grad():
            for i, l in enumerate(self.first_linears):
                wst[i].copy_(l.weight)
                bst[i].copy_(l.bias)
                l.weight.data = wst[i]
                l.bias.data = bst[i]
        self._fl_w, self._fl_b = wst, bst

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)   # .to()/.cuda() re-allocates each parameter separately
        if hasattr(self, "first_linears") and hasattr(self, "_fl_w"):
            self._restack_first_linears()
        return out

    def _stacked(self):
        """(n,W,W) / (n,W) views over the per-layer parameters.  When an arena optimiser (FlatAdamW) has re-homed the
        parameters, they are still laid out back to back inside its arena (it moves each shared storage as one block):
        the stacked tensors are rebuilt as strided VIEWS of that memory — never copied out of the arena, which would
        detach the parameters from the optimiser's updates."""
        l0 = self.first_linears[0].weight
        if l0.data_ptr() != self._fl_w.data_ptr() or l0.device != self._fl_w.device:
            n, W = len(self.first_linears), l0.shape[0]
            b0 = self.first_linears[0].bias
            es = l0.element_size()
            contiguous = all(l.weight.data_ptr() == l0.data_ptr() + i * W * W * es
                             and l.bias.data_ptr() == b0.data_ptr() + i * W * es
                             and l.weight.is_contiguous() and l.bias.is_contiguous()
                             for i, l in enumerate(self.first_linears))
            if contiguous:
                self._fl_w = torch.as_strided(l0.data, (n, W, W), (W * W, W, 1))
                self._fl_b = torch.as_strided(b0.data, (n, W), (W, 1))
            else:
                self._restack_first_linears()
        return self._fl_w, self._fl_b

    def preprocess(self, x):
        # kornia.geometry.resize(bicubic, align_corners=True, antialias=False) + CLIP normalisation (encoder.py:131-139)
        if self.antialias:
            raise NotImplementedError("antialias=True")
        x = F.interpolate(x.float(), size=(self.image_size, self.image_size), mode="bicubic", align_corners=True)
        x = (x + 1.0) / 2.0
        return (x - self.mean.view(1, 3, 1, 1)) / self.std.view(1, 3, 1, 1)

    def forward(self, x, unet_down_block_samples: tuple):
        maps = []
        for m in unet_down_block_samples:
            if m.dim() == 4 and m.is_cuda and m.dtype == torch.bfloat16 and m.permute(0, 2, 3, 1).is_contiguous():
                maps.append(m.permute(0, 2, 3, 1))      # channels-last view produced by the e4t UNet
            else:
                maps = None
                break
        if maps is not None:
            u = FN.MeanPoolCatFn.apply(*maps)                                                    # encoder.py:147-148
        else:   # foreign tensors (e.g. NCHW fp32 from another UNet implementation)
            u = torch.cat([m.float().mean(dim=(2, 3)) for m in unet_down_block_samples], dim=-1)
        u = self.unet_feature_embedder(u)                                                        # :149
        pooled, tokens = self.clip_vision(self.preprocess(x))                                    # :153-154
        hs = torch.cat([pooled.unsqueeze(1), tokens[:, 1::2, :]], dim=1).float()                 # :155-156  (B,n,W)
        n, W = hs.shape[1], hs.shape[2]
        if n != len(self.first_linears):
            raise ValueError(f"{n} CLIP states but {len(self.first_linears)} first_linears")
        wf, bf = self.feature_linear.weight, self.feature_linear.bias
        h = F.linear(hs, wf[:, :W]) + (F.linear(u, wf[:, W:]) + bf).unsqueeze(1)                 # :160 for all i
        wst, bst = self._stacked()
        if torch.is_grad_enabled() and self.first_linears[0].weight.requires_grad:
            wst, bst = _StackedParams.apply(self, wst, bst, *[p for l in self.first_linears for p in (l.weight, l.bias)])
        out = torch.baddbmm(bst.unsqueeze(1), h.transpose(0, 1), wst.transpose(1, 2))            # :161 (n,B,W)
        out = self.act(out.mean(dim=0))                                                          # :163-166
        return self.final_linear(out)                                                            # :168


class _StackedParams(torch.autograd.Function):
    """Identity on the stacked storage whose backward hands each per-layer parameter its slice of the stacked grad.
    With an arena optimiser (contiguous `.grad` views) the stacked gradient is added in ONE kernel per stack instead
    of 258 per-parameter accumulations."""

    @staticmethod
    def forward(ctx, module, wst, bst, *params):
        ctx.n = len(params) // 2
        ctx.module = module
        return wst.view_as(wst), bst.view_as(bst)

    @staticmethod
    def _arena_stack(lins, attr, shape):
        g0 = getattr(lins[0], attr).grad
        if g0 is None or not getattr(getattr(lins[0], attr), "_e4t_arena", False):
            return None
        step = g0.numel() * g0.element_size()
        for i, l in enumerate(lins):
            g = getattr(l, attr).grad
            if g is None or g.data_ptr() != g0.data_ptr() + i * step:
                return None
        return torch.as_strided(g0, shape, tuple([g0.numel()] + list(g0.stride())))

    @staticmethod
    def backward(ctx, dw, db):
        lins = ctx.module.first_linears
        if FN.DIRECT_GRAD_WRITE:
            gw = _StackedParams._arena_stack(lins, "weight", tuple(dw.shape))
            gb = _StackedParams._arena_stack(lins, "bias", tuple(db.shape))
            if gw is not None and gb is not None:
                gw.add_(dw)
                gb.add_(db)
                return (None, None, None) + (None,) * (2 * ctx.n)
        grads = []
        for i in range(ctx.n):
            grads += [dw[i], db[i]]
        return (None, None, None) + tuple(grads)
