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

Everything runs on the sm_100a kernels (round 2): the ViT tower (frozen by default; with freeze_clip_vision=False —
`--unfreeze_clip_vision` — every tower weight receives its gradient from the same kernels), and the head.  Random-init
construction never touches the network."""
import torch
import torch.nn.functional as F
from torch import nn

from e4t._mixins import ConfigMixin, ModelMixin, register_to_config
from e4t_b200 import functional as FN
from e4t_b200 import ops

_VIT_ARCHS = {
    # arch: (width, layers, heads, mlp, patch, image)
    "ViT-H-14": (1280, 32, 16, 5120, 14, 224),
    "ViT-L-14": (1024, 24, 16, 4096, 14, 224),
    "ViT-B-32": (768, 12, 12, 3072, 32, 224),
    "ViT-tiny-test": (64, 2, 4, 128, 14, 224),
}


def _bf16(p):
    """bf16 operand copy of a (possibly trainable) fp32 master, cached until the master changes."""
    return FN.prepared(p, "bf16", lambda t: t.to(torch.bfloat16).contiguous())


def _f32(p):
    return p if p.dtype == torch.float32 else FN.prepared(p, "f32", lambda t: t.float().contiguous())


class _MHA(nn.Module):
    """nn.MultiheadAttention-compatible parameter names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        nn.init.xavier_uniform_(self.in_proj_weight)

    def forward(self, h, residual):
        """in_proj (one fused QKV GEMM, bias in the epilogue) -> fused attention core -> out_proj (+bias +residual)."""
        W = h.shape[-1]
        qkv = FN.LinearFn.apply(h, _bf16(self.in_proj_weight), _f32(self.in_proj_bias), None, self.in_proj_weight)
        o = FN.AttentionFn.apply(qkv, None, self.heads, (W // self.heads) ** -0.5)
        return FN.LinearFn.apply(o, _bf16(self.out_proj.weight), _f32(self.out_proj.bias), residual, self.out_proj.weight)


class _MLP(nn.Module):
    def __init__(self, width, mlp):
        super().__init__()
        self.c_fc = nn.Linear(width, mlp)
        self.gelu = nn.GELU()
        self.c_proj = nn.Linear(mlp, width)

    def forward(self, h, residual):
        h = FN.LinearFn.apply(h, _bf16(self.c_fc.weight), _f32(self.c_fc.bias), None, self.c_fc.weight)
        h = FN.ActFn.apply(h, ops.ACT_GELU)
        return FN.LinearFn.apply(h, _bf16(self.c_proj.weight), _f32(self.c_proj.bias), residual, self.c_proj.weight)


def _ln(norm, x):
    return FN.LayerNormFn.apply(x, _f32(norm.weight), _f32(norm.bias), norm.eps)


class _ResBlock(nn.Module):
    def __init__(self, width, heads, mlp):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = _MHA(width, heads)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = _MLP(width, mlp)

    def forward(self, x):
        x = self.attn(_ln(self.ln_1, x), x)            # x + attn(ln_1(x)): the residual add rides in out_proj's epilogue
        return self.mlp(_ln(self.ln_2, x), x)          # x + mlp(ln_2(x))


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(width, heads, mlp) for _ in range(layers)])


class VisionTransformer(nn.Module):
    """open_clip VisionTransformer with proj=None, output_tokens=True: returns (ln_post(cls), patch tokens), on the
    sm_100a kernels: patch embedding as one GEMM over unfolded 14x14 patches, LayerNorm kernel, fused-QKV tcgen05 GEMMs
    with bias/residual epilogues, the two-tile attention forward (N = 257, 16 heads x 80), GELU kernel.

    `ln_post_on_tokens`: open_clip changed what `output_tokens=True` returns — up to 2.2x the patch tokens come back
    WITHOUT ln_post, later releases apply ln_post to them first.  The reference pins no version;
    the default is the contemporaneous behaviour (no ln_post on the tokens)."""

    def __init__(self, width, layers, heads, mlp, patch, image, ln_post_on_tokens=False):
        super().__init__()
        self.output_tokens = True
        self.proj = None
        self.patch = patch
        self.ln_post_on_tokens = ln_post_on_tokens
        grid = image // patch
        self.grid = grid
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(grid * grid + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads, mlp)
        self.ln_post = nn.LayerNorm(width)

    def _patch_weight(self):
        """conv1 weight as a (W, 3*p*p padded to a multiple of 8) bf16 GEMM operand (TMA row strides are 16-byte units)."""
        def prep(w):
            w2 = w.reshape(w.shape[0], -1)
            kp = (w2.shape[1] + 7) // 8 * 8
            out = torch.zeros((w2.shape[0], kp), device=w.device, dtype=torch.bfloat16)
            out[:, :w2.shape[1]] = w2.to(torch.bfloat16)
            return out
        return FN.prepared(self.conv1.weight, "patch_bf16", prep)

    def forward(self, x, compute_dtype=torch.bfloat16):
        if not x.is_cuda:
            from e4t_b200._lib import E4TError
            raise E4TError("e4t VisionTransformer runs on the sm_100a kernels only (no CPU fallback)")
        trainable = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        W, p, g = self.conv1.out_channels, self.patch, self.grid
        B = x.shape[0]
        with torch.set_grad_enabled(trainable):
            wp = self._patch_weight()
            # conv 14x14 / stride 14 == GEMM over unfolded patches (pixels have no gradient)
            cols = x.detach().reshape(B, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * p * p)
            a = torch.zeros((B, g * g, wp.shape[1]), device=x.device, dtype=torch.bfloat16)
            a[..., :3 * p * p] = cols
            tok = FN.LinearFn.apply(a, wp, None, None, self.conv1.weight)                       # (B, g*g, W)
            cls = self.class_embedding.to(torch.bfloat16).expand(B, 1, W)
            h = torch.cat([cls, tok], dim=1) + self.positional_embedding.to(torch.bfloat16)
            h = _ln(self.ln_pre, h)
            for blk in self.transformer.resblocks:
                h = blk(h)
            pooled = _ln(self.ln_post, h[:, 0])
            tokens = h[:, 1:]
            if self.ln_post_on_tokens:
                tokens = _ln(self.ln_post, tokens)
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
        self.clip_vision = VisionTransformer(width, layers, heads, mlp, patch, image,
                                             ln_post_on_tokens=bool(kwargs.get("ln_post_on_tokens", False)))
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

    def image_features(self, x):
        """(pooled, tokens) of the conditioning image: step-invariant in the sampling loop, where the pipeline computes
        them once per call and passes them back through forward(clip_features=...)."""
        with torch.no_grad():
            return self.clip_vision(self.preprocess(x))

    def forward(self, x, unet_down_block_samples: tuple, clip_features=None):
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
        bf = torch.bfloat16
        fe0, fe2 = self.unet_feature_embedder[0], self.unet_feature_embedder[2]
        u = FN.LinearFn.apply(u.to(bf), _bf16(fe0.weight), fe0.bias, None, fe0.weight)               # :149  (B,W)
        u = FN.LinearFn.apply(FN.ActFn.apply(u, ops.ACT_LEAKY_RELU), _bf16(fe2.weight), fe2.bias, None, fe2.weight)
        pooled, tokens = clip_features if clip_features is not None else self.clip_vision(self.preprocess(x))  # :153-154
        hs = torch.cat([pooled.unsqueeze(1), tokens[:, 1::2, :]], dim=1)                         # :155-156  (B,n,W)
        B, n, W = hs.shape
        if n != len(self.first_linears):
            raise ValueError(f"{n} CLIP states but {len(self.first_linears)} first_linears")
        fl = self.feature_linear
        # feature_linear(cat[hs_i, u]) for all i at once: one (B*n, 2W) x (2W, W) GEMM                       :160
        h = FN.LinearFn.apply(torch.cat([hs, u.unsqueeze(1).expand(B, n, W)], dim=-1), _bf16(fl.weight), fl.bias, None,
                              fl.weight)
        wst, bst = self._stacked()
        if torch.is_grad_enabled() and self.first_linears[0].weight.requires_grad:
            wst, bst = _StackedParams.apply(self, wst, bst, *[p for l in self.first_linears for p in (l.weight, l.bias)])
        # the 129 first_linears as ONE batched GEMM; their biases enter through the mean (mean_i(b_i) is exact)  :161-166
        w16 = FN.prepared(self.first_linears[0].weight, ("stack_bf16", n),
                          lambda t: self._stacked()[0].detach().to(bf).contiguous())
        out = FN.BatchedLinearFn.apply(h.transpose(0, 1), w16, wst)                               # (n,B,W)
        out = out.float().mean(dim=0) + bst.float().mean(dim=0)
        out = FN.ActFn.apply(out.to(bf), ops.ACT_LEAKY_RELU)                                      # :167
        fin = self.final_linear
        return FN.LinearFn.apply(out, _bf16(fin.weight), fin.bias, None, fin.weight).float()      # :168


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
