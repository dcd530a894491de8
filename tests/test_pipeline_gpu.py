"""Inference pipeline (BASELINE.json configs[4]): StableDiffusionE4TPipeline mirror on the sm_100a kernels
against the fp32 CPU oracle restatement of pipeline_stable_diffusion_e4t.py:181-216 + the SD-v1.x DDIM step."""
import types

import pytest
import torch

from oracle import e4t_oracle as O

pytestmark = pytest.mark.gpu


class _Tok:
    """Whitespace tokenizer over the oracle's fixed word ids (a real CLIP tokenizer needs its vocabulary file)."""
    model_max_length = 77

    def __init__(self):
        self.extra = {}

    def add_tokens(self, tok):
        if tok in self.extra:
            return 0
        self.extra[tok] = O.PLACEHOLDER_ID
        return 1

    def __len__(self):
        return 49408 + len(self.extra)

    def convert_tokens_to_ids(self, tok):
        return self.extra[tok]

    def __call__(self, text, padding=None, truncation=None, max_length=77, return_tensors=None, add_special_tokens=True):
        texts = [text] if isinstance(text, str) else text
        rows = []
        for s in texts:
            ids = [self.extra.get(w, O._WORD_IDS.get(w)) for w in s.split()]
            if add_special_tokens:
                ids = [O.BOS] + ids
                ids = ids + [O.EOS] * (max_length - len(ids))
            rows.append(ids)
        return types.SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.int64))


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("guidance", [7.5, 1.0])
def test_pipeline_denoising_loop_vs_oracle(guidance):
    from e4t.encoder import E4TEncoder
    from e4t.models.modeling_clip import CLIPTextConfig, CLIPTextModel
    from e4t.models.unet_2d_condition import UNet2DConditionModel
    from e4t.pipeline_stable_diffusion_e4t import DDIMScheduler, StableDiffusionE4TPipeline
    ucfg, vcfg, tcfg = O.TINY_UNET, O.VIT_TINY, O.CLIP_TEXT_TINY
    fd = O.pooled_feature_dim(ucfg)
    sd_u = O.synth_state_dict(O.unet_param_shapes(ucfg), 41)
    sd_e = O.synth_state_dict(O.encoder_param_shapes(vcfg, fd, tcfg["width"], 129), 42)
    sd_t = O.synth_state_dict(O.text_param_shapes(tcfg), 43)
    unet = UNet2DConditionModel(**O.ref_unet_kwargs(ucfg)); unet.load_state_dict(sd_u)
    enc = E4TEncoder(arch="ViT-tiny-test", word_embedding_dim=tcfg["width"], n_odd_layers=129, unet_feature_dim=fd)
    enc.load_state_dict(sd_e)
    text = CLIPTextModel(CLIPTextConfig(vocab_size=tcfg["vocab"] - 1, hidden_size=tcfg["width"],
                                        intermediate_size=tcfg["mlp"], num_hidden_layers=tcfg["layers"],
                                        num_attention_heads=tcfg["heads"]))
    sd_t_small = dict(sd_t)
    sd_t_small["text_model.embeddings.token_embedding.weight"] = sd_t["text_model.embeddings.token_embedding.weight"][:-1]
    text.load_state_dict(sd_t_small)
    cfg = types.SimpleNamespace(placeholder_token="*s", domain_class_token="a", domain_embed_scale=0.1)
    pipe = StableDiffusionE4TPipeline(None, text.cuda(), _Tok(), unet.cuda(), enc.cuda(), DDIMScheduler(), e4t_config=cfg)
    assert text.get_input_embeddings().num_embeddings == tcfg["vocab"]          # resized for the placeholder (:53)
    with torch.no_grad():   # the new row is random-initialised by resize_token_embeddings; pin it to the oracle's
        text.get_input_embeddings().weight[-1] = sd_t["text_model.embeddings.token_embedding.weight"][-1].cuda()
    g = torch.Generator().manual_seed(3)
    image = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    latents = torch.randn(2, 4, 16, 16, generator=g)
    prompt = ["a photo of *s", "a photo of *s"]
    out = pipe(prompt, num_inference_steps=4, guidance_scale=guidance, latents=latents.clone(), image=image,
               output_type="latent").images
    ids = pipe.tokenizer(prompt, max_length=77).input_ids
    ref = O.pipeline_sample(sd_u, ucfg, sd_e, vcfg, sd_t, tcfg, image, ids, latents, num_inference_steps=4,
                            guidance_scale=guidance, class_token_id=O._WORD_IDS["a"])
    e = _rel(out, ref)
    print(f"[pipeline] guidance {guidance}: latents after 4 DDIM steps rel err {e:.3e}")
    assert out.shape == (2, 4, 16, 16) and e < (8e-2 if guidance > 1 else 4e-2)
