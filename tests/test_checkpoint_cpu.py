"""Checkpoint formats and loader parity: weight_offsets.pt / encoder.pt / unet.pt / config.json of
e4t/utils.py:76-159, pretrain_e4t.py:515-528, tuning_e4t.py:225-240 — module surface only, runs on CPU."""
import json
import os
import types

import pytest
import torch

from oracle import e4t_oracle as O


def _tiny_unet():
    from e4t.models.unet_2d_condition import UNet2DConditionModel
    m = UNet2DConditionModel(**O.ref_unet_kwargs(O.TINY_UNET))
    m.load_state_dict(O.synth_state_dict(O.unet_param_shapes(O.TINY_UNET), 61))
    return m


def _write_sd_dir(root, unet, fmt):
    d = os.path.join(root, "unet")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(dict(O.ref_unet_kwargs(O.TINY_UNET), _class_name="UNet2DConditionModel"), f)
    base = {k: v.clone() for k, v in unet.state_dict().items() if "wo" not in k}     # what a stock SD checkpoint holds
    if fmt == "safetensors":
        from safetensors.torch import save_file
        save_file(base, os.path.join(d, "diffusion_pytorch_model.safetensors"))
    elif fmt == "bin":
        torch.save(base, os.path.join(d, "diffusion_pytorch_model.bin"))
    elif fmt == "sharded":
        from safetensors.torch import save_file
        keys = sorted(base)
        half = len(keys) // 2
        wm = {}
        for i, ks in enumerate((keys[:half], keys[half:])):
            name = f"diffusion_pytorch_model-0000{i + 1}-of-00002.safetensors"
            save_file({k: base[k] for k in ks}, os.path.join(d, name))
            wm.update({k: name for k in ks})
        with open(os.path.join(d, "diffusion_pytorch_model.safetensors.index.json"), "w") as f:
            json.dump({"weight_map": wm}, f)


@pytest.mark.parametrize("fmt", ["safetensors", "bin", "sharded"])
def test_load_e4t_unet_from_local_diffusers_dir(tmp_path, fmt):
    from e4t import utils
    src = _tiny_unet()
    _write_sd_dir(str(tmp_path), src, fmt)
    m = utils.load_e4t_unet(str(tmp_path))                       # fresh run: WO params keep their default init
    sd, ref = m.state_dict(), src.state_dict()
    assert all(torch.equal(sd[k], ref[k]) for k in sd if "wo" not in k)
    # weight_offsets.pt round trip + config.json resolution from the checkpoint directory alone (utils.py:98-106)
    run = tmp_path / "run" / "1000"
    utils.save_e4t_unet(src, str(run))
    wo = torch.load(run / "weight_offsets.pt")
    assert wo and all("wo" in k for k in wo) and all(v.untyped_storage().nbytes() == v.numel() * 4 for v in wo.values())
    utils.save_config(types.SimpleNamespace(pretrained_model_name_or_path=str(tmp_path), placeholder_token="*s"), str(run))
    m2 = utils.load_e4t_unet(ckpt_path=str(run / "weight_offsets.pt"))
    assert all(torch.equal(v, ref[k]) for k, v in m2.state_dict().items())
    # tuning: full unet.pt + nested pretrained_args (tuning_e4t.py:225-232, utils.py:104-106)
    tune = tmp_path / "tune" / "15"
    utils.save_e4t_unet(src, str(tune), save_all=True)
    utils.save_config({"learning_rate": 1e-6}, str(tune), pretrained_args={"pretrained_model_name_or_path": str(tmp_path)})
    cfg = utils.load_config_from_pretrained(str(tune))
    assert cfg.pretrained_args["pretrained_model_name_or_path"] == str(tmp_path) and cfg.not_there is None
    m3 = utils.load_e4t_unet(ckpt_path=str(tune / "unet.pt"))
    assert all(torch.equal(v, ref[k]) for k, v in m3.state_dict().items())


def test_load_e4t_unet_refuses_missing_or_foreign_weights(tmp_path):
    from e4t import utils
    src = _tiny_unet()
    _write_sd_dir(str(tmp_path), src, "bin")
    os.remove(tmp_path / "unet" / "diffusion_pytorch_model.bin")
    with pytest.raises(FileNotFoundError):                       # never a silent random init
        utils.load_e4t_unet(str(tmp_path))
    _write_sd_dir(str(tmp_path), src, "bin")
    sd = torch.load(tmp_path / "unet" / "diffusion_pytorch_model.bin")
    sd["not.a.key"] = torch.zeros(1)
    torch.save(sd, tmp_path / "unet" / "diffusion_pytorch_model.bin")
    with pytest.raises(RuntimeError, match="unexpected keys"):
        utils.load_e4t_unet(str(tmp_path))
    with pytest.raises(AssertionError, match="specify the filename"):
        utils.load_e4t_unet(ckpt_path=str(tmp_path / "something.pt"))


def test_encoder_checkpoint_roundtrip_and_strictness(tmp_path):
    from e4t import utils
    kw = dict(arch="ViT-tiny-test", word_embedding_dim=64, n_odd_layers=129, unet_feature_dim=448)
    enc = utils.load_e4t_encoder(**kw)
    utils.save_e4t_encoder(enc, str(tmp_path))
    sd = torch.load(tmp_path / "encoder.pt")
    assert set(sd) == set(enc.state_dict())
    assert all(v.untyped_storage().nbytes() == v.numel() * v.element_size() for v in sd.values())   # clones, not arena views
    enc2 = utils.load_e4t_encoder(ckpt_path=str(tmp_path), **kw)
    assert all(torch.equal(v, sd[k]) for k, v in enc2.state_dict().items())
    del sd["final_linear.bias"]
    torch.save(sd, tmp_path / "encoder.pt")
    with pytest.raises(RuntimeError, match="missing keys"):
        utils.load_e4t_encoder(ckpt_path=str(tmp_path), **kw)


def test_attention_api_surface():
    """get_attention_scores / prepare_attention_mask / head reshapes of CrossAttention (cross_attention.py:208-282)."""
    from e4t.models.cross_attention import CrossAttention
    attn = CrossAttention(query_dim=32, heads=4, dim_head=8)
    x = torch.randn(2, 5, 32)
    q = attn.head_to_batch_dim(x)
    assert q.shape == (8, 5, 8) and torch.equal(attn.batch_to_head_dim(q), x)
    p = attn.get_attention_scores(q, q)
    ref = (torch.bmm(q, q.transpose(1, 2)) * attn.scale).softmax(-1)
    assert torch.allclose(p, ref, atol=1e-6) and torch.allclose(p.sum(-1), torch.ones(8, 5), atol=1e-6)
    mask = torch.zeros(2, 1, 3)
    m = attn.prepare_attention_mask(mask, target_length=4, batch_size=2)
    assert m.shape == (8, 1, 7) and attn.prepare_attention_mask(None, 4, 2) is None
    pm = attn.get_attention_scores(q, q, torch.full((8, 5, 5), float("-inf")).triu(1))
    assert torch.allclose(pm[:, 0, 1:], torch.zeros(8, 4))


def test_load_image(tmp_path):
    from PIL import Image
    from e4t import utils
    Image.new("RGB", (300, 200), (10, 200, 30)).save(tmp_path / "a.png")
    assert utils.load_image(str(tmp_path / "a.png")).size == (300, 200)
    assert utils.load_image(str(tmp_path / "a.png"), resolution=64).size == (64, 64)
