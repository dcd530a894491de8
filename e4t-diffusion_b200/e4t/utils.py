"""Checkpoint / config helpers — the state-dict key contract of e4t/utils.py:76-159:
  weight_offsets.pt = {k: v for k in unet.state_dict() if "wo" in k}   (utils.py:129-131)
  encoder.pt        = encoder.state_dict()                               (utils.py:158-159)
  config.json       = argparse namespace, read back as AttributeDict     (utils.py:17-40,76-89)
Hub download (utils.py:43-73) needs network and is out of scope; local paths work the same way."""
import json
import os

import torch

from e4t.encoder import E4TEncoder
from e4t.models.unet_2d_condition import UNet2DConditionModel


class AttributeDict(dict):
    """dict with attribute access; a missing attribute reads as None (utils.py:17-40)."""

    def __getattr__(self, k):
        return self.get(k, None)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def load_config_from_pretrained(pretrained_model_name_or_path):
    if not os.path.exists(pretrained_model_name_or_path):
        raise FileNotFoundError(f"{pretrained_model_name_or_path} (hub download is unavailable offline)")
    if "config.json" not in pretrained_model_name_or_path:
        pretrained_model_name_or_path = os.path.join(pretrained_model_name_or_path, "config.json")
    with open(pretrained_model_name_or_path, "r", encoding="utf-8") as f:
        return AttributeDict(json.load(f))


def load_e4t_unet(pretrained_model_name_or_path=None, ckpt_path=None, **kwargs):
    """Base SD UNet weights from a local diffusers-format directory (<path>/unet/{config.json,*.bin}) overlaid with
    weight_offsets.pt / unet.pt; missing keys are tolerated only for a fresh WO init, unexpected keys are fatal
    (utils.py:92-126)."""
    assert pretrained_model_name_or_path is not None or ckpt_path is not None
    if pretrained_model_name_or_path is None:
        assert os.path.basename(ckpt_path) in ("unet.pt", "weight_offsets.pt"), \
            "You must specify the filename! (`unet.pt` or `weight_offsets.pt`)"
        config = load_config_from_pretrained(os.path.dirname(ckpt_path))
        pretrained_model_name_or_path = config.pretrained_model_name_or_path if config.pretrained_args is None \
            else config.pretrained_args["pretrained_model_name_or_path"]
    d = os.path.join(pretrained_model_name_or_path, "unet")
    with open(os.path.join(d, "config.json")) as f:
        cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
    unet = UNet2DConditionModel(**cfg)
    state_dict = {}
    for name in ("diffusion_pytorch_model.bin",):
        if os.path.exists(os.path.join(d, name)):
            state_dict.update(torch.load(os.path.join(d, name), map_location="cpu"))
    if ckpt_path:
        state_dict.update(torch.load(ckpt_path, map_location="cpu"))
        print(f"Resuming from {ckpt_path}")
    m, u = unet.load_state_dict(state_dict, strict=False)
    if len(m) > 0 and ckpt_path:
        raise RuntimeError(f"missing keys:\n{m}")
    if len(u) > 0:
        raise RuntimeError(f"unexpected keys:\n{u}")
    return unet


def save_e4t_unet(model, save_dir):
    os.makedirs(save_dir, exist_ok=True)
    torch.save({k: v for k, v in model.state_dict().items() if "wo" in k}, os.path.join(save_dir, "weight_offsets.pt"))


def load_e4t_encoder(ckpt_path=None, **kwargs):
    encoder = E4TEncoder(**kwargs)
    if ckpt_path:
        if not os.path.exists(ckpt_path):
            raise FileNotFoundError(f"{ckpt_path} (hub download is unavailable offline)")
        if "encoder.pt" not in ckpt_path:
            ckpt_path = os.path.join(ckpt_path, "encoder.pt")
        state_dict = torch.load(ckpt_path, map_location="cpu")
        print(f"Resuming from {ckpt_path}")
        m, u = encoder.load_state_dict(state_dict, strict=False)
        if len(m) > 0:
            raise RuntimeError(f"missing keys:\n{m}")
        if len(u) > 0:
            raise RuntimeError(f"unexpected keys:\n{u}")
    return encoder


def save_e4t_encoder(model, save_dir):
    os.makedirs(save_dir, exist_ok=True)
    torch.save(model.state_dict(), os.path.join(save_dir, "encoder.pt"))


def image_grid(imgs, rows, cols):
    from PIL import Image
    assert len(imgs) == rows * cols
    w, h = imgs[0].size
    grid = Image.new("RGB", size=(cols * w, rows * h))
    for i, img in enumerate(imgs):
        grid.paste(img, box=(i % cols * w, i // cols * h))
    return grid
