"""Minimal stand-ins for the diffusers plumbing the reference modules inherit from
(diffusers.configuration_utils.ConfigMixin/register_to_config, diffusers.models.modeling_utils.ModelMixin,
diffusers.utils.BaseOutput): `.config` with attribute + item access, `.dtype`, `.device`, no-op xformers toggles."""
import functools
import inspect
import json
import os
from collections import OrderedDict
from dataclasses import fields

import torch


class FrozenDict(OrderedDict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict

    def register_to_config(self, **kwargs):
        self._internal_dict = FrozenDict(kwargs)

    def save_config(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}
        cfg["_class_name"] = type(self).__name__
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(cfg, f, indent=2)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for p in list(sig.parameters.values())[1:] if p.kind != p.VAR_KEYWORD]
        cfg = OrderedDict((p.name, p.default) for p in params)
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        self._internal_dict = FrozenDict(cfg)
        init(self, *args, **kwargs)

    return inner


class ModelMixin(torch.nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.float32

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    # the training scripts call these (pretrain_e4t.py:264-272); attention already runs on the fused sm_100a kernel
    def enable_xformers_memory_efficient_attention(self, attention_op=None):
        return None

    def disable_xformers_memory_efficient_attention(self):
        return None

    def enable_gradient_checkpointing(self):
        if not self._supports_gradient_checkpointing:
            raise ValueError(f"{type(self).__name__} does not support gradient checkpointing.")
        self.gradient_checkpointing = True

    def save_pretrained(self, save_directory):
        self.save_config(save_directory)
        torch.save(self.state_dict(), os.path.join(save_directory, "diffusion_pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kwargs):
        d = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(d, cls.config_name)) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        model = cls(**cfg)
        w = os.path.join(d, "diffusion_pytorch_model.bin")
        if os.path.exists(w):
            model.load_state_dict(torch.load(w, map_location="cpu"), strict=False)
        return model


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())
