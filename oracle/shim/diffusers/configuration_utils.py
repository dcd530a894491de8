import functools
import inspect


class FrozenDict(dict):
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


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for p in list(sig.parameters.values())[1:] if p.kind != p.VAR_KEYWORD]
        cfg = {p.name: p.default for p in params}
        for name, a in zip([p.name for p in params], args):
            cfg[name] = a
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        self._internal_dict = FrozenDict(cfg)
        init(self, *args, **kwargs)

    return inner
