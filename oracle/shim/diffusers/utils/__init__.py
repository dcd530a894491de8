import logging as _pylog
from collections import OrderedDict
from dataclasses import fields


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylog.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    return None


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
