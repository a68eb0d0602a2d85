"""Small host-side helpers (attribute dicts for options / batches)."""
from __future__ import annotations

from typing import Any


class AttrDict(dict):
    """dict with attribute access and recursive wrapping -- stands in for ``easydict.EasyDict``,
    which the reference uses for options and batches (options.py, coach.py) but is not in this image."""

    def __init__(self, d=None, **kw):
        super().__init__()
        src = dict(d or {})
        src.update(kw)
        for k, v in src.items():
            self[k] = v

    @staticmethod
    def _wrap(v: Any) -> Any:
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __delattr__(self, k):
        del self[k]

    def update(self, e=None, **f):
        src = dict(e or {})
        src.update(f)
        for k, v in src.items():
            self[k] = v


def get_opt(node: Any, path: str, default: Any = None) -> Any:
    """``get_opt(opts, 'decoder.raytrans_act', 'ReLU')`` on attribute- or key-style option trees."""
    cur = node
    for part in path.split("."):
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(part, None)
        else:
            cur = getattr(cur, part, None)
    return default if cur is None else cur
