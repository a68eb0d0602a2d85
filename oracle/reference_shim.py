"""Import shim for the UNMODIFIED reference (test / benchmark infrastructure; never imported by the product package).

The reference's import chain needs four third-party modules that this image lacks: ``easydict``, ``ipdb``,
``termcolor`` and ``skvideo.io``.  They are stubbed in ``sys.modules`` (none of them does arithmetic) and the reference tree
is put on ``sys.path``: ``/root/reference`` in the dev container, ``baseline/_ref`` (a verbatim copy made by
``baseline/install_reference.py``; git-ignored, travels with gpurun) on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EasyDict(dict):
    """Minimal stand-in for easydict.EasyDict: attribute access, recursive wrapping of dicts, ``update``."""

    def __init__(self, d=None, **kw):
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k in d:
            setattr(self, k, d[k])


def reference_root():
    """Directory of the reference checkout, or None: $MATCHNERF_REFERENCE, /root/reference, baseline/_ref."""
    for cand in (os.environ.get("MATCHNERF_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models")) and os.path.isfile(os.path.join(cand, "configs", "base.yaml")):
            return cand
    return None


def install_shim(ref: str | None = None) -> str:
    """Stub the absent modules and put the reference on sys.path.  Returns the reference root (raises if absent)."""
    ref = ref or reference_root()
    if ref is None:
        raise RuntimeError("the reference checkout is not available (neither /root/reference nor baseline/_ref)")
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    for name in ("ipdb", "termcolor", "skvideo", "skvideo.io"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ipdb"].set_trace = lambda *a, **k: None
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    sys.modules["skvideo"].io = sys.modules["skvideo.io"]
    if ref not in sys.path:
        sys.path.insert(0, ref)
    return ref


def reference_options(S: int, device: str = "cpu", ref: str | None = None, **over):
    """configs/base.yaml of the reference as an EasyDict, with ``nerf.sample_intvs = S`` and dotted-key overrides."""
    import yaml
    ref = ref or reference_root()
    opt = EasyDict(yaml.safe_load(open(os.path.join(ref, "configs/base.yaml"))))
    opt.device = device
    opt.nerf.sample_intvs = S
    for k, v in over.items():
        node = opt
        ks = k.split(".")
        for kk in ks[:-1]:
            node = getattr(node, kk)
        setattr(node, ks[-1], v)
    return opt
