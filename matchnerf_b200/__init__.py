"""matchnerf_b200: B200-native (sm_100a) implementation of MatchNeRF's per-ray hot path.

The product is ``libmatchnerf_b200.so`` (hand-written CUDA behind the C ABI in ``include/matchnerf_b200.h``);
this package is the Python host side mirroring the reference's model interface.
"""
from .utils import AttrDict  # noqa: F401

__all__ = ["AttrDict", "models_dict", "MatchNeRF", "datas_dict", "EvalTools"]


def __getattr__(name):
    if name in ("models_dict", "MatchNeRF"):
        from . import matchnerf as _m
        return getattr(_m, name)
    if name == "datas_dict":                    # datasets/__init__.py of the reference (the LLFF / COLMAP-layout loader)
        from . import datasets as _d
        return _d.datas_dict
    if name == "EvalTools":                     # misc/metrics.py of the reference, on the GPU
        from . import metrics as _e
        return _e.EvalTools
    raise AttributeError(name)
