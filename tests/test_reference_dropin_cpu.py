"""CPU, dev container only (skipped where /root/reference is absent, e.g. on the GPU box): the drop-in surface against the
UNMODIFIED reference -- checkpoint compatibility of the whole module tree and the reference Coach building this repo's
model through its own registry lookup (coach.py:75-85)."""
import os
import sys

import pytest
import torch

REF = os.environ.get("MATCHNERF_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref_env():
    from oracle.make_golden import install_shim, ref_options
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("models", "misc", "datasets", "coach", "options")}
    install_shim()
    import types
    for name in ("imageio", "lpips", "skimage", "skimage.metrics"):             # coach.py / misc/metrics.py imports absent from this image
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            saved_mods.setdefault(name, None)
    sys.modules["skimage"].metrics = sys.modules["skimage.metrics"]
    sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: 0.0
    sys.modules["lpips"].LPIPS = object
    yield ref_options
    sys.path[:] = saved_path
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "misc", "datasets", "coach", "options")]:
        del sys.modules[k]
    for k, v in saved_mods.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_state_dict_is_interchangeable_with_the_reference(ref_env):
    """misc/utils.py:183-205 restore_checkpoint loads per child with strict=True: every key and shape must match, both ways."""
    from models.matchnerf import MatchNeRF as RefMatchNeRF                      # the reference
    from matchnerf_b200.matchnerf import MatchNeRF
    opt = ref_env(64)
    torch.manual_seed(0)
    ref = RefMatchNeRF(opt)
    ours = MatchNeRF(opt)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())
    assert all(sd_ref[k].shape == sd_ours[k].shape for k in sd_ref)
    for name, child in ours.named_children():                                   # the way restore_checkpoint does it
        child.load_state_dict(dict(getattr(ref, name).state_dict()), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    assert {n for n, _ in ours.feat_enc.named_children()} >= {"backbone", "transformer", "featup_net"}      # misc/utils.py:177-180
    assert sum(p.numel() for p in ours.feat_enc.parameters()) == 4642208 and sum(p.numel() for p in ours.nerf_dec.parameters()) == 130324


def test_reference_coach_builds_this_model(ref_env, monkeypatch):
    """coach.py:77 ``models_dict[opts.model](opts).to(device)`` with the registry pointed at this package (INTEGRATION.md
    option A), then the attribute pokes coach does (coach.py:79-102, :383)."""
    import models as ref_models                                                 # the reference's registry module
    from matchnerf_b200.matchnerf import models_dict
    monkeypatch.setattr(ref_models, "models_dict", models_dict)
    import coach as ref_coach
    monkeypatch.setattr(ref_coach, "models_dict", models_dict, raising=False)
    opt = ref_env(64)
    opt.model = "matchnerf"
    opt.gpu_ids = [0]
    opt.encoder.pretrain_weight = None
    opt.load, opt.resume = None, False
    c = ref_coach.Coach.__new__(ref_coach.Coach)
    c.opts = opt
    c.build_networks()
    c.setup_optimizer() if hasattr(opt, "optim") and hasattr(opt, "max_epoch") else None
    m = c.model
    assert type(m).__module__.startswith("matchnerf_b200") and hasattr(m, "nerf_setbg_opaque")
    m.nerf_setbg_opaque = True
    assert len(list(m.feat_enc.parameters())) > 0 and len(list(m.nerf_dec.parameters())) > 0
    assert callable(getattr(m.nerf_dec, "composite"))
