"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is the pin that lets the GPU tests trust the oracle."""
import numpy as np
import pytest
import torch

from oracle import encoder_oracle as EO
from oracle import render_oracle as RO
from oracle import synth
from tests.helpers import config1_inputs, dec_from_npz, load_npz, oracle_render, rms


@pytest.mark.parametrize("S", [64, 128])
def test_known_answer_reference_init(golden_dir, S):
    """SURVEY.md 8(c) recipe: reference default init (seed 0), 1024 rays of the 512x640 synthetic scene."""
    z = load_npz(golden_dir, f"config1_refinit_S{S}.npz")
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    assert np.array_equal(ray_idx.numpy(), z["ray_idx"])
    rgb, depth, op = oracle_render(dec_from_npz(z), feats, imgs, extr, intr, nf, ray_idx, S)
    assert rms(rgb, z["rgb"]) < 1e-6 and rms(depth, z["depth"]) < 2e-6 and rms(op, z["opacity"]) < 1e-6
    exp_mean = {64: 0.1042879, 128: 0.1714871}[S]
    assert abs(float(z["rgb"].mean()) - exp_mean) < 2e-6


@pytest.mark.parametrize("S", [64, 128])
def test_config1_synthetic_weights(golden_dir, S):
    z = load_npz(golden_dir, f"config1_synth_S{S}.npz")
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    dec = synth.synthetic_decoder(seed=0)
    rgb, depth, op = oracle_render(dec, feats, imgs, extr, intr, nf, ray_idx, S)
    assert 0.05 < float(z["opacity"].mean()) < 0.95           # non-degenerate (SURVEY 8c warning)
    assert rms(rgb, z["rgb"]) < 1e-6 and rms(depth, z["depth"]) < 3e-6 and rms(op, z["opacity"]) < 1e-6


SMALL_CASES = ["small_base", "small_elu_maskfill_posenc_bg", "small_s24_elu_maskfill_posenc_bg", "small_wide_baseline",
               "demo_own_S128", "video_own_S256", "small_local_radius1"]


@pytest.mark.parametrize("name", SMALL_CASES)
def test_small_cases_all_options(golden_dir, name):
    z = load_npz(golden_dir, name + ".npz")
    dec = dec_from_npz(z)
    S = int(z["S"])
    feats = [torch.from_numpy(z["feat8"]), torch.from_numpy(z["feat4"])]
    out = oracle_render(dec, feats, torch.from_numpy(z["images"]), torch.from_numpy(z["extrinsics"]),
                        torch.from_numpy(z["intrinsics"]), torch.from_numpy(z["near_fars"]),
                        torch.from_numpy(z["ray_idx"]), S, setbg_opaque=bool(z["setbg_opaque"]),
                        raytrans_act=str(z["raytrans_act"]), raytrans_posenc=bool(z["raytrans_posenc"]),
                        density_maskfill=bool(z["density_maskfill"]), return_aux=True,
                        local_radius=int(z["local_radius"]), local_dilation=int(z["local_dilation"]))
    assert rms(out[3]["cond"], z["cond"]) < 1e-6
    assert rms(out[0], z["rgb"]) < 1e-6 and rms(out[1], z["depth"]) < 3e-6 and rms(out[2], z["opacity"]) < 1e-6
    # the wide-baseline case must exercise out-of-view samples (mask = 0)
    if name in ("small_wide_baseline", "demo_own_S128", "video_own_S256"):
        assert float(z["cond"][:, 19:].mean()) < 0.999
    if bool(z["density_maskfill"]) and name != "small_elu_maskfill_posenc_bg" and name != "small_s24_elu_maskfill_posenc_bg":
        assert float((z["cond"][:, 19:].sum(1) < 1).mean()) > 0.01     # density_maskfill has samples to act on


@pytest.mark.parametrize("name", ["window_attn_8x12_k2_s0", "window_attn_8x12_k2_s1", "window_attn_12x16_k4_s1",
                                  "window_attn_6x10_k1_s0"])
def test_window_attention(golden_dir, name):
    z = load_npz(golden_dir, name + ".npz")
    out = EO.window_attention(torch.from_numpy(z["q"]), torch.from_numpy(z["k"]), torch.from_numpy(z["v"]),
                              int(z["h"]), int(z["w"]), int(z["num_splits"]), bool(z["with_shift"]))
    assert rms(out, z["out"]) < 1e-6


def test_encoder_end_to_end(golden_dir):
    z = load_npz(golden_dir, "encoder_64x96.npz")
    sd = synth.synthetic_encoder(seed=1)
    f8, f4 = EO.encode_views(sd, torch.from_numpy(z["images"])[0])
    assert rms(f8, z["feat8"]) < 5e-5
    assert rms(f4[:, ::8], z["feat4_ch0mod8"]) < 5e-5


def test_composite_properties():
    """Size-independent properties of the compositing step: weights are a sub-probability distribution,
    zero density renders nothing, huge density at sample 0 renders exactly that sample."""
    g = torch.Generator().manual_seed(0)
    sigma = torch.rand(7, 33, generator=g)
    rgb = torch.rand(7, 33, 3, generator=g)
    depth = torch.linspace(2, 4, 33)[None].expand(7, 33)
    c, d, o, w = RO.composite(sigma, rgb, depth)
    assert torch.all(w >= 0) and torch.all(o <= 1 + 1e-6)
    assert torch.allclose(o[:, 0], 1 - torch.exp(-sigma.sum(1)), atol=1e-5)      # telescoping sum
    c0, d0, o0, _ = RO.composite(torch.zeros(2, 5), torch.rand(2, 5, 3), torch.ones(2, 5))
    assert float(o0.abs().max()) == 0 and float(c0.abs().max()) == 0
    s = torch.zeros(1, 4)
    s[0, 0] = 1e4
    c1, d1, o1, _ = RO.composite(s, rgb[:1, :4], depth[:1, :4])
    assert torch.allclose(c1[0], rgb[0, 0]) and abs(float(o1) - 1) < 1e-6
    cb, _, ob, _ = RO.composite(torch.zeros(1, 4), rgb[:1, :4], depth[:1, :4], setbg_opaque=True)
    assert torch.allclose(cb, torch.ones(1, 3))
