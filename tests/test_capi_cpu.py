"""CPU: the C-ABI library builds, loads, and exports every symbol include/matchnerf_b200.h declares; the ctypes
mirrors of its structs have the layout the C compiler gives them; argument validation works without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "matchnerf_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("mnf_gather_cossim_fwd", "mnf_decoder_composite_fwd", "mnf_render_rays_fwd", "mnf_window_attn_fwd",
                 "mnf_ctx_create", "mnf_decoder_load_host", "mnf_pack_features", "mnf_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_struct_layout_matches_c(lib_path, tmp_path):
    from matchnerf_b200 import capi
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "matchnerf_b200.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(mnf_decoder_cfg), sizeof(mnf_scene), sizeof(mnf_rays));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(mnf_scene, feat0), offsetof(mnf_scene, images), offsetof(mnf_scene, src_w2c),
         offsetof(mnf_scene, src_K), offsetof(mnf_scene, tgt_c2w), offsetof(mnf_scene, tgt_near_far));
  printf("%zu %zu\n", offsetof(mnf_scene, sample_local_radius), offsetof(mnf_scene, sample_local_dilation));
  printf("%zu %zu %zu\n", offsetof(mnf_rays, ray_idx), offsetof(mnf_rays, first_ray), offsetof(mnf_rays, jitter));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    got = [int(x) for x in out]
    S, R = capi.Scene, capi.Rays
    exp = [C.sizeof(capi.DecoderCfg), C.sizeof(S), C.sizeof(R),
           S.feat0.offset, S.images.offset, S.src_w2c.offset, S.src_K.offset, S.tgt_c2w.offset, S.tgt_near_far.offset,
           S.sample_local_radius.offset, S.sample_local_dilation.offset,
           R.ray_idx.offset, R.first_ray.offset, R.jitter.offset]
    assert got == exp


def test_param_count_and_order(lib_path, golden_dir):
    from matchnerf_b200 import capi
    from oracle import synth
    from tests.helpers import load_npz
    lib = capi.load()
    shapes = synth.decoder_param_shapes()
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert n == 130324 == lib.mnf_decoder_param_count()
    assert list(shapes.keys()) == list(capi.DECODER_PARAM_ORDER)
    # the order of the unmodified reference's nerf_dec.state_dict() (stored by make_golden.py)
    z = load_npz(golden_dir, "config1_refinit_S64.npz")
    ref_order = [k[4:] for k in z if k.startswith("dec.")]
    assert ref_order == list(capi.DECODER_PARAM_ORDER)
    blob = capi.flatten_decoder_state(synth.synthetic_decoder())
    assert blob.numel() == n and blob.dtype == torch.float32


def test_calls_fail_loudly_without_gpu(lib_path):
    """No GPU in the dev container: context creation must fail with a CUDA error, never fall back."""
    from matchnerf_b200 import capi
    lib = capi.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.mnf_ctx_create(0, C.byref(h))
    assert rc != 0 and len(lib.mnf_last_error()) > 0
    with pytest.raises(RuntimeError):
        capi.Context()
    assert lib.mnf_render_workspace_bytes(1024, 64, 1) >= 1024 * 64 * 22 * 4          # fp32 rows for the fp32 kernel
    assert 1024 * 64 * 32 * 2 <= lib.mnf_render_workspace_bytes(1024, 64, 2) < 1024 * 64 * 32 * 2 + 256   # fp16 rows only
    assert lib.mnf_render_workspace_bytes(1024, 64, 0) == lib.mnf_render_workspace_bytes(1024, 64, 2)       # auto = tcgen05 at S=64
    assert lib.mnf_render_workspace_bytes(1024, 24, 0) == lib.mnf_render_workspace_bytes(1024, 24, 1)       # S=24: fp32 kernel


def test_integration_doc_covers_every_compute_entry_point():
    """INTEGRATION.md must say, for every compute entry point of the header, which reference code it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    compute = [n for n in declared_functions() if n.endswith("_fwd") or n in ("mnf_pack_features", "mnf_pack_images", "mnf_decoder_load_host")]
    assert len(compute) >= 10
    missing = [n for n in compute if f"`{n}`" not in doc]
    assert not missing, f"INTEGRATION.md lacks {missing}"
