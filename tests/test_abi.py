"""The C-ABI library loads, exports every symbol include/rt1w.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rt1w.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt1w_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(rt):
    lib = rt.api.load_library()
    names = _declared_symbols()
    assert set(names) == set(rt.api.ABI_SYMBOLS), "api.ABI_SYMBOLS out of date with include/rt1w.h"
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/rt1w.h but not exported by librt1w.so"
    assert lib.rt1w_abi_version() == 2


def test_host_library_exports(rt):
    h = rt.api.load_host_library()
    text = open(os.path.join(ROOT, "raytracing-1w_b200", "host", "host_api.h")).read()
    for name in set(re.findall(r"\b(rt1w_host_[a-z0-9_]+)\s*\(", text)):
        assert hasattr(h, name)


def test_struct_layouts_match_the_header(rt):
    """ctypes mirrors vs the C compiler's view of include/rt1w.h (sizes probed through a tiny C program)."""
    import subprocess
    import tempfile

    api = rt.api
    src = r'''
#include <stdio.h>
#include "rt1w.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rt1w_node), sizeof(rt1w_material), sizeof(rt1w_texture),
         sizeof(rt1w_perlin), sizeof(rt1w_image), sizeof(rt1w_scene_desc), sizeof(rt1w_camera), sizeof(rt1w_render_params),
         sizeof(rt1w_render_stats), sizeof(rt1w_scene_info), sizeof(rt1w_flat_prim), sizeof(rt1w_ray));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    mirrors = [api.Node, api.Material, api.Texture, api.Perlin, api.Image, api.SceneDesc, api.Camera, api.RenderParams,
               api.RenderStats, api.SceneInfo, api.FlatPrim, api.Ray]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_cpu_fallback(rt):
    """Without a usable CUDA device the product refuses to run (there is no CPU path to fall back to)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rt.api.Rt1wError) as e:
        rt.api.Context(0)
    assert e.value.status == rt.api.ERR_NO_DEVICE


def test_product_does_not_link_the_oracle(rt):
    """The shipped library must not depend on the checker."""
    import subprocess

    out = subprocess.check_output(["ldd", rt.api.LIB_PATH], text=True)
    assert "oracle" not in out
    syms = subprocess.check_output(["nm", "-D", "--defined-only", rt.api.LIB_PATH], text=True)
    assert "oracle_" not in syms


def test_demo_driver_fails_loudly_without_a_gpu(rt):
    """rt1w_main (the reference's `main` on top of the C ABI) exists and has no CPU path either."""
    import subprocess

    import torch

    exe = os.path.join(os.path.dirname(rt.api.LIB_PATH), "rt1w_main")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "cornel_box", "--width", "8", "--spp", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "rt1w_context_create" in r.stderr
    assert not r.stdout.startswith("P3")


def test_library_shard_rule_matches_the_host_side(rt):
    """rt1w_shard_sample_range (what a communicator's ranks render) == shard.sample_range (host-side rule): a partition of
    the sample range into contiguous parts whose sizes differ by at most one."""
    import importlib

    shard = importlib.import_module("raytracing-1w_b200.shard")
    for world in (1, 2, 3, 4, 8):
        for begin, end in ((0, 1), (0, 7), (0, 100), (0, 4096), (300, 400), (5, 6)):
            ranges = [rt.api.shard_sample_range(r, world, begin, end) for r in range(world)]
            assert ranges[0][0] == begin and ranges[-1][1] == end
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0
            assert [(a - begin, b - begin) for a, b in ranges] == [shard.sample_range(r, world, end - begin) for r in range(world)]


def test_multi_gpu_entry_points_fail_loudly_without_a_gpu(rt):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rt.api.Rt1wError) as e:
        rt.api.Context([0, 1])
    assert e.value.status == rt.api.ERR_NO_DEVICE
    with pytest.raises(rt.api.Rt1wError) as e:
        rt.api.Context([])
    assert e.value.status == rt.api.ERR_INVALID
