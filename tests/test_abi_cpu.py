"""CPU-side checks of the drop-in boundary: libtrack2d.so builds, loads, and exports every symbol that
include/track2d.h declares; the gym id grammar; loud failure without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from active_tracking_rl_b200 import _lib, build as t2d_build


def header_symbols():
    text = open(os.path.join(ROOT, "include", "track2d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(track2d_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = t2d_build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libtrack2d.so does not export %s" % n
    # the python binding covers the same set (no drift between header and binding)
    assert sorted(_lib.SYMBOLS) == names


def test_binding_loads_and_reports_abi():
    lib = _lib.load()
    assert lib.track2d_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.track2d_last_error(), bytes)


def test_env_id_grammar_matches_reference_registry():
    ids = _lib.all_env_ids()
    assert len(ids) == 72 and len(set(ids)) == 72  # 3 maps x 2 obs x 6 targets x 2 levels
    assert _lib.parse_env_id("Track2D-BlockPartialPZR-v0") == ("Block", "Partial", "PZR", 0)
    assert _lib.parse_env_id("Track2D-MazeFullRPF-v1") == ("Maze", "Full", "RPF", 1)
    for bad in ("Track2D-BlockPartialPZR-v2", "Track2D-CavePartialPZR-v0", "Track3D-BlockPartialPZR-v0", "Track2D-BlockPartialXYZ-v0"):
        with pytest.raises(KeyError):
            _lib.parse_env_id(bad)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from active_tracking_rl_b200.envs import Track2DVecEnv
    with pytest.raises(_lib.Track2DError):
        Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=4)


def test_create_rejects_bad_config_without_touching_a_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p(0)
    cfg = _lib.Config(_lib.ABI_VERSION + 7, 4, 0, 0, 1, 0, 0, 0, 500, 0, 1)
    assert lib.track2d_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"ABI" in lib.track2d_last_error()
    cfg = _lib.Config(_lib.ABI_VERSION, 0, 0, 0, 1, 0, 0, 0, 500, 0, 1)
    assert lib.track2d_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    cfg = _lib.Config(_lib.ABI_VERSION, 4, 0, 7, 1, 0, 0, 0, 500, 0, 1)
    assert lib.track2d_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"obs_type" in lib.track2d_last_error()


def test_gemm_plan_and_argument_checks_without_a_gpu():
    lib = _lib.load()
    # the launch plan is a pure function of the shape: many super-tiles -> no split-K workspace; weight gradients
    # (few output tiles, reduction over the env axis) -> split-K partials of M * N floats each
    assert lib.track2d_gemm_workspace_floats(65536, 256, 1024) == 0
    ws = lib.track2d_gemm_workspace_floats(256, 1024, 65536)
    assert ws > 0 and ws % (256 * 1024) == 0 and ws // (256 * 1024) <= 148
    assert lib.track2d_gemm_workspace_floats(0, 256, 16) == 0
    null = ctypes.c_void_p(0)
    assert lib.track2d_gemm_tf32x3(null, 0, 16, null, 0, 16, null, 16, 16, 16, 16, null, 0, null, 0, null) == -1
    assert b"track2d_gemm_tf32x3" in lib.track2d_last_error()
    # extents that are not multiples of 4 are refused before any launch (pointers are never dereferenced on the host)
    fake = ctypes.c_void_p(4096)
    assert lib.track2d_gemm_tf32x3(fake, 0, 8, fake, 0, 8, fake, 8, 8, 8, 6, null, 0, null, 0, null) == -1
    assert b"multiples of 4" in lib.track2d_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "active_tracking_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), "%s mentions the oracle" % f

