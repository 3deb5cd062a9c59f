"""The C-ABI library must load without a GPU and export every symbol include/*.h declares."""
import ctypes
import glob
import os
import re
import subprocess

import pytest

import smoothmesh_b200 as sm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(sm(?:gpu|mesh)_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_headers_declare_something():
    names = declared_functions()
    assert "smgpu_create" in names and "smgpu_iterate" in names and "smmesh_read" in names
    assert len(names) > 40


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(sm.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", sm.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sm.SmoothMeshError, match="no CUDA device"):
        sm.Smoother(sm.Mesh.hex_block(2, 2, 2))


def test_product_does_not_import_the_oracle():
    # the oracle is test infrastructure: nothing under smoothmesh_b200/ may reference it
    for path in glob.glob(os.path.join(ROOT, "smoothmesh_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
            txt = open(path, errors="ignore").read()
            for needle in ("import oracle", "from oracle", "liboracle", "oracle/_build", '#include "../../oracle', "orc_create"):
                assert needle not in txt, (path, needle)
