"""The C-ABI library loads and exports every symbol include/fsvc.h declares (no compute: CPU only)."""
import ctypes
import os
import re

import pytest

from conftest import REPO
from svcc23_fastsvc_b200 import abi, build


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_symbols_are_exported(lib_path):
    header = open(os.path.join(REPO, "include", "fsvc.h")).read()
    declared = set(re.findall(r"\b(fsvc_[a-z_0-9]+)\s*\(", header))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_abi_version_and_error_paths(lib_path):
    lib = abi.load()
    assert lib.fsvc_abi_version() == 1
    # argument validation happens before any device work
    h = ctypes.c_void_p()
    rc = lib.fsvc_create(None, ctypes.byref(h))
    assert rc == -1 and b"null" in lib.fsvc_last_error()
    cfg = abi.FsvcConfig()
    cfg.num_stages = 99
    assert lib.fsvc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(abi.FsvcError, match="no CUDA device|CUDA"):
        abi.Handle(144, [192, 96, 48, 24], [2, 4, 4, 5], 1, 512, True)


def test_library_is_sm100a_only(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
