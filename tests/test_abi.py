"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/sb2gpu.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from starbeast2_b200 import _lib, build
from starbeast2_b200.engine import shard_loci


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    names = _lib.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sb2gpu.h but not exported"
        assert n in lib._sb2_signatures, f"{n} has no ctypes signature"


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert all("sm_100a" in line for line in out.splitlines() if "sm_" in line)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.Sb2Config(device=0)
    h = C.c_void_p()
    rc = lib.sb2_create(C.byref(cfg), C.byref(h))
    assert rc == -3  # SB2_ERR_CUDA
    assert b"no CPU fallback" in lib.sb2_last_error(h)
    lib.sb2_destroy(h)
    from starbeast2_b200.engine import Engine
    with pytest.raises(_lib.Sb2Error):
        Engine()


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "starbeast2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "sb2_oracle" not in src, f


def test_shard_loci_partitions():
    w = np.arange(1, 51, dtype=float)
    for ws in (1, 2, 3, 8):
        parts = shard_loci(w, ws)
        assert len(parts) == ws and parts[0][0] == 0 and parts[-1][1] == 50
        assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
        loads = [w[a:b].sum() for a, b in parts]
        assert max(loads) <= w.sum() / ws + w.max()
