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


def test_java_jni_sources_stay_consistent_with_the_header():
    """java/ cannot be compiled here (no JDK), so at least keep it honest: every `native` method of Sb2Gpu.java has a JNI
    implementation and vice versa, every sb2_* function the shim calls is declared in include/sb2gpu.h, and every entry point
    a JVM host needs is bound (the unbound remainder is the process-per-GPU / torch / introspection surface)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    java = open(os.path.join(root, "java", "src", "starbeast2", "gpu", "Sb2Gpu.java")).read()
    jni = open(os.path.join(root, "java", "jni", "sb2gpu_jni.c")).read()
    natives = set(re.findall(r"native\s+[\w\[\]]+\s+(\w+)\s*\(", java))
    impls = set(re.findall(r"JNICALL FN\((\w+)\)", jni))
    assert natives == impls, (sorted(natives - impls), sorted(impls - natives))
    called = set(re.findall(r"\b(sb2_\w+)\s*\(", jni))
    declared = set(_lib.header_symbols())
    assert called <= declared, sorted(called - declared)
    not_for_a_jvm = {"sb2_bind_reduce_vector", "sb2_comm_connect", "sb2_comm_handle", "sb2_set_stream",          # torch / one process per GPU
                     "sb2_device_bytes", "sb2_get_matrices", "sb2_get_partials", "sb2_get_pattern_logl", "sb2_kernel_time",
                     "sb2_last_node_updates", "sb2_set_profile", "sb2_exchange_wait_us", "sb2_step_phase_times",                                                    # introspection for tests / bench
                     "sb2_compress_alignment", "sb2_compress_last_error"}                                             # BEAST's own Alignment compresses on the host
    assert declared - called == not_for_a_jvm, sorted((declared - called) ^ not_for_a_jvm)


def test_java_ffm_twin_binds_the_same_surface_with_matching_arity():
    """Sb2GpuFfm.java (Panama FFM, JDK >= 22; uncompiled here) must expose exactly the static methods of the JNI class
    Sb2Gpu.java, bind only symbols include/sb2gpu.h declares, and give every downcall as many arguments as the C prototype."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jni = open(os.path.join(root, "java", "src", "starbeast2", "gpu", "Sb2Gpu.java")).read()
    ffm = open(os.path.join(root, "java", "src", "starbeast2", "gpu", "Sb2GpuFfm.java")).read()
    hdr = open(os.path.join(root, "include", "sb2gpu.h")).read()
    natives = set(re.findall(r"native\s+[\w\[\]]+\s+(\w+)\s*\(", jni))
    publics = set(re.findall(r"public static [\w\[\]]+\s+(\w+)\s*\(", ffm))
    assert natives == publics, sorted(natives ^ publics)
    protos = {m.group(1): m.group(2) for m in re.finditer(r"SB2_API\s+[\w\s\*]+?\b(sb2_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S)}
    binds = re.findall(r'bind\("(sb2_\w+)",\s*(\w+)((?:,\s*\w+)*)\)', ffm)
    assert len(binds) >= 30
    for name, _ret, args in binds:
        assert name in protos, name
        n_c = 0 if protos[name].strip() in ("", "void") else protos[name].count(",") + 1
        n_j = len([a for a in args.split(",") if a.strip()])
        assert n_c == n_j, (name, n_c, n_j)
