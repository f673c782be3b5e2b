"""java/jni/sb2gpu_jni.c through a real C compiler and -- on the GPU box -- through a mock JNIEnv.

The image has no JDK, so the shim cannot be loaded by a JVM here.  What CAN be done, and is done:
  * compile it with gcc -Wall -Wextra -Werror against a stand-in <jni.h> that declares exactly the JNI surface it uses
    (tests/jni_stub/jni.h) and against include/sb2gpu.h -- every call into the C ABI is type-checked;
  * link it with libsb2gpu.so into a shared object and check that every `native` method of Sb2Gpu.java is exported
    under its JNI-mangled name;
  * EXECUTE it: tests/jni_stub/mock_jvm.c is a mock JNIEnv ("Java arrays" wrap numpy memory, pin / unpin pairs are
    counted, ThrowNew is recorded); a posterior evaluation driven entirely through the Java_starbeast2_gpu_Sb2Gpu_*
    entry points must agree with the CPU oracle, and every pinned array must have been released.
What remains untested is stated in INTEGRATION.md: the JVM's own array pinning, System.loadLibrary, and the Java classes.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "jni_stub")
SHIM = os.path.join(ROOT, "java", "jni", "sb2gpu_jni.c")
OUT = os.path.join(STUB, "_build")
PREFIX = "Java_starbeast2_gpu_Sb2Gpu_"
CFLAGS = ["-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fPIC",
          "-I" + STUB, "-I" + os.path.join(ROOT, "include")]


def _natives():
    java = open(os.path.join(ROOT, "java", "src", "starbeast2", "gpu", "Sb2Gpu.java")).read()
    return sorted(set(re.findall(r"native\s+[\w\[\]]+\s+(\w+)\s*\(", java)))


def _build_mock():
    from starbeast2_b200 import build as b
    lib = b.build()
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libsb2gpu_jni_mock.so")
    srcs = [SHIM, os.path.join(STUB, "mock_jvm.c")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs + [lib, os.path.join(STUB, "jni.h")]):
        subprocess.check_call(["gcc", "-shared"] + CFLAGS + srcs + ["-L" + os.path.dirname(lib), "-lsb2gpu",
                                                                    "-Wl,-rpath," + os.path.dirname(lib), "-o", so])
    return so


def test_shim_compiles_warning_free():
    subprocess.check_call(["gcc", "-fsyntax-only"] + CFLAGS + [SHIM])


def test_shim_links_and_exports_every_native_method():
    J = C.CDLL(_build_mock())
    for name in _natives():
        assert hasattr(J, PREFIX + name), name


def _mock():
    J = C.CDLL(_build_mock())
    J.mock_env.restype = C.c_void_p
    J.mock_array.restype = C.c_void_p
    J.mock_array.argtypes = [C.c_void_p]
    J.mock_free.argtypes = [C.c_void_p]
    for f in ("mock_exception_class", "mock_exception_message"):
        getattr(J, f).restype = C.c_char_p
    J.mock_string.restype = C.c_char_p
    J.mock_string.argtypes = [C.c_void_p]
    return J


def test_create_without_a_device_throws_into_the_jvm():
    """No CPU fallback: on a box without an sm_100 GPU `create` returns 0 and leaves an IllegalStateException pending."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    J = _mock()
    f = getattr(J, PREFIX + "create")
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint8]
    assert f(J.mock_env(), None, 0, 0) == 0
    assert J.mock_exception_class() == b"java/lang/IllegalStateException"
    assert b"no CPU fallback" in J.mock_exception_message() or b"CUDA" in J.mock_exception_message()


@pytest.mark.gpu
def test_posterior_through_the_jni_entry_points_matches_the_oracle():
    import oracle
    from starbeast2_b200 import synth

    pb = synth.make_problem("C3", n_loci=5, n_sites=150, seed=11)
    want = oracle.posterior(pb, n_threads=2)
    J = _mock()
    env = J.mock_env()
    keep = []

    def arr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return C.c_void_p(J.mock_array(a.ctypes.data))

    def call(name, *args, restype=C.c_int32):
        f = getattr(J, PREFIX + name)
        f.restype = restype
        rc = f(C.c_void_p(env), None, *args)
        if restype is C.c_int32:
            assert rc >= 0, (name, rc)
        return rc

    h = C.c_int64(call("create", C.c_int32(0), C.c_uint8(0), restype=C.c_int64))
    assert h.value != 0, J.mock_exception_message()
    for loc in pb.loci:
        call("addLocus", h, C.c_int32(loc.tip_states.shape[0]), C.c_int32(loc.tip_states.shape[1]), arr(loc.tip_states, np.uint8),
             arr(loc.weights, np.int32), C.c_int32(loc.n_cat), arr(loc.tip_species, np.int32), C.c_double(loc.ploidy))
    sp = pb.species
    call("finalizeEngine", h, C.c_int32(sp.n_nodes), C.c_int32(sp.n_leaves))
    call("setSpeciesTree", h, arr(sp.parent, np.int32), arr(sp.left, np.int32), arr(sp.right, np.int32), arr(sp.height, np.float64))
    call("setSpeciesRates", h, arr(pb.species_rates, np.float64))
    call("setPopModel", h, C.c_int32(pb.pop_kind), arr(pb.pop_a, np.float64), arr(pb.pop_b, np.float64),
         C.c_double(pb.alpha), C.c_double(pb.beta))
    for i, loc in enumerate(pb.loci):
        t = loc.tree
        call("setGeneTree", h, C.c_int32(i), arr(t.parent, np.int32), arr(t.left, np.int32), arr(t.right, np.int32), arr(t.height, np.float64))
        p = np.zeros(40)
        p[: len(np.ravel(loc.subst_params))] = np.ravel(loc.subst_params)
        call("setSubstModel", h, C.c_int32(i), C.c_int32(loc.subst_kind), arr(p, np.float64))
        call("setSiteModel", h, C.c_int32(i), arr(loc.cat_rates, np.float64), arr(loc.cat_props, np.float64), C.c_double(loc.p_inv))
        call("setClock", h, C.c_int32(i), C.c_int32(loc.clock_kind), C.c_double(loc.clock_rate), arr(loc.explicit_rates, np.float64))
    L, S = pb.n_loci, sp.n_nodes
    coal, lik, br, tot = np.zeros(L), np.zeros(L), np.zeros(S), np.zeros(3)
    call("evalPosterior", h, arr(coal, np.float64), arr(lik, np.float64), arr(br, np.float64), arr(tot, np.float64))
    # `arr` made no copies (already contiguous float64): the shim wrote straight into these buffers
    assert np.allclose(lik, want.per_locus_lik, rtol=1e-9, atol=0)
    assert np.allclose(coal, want.per_locus_coal, rtol=1e-9, atol=0)
    assert abs(tot[2] - want.total) <= 1e-9 * abs(want.total)
    # the two-phase protocol of the reference's CompoundDistribution, then store / restore
    call("store", h)
    t1 = np.zeros(1)
    call("evalCoalescent", h, None, None, arr(t1, np.float64))
    assert abs(t1[0] - want.coalescent) <= 1e-9 * abs(want.coalescent)
    call("evalLikelihood", h, None, arr(t1, np.float64))
    assert abs(t1[0] - want.likelihood) <= 1e-9 * abs(want.likelihood)
    call("restore", h)
    assert call("numLoci", h) == L
    msg = getattr(J, PREFIX + "lastError")
    msg.restype = C.c_void_p
    assert J.mock_string(C.c_void_p(msg(C.c_void_p(env), None, h))) is not None
    call("destroy", h, restype=None)
    assert J.mock_pins() > 0 and J.mock_live_pins() == 0 and J.mock_pins() == J.mock_unpins()
