"""No-GPU checks of the drop-in boundary: the CUDA library loads, exports every symbol include/ndfft_b200.h
declares, plans build without a device, and compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ndrustfft_b200", "lib", "libndfft_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", ROOT, "lib"], stdout=subprocess.DEVNULL)
    from ndrustfft_b200 import _lib
    return _lib.CLib(LIB)


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ndfft_b200.h")).read()
    declared = sorted(set(re.findall(r"NDFB_API[^;(]*?\b(ndfb_\w+)\s*\(", hdr)))
    assert len(declared) >= 9
    for sym in declared:
        assert hasattr(lib.dll, sym), sym
    from ndrustfft_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared


def test_version_is_cuda_build(lib):
    assert "sm_100a" in lib.version() and "emu" not in lib.version()


def test_library_contains_sm100a_sass():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("n,kind,family", [(8192, 0, "direct"), (1009, 0, "bluestein"), (1 << 24, 0, "four-step"),
                                           (360, 0, "direct"), (512, 1, "direct"), (4096, 2, "direct")])
def test_plan_builds_without_gpu(lib, n, kind, family):
    import json
    p = ctypes.c_void_p()
    assert lib.dll.ndfb_plan_create(ctypes.byref(p), kind, 0, n, 0) == 0
    need = lib.dll.ndfb_plan_describe(p, None, 0)
    buf = ctypes.create_string_buffer(int(need))
    lib.dll.ndfb_plan_describe(p, buf, need)
    d = json.loads(buf.value.decode())
    assert d["n"] == n
    assert d["ops"][0]["family"] == family
    if family != "four-step":
        N, rad = d["ops"][0]["N"], d["ops"][0]["radix"]
        B = d["ops"][0]["M"] or N
        assert int(np.prod(rad)) == B
        assert all(r in (2, 3, 4, 5, 7, 8, 11, 13, 16) for r in rad)
    lib.dll.ndfb_plan_destroy(p)


def test_dct_plan_schedules(lib):
    import json
    p = ctypes.c_void_p()
    assert lib.dll.ndfb_plan_create(ctypes.byref(p), 2, 1, 4096, 0) == 0
    need = lib.dll.ndfb_plan_describe(p, None, 0)
    buf = ctypes.create_string_buffer(int(need))
    lib.dll.ndfb_plan_describe(p, buf, need)
    ops = json.loads(buf.value.decode())["ops"]
    assert [o["N"] for o in ops] == [4095, 2048, 2048, 2048]     # DCT-I needs the 3^2*5*7*13 schedule
    assert sorted(ops[0]["radix"]) == [3, 3, 5, 7, 13]
    lib.dll.ndfb_plan_destroy(p)


def test_no_cpu_fallback_without_gpu(lib):
    """Without a device the product path must fail loudly, never compute on the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import ndrustfft_b200 as nb
    be = nb.Backend(lib)
    x = np.ones((2, 8), complex)
    y = np.zeros((2, 8), complex)
    with pytest.raises(nb.NdfftError) as e:
        be.ndfft(x, y, be.FftHandler(8), 1)
    assert e.value.code == -6            # NDFB_E_CUDA
    assert not y.any()


def test_argument_errors_need_no_gpu(lib):
    import ndrustfft_b200 as nb
    be = nb.Backend(lib)
    with pytest.raises(AssertionError, match="Size mismatch in fft, got 5 expected 6"):
        be.ndfft(np.zeros((2, 5), complex), np.zeros((2, 5), complex), be.FftHandler(6), 1)
    with pytest.raises(IndexError):
        be.ndfft(np.zeros((2, 6), complex), np.zeros((2, 6), complex), be.FftHandler(6), 3)
    p = ctypes.c_void_p()
    assert lib.dll.ndfb_plan_create(ctypes.byref(p), 7, 0, 8, 0) == -1
    assert "kind" in lib.last_error()


def test_package_never_references_oracle_or_emulator():
    pkg = os.path.join(ROOT, "ndrustfft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "scipy" not in txt and "numpy.fft" not in txt and "np.fft" not in txt, f
                if f.endswith(".py"):
                    assert "libndfft_b200_emu" not in txt, f
