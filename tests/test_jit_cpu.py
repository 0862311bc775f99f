"""Run-time schedule compilation (csrc/jit.h) without a GPU: the schedule planner and NVRTC -> sm_100a cubin."""
import ctypes
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ndrustfft_b200 import _lib
    return _lib.default_lib()


@pytest.mark.parametrize("dtype,rkind,n,cols", [(1, -1, 96, 0), (0, -1, 768, 1), (1, -1, 720, 1), (0, -1, 6561, 0), (1, 0, 1536, 1), (1, 3, 360, 0)])
def test_compile_to_sm100a_cubin(lib, dtype, rkind, n, cols, tmp_path):
    os.environ["NDFB_JIT_CACHE"] = str(tmp_path)
    try:
        buf = ctypes.create_string_buffer(800)
        rc = lib.dll.ndfb_jit_compile_check(dtype, rkind, n, cols, buf, 800)
        if rc != 0 and "libnvrtc not available" in lib.last_error():
            pytest.skip("no libnvrtc here")
        assert rc == 0, lib.last_error()
        info = json.loads(buf.value.decode())
        assert info["cubin_bytes"] > 1000 and str(n) in info["kernel"]
        assert info["E"] <= (16 if dtype == 1 else 24)
        # second call comes from the disk cache
        assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))
        assert lib.dll.ndfb_jit_compile_check(dtype, rkind, n, cols, buf, 800) == 0
    finally:
        del os.environ["NDFB_JIT_CACHE"]


def test_lengths_without_a_schedule_are_refused(lib):
    buf = ctypes.create_string_buffer(200)
    assert lib.dll.ndfb_jit_compile_check(1, -1, 1009, 0, buf, 200) != 0       # prime: Bluestein, not a radix schedule
    assert lib.dll.ndfb_jit_compile_check(1, -1, 17 * 64, 0, buf, 200) != 0    # 17 is not a butterfly
    assert lib.dll.ndfb_jit_compile_check(1, -1, 7 ** 5, 0, buf, 200) != 0     # needs five passes
