"""Compiles the header-only C++ mirror (include/ndrustfft_b200.hpp) against the CUDA library and runs the
reference's unit tests transliterated to C++ (tests/cpp/mirror_test.cpp)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ndrustfft_b200", "lib")
EXE = os.path.join(ROOT, "build", "mirror_test")


def _build(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libndfft_b200.so")):
        subprocess.check_call(["make", "-C", ROOT, "lib"], stdout=subprocess.DEVNULL)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"),
                           "-L", LIBDIR, "-lndfft_b200", f"-Wl,-rpath,{LIBDIR}", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", EXE])
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")))
    txt = tmp_path / "goldens.txt"

    def line(name, arr):
        a = np.asarray(arr, dtype=float).ravel()
        return f"{name} {a.size} " + " ".join(repr(float(v)) for v in a) + "\n"

    with open(txt, "w") as f:
        f.write(line("test_matrix", g["test_matrix"]["values"]))
        f.write(line("fft_re", g["test_fft"]["solution_re"])); f.write(line("fft_im", g["test_fft"]["solution_im"]))
        f.write(line("r2c_re", g["test_fft_r2c"]["solution_re"])); f.write(line("r2c_im", g["test_fft_r2c"]["solution_im"]))
        for k in (1, 2, 3, 4):
            f.write(line(f"dct{k}", g[f"test_dct{k}"]["solution"]))
    return str(txt)


def test_cpp_mirror_compiles_and_reports_reference_errors(tmp_path):
    txt = _build(tmp_path)
    p = subprocess.run([EXE, txt, "errors"], capture_output=True, text=True)
    assert p.returncode == 0 and "MIRROR_OK" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_cpp_mirror_reference_unit_tests(tmp_path):
    txt = _build(tmp_path)
    p = subprocess.run([EXE, txt, "gpu"], capture_output=True, text=True)
    assert p.returncode == 0 and "MIRROR_OK" in p.stdout, p.stdout + p.stderr


def test_cpp_mirror_full_suite_against_the_emulation_build(tmp_path):
    """The same C ABI is exported by the CPU emulation build of the kernels (tests/emu): run the whole C++ suite,
    including the one-call fft2 / rfft2 compositions, without a GPU."""
    emu_dir = os.path.join(ROOT, "tests", "emu")
    if not os.path.exists(os.path.join(emu_dir, "libndfft_b200_emu.so")):
        subprocess.check_call(["make", "-C", ROOT, "emu"], stdout=subprocess.DEVNULL)
    txt = _build(tmp_path)
    exe = os.path.join(ROOT, "build", "mirror_test_emu")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"),
                           "-L", emu_dir, "-lndfft_b200_emu", f"-Wl,-rpath,{emu_dir}", "-o", exe])
    p = subprocess.run([exe, txt, "gpu"], capture_output=True, text=True)
    assert p.returncode == 0 and "MIRROR_OK" in p.stdout, p.stdout + p.stderr
