"""Loads the SIMT-emulation build of the kernel sources for CPU-side kernel-logic tests (never the product)."""
import os
import subprocess

import ndrustfft_b200 as nb
from ndrustfft_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libndfft_b200_emu.so")
_be = None


def emu_backend():
    global _be
    if _be is None:
        srcs = [os.path.join(ROOT, "ndrustfft_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "ndrustfft_b200", "csrc"))]
        srcs.append(os.path.join(ROOT, "tests", "emu", "simt_emu.h"))
        if not os.path.exists(EMU) or any(os.path.getmtime(s) > os.path.getmtime(EMU) for s in srcs):
            subprocess.check_call(["make", "-C", ROOT, "emu"], stdout=subprocess.DEVNULL)
        _be = nb.Backend(_lib.CLib(EMU))
        assert "emu" in _be.lib.version()
    return _be
