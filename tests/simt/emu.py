"""Loader of the CPU-emulated build of the kernels (tests/simt/_build/libvcfdist_emu.so, see simt_emu.h).
TEST INFRASTRUCTURE: a debugging aid for kernel logic in the GPU-less container.  The product path
(vcfdist_b200.capi.Engine) never loads it."""
import ctypes as C
import os
import subprocess

from vcfdist_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_LIB = os.path.join(HERE, "_build", "libvcfdist_emu.so")


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


class EmuEngine(capi.Engine):
    def __init__(self, **env):
        build()
        os.environ.update({k: str(v) for k, v in env.items()})
        try:
            self.lib = capi.load_library(EMU_LIB)
            h = C.c_void_p()
            rc = self.lib.vd_create(0, 0, C.byref(h))
            if rc != 0:
                raise capi.VdError(rc, "vd_create (emulated)")
            self.h = h
            self.device = 0
        finally:
            for k in env:
                del os.environ[k]
