"""TEST INFRASTRUCTURE: loads the g++/SIMT-emulator build of the kernel sources so the kernel arithmetic can be
compared with the oracle on a machine without a GPU.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C

from levels2fm_b200 import _C

from . import build as _build


class HostLib(_C.Lib):
    def _check_device(self, t):
        if t.is_cuda:
            raise RuntimeError("hostsim takes CPU tensors")

    def stream(self):
        return C.c_void_p(0)


_host = None


def get() -> HostLib:
    global _host
    if _host is None:
        _host = HostLib(_build.build())
    return _host
