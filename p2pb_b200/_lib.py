"""ctypes binding of ``libp2pb_b200.so`` (the C ABI declared in ``include/p2pb_b200.h``).

There is no CPU fallback: if the library is missing this raises, and every entry point raises when its
return code is non-zero (the reference ``exit(-1)``s on launch errors, ``cuda_utils.cuh:30-40``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# P2PB_LIB: development aid (A/B of two builds of the library on the same box, tools/gpu_ab.sh)
LIB_PATH = os.environ.get("P2PB_LIB") or os.path.join(_HERE, "libp2pb_b200.so")

_lib = None


class P2PBError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise P2PBError(
                f"{LIB_PATH} not found: build it with `python -m p2pb_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.p2pb_last_error.restype = ctypes.c_char_p
        if os.environ.get("P2PB_PDL") is not None:
            check(_lib.p2pb_set_pdl(int(os.environ["P2PB_PDL"])), "p2pb_set_pdl")
        kb = os.environ.get("P2PB_SMEM_KB")
        if kb:
            check(_lib.p2pb_set_smem_budget_kb(int(kb)), "p2pb_set_smem_budget_kb")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().p2pb_last_error().decode("utf-8", "replace")
        raise P2PBError(f"{what} failed (rc={rc}): {msg}")


def call(name: str, *args) -> None:
    fn = getattr(lib(), name)
    check(fn(*args), name)


def launch_count() -> int:
    """Kernels launched (or captured into a graph) through the library so far."""
    fn = lib().p2pb_launch_count
    fn.restype = ctypes.c_ulonglong
    return int(fn())
