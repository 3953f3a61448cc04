"""ORACLE -- test infrastructure only (see oracle/step_oracle.cpp header).

Python shim over oracle/_build/liboracle.so.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """g++ -O2 -ffp-contract=off -fopenmp -shared oracle/step_oracle.cpp (oracle/Makefile)."""
    args = ["make", "-C", _HERE] + (["-B"] if force else [])
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.orc_run.restype = C.c_int
    return _lib


def run(world, nt_begin, nt_end, time_elapsed=0.0, ranks=None, threads=None):
    """Advance every sub-domain of `world` (eqdyna_b200.host.World, all ranks built
    in-process) through steps nt_begin..nt_end of driver.f90's loop, in place.
    Returns the accumulated timeElapsed."""
    from eqdyna_b200.host import EqhView
    ranks = list(range(world.size)) if ranks is None else ranks
    arr = (EqhView * len(ranks))()
    for k, r in enumerate(ranks):
        arr[k] = world.raw_view(r)
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    t = C.c_double(time_elapsed)
    rc = lib().orc_run(arr, len(ranks), int(nt_begin), int(nt_end), C.byref(t))
    if rc:
        raise RuntimeError("oracle run failed with code %d" % rc)
    return t.value


class RankStepper:
    """One sub-domain stepped by this process; the caller moves the face buffers
    (tests/test_multiproc_gloo.py does it over torch.distributed/gloo)."""

    def __init__(self, world, rank):
        self.v = world.raw_view(rank)
        self.t = C.c_double(0.0)
        self.L = lib()
        self.L.orc_face_pack.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.L.orc_face_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        self.L.orc_step_pre.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        self.L.orc_step_post.argtypes = [C.c_void_p, C.c_int, C.c_double]

    def pre(self, nt):
        rc = self.L.orc_step_pre(C.byref(self.v), nt, C.byref(self.t))
        if rc:
            raise RuntimeError("orc_step_pre failed: %d" % rc)

    def pack(self, axis, side):
        import numpy as np
        n = self.L.orc_face_pack(C.byref(self.v), axis, side, None)
        buf = np.zeros(n)
        self.L.orc_face_pack(C.byref(self.v), axis, side, buf.ctypes.data)
        return buf

    def add(self, axis, side, buf):
        rc = self.L.orc_face_add(C.byref(self.v), axis, side, buf.ctypes.data, int(buf.size))
        if rc:
            raise RuntimeError("orc_face_add: face size mismatch")

    def post(self, nt):
        self.L.orc_step_post(C.byref(self.v), nt, self.t)
