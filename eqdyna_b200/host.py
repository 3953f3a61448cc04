"""ctypes binding of libeqdyna_host.so (include/eqdyna_host.h): the stand-in for
EQdyna's Fortran host.  `World` = the sub-domains of one case; `World.view(r)`
returns the globalvar arrays of rank r as numpy views over host memory."""
import ctypes as C
import os

import numpy as np

from . import build

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class EqdParams(C.Structure):
    """struct eqd_params of include/eqdyna_b200.h (field order is the ABI)."""
    _fields_ = [
        ("dt", C.c_double), ("nstep", C.c_int32), ("me", C.c_int32),
        ("npx", C.c_int32), ("npy", C.c_int32), ("npz", C.c_int32),
        ("rdampk", C.c_double), ("rdampm", C.c_double), ("w", C.c_double),
        ("grav", C.c_double), ("roumax", C.c_double), ("rhow", C.c_double), ("gamar", C.c_double),
        ("ccosphi", C.c_double), ("sinphi", C.c_double), ("tv", C.c_double),
        ("kapa_hg", C.c_double), ("dx", C.c_double),
        ("C_elastic", C.c_int32), ("C_Q", C.c_int32), ("C_hg", C.c_int32),
        ("PMLb", C.c_double * 8), ("nPML", C.c_int32), ("R", C.c_double), ("vmaxPML", C.c_double),
        ("friclaw", C.c_int32), ("C_nuclea", C.c_int32), ("nucfault", C.c_int32), ("TPV", C.c_int32),
        ("insertFaultType", C.c_int32), ("ntotft", C.c_int32),
        ("nucR", C.c_double), ("nucT", C.c_double), ("nucRuptVel", C.c_double), ("nucdtau0", C.c_double),
        ("xsource", C.c_double), ("ysource", C.c_double), ("zsource", C.c_double),
        ("slipRateThres", C.c_double), ("tol", C.c_double), ("fric_tp_h", C.c_double),
        ("outputGroundMotion", C.c_int32), ("reserved_i", C.c_int32 * 7), ("reserved_d", C.c_double * 8),
    ]


class EqhView(C.Structure):
    """struct eqh_view of include/eqdyna_host.h."""
    _fields_ = [
        ("params", EqdParams),
        ("Nn", C.c_int32), ("Ne", C.c_int32), ("Neq", C.c_int32), ("sizeEq", C.c_int32), ("sizeStress", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("nftmx", C.c_int32), ("ntotft", C.c_int32), ("nOn", C.c_int32), ("nOnAlloc", C.c_int32),
        ("nOff", C.c_int32), ("nSurf", C.c_int32), ("nstep", C.c_int32),
        ("stressUsed", C.c_int32), ("pad_", C.c_int32),
        ("meshCoor", c_f64p), ("nodeElemIdRelation", c_i32p), ("elemTypeArr", c_i32p),
        ("numOfDofPerNodeArr", c_i32p), ("eqNumStartIndexLoc", c_i32p), ("eqNumIndexArr", c_i32p),
        ("stressCompIndexArr", c_i32p),
        ("eleshp", c_f64p), ("eledet", c_f64p), ("elemass", c_f64p), ("mat", c_f64p), ("ss", c_f64p),
        ("phi", c_f64p), ("eleporep", c_f64p), ("stressArr", c_f64p), ("pstrain", c_f64p),
        ("nodalMassArr", c_f64p), ("fnms", c_f64p), ("v1", c_f64p), ("velArr", c_f64p), ("dispArr", c_f64p),
        ("nodalForceArr", c_f64p),
        ("nftnd", c_i32p), ("nsmp", c_i32p), ("un", c_f64p), ("us", c_f64p), ("ud", c_f64p), ("arn", c_f64p),
        ("fric", c_f64p), ("fnft", c_f64p),
        ("numcount", c_i32p), ("fltnum", c_i32p), ("fltMPI", c_i32p), ("fltface", c_i32p * 6),
        ("idhist", c_i32p), ("anonfs", c_i32p), ("surfaceNodeIdArr", c_i32p),
        ("onFaultQuantHistSCECForm", c_f64p), ("OffFaultStGramSCEC", c_f64p), ("hypoLog", c_f64p),
        ("onFaultTPHist", c_f64p),
        ("gmHist", c_f64p), ("srcEvolHist", c_f64p), ("nGmSamples", c_i32p), ("nGmAlloc", C.c_int32), ("pad2_", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build.host_lib_path()
        if not os.path.exists(path):
            build.build_host()
        L = C.CDLL(path)
        L.eqh_last_error.restype = C.c_char_p
        L.eqh_world_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.eqh_world_destroy.argtypes = [C.c_void_p]
        L.eqh_world_set_switch.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.eqh_world_size.argtypes = [C.c_void_p]
        L.eqh_world_build.argtypes = [C.c_void_p, C.c_int]
        L.eqh_world_sum_shared.argtypes = [C.c_void_p]
        L.eqh_get_view.argtypes = [C.c_void_p, C.c_int, C.POINTER(EqhView)]
        L.eqh_write_outputs.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.eqh_release_operators.argtypes = [C.c_void_p, C.c_int]
        L.eqh_set_comp_time.argtypes = [C.c_void_p, C.c_int, c_f64p]
        _lib = L
    return _lib


def _np(ptr, shape, dtype):
    """numpy view (Fortran order) over a host pointer; None for NULL."""
    if not ptr:
        return None
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype, order="F")
    a = np.ctypeslib.as_array(ptr, shape=(n,))
    return a.reshape(shape, order="F")


class View:
    """numpy views of one sub-domain's globalvar arrays (Fortran shapes)."""

    def __init__(self, raw):
        self.raw = raw
        v = raw
        self.params = v.params
        Nn, Ne, Neq = v.Nn, v.Ne, v.Neq
        self.Nn, self.Ne, self.Neq = Nn, Ne, Neq
        self.nftmx, self.ntotft, self.nstep = v.nftmx, v.ntotft, v.nstep
        self.nOn, self.nOff = v.nOn, v.nOff
        f64, i32 = np.float64, np.int32
        self.meshCoor = _np(v.meshCoor, (3, Nn), f64)
        self.nodeElemIdRelation = _np(v.nodeElemIdRelation, (8, Ne), i32)
        self.elemTypeArr = _np(v.elemTypeArr, (Ne,), i32)
        self.numOfDofPerNodeArr = _np(v.numOfDofPerNodeArr, (Nn,), i32)
        self.eqNumStartIndexLoc = _np(v.eqNumStartIndexLoc, (Nn,), i32)
        self.eqNumIndexArr = _np(v.eqNumIndexArr, (v.sizeEq,), i32)
        self.stressCompIndexArr = _np(v.stressCompIndexArr, (Ne,), i32)
        self.eleshp = _np(v.eleshp, (3, 8, Ne), f64)
        self.eledet = _np(v.eledet, (Ne,), f64)
        self.elemass = _np(v.elemass, (24, Ne), f64)
        self.mat = _np(v.mat, (Ne, 5), f64)
        self.ss = _np(v.ss, (6, Ne), f64)
        self.phi = _np(v.phi, (8, 4, Ne), f64)
        self.eleporep = _np(v.eleporep, (Ne,), f64)
        self.stressArr = _np(v.stressArr, (v.sizeStress,), f64)
        self.pstrain = _np(v.pstrain, (Ne,), f64)
        self.nodalMassArr = _np(v.nodalMassArr, (Neq,), f64)
        self.fnms = _np(v.fnms, (Nn,), f64)
        self.v1 = _np(v.v1, (Neq,), f64)
        self.velArr = _np(v.velArr, (3, Nn), f64)
        self.dispArr = _np(v.dispArr, (3, Nn), f64)
        self.nodalForceArr = _np(v.nodalForceArr, (Neq,), f64)
        self.nftnd = _np(v.nftnd, (v.ntotft,), i32)
        self.nsmp = _np(v.nsmp, (2, v.nftmx, v.ntotft), i32)
        self.un = _np(v.un, (3, v.nftmx, v.ntotft), f64)
        self.us = _np(v.us, (3, v.nftmx, v.ntotft), f64)
        self.ud = _np(v.ud, (3, v.nftmx, v.ntotft), f64)
        self.arn = _np(v.arn, (v.nftmx, v.ntotft), f64)
        self.fric = _np(v.fric, (100, v.nftmx, v.ntotft), f64)
        self.fnft = _np(v.fnft, (v.nftmx, v.ntotft), f64)
        self.numcount = _np(v.numcount, (9,), i32)
        self.fltnum = _np(v.fltnum, (6,), i32)
        self.fltMPI = _np(v.fltMPI, (6,), i32)
        self.fltface = [_np(v.fltface[k], (int(self.fltnum[k]),), i32) for k in range(6)]
        self.idhist = _np(v.idhist, (3, 6 * v.nOff), i32)
        self.anonfs = _np(v.anonfs, (3, max(v.nOn, 1)), i32)
        self.surfaceNodeIdArr = _np(v.surfaceNodeIdArr, (v.nSurf,), i32)
        self.onFaultQuantHistSCECForm = _np(v.onFaultQuantHistSCECForm, (12, v.nstep, v.nOnAlloc), f64)
        self.OffFaultStGramSCEC = _np(v.OffFaultStGramSCEC, (6 * v.nOff + 1, v.nstep), f64)
        self.hypoLog = _np(v.hypoLog, (13, v.nstep), f64)
        self.onFaultTPHist = _np(v.onFaultTPHist, (2, v.nftmx, v.nstep, v.ntotft), f64)
        # samples of output_gm / output_src_evol (every step with mod(nt,10) == 1)
        self.nGmAlloc = v.nGmAlloc
        self.gmHist = _np(v.gmHist, (3, v.nSurf, v.nGmAlloc), f64) if v.nGmAlloc and v.nSurf else None
        self.srcEvolHist = _np(v.srcEvolHist, (int(self.nftnd[0]), v.nGmAlloc), f64) if v.nGmAlloc and int(self.nftnd[0]) else None
        self.nGmSamples = _np(v.nGmSamples, (1,), i32)


class World:
    """All (or some) sub-domains of a case directory, built by the stand-in host."""

    def __init__(self, case_dir, np_xyz=None, nstep=0):
        self._h = C.c_void_p()
        npx, npy, npz = np_xyz if np_xyz else (0, 0, 0)
        rc = lib().eqh_world_create(os.fsencode(case_dir), npx, npy, npz, int(nstep), C.byref(self._h))
        if rc:
            raise RuntimeError("eqh_world_create: " + lib().eqh_last_error().decode())
        self.size = lib().eqh_world_size(self._h)
        self.case_dir = case_dir

    def build(self, rank=-1, sum_shared=True):
        rc = lib().eqh_world_build(self._h, rank)
        if rc:
            raise RuntimeError("eqh_world_build: " + lib().eqh_last_error().decode())
        if rank < 0 and sum_shared:
            self.sum_shared()
        return self

    def set_switch(self, name, value):
        """Emulate another setting of a compile-time switch of globalvar.f90 (C_Q, C_hg, ...)."""
        rc = lib().eqh_world_set_switch(self._h, name.encode(), float(value))
        if rc:
            raise RuntimeError("eqh_world_set_switch: " + lib().eqh_last_error().decode())
        return self

    def sum_shared(self):
        rc = lib().eqh_world_sum_shared(self._h)
        if rc:
            raise RuntimeError("eqh_world_sum_shared: " + lib().eqh_last_error().decode())

    def raw_view(self, rank):
        v = EqhView()
        rc = lib().eqh_get_view(self._h, rank, C.byref(v))
        if rc:
            raise RuntimeError("eqh_get_view: " + lib().eqh_last_error().decode())
        return v

    def view(self, rank):
        return View(self.raw_view(rank))

    def write_outputs(self, rank, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        rc = lib().eqh_write_outputs(self._h, rank, os.fsencode(out_dir))
        if rc:
            raise RuntimeError("eqh_write_outputs: " + lib().eqh_last_error().decode())

    def set_comp_time(self, rank, t10):
        """compTimeInSeconds(1:9) + MPICommTimeInSeconds for compTime<me> (library_output.f90:208-218)."""
        a = (C.c_double * 10)(*[float(x) for x in t10])
        rc = lib().eqh_set_comp_time(self._h, rank, a)
        if rc:
            raise RuntimeError("eqh_set_comp_time: " + lib().eqh_last_error().decode())

    def release_operators(self, rank):
        lib().eqh_release_operators(self._h, rank)

    def close(self):
        if self._h:
            lib().eqh_world_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
