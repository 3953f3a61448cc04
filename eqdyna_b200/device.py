"""ctypes binding of libeqdyna_b200.so (include/eqdyna_b200.h), the CUDA step
library.  `Domain` = one sub-domain on one GPU, fed from the globalvar arrays
of the host (eqdyna_b200.host.View), exactly as the Fortran host would feed it
through ISO_C_BINDING (eqdyna_b200/csrc/fortran/eqdyna_cuda_iface.f90).

There is no CPU path: if the library is missing or no GPU is visible every
entry point raises.
"""
import ctypes as C
import os

import numpy as np

from . import build
from .host import EqdParams

F_DISP, F_VEL, F_V1, F_FORCE, F_FRIC, F_FNFT, F_PSTRAIN, F_STRESS = 1, 2, 3, 4, 5, 6, 7, 8
F_ONFAULT_HIST, F_OFFFAULT_HIST, F_HYPO_LOG, F_GM, F_SRC_EVOL, F_TPHIST, F_MASS, F_FNMS, F_ARN = 9, 10, 11, 12, 13, 14, 15, 16, 17
F_ELEDET, F_ELESHP, F_SS, F_PHI = 18, 19, 20, 21
T_TOTAL, T_NODE, T_ELEM, T_ASSEMBLE, T_HALO, T_FAULT, T_ELEM_PML, T_ELEM_REGX, T_MARCH, T_MARCH_PML, T_NSLOTS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10

EXPORTS = [
    "eqd_create", "eqd_destroy", "eqd_last_error", "eqd_set_mesh", "eqd_set_elem_ops", "eqd_compute_elem_ops", "eqd_set_nodal",
    "eqd_set_fault", "eqd_set_halo", "eqd_set_stations", "eqd_get_unique_id", "eqd_set_comm", "eqd_set_host_comm", "eqd_sum_shared",
    "eqd_run", "eqd_run_group", "eqd_fetch", "eqd_get_counts", "eqd_get_timing", "eqd_set_option", "eqd_plan_check",
    "eqd_box_check", "eqd_get_box_counts", "eqd_plan_bank_model", "eqd_march_emulate", "eqd_get_march_counts", "eqd_march_pml_emulate", "eqd_get_halo_mode",
]

_lib = None
EQD_ERR_ARG = 4
# eqd_allgather_fn of include/eqdyna_b200.h: fn(ctx, send, bytes, recv) -> 0 on success
ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class StepError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("eqdyna_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libeqdyna_b200.so (built in-tree by eqdyna_b200.build).  Raises if absent."""
    global _lib
    if _lib is None:
        path = build.cuda_lib_path()
        if not os.path.exists(path):
            raise RuntimeError("libeqdyna_b200.so is not built (python -m eqdyna_b200.build cuda); there is no CPU fallback")
        L = C.CDLL(path)
        vp, i32, i64, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_char_p
        L.eqd_create.argtypes = [C.POINTER(EqdParams), C.c_int, C.POINTER(vp)]
        L.eqd_destroy.argtypes = [vp]
        L.eqd_last_error.argtypes = [vp, cp, C.c_int]
        L.eqd_set_mesh.argtypes = [vp, i32, i32, i32, i32] + [vp] * 7 + [i32]
        L.eqd_set_elem_ops.argtypes = [vp] + [vp] * 9
        L.eqd_compute_elem_ops.argtypes = [vp] + [vp] * 4
        L.eqd_set_nodal.argtypes = [vp] + [vp] * 6
        L.eqd_set_fault.argtypes = [vp, i32] + [vp] * 8
        L.eqd_set_halo.argtypes = [vp] + [vp] * 9
        L.eqd_set_stations.argtypes = [vp, vp, i32, vp, i32, vp, i32]
        L.eqd_get_unique_id.argtypes = [vp]
        L.eqd_set_comm.argtypes = [vp, vp, i32, i32]
        L.eqd_set_host_comm.argtypes = [vp, i32, i32, ALLGATHER_FN, vp]
        L.eqd_sum_shared.argtypes = [vp]
        L.eqd_run.argtypes = [vp, i32, i32]
        L.eqd_run_group.argtypes = [C.POINTER(vp), i32, i32, i32]
        L.eqd_fetch.argtypes = [vp, i32, vp, i64]
        L.eqd_get_counts.argtypes = [vp] + [C.POINTER(i64)] * 4
        L.eqd_get_timing.argtypes = [vp, C.POINTER(C.c_double)]
        L.eqd_get_box_counts.argtypes = [vp] + [C.POINTER(i64)] * 2
        L.eqd_get_march_counts.argtypes = [vp, C.POINTER(i64)]
        L.eqd_set_option.argtypes = [vp, cp, i32]
        L.eqd_plan_check.argtypes = [i32, i32, vp, vp, vp, vp]
        L.eqd_box_check.argtypes = [i32, i32] + [vp] * 8
        L.eqd_plan_bank_model.argtypes = [i32, i32, vp, vp, vp, i32, vp]
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)


def unique_id():
    buf = (C.c_char * 128)()
    rc = lib().eqd_get_unique_id(buf)
    if rc:
        raise StepError(rc, "eqd_get_unique_id failed (NCCL not loadable?)")
    return bytes(buf)


def plan_check(view):
    """Run the tile planner on a sub-domain's connectivity on the host (no GPU) and
    verify its invariants; returns per-class statistics."""
    r = view.raw
    st = np.zeros(24, dtype=np.int64)
    rc = lib().eqd_plan_check(r.Nn, r.Ne, _ptr(view.nodeElemIdRelation), _ptr(view.elemTypeArr),
                              _ptr(view.numOfDofPerNodeArr), _ptr(st))
    if rc:
        raise StepError(EQD_ERR_ARG, "tile planner invariant violated (eqd_tiles.cu:%d)" % rc)
    keys = ("tiles", "elements", "slots", "node_slots", "max_tile_nodes", "max_colours", "multi_colour_tiles", "grid")
    return {name: dict(zip(keys, (int(x) for x in st[8 * c:8 * c + 8]))) for c, name in enumerate(("reg", "regx", "pml"))}


def plan_bank_model(view, bank_order=0):
    """Modelled shared-memory wavefronts of the tile kernels' corner accesses on this sub-domain (host only):
    {class: (conflict-free, ascending element order, order chosen under bank_order)}."""
    r = view.raw
    out = np.zeros(9, dtype=np.int64)
    rc = lib().eqd_plan_bank_model(r.Nn, r.Ne, _ptr(view.nodeElemIdRelation), _ptr(view.elemTypeArr),
                                   _ptr(view.numOfDofPerNodeArr), int(bank_order), _ptr(out))
    if rc:
        raise StepError(EQD_ERR_ARG, "eqd_plan_bank_model: planner invariant violated (eqd_tiles.cu:%d)" % rc)
    return {name: tuple(int(x) for x in out[3 * c:3 * c + 3]) for c, name in enumerate(("reg", "regx", "pml"))}


def box_check(view):
    """Compare the closed-form operators of axis-aligned hexahedra (option "box", eqd_box.h)
    with the host's precomputed eleshp / phi / ss on every box element of a sub-domain
    (host only, no GPU).  Returns (box elements, [strain, force, hourglass] relative deviation)."""
    r = view.raw
    n = C.c_int64()
    dev = np.zeros(3)
    rc = lib().eqd_box_check(r.Nn, r.Ne, _ptr(view.meshCoor), _ptr(view.nodeElemIdRelation), _ptr(view.elemTypeArr),
                             _ptr(view.eleshp), _ptr(view.phi), _ptr(view.ss), C.byref(n), _ptr(dev))
    if rc:
        raise StepError(EQD_ERR_ARG, "eqd_box_check: bad input (eqd_tiles.cu:%d)" % rc)
    return n.value, dev


class Domain:
    """One sub-domain of a case on one GPU."""

    def __init__(self, view, device=0, compute_ops=False, options=None, comm=None, host_comm=None):
        """compute_ops: let the device compute the element operators and the lumped mass from the
        mesh (eqd_compute_elem_ops) instead of uploading the host's (eqd_set_elem_ops).
        comm = (id128, nranks, rank): eqd_set_comm right after eqd_create, so that the communicator
        starts up while the state is uploaded.
        host_comm = (nranks, rank, allgather): eqd_set_host_comm -- allgather(send: bytes-like, nranks) -> bytes-like
        of nranks * len(send) bytes in rank order is the host's own all-gather (MPI_Allgather / torch.distributed);
        the library then creates no NCCL communicator."""
        self.view = view
        self.compute_ops = compute_ops
        self._h = C.c_void_p()
        p = EqdParams.from_buffer_copy(view.params)
        rc = lib().eqd_create(C.byref(p), int(device), C.byref(self._h))
        if rc:
            raise StepError(rc, "eqd_create failed (no CUDA device? see stderr)")
        self._keep = []
        self._nt_done = 0
        for k, val in (options or {}).items():   # options that shape the upload (tile bricks)
            self.set_option(k, val)
        if comm is not None:
            self.set_comm(*comm)
        if host_comm is not None:
            self.set_host_comm(*host_comm)
        self._upload(view)

    def _check(self, rc):
        if rc:
            buf = C.create_string_buffer(1024)
            lib().eqd_last_error(self._h, buf, 1024)
            raise StepError(rc, buf.value.decode(errors="replace"))

    def _upload(self, v):
        L = lib()
        r = v.raw
        self._check(L.eqd_set_mesh(self._h, r.Nn, r.Ne, r.Neq, r.sizeEq, _ptr(v.meshCoor), _ptr(v.nodeElemIdRelation),
                                   _ptr(v.elemTypeArr), _ptr(v.numOfDofPerNodeArr), _ptr(v.eqNumStartIndexLoc),
                                   _ptr(v.eqNumIndexArr), _ptr(v.stressCompIndexArr), r.sizeStress))
        if self.compute_ops:
            self._check(L.eqd_compute_elem_ops(self._h, _ptr(v.mat), _ptr(v.eleporep), _ptr(v.stressArr), _ptr(v.pstrain)))
            self._check(L.eqd_set_nodal(self._h, None, None, _ptr(v.v1), _ptr(v.velArr), _ptr(v.dispArr), _ptr(v.nodalForceArr)))
        else:
            self._check(L.eqd_set_elem_ops(self._h, _ptr(v.eleshp), _ptr(v.eledet), _ptr(v.elemass), _ptr(v.mat), _ptr(v.ss),
                                           _ptr(v.phi), _ptr(v.eleporep), _ptr(v.stressArr), _ptr(v.pstrain)))
            self._check(L.eqd_set_nodal(self._h, _ptr(v.nodalMassArr), _ptr(v.fnms), _ptr(v.v1), _ptr(v.velArr),
                                        _ptr(v.dispArr), _ptr(v.nodalForceArr)))
        if int(np.sum(v.nftnd)) > 0:
            self._check(L.eqd_set_fault(self._h, r.nftmx, _ptr(v.nftnd), _ptr(v.nsmp), _ptr(v.un), _ptr(v.us), _ptr(v.ud),
                                        _ptr(v.arn), _ptr(v.fric), _ptr(v.fnft)))
        fl = [(_ptr(a) if a is not None and a.size else None) for a in v.fltface]
        self._check(L.eqd_set_halo(self._h, _ptr(v.numcount), _ptr(v.fltnum), _ptr(v.fltMPI), *fl))
        self._check(L.eqd_set_stations(self._h, _ptr(v.idhist) if r.nOff else None, r.nOff,
                                       _ptr(v.anonfs) if r.nOn else None, r.nOn,
                                       _ptr(v.surfaceNodeIdArr) if r.nSurf else None, r.nSurf))

    # -- multi-process plumbing
    def set_comm(self, id128, nranks, rank):
        self._check(lib().eqd_set_comm(self._h, C.c_char_p(id128), nranks, rank))

    def set_host_comm(self, nranks, rank, allgather):
        def fn(ctx, send, nbytes, recv):
            try:
                out = allgather((C.c_char * nbytes).from_address(send), nranks)
                C.memmove(recv, bytes(out) if not isinstance(out, (bytes, bytearray)) else out, nbytes * nranks)
                return 0
            except Exception:   # an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1
        self._host_ag = ALLGATHER_FN(fn)   # keep the trampoline alive as long as the handle
        self._check(lib().eqd_set_host_comm(self._h, int(nranks), int(rank), self._host_ag, None))

    def sum_shared(self):
        self._check(lib().eqd_sum_shared(self._h))

    def set_option(self, key, value):
        self._check(lib().eqd_set_option(self._h, key.encode(), int(value)))

    def run(self, nt_begin, nt_end):
        self._check(lib().eqd_run(self._h, int(nt_begin), int(nt_end)))
        self._nt_done = max(self._nt_done, int(nt_end))

    def counts(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        lib().eqd_get_counts(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return {"regular": a.value, "pml": b.value, "pairs": c.value, "launches": d.value}

    def box_counts(self):
        """Elements swept with closed-form box operators (option "box"), after the first run."""
        a, b = C.c_int64(), C.c_int64()
        lib().eqd_get_box_counts(self._h, C.byref(a), C.byref(b))
        return {"regular": a.value, "pml": b.value}

    def halo_mode(self):
        """0 = no rank neighbours, 1 = ncclSend / ncclRecv, 2 = peer memory (valid after the first run / sum_shared)."""
        lib().eqd_get_halo_mode.argtypes = [C.c_void_p]
        return int(lib().eqd_get_halo_mode(self._h))

    def march_counts(self):
        """Marching class (option "march"): elements in bundles, bundles, node slots, fused nodes, CTAs."""
        a = (C.c_int64 * 8)()
        lib().eqd_get_march_counts(self._h, a)
        return {"elements": a[0], "bundles": a[1], "node_slots": a[2], "fused_nodes": a[3], "grid": a[4],
                "pml_elements": a[5], "pml_bundles": a[6], "pml_node_slots": a[7]}

    def timing(self):
        t = (C.c_double * T_NSLOTS)()
        lib().eqd_get_timing(self._h, t)
        return {"total": t[T_TOTAL], "node": t[T_NODE], "elem": t[T_ELEM], "assemble": t[T_ASSEMBLE],
                "halo": t[T_HALO], "fault": t[T_FAULT], "elem_pml": t[T_ELEM_PML], "elem_regx": t[T_ELEM_REGX], "march": t[T_MARCH], "march_pml": t[T_MARCH_PML]}

    def fetch(self, which, shape, dtype=np.float64):
        out = np.zeros(shape, dtype=dtype, order="F")
        self._check(lib().eqd_fetch(self._h, which, _ptr(out), out.nbytes))
        return out

    def fetch_into_view(self):
        """Copy the post-run state back into the host arrays (what the Fortran
        host's output routines read, eqdyna3d.f90:75-79)."""
        v = self.view
        r = v.raw
        v.dispArr[...] = self.fetch(F_DISP, (3, r.Nn))
        v.velArr[...] = self.fetch(F_VEL, (3, r.Nn))
        v.v1[...] = self.fetch(F_V1, (r.Neq,))
        v.nodalForceArr[...] = self.fetch(F_FORCE, (r.Neq,))
        v.stressArr[...] = self.fetch(F_STRESS, (r.sizeStress,))
        if v.params.C_elastic == 0:
            v.pstrain[...] = self.fetch(F_PSTRAIN, (r.Ne,))
        if int(np.sum(v.nftnd)) > 0:
            v.fric[...] = self.fetch(F_FRIC, (100, r.nftmx, r.ntotft))
            v.fnft[...] = self.fetch(F_FNFT, (r.nftmx, r.ntotft))
            v.onFaultQuantHistSCECForm[...] = self.fetch(F_ONFAULT_HIST, (12, r.nstep, r.nOnAlloc))
            v.hypoLog[...] = self.fetch(F_HYPO_LOG, (13, r.nstep))
            if v.onFaultTPHist is not None and v.params.friclaw == 5:
                v.onFaultTPHist[...] = self.fetch(F_TPHIST, (2, r.nftmx, r.nstep, r.ntotft))
        if r.nOff:
            v.OffFaultStGramSCEC[...] = self.fetch(F_OFFFAULT_HIST, (6 * r.nOff + 1, r.nstep))
        if v.params.outputGroundMotion == 1 and self._nt_done > 0:
            # output_gm / output_src_evol samples: one per step with mod(nt,10) == 1 (driver.f90:30-33)
            k = (self._nt_done - 1) // 10 + 1
            if v.gmHist is not None and k <= v.nGmAlloc:
                v.gmHist[:, :, :k] = self.fetch(F_GM, (3, r.nSurf, k))
            if v.srcEvolHist is not None and k <= v.nGmAlloc:
                v.srcEvolHist[:, :k] = self.fetch(F_SRC_EVOL, (int(v.nftnd[0]), k))
            v.nGmSamples[0] = k

    def close(self):
        if self._h:
            lib().eqd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_group(domains, nt_begin, nt_end):
    """Lock-step run of all sub-domains of a decomposition from one process."""
    arr = (C.c_void_p * len(domains))(*[d._h for d in domains])
    rc = lib().eqd_run_group(arr, len(domains), int(nt_begin), int(nt_end))
    if rc:
        domains[0]._check(rc)
    for d in domains:
        d._nt_done = max(d._nt_done, int(nt_end))
