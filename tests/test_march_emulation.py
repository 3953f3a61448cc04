"""The marching kernel (eqdyna_b200/csrc/cuda/eqd_march.h) on the CPU: its body is written as
barrier-separated phases, and `eqd_march_emulate` (host code of libeqdyna_b200.so, no GPU) runs the
planner and then every phase over all thread ids, CTA after CTA, with the asynchronous copies done at
issue time.  What is checked here is therefore the source the GPU executes: bundle planning against
the reference's connectivity, the lattice indexing, the register carries along x, the transformed
(Walsh-Hadamard) form of strain / hourglass / nodal forces, the fused node update.
Reference: the same elements evaluated one by one from the host's stored operators eleshp / phi / ss
(calcElemKU.f90:44-189, hrglss.f90:20-54) and scatter-added (assembleGlobalKU.f90:28-35), then
driver.f90:29,102-104 for the nodes the bundles update themselves."""
import ctypes as C

import numpy as np
import pytest

import parity


def _emulate(v, vel, disp, stress6, mass, dt, update, grid):
    from eqdyna_b200 import device
    L = device.lib()
    L.eqd_march_emulate.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 8 + [C.c_double] * 3 + \
                                  [C.c_int32] + [C.c_void_p] * 4
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    fsum = np.zeros((3, v.Nn), order="F")
    fused = np.zeros(v.Nn, dtype=np.int32)
    inb = np.zeros(v.Ne, dtype=np.int32)
    st = np.zeros(8, dtype=np.int64)
    rc = L.eqd_march_emulate(v.Nn, v.Ne, P(v.meshCoor), P(v.nodeElemIdRelation), P(v.elemTypeArr), P(v.numOfDofPerNodeArr), grid,
                             P(v.eleshp), P(v.ss), P(v.eledet), P(v.mat), P(stress6), P(vel), P(disp), P(mass),
                             dt, v.params.rdampk, v.params.w, update, P(fsum), P(fused), P(inb), P(st))
    assert rc == 0, "eqd_march_emulate failed at eqd_march.cu:%d" % rc
    return fsum, fused.astype(bool), inb.astype(bool), st


def _reference(v, E, vel, disp, stress6, dt):
    """element-by-element forces and stresses of elements E from the stored operators"""
    conn = v.nodeElemIdRelation[:, E] - 1                         # (8,nE)
    shp = v.eleshp[:, :, E]                                        # (3,8,nE)
    phi = v.phi[:, :, E]                                           # (8,4,nE)
    ss = v.ss[:, E]
    lam, mu, det = v.mat[E, 3], v.mat[E, 4], v.eledet[E]
    rk, w = v.params.rdampk, v.params.w
    ve = vel[:, conn]                                              # (3,8,nE)
    le = disp[:, conn] + rk * ve
    sr = np.zeros((6, len(E)))
    for i in range(8):
        s1, s2, s3 = shp[0, i], shp[1, i], shp[2, i]
        vx, vy, vz = ve[0, i], ve[1, i], ve[2, i]
        sr[0] += s1 * vx; sr[1] += s2 * vy; sr[2] += s3 * vz
        sr[3] += s3 * vy + s2 * vz; sr[4] += s3 * vx + s1 * vz; sr[5] += s2 * vx + s1 * vy
    l2m = lam + 2 * mu
    rate = np.stack([l2m * sr[0] + lam * sr[1] + lam * sr[2], lam * sr[0] + l2m * sr[1] + lam * sr[2],
                     lam * sr[0] + lam * sr[1] + l2m * sr[2], mu * sr[3], mu * sr[4], mu * sr[5]])
    sg = stress6[:, E] + rate * dt
    t = (-det * w) * (sg + rk * rate)
    phid = np.einsum("ime,cie->mce", phi, le)                      # (4,3,nE)
    S = np.array([[0, 1, 2], [1, 3, 4], [2, 4, 5]])
    hv = np.stack([np.stack([sum(ss[S[c, k]] * phid[m, k] for k in range(3)) for c in range(3)]) for m in range(4)])   # (4,3,nE)
    f = np.zeros((3, vel.shape[1]))
    for i in range(8):
        s1, s2, s3 = shp[0, i], shp[1, i], shp[2, i]
        fi = np.stack([s1 * t[0] + s3 * t[4] + s2 * t[5], s2 * t[1] + s3 * t[3] + s1 * t[5], s3 * t[2] + s2 * t[3] + s1 * t[4]])
        fi -= np.einsum("me,mce->ce", phi[i], hv)
        for c in range(3):
            np.add.at(f[c], conn[i], fi[c])
    return f, sg


SHARE = 3 << 16   # flags in the grid argument of eqd_march_emulate: neighbouring strips share ghost columns (bit 16) / rows (bit 17)
SHARE_Y = 1 << 16


@pytest.mark.parametrize("case,grid,min_cover,min_fused", [("test.tpv104", 444, 0.95, 0.3), ("test.tpv8", 37, 0.95, 0.3), ("test.tpv104", 7, 0.95, 0.3),
                                                           ("test.tpv104", 444 | SHARE, 0.95, 0.75), ("test.tpv8", 37 | SHARE, 0.95, 0.7),
                                                           ("test.tpv36", 100 | SHARE, 0.6, 0.5), ("test.tpv104", 444 | SHARE_Y, 0.95, 0.5)],
                         ids=["tpv104-444ctas", "tpv8-37ctas", "tpv104-7ctas", "tpv104-444ctas-ghosts", "tpv8-37ctas-ghosts", "tpv36-dipping-fault-ghosts",
                              "tpv104-444ctas-ghost-columns-only"])
def test_marching_kernel_phases_on_the_host(case, grid, min_cover, min_fused):
    w = parity.build_world(case, (1, 1, 1), 2)
    v = w.view(0)
    rng = np.random.default_rng(5)
    vel0 = np.asfortranarray(rng.standard_normal((3, v.Nn)))
    disp0 = np.asfortranarray(rng.standard_normal((3, v.Nn)) * 1e-2)
    sidx = v.stressCompIndexArr
    stress0 = np.asfortranarray(rng.standard_normal((6, v.Ne)) * 1e6)
    mass = rng.uniform(1.0e9, 2.0e9, v.Nn)
    dt = 0.004
    reg = np.array([(v.elemTypeArr[e] != 2) for e in range(v.Ne)])
    for update in (1, 0):
        vel, disp, stress = vel0.copy(order="F"), disp0.copy(order="F"), stress0.copy(order="F")
        fsum, fused, inb, st = _emulate(v, vel, disp, stress, mass, dt, update, grid)
        fused_set = fused.copy()
        E = np.nonzero(inb)[0]
        assert st[0] == len(E) and st[5] == grid & 0xffff
        # coverage: nearly every regular element on 3-dof nodes of these rectilinear meshes marches
        n3 = v.numOfDofPerNodeArr[v.nodeElemIdRelation - 1].max(axis=0) == 3
        cand = int((reg & n3).sum())
        assert len(E) >= min_cover * cand, (len(E), cand)
        assert st[0] + st[4] == cand
        fref, sref = _reference(v, E, vel0, disp0, stress0, dt)
        assert np.abs(fref).max() > 0
        # (with update = 1 the force of a fused node never leaves the CTA: only its new v, d are visible)
        seen = ~fused if update else np.ones(v.Nn, bool)
        assert np.abs(fsum[:, seen] - fref[:, seen]).max() <= 1e-12 * np.abs(fref).max()
        assert np.abs(stress[:, E] - sref).max() <= 1e-13 * np.abs(sref).max()
        out = np.ones(v.Ne, bool); out[E] = False
        assert np.array_equal(stress[:, out], stress0[:, out])
        # fused nodes: all eight elements around them are bundle elements, no other element touches them
        val = np.bincount((v.nodeElemIdRelation - 1).ravel(), minlength=v.Nn)
        valb = np.bincount((v.nodeElemIdRelation[:, E] - 1).ravel(), minlength=v.Nn)
        assert fused.sum() == st[3] > min_fused * v.Nn * len(E) / v.Ne, (fused.sum(), st[3], v.Nn * len(E) / v.Ne)
        assert np.all(val[fused] == 8) and np.all(valb[fused] == 8)
        assert np.all(v.numOfDofPerNodeArr[fused] == 3)
        if update:
            vn = vel0[:, fused] + (fref[:, fused] / mass[fused]) * dt
            dn = disp0[:, fused] + vn * dt
            assert np.abs(vel[:, fused] - vn).max() <= 1e-12 * np.abs(vn).max()
            assert np.abs(disp[:, fused] - dn).max() <= 1e-12 * np.abs(dn).max()
            assert np.array_equal(vel[:, ~fused], vel0[:, ~fused]) and np.array_equal(disp[:, ~fused], disp0[:, ~fused])
        else:
            assert np.array_equal(vel, vel0) and np.array_equal(disp, disp0)
    w.close()


def _emulate_pml(v, vel, disp, stress21, damps, dt, grid):
    from eqdyna_b200 import device
    L = device.lib()
    L.eqd_march_pml_emulate.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 8 + [C.c_double] * 3 + \
                                      [C.c_void_p] * 3
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    f12 = np.zeros((12, v.Nn), order="F")
    inb = np.zeros(v.Ne, dtype=np.int32)
    st = np.zeros(8, dtype=np.int64)
    rc = L.eqd_march_pml_emulate(v.Nn, v.Ne, P(v.meshCoor), P(v.nodeElemIdRelation), P(v.elemTypeArr), P(v.numOfDofPerNodeArr), grid,
                                 P(v.eleshp), P(v.ss), P(v.eledet), P(v.mat), P(damps), P(stress21), P(vel), P(disp),
                                 dt, v.params.rdampk, v.params.w, P(f12), P(inb), P(st))
    assert rc == 0, "eqd_march_pml_emulate failed at eqd_march.cu:%d" % rc
    return f12, inb.astype(bool), st


def _reference_pml(v, E, vel, disp, stress21, damps, dt):
    """calcPMLElemKU (assembleGlobalKU.f90:215-344) + hrglss.f90:20-54 of elements E, from the stored operators"""
    conn = v.nodeElemIdRelation[:, E] - 1
    shp = v.eleshp[:, :, E]
    phi = v.phi[:, :, E]
    ss = v.ss[:, E]
    lam, mu, det = v.mat[E, 3], v.mat[E, 4], v.eledet[E]
    rk, w = v.params.rdampk, v.params.w
    ve = vel[:, conn]                                              # (3,8,nE)
    le = disp[:, conn] + rk * ve
    D = lambda d, c: np.einsum("ie,ie->e", shp[d], ve[c])        # noqa: E731  D_d v_c
    l2m = lam + 2 * mu
    s = stress21[:, E].copy()
    rdt = 1.0 / dt
    dm = damps[:, E]

    def upd(k, cf, Dv, a):
        s[k] = (cf * Dv + (rdt - dm[a] / 2) * s[k]) / (rdt + dm[a] / 2)
    Dxx, Dyy, Dzz = D(0, 0), D(1, 1), D(2, 2)
    upd(0, l2m, Dxx, 0); upd(1, lam, Dyy, 1); upd(2, lam, Dzz, 2)
    upd(3, lam, Dxx, 0); upd(4, l2m, Dyy, 1); upd(5, lam, Dzz, 2)
    upd(6, lam, Dxx, 0); upd(7, lam, Dyy, 1); upd(8, l2m, Dzz, 2)
    upd(9, mu, D(0, 1), 0); upd(10, mu, D(1, 0), 1)
    upd(11, mu, D(0, 2), 0); upd(12, mu, D(2, 0), 2)
    upd(13, mu, D(1, 2), 1); upd(14, mu, D(2, 1), 2)
    sxx, syy, szz = s[0] + s[1] + s[2], s[3] + s[4] + s[5], s[6] + s[7] + s[8]
    sxy, sxz, syz = s[9] + s[10], s[11] + s[12], s[13] + s[14]
    sr = np.stack([Dxx, Dyy, Dzz, D(2, 1) + D(1, 2), D(2, 0) + D(0, 2), D(1, 0) + D(0, 1)])
    rate = np.stack([l2m * sr[0] + lam * sr[1] + lam * sr[2], lam * sr[0] + l2m * sr[1] + lam * sr[2],
                     lam * sr[0] + lam * sr[1] + l2m * sr[2], mu * sr[3], mu * sr[4], mu * sr[5]])
    s0 = s[15:21] + rk * rate
    detw = det * w
    phid = np.einsum("ime,cie->mce", phi, le)
    S = np.array([[0, 1, 2], [1, 3, 4], [2, 4, 5]])
    hv = np.stack([np.stack([sum(ss[S[c, k]] * phid[m, k] for k in range(3)) for c in range(3)]) for m in range(4)])
    f = np.zeros((12, vel.shape[1]))
    for i in range(8):
        s1, s2, s3 = shp[0, i], shp[1, i], shp[2, i]
        rows = [s1 * sxx, s2 * sxy, s3 * sxz, s1 * sxy, s2 * syy, s3 * syz, s1 * sxz, s2 * syz, s3 * szz,
                s1 * s0[0] + s3 * s0[4] + s2 * s0[5], s2 * s0[1] + s3 * s0[3] + s1 * s0[5], s3 * s0[2] + s2 * s0[3] + s1 * s0[4]]
        hg = np.einsum("me,mce->ce", phi[i], hv)
        for r in range(12):
            val = -detw * rows[r]
            if r >= 9:
                val = val - hg[r - 9]
            np.add.at(f[r], conn[i], val)
    return f, s


@pytest.mark.parametrize("case,grid,min_cover", [("test.tpv104", 296, 0.9), ("test.tpv8", 17, 0.9)], ids=["tpv104-296ctas", "tpv8-17ctas"])
def test_pml_marching_kernel_phases_on_the_host(case, grid, min_cover):
    """eqd_march_pml.h on the CPU: the 15 split stresses, the twelve force rows in their one-number-per-column form and
    the regular + hourglass rows, against calcPMLElemKU evaluated element by element from the stored operators."""
    w = parity.build_world(case, (1, 1, 1), 2)
    v = w.view(0)
    rng = np.random.default_rng(11)
    vel = np.asfortranarray(rng.standard_normal((3, v.Nn)))
    disp = np.asfortranarray(rng.standard_normal((3, v.Nn)) * 1e-2)
    stress0 = np.asfortranarray(rng.standard_normal((21, v.Ne)) * 1e6)
    damps = np.asfortranarray(rng.uniform(0.0, 30.0, (3, v.Ne)))
    dt = 0.004
    stress = stress0.copy(order="F")
    f12, inb, st = _emulate_pml(v, vel, disp, stress, damps, dt, grid)
    E = np.nonzero(inb)[0]
    npml = int((v.elemTypeArr == 2).sum())
    assert st[0] == len(E) >= min_cover * npml and st[0] + st[4] == npml and st[3] == 0
    assert np.all(v.elemTypeArr[E] == 2)
    fref, sref = _reference_pml(v, E, vel, disp, stress0, damps, dt)
    assert np.abs(fref).max() > 0
    for r in range(12):
        assert np.abs(f12[r] - fref[r]).max() <= 1e-12 * np.abs(fref).max(), r
    assert np.abs(stress[:15, E] - sref[:15]).max() <= 1e-13 * np.abs(sref[:15]).max()
    assert np.array_equal(stress[15:, E], stress0[15:, E])           # the regular slots are read, never written
    out = np.ones(v.Ne, bool); out[E] = False
    assert np.array_equal(stress[:, out], stress0[:, out])
    w.close()
