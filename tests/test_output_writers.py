"""The host stand-in's writers for the reference's remaining output files
(library_output.f90:208-312): gm<me> / src_evol<me> streams sampled at
every step with mod(nt,10) == 1 (driver.f90:30-33), surface_coor.txt<me>, finalSurfDisp.txt<me>,
pstr.txt<me>, compTime<me>.  Driven by the CPU oracle (no GPU)."""
import os

import numpy as np

import parity


def _world(case, np_xyz, nstep, switches):
    from eqdyna_b200 import cases
    from eqdyna_b200.host import World
    w = World(cases.materialize(case), np_xyz=np_xyz, nstep=nstep)
    for k, v in switches.items():          # before the build: they size the sample arrays
        w.set_switch(k, v)
    return w.build()


def _rows(path):
    return np.array([[float(x) for x in l.split()] for l in open(path) if l.strip()])


def test_ground_motion_and_source_evolution_streams(tmp_path):
    import oracle
    nstep = 45
    w = _world("test.tpv8", (2, 2, 1), nstep, {"outputGroundMotion": 1, "outputFinalSurfDisp": 1})
    views = [w.view(r) for r in range(w.size)]
    assert sum(v.raw.nSurf for v in views) > 0 and all(v.nGmAlloc == nstep // 10 + 1 for v in views)
    # independent snapshots: stop after every sampled step (nt = 1, 11, 21) and copy what the
    # reference would append there (the P wave reaches the surface above the hypocentre by nt = 41)
    want_gm = [[] for _ in views]
    want_src = [[] for _ in views]
    t, done = 0.0, 0
    for stop in (1, 11, 21, 31, 41, nstep):
        t = oracle.run(w, done + 1, stop, time_elapsed=t)
        done = stop
        if stop % 10 == 1:
            for r, v in enumerate(views):
                if v.raw.nSurf:
                    want_gm[r].append(v.velArr[:, v.surfaceNodeIdArr - 1].T.copy())      # (nSurf,3): node-major triples
                if int(v.nftnd[0]):
                    want_src[r].append(v.fric[46, :int(v.nftnd[0]), 0].copy())
    out = str(tmp_path)
    moved = False
    for r, v in enumerate(views):
        assert int(v.nGmSamples[0]) == 5
        w.set_comp_time(r, [0.5 * k for k in range(10)])
        w.write_outputs(r, out)
        if v.raw.nSurf:
            gm = np.fromfile(os.path.join(out, "gm%d" % r), dtype="<f8")
            assert gm.size == 3 * v.raw.nSurf * 5
            np.testing.assert_array_equal(gm, np.concatenate([x.ravel() for x in want_gm[r]]))
            moved = moved or np.abs(gm).max() > 0
            sc = _rows(os.path.join(out, "surface_coor.txt%d" % r))
            assert sc.shape == (v.raw.nSurf, 3)
            np.testing.assert_allclose(sc, v.meshCoor[:, v.surfaceNodeIdArr - 1].T, rtol=6e-7, atol=1e-30)
            assert np.all(np.abs(sc[:, 2]) < 500.0 / 1000)                                # free-surface nodes only
            fd = _rows(os.path.join(out, "finalSurfDisp.txt%d" % r))
            np.testing.assert_allclose(fd, v.dispArr[:, v.surfaceNodeIdArr - 1].T, rtol=6e-7, atol=1e-30)
        else:
            assert not os.path.exists(os.path.join(out, "gm%d" % r))
        if int(v.nftnd[0]):
            src = np.fromfile(os.path.join(out, "src_evol%d" % r), dtype="<f8")
            np.testing.assert_array_equal(src, np.concatenate(want_src[r]))
        else:
            assert not os.path.exists(os.path.join(out, "src_evol%d" % r))
        ct = open(os.path.join(out, "compTime%d" % r)).read().split()
        assert len(ct) == 12 and int(ct[10]) == v.Ne and int(ct[11]) == v.Neq and float(ct[3]) == 1.5
    assert moved, "no ground motion reached the surface: the samples would be trivially equal"
    # the first line of surface_coor uses the reference's (1x,3e18.7e4) editing
    first = [r for r, v in enumerate(views) if v.raw.nSurf][0]
    line = open(os.path.join(out, "surface_coor.txt%d" % first)).readline().rstrip("\n")
    assert len(line) == 1 + 3 * 18 and line[1 + 18 - 6] == "E"
    w.close()


def test_plastic_strain_file(tmp_path):
    """pstr.txt<me> (library_output.f90:221-244): the filter on pstrain and on the first node's
    position, centroid + pstrain + 12 stress slots per line."""
    w = _world("test.drv.a6", (1, 1, 1), 2, {"output_plastic": 1})
    v = w.view(0)
    x1 = v.meshCoor[:, v.nodeElemIdRelation[0] - 1]                                      # first node of every element
    inside = (np.abs(x1[0]) < 5.0e3) & (np.abs(x1[1]) < 2.0e3) & (np.abs(x1[2]) < 8.0e3)
    reg = v.elemTypeArr != 2
    pick_in = np.nonzero(inside & reg)[0][[3, 40, 500]]
    pick_out = np.nonzero(~inside & reg)[0][:2]
    v.pstrain[pick_in] = [2.0e-4, 3.5e-3, 0.9e-4]                                        # the last one is below the 1e-4 threshold
    v.pstrain[pick_out] = 1.0e-2                                                          # outside the window
    out = str(tmp_path)
    w.write_outputs(0, out)
    rows = _rows(os.path.join(out, "pstr.txt0"))
    assert rows.shape == (2, 16)
    for row, e in zip(rows, pick_in[:2]):
        c = v.meshCoor[:, v.nodeElemIdRelation[:, e] - 1].sum(axis=1) / 8.0
        np.testing.assert_allclose(row[:3], c, rtol=6e-7, atol=1e-6)
        assert abs(row[3] - v.pstrain[e]) <= 1e-7 * v.pstrain[e]
        s0 = int(v.stressCompIndexArr[e])
        np.testing.assert_allclose(row[4:], v.stressArr[s0:s0 + 12], rtol=6e-7, atol=1e-30)
    w.close()
