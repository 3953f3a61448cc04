"""Operator precompute on the device (-m gpu): eqd_compute_elem_ops against the
host's assembleGlobalMass restatement (src/assembleGlobalMass.f90:3-56,283-406,
calcGlobalShapeFunc.f90, library.f90:60-93) on the same meshes, and the step loop
run from device-computed operators against the CPU oracle."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["test.tpv8", "test.tpv10", "test.tpv36", "test.drv.a6"],
                         ids=["tpv8-bricks+pml", "tpv10-warped", "tpv36-wedges", "drv.a6-fractal-fault"])
def test_device_operators_equal_host_operators(case):
    """Per-element operators are bit-identical (same operation order, no FMA contraction);
    the lumped nodal mass agrees to rounding (different, but fixed, summation order)."""
    from eqdyna_b200 import device as dev
    w = parity.build_world(case, (1, 1, 1), 4)
    v = w.view(0)
    r = v.raw
    d = dev.Domain(v, device=0, compute_ops=True)
    try:
        for which, host, shape in ((dev.F_ELEDET, v.eledet, (r.Ne,)), (dev.F_ELESHP, v.eleshp, (3, 8, r.Ne)),
                                   (dev.F_SS, v.ss, (6, r.Ne)), (dev.F_PHI, v.phi, (8, 4, r.Ne))):
            got = d.fetch(which, shape)
            ref = np.asarray(host)
            assert ref.shape == tuple(shape)
            assert np.array_equal(got, ref), "operator %d differs from the host's: max rel %.3e" % (
                which, float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)))
        mass = d.fetch(dev.F_MASS, (r.Neq,))
        ref = np.asarray(v.nodalMassArr)
        assert np.max(np.abs(mass - ref) / np.abs(ref)) <= 1.0e-14
    finally:
        d.close()
        w.close()


@pytest.mark.parametrize("case,nstep", [("test.tpv8", 60), ("test.tpv36", 120), ("test.tpv104", 60)],
                         ids=["tpv8", "tpv36-wedges", "tpv104-rsf"])
def test_step_loop_from_device_operators_matches_oracle(case, nstep):
    wg = parity.build_world(case, (1, 1, 1), nstep)
    wo = parity.build_world(case, (1, 1, 1), nstep)
    parity.run_gpu(wg, compute_ops=True)
    parity.run_oracle(wo)
    parity.assert_parity(parity.compare_worlds(wg, wo))
    wg.close()
    wo.close()


def test_non_positive_determinant_is_reported():
    """calcGlobalShapeFunc.f90:57-61 stops on det <= 0; the library returns the element number."""
    from eqdyna_b200 import device as dev
    w = parity.build_world("test.tpv8", (1, 1, 1), 4)
    v = w.view(0)
    keep = v.meshCoor.copy()
    try:
        n0, n1 = int(v.nodeElemIdRelation[0, 100]) - 1, int(v.nodeElemIdRelation[6, 100]) - 1
        a, b = v.meshCoor[:, n0].copy(), v.meshCoor[:, n1].copy()
        v.meshCoor[:, n0], v.meshCoor[:, n1] = b, a   # invert one brick
        with pytest.raises(dev.StepError) as ei:
            dev.Domain(v, device=0, compute_ops=True)
        assert "determinant" in str(ei.value)
    finally:
        v.meshCoor[...] = keep
        w.close()
