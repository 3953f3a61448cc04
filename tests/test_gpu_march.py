"""Parity of the marching kernel (eqd_set_option "march", eqdyna_b200/csrc/cuda/eqd_march.h) against the CPU
oracle, through the C ABI: bundles of box elements whose inner nodes are updated by the element sweep
itself, the rest of the mesh (PML, elements next to a dipping fault, wedges) on the tile kernels in the
same step.  Same tolerances as test_gpu_parity.py (BASELINE.json: 1e-6 relative L2 on fault / station
series, rupture time within one step)."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

OPTS = {"box": 2, "box_compact": 1}

# (case, decomposition, steps [0 = the case's own], chunks, device operators, what)
CASES = [
    ("test.tpv8", (1, 1, 1), 0, 1, False, "slip weakening + PML, full 114 steps, host operators"),
    ("test.tpv8", (1, 1, 1), 40, 4, True, "four eqd_run calls (the last step of each leaves forces instead of updating), device operators + mass"),
    ("test.tpv104", (2, 2, 1), 0, 1, False, "RSF, the reference's decomposition, full 120 steps"),
    ("test.tpv104", (1, 1, 1), 60, 1, True, "device operators + lumped mass of the bundles (k_march_mass)"),
    ("test.tpv104", (4, 1, 2), 60, 2, False, "the benchmark's 8-GPU decomposition"),
    ("test.tpv36", (2, 2, 2), 60, 1, False, "dipping fault: bundles only away from it, wedges and type-13 bricks on the tile kernels"),
    ("test.tpv10", (1, 1, 1), 40, 1, False, "warped mesh below the fault: box region marches, the rest does not"),
]


@pytest.mark.parametrize("case,np_xyz,nstep,chunks,cops,what", CASES,
                         ids=["%s-%dx%dx%d-%d-c%d%s" % (c, *d, n, k, "-devops" if o else "") for c, d, n, k, o, _ in CASES])
def test_marching_kernel_matches_oracle(case, np_xyz, nstep, chunks, cops, what):
    wg = parity.build_world(case, np_xyz, nstep)
    wo = parity.build_world(case, np_xyz, nstep)
    doms = parity.run_gpu(wg, options=OPTS, pre_options={"march": 1}, chunks=chunks, compute_ops=cops)
    parity.run_oracle(wo)
    res = parity.compare_worlds(wg, wo)
    parity.assert_parity(res)
    t = [d.timing() for d in doms]
    assert sum(d.box_counts()["regular"] for d in doms) > 0


def test_march_timing_slot_and_determinism():
    """the marching kernel really ran (its timing slot is non-zero), twice the same bits; one run in three
    chunks agrees to rounding (the chunk ends take the unfused path: same operations, another kernel)."""
    n = 30
    runs = []
    for chunks in (1, 1, 3):
        w = parity.build_world("test.tpv104", (1, 1, 1), n)
        from eqdyna_b200 import device as dev
        d = dev.Domain(w.view(0), options={"march": 1})
        for k, v in OPTS.items():
            d.set_option(k, v)
        d.set_option("timing", 1)
        b = np.linspace(0, n, chunks + 1).astype(int)
        for a, e in zip(b[:-1], b[1:]):
            d.run(a + 1, e)
        assert d.timing()["march"] > 0.0 and d.timing()["march_pml"] > 0.0
        mc = d.march_counts()
        assert mc["elements"] > 0 and mc["pml_elements"] > 0.8 * d.counts()["pml"]
        d.fetch_into_view()
        runs.append(w)
    a, b, c = (w.view(0) for w in runs)
    for name in ("dispArr", "velArr", "v1", "nodalForceArr", "fric", "stressArr", "onFaultQuantHistSCECForm"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert parity.rel_l2(getattr(c, name), getattr(a, name)) <= 1e-12, name


def test_march_and_tiles_agree_to_rounding():
    """tile kernels only / bundles / bundles with shared ghost rows and columns: the same physics to rounding"""
    n = 20
    a = parity.build_world("test.tpv8", (1, 1, 1), n); parity.run_gpu(a, options=OPTS)
    for march in (1, 2):
        b = parity.build_world("test.tpv8", (1, 1, 1), n); parity.run_gpu(b, options=OPTS, pre_options={"march": march})
        for name in ("dispArr", "velArr", "stressArr", "nodalForceArr"):
            assert parity.rel_l2(getattr(b.view(0), name), getattr(a.view(0), name)) < 1e-9, (march, name)


@pytest.mark.parametrize("case,np_xyz,nstep,march", [("test.tpv104", (2, 2, 1), 60, 2), ("test.tpv8", (1, 1, 1), 40, 2), ("test.tpv36", (2, 2, 2), 40, 2),
                                                     ("test.tpv104", (2, 2, 1), 60, 3)])
def test_march_with_ghost_sharing_matches_oracle(case, np_xyz, nstep, march):
    """option march = 2 / 3: neighbouring strips share ghost rows and columns / columns only, v and d double-buffered"""
    wg = parity.build_world(case, np_xyz, nstep)
    wo = parity.build_world(case, np_xyz, nstep)
    parity.run_gpu(wg, options=OPTS, pre_options={"march": march}, chunks=2)
    parity.run_oracle(wo)
    parity.assert_parity(parity.compare_worlds(wg, wo))
