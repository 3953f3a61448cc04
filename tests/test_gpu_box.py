"""Parity of the closed-form box operators (eqd_set_option "box" / "box_compact",
eqdyna_b200/csrc/cuda/eqd_box.h) against the CPU oracle, through the C ABI.  Same
tolerances as test_gpu_parity.py (BASELINE.json: 1e-6 relative L2, rupture time within
one step).  The oracle uses the reference's precomputed eleshp / phi / ss; the box tiles
rebuild them as sign*a_d / ha / diag(ss), which differs from the reference's own rounding
by <= 1.4e-13 (tests/test_host_and_abi.py::test_box_operators)."""
import pytest

import parity

pytestmark = pytest.mark.gpu

# (case, decomposition, steps, options, switches, what)
CASES = [
    ("test.tpv8", (1, 1, 1), 20, {"box": 2, "box_compact": 1}, None, "every tile a box tile: compact stage buffer, three CTAs per SM"),
    ("test.tpv104", (2, 2, 2), 60, {"box": 2, "box_compact": 1}, None, "RSF; halo: box flags follow the rank-face-first tile order"),
    ("test.tpv8", (1, 1, 1), 20, {"box": 1}, None, "regular classes only, PML tiles stream every row"),
    ("test.tpv10", (2, 2, 2), 30, {"box": 2}, None, "warped mesh: box and general tiles mixed in one launch, REGX class"),
    ("test.tpv36", (2, 2, 2), 40, {"box": 2}, None, "wedges (never box) next to box tiles"),
    ("test.tpv8", (2, 2, 1), 20, {"box": 2}, {"C_Q": 1}, "Q path: displacement strains in closed form"),
]


@pytest.mark.parametrize("case,np_xyz,nstep,options,switches,what", CASES,
                         ids=["%s-%dx%dx%d-%s%s" % (c, *d, "-".join("%s%d" % kv for kv in o.items()), "-Q" if s else "")
                              for c, d, _, o, s, _ in CASES])
def test_box_operators_match_oracle(case, np_xyz, nstep, options, switches, what):
    wg = parity.build_world(case, np_xyz, nstep, switches)
    wo = parity.build_world(case, np_xyz, nstep, switches)
    doms = parity.run_gpu(wg, options=options)
    parity.run_oracle(wo)
    res = parity.compare_worlds(wg, wo)
    parity.assert_parity(res)
    nbox = sum(d.box_counts()["regular"] for d in doms)
    assert nbox > 0, "no tile took the box path"
    if case in ("test.tpv8", "test.tpv104"):
        assert nbox == sum(d.counts()["regular"] for d in doms)     # rectilinear meshes: all of them


def test_box_and_full_rows_agree_to_rounding():
    """Same run with and without the option: bulk fields agree to ~1e-12 (the operators differ
    by the reference's own rounding only)."""
    n = 20
    a = parity.build_world("test.tpv8", (1, 1, 1), n); parity.run_gpu(a)
    b = parity.build_world("test.tpv8", (1, 1, 1), n); parity.run_gpu(b, options={"box": 2, "box_compact": 1})
    for name in ("dispArr", "velArr", "stressArr"):
        x, y = getattr(a.view(0), name), getattr(b.view(0), name)
        assert parity.rel_l2(y, x) < 1e-9, name
