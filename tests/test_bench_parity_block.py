"""bench.py's `parity` block (CPU part): the position-keyed extraction of fault / station fields and
their comparison must be independent of the decomposition, because the benchmark compares the GPU
run (1, 2, 4 or 8 sub-domains) with an oracle run split over the host cores.  Two oracle runs of
test.tpv104 with different decompositions play both roles here (no GPU)."""
import numpy as np

import parity


def _fields(decomp, nstep):
    import bench
    w = parity.build_world("test.tpv104", decomp, nstep)
    parity.run_oracle(w, nstep, threads=4)
    f = bench.merge_fields([bench.fault_and_station_fields(w.view(r), nstep) for r in range(w.size)])
    dt = float(w.view(0).params.dt)
    w.close()
    return f, dt


def test_parity_block_is_decomposition_independent_and_detects_errors():
    import bench
    nstep = 40
    a, dt = _fields((2, 2, 1), nstep)
    b, _ = _fields((1, 2, 2), nstep)
    assert len(a["pairs"]) == len(b["pairs"]) == 2701 and len(a["onst"]) == len(b["onst"]) > 0 and len(a["offst"]) == len(b["offst"]) > 0
    blk = bench.parity_block(a, b, nstep, dt, "oracle 2x2x1 vs oracle 1x2x2")
    assert blk["ok"] and blk["worst"] <= 1e-9 and blk["unmatched_pairs"] == 0, blk
    assert set(blk["fields"]) >= {"fric.sliprate(74:75)", "fric.traction(78:80)", "onfault.sliprate", "onfault.shear", "station.vel"}
    # a perturbation at the contract level must show
    k = next(iter(a["pairs"]))
    big = max(abs(x[7:10]).max() for x in a["pairs"].values())
    a["pairs"][k] = a["pairs"][k].copy()
    a["pairs"][k][8] += 1e-3 * big
    bad = bench.parity_block(a, b, nstep, dt, "perturbed")
    assert not bad["ok"] and bad["fields"]["fric.traction(78:80)"] > 1e-6
