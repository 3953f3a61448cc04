#!/usr/bin/env python3
"""GPU triage of the marching kernel (run on a GPU box): steps a small case with option "march" for 1, 2, 3, 10
steps and prints, against the CPU oracle, the relative error of velocity / displacement separately for the
nodes the bundles update themselves and for the others, and of the stresses for bundle / other elements."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import test_march_emulation as T  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "test.tpv8"
decomp = tuple(int(x) for x in sys.argv[2].split("x")) if len(sys.argv) > 2 else (1, 1, 1)
for n in (1, 2, 3, 10):
    wg = parity.build_world(case, decomp, n)
    wo = parity.build_world(case, decomp, n)
    v0 = wg.view(0)
    z = np.zeros((3, v0.Nn), order="F")
    _, fused, inb, st = T._emulate(v0, z.copy(order="F"), z.copy(order="F"), np.zeros((6, v0.Ne), order="F"), np.ones(v0.Nn), 1e-3, 0, 444)
    doms = parity.run_gpu(wg, options={"box": 2, "box_compact": 1}, pre_options={"march": 2})
    parity.run_oracle(wo)
    g, o = wg.view(0), wo.view(0)

    def err(a, b, m):
        d = np.abs(a[..., m] - b[..., m]).max() if m.any() else 0.0
        return d / max(np.abs(b).max(), 1e-300)
    s6 = lambda v: np.stack([v.stressArr[v.stressCompIndexArr + k] for k in range(6)])  # noqa: E731
    reg = g.elemTypeArr != 2
    print("steps %2d: vel fused %.2e other %.2e | disp fused %.2e other %.2e | stress bundle %.2e other-regular %.2e | march %s" % (
        n, err(g.velArr, o.velArr, fused), err(g.velArr, o.velArr, ~fused), err(g.dispArr, o.dispArr, fused), err(g.dispArr, o.dispArr, ~fused),
        err(s6(g), s6(o), inb), err(s6(g), s6(o), reg & ~inb), doms[0].march_counts()), flush=True)
    res = parity.compare_worlds(wg, wo)
    print("          ", {k: "%.1e" % x for k, x in res.items() if x > 1e-9}, flush=True)
    for d in doms:
        d.close()
    wg.close(); wo.close()
