#!/usr/bin/env python3
"""Time tile-brick shapes of the element kernels on the benchmark mesh (run on a GPU box).
usage: tune_tiles.py [CASE]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqdyna_b200 import cases, device as dev  # noqa: E402
from eqdyna_b200.host import World  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "bench.tpv104_100m"
w = World(cases.materialize(case), np_xyz=(1, 1, 1), nstep=100)
w.build(rank=0, sum_shared=False)
CONFIGS = [
    ((4, 4, 16), (6, 3, 12)),
    ((4, 4, 16), (3, 6, 12)),
    ((4, 4, 16), (6, 4, 9)),
    ((4, 4, 16), (6, 2, 16)),
    ((4, 4, 16), (3, 4, 16)),
    ((4, 4, 16), (6, 8, 5)),
    ((2, 4, 21), (6, 7, 6)),
    ((3, 3, 14), (6, 7, 6)),
    # worth re-timing with EQD_TUNE_BANK_ORDER=2 (conflict free under the residue numbering, fewer node slots per element):
    ((4, 4, 14), (6, 5, 6)),
    ((5, 4, 12), (4, 7, 6)),
]
if len(sys.argv) > 2:            # tune_tiles.py CASE N: only the first N configurations
    CONFIGS = CONFIGS[:int(sys.argv[2])]
for reg, pml in CONFIGS:
    opts = dict(zip(("reg_bx", "reg_bz", "reg_by", "pml_bx", "pml_bz", "pml_by"), reg + pml))
    opts["bank_order"] = int(os.environ.get("EQD_TUNE_BANK_ORDER", "0"))   # 1: element order, 2: residue node numbering
    try:
        d = dev.Domain(w.view(0), device=0, compute_ops=True, options=opts)
    except Exception as e:  # noqa: BLE001 -- a brick that does not fit the node cap: report and go on
        print("reg %s pml %s: rejected (%s)" % (reg, pml, str(e)[:120]), flush=True)
        continue
    box = int(os.environ.get("EQD_TUNE_BOX", "2"))     # the bench default; 0 = every operator row streamed
    d.set_option("box", box)
    d.set_option("box_compact", 1 if box else 0)
    d.set_option("timing", 1)
    d.run(1, 10)
    d.set_option("timing", 2)
    d.run(11, 40)
    t = d.timing()
    print("reg %s pml %s: elem %.4f pml %.4f node %.4f total %.4f ms/step" % (
        reg, pml, t["elem"] / 30, t["elem_pml"] / 30, t["node"] / 30, t["total"] / 30), flush=True)
    d.close()
w.close()
