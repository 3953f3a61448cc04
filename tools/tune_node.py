#!/usr/bin/env python3
"""Time the launch-bounds variants of k_node_update3 on the benchmark mesh (run on a GPU box).
usage: tune_node.py [CASE]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqdyna_b200 import cases, device as dev  # noqa: E402
from eqdyna_b200.host import World  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "bench.tpv104_100m"
w = World(cases.materialize(case), np_xyz=(1, 1, 1), nstep=600)
w.build(rank=0, sum_shared=False)
d = dev.Domain(w.view(0), device=0, compute_ops=True)
box = int(os.environ.get("EQD_TUNE_BOX", "2"))     # the bench default; 0 = every operator row streamed
d.set_option("box", box)
d.set_option("box_compact", 1 if box else 0)
d.set_option("timing", 1)
d.run(1, 10)
nt = 10
for rep in range(2):
    for variant in (4,):
        d.set_option("node_variant", variant)
        d.set_option("timing", 2)
        d.run(nt + 1, nt + 25)
        nt += 25
        t = d.timing()
        print("variant %d: node %.4f elem %.4f pml %.4f fault %.4f total %.4f ms/step" % (variant, t["node"] / 25, t["elem"] / 25, t["elem_pml"] / 25, t["fault"] / 25, t["total"] / 25), flush=True)
d.close()
w.close()
