#!/usr/bin/env python3
"""Target of the ncu capture of the box tile kernels: TPV104 at dx = 200 m (5.3 M
elements), options box = 2 + box_compact, 8 steps on cuda:0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from eqdyna_b200 import device as dev  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "bench.tpv104_200m"
w = parity.build_world(case, (1, 1, 1), 10)
d = dev.Domain(w.view(0), compute_ops=True)
d.set_option("box", 2)
d.set_option("box_compact", int(os.environ.get("EQD_BOX_COMPACT", "1")))
d.run(1, 8)
print(d.counts(), d.box_counts())
d.close()
