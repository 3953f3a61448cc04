#!/usr/bin/env python3
"""GPU-vs-oracle parity report for one or more cases (run on a GPU box).
usage: gpu_parity.py CASE[:npx,npy,npz][:nstep] ..."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import parity  # noqa: E402

out = {}
for spec in sys.argv[1:]:
    parts = spec.split(":")
    case = parts[0]
    npx = tuple(int(x) for x in parts[1].split(",")) if len(parts) > 1 and parts[1] else None
    nstep = int(parts[2]) if len(parts) > 2 else 0
    t = time.time()
    wg = parity.build_world(case, npx, nstep)
    wo = parity.build_world(case, npx, nstep)
    tb = time.time() - t
    t = time.time()
    doms = parity.run_gpu(wg, options={"timing": 1})
    tg = time.time() - t
    t = time.time()
    parity.run_oracle(wo)
    to = time.time() - t
    print("== %s  build %.1fs  gpu %.2fs  oracle %.1fs  timing %s counts %s" % (spec, tb, tg, to, doms[0].timing(), doms[0].counts()))
    res = parity.compare_worlds(wg, wo, verbose=True)
    out[spec] = res
    try:
        parity.assert_parity(res)
        print("  PARITY OK")
    except AssertionError as e:
        print("  " + str(e))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/parity.json", "w"), indent=1)
