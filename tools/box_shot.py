#!/usr/bin/env python3
"""One short GPU session for the closed-form box operators (eqd_set_option "box"):
parity against the CPU oracle on a few cases, then per-kernel timings with the option
off / on.  Results are appended to gpurun_out/box_shot.jsonl after every item so that a
run cut short still leaves what it measured.

  python tools/box_shot.py [budget_seconds]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "box_shot.jsonl")
T0 = time.time()
BUDGET = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0


def emit(rec):
    rec["t"] = round(time.time() - T0, 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def left():
    return BUDGET - (time.time() - T0)


def parity_case(case, np_xyz, nstep, box, switches=None):
    rec = {"kind": "parity", "case": case, "np": list(np_xyz), "nstep": nstep, "box": box, "switches": switches or {}}
    try:
        wg = parity.build_world(case, np_xyz, nstep, switches)
        wo = parity.build_world(case, np_xyz, nstep, switches)
        doms = parity.run_gpu(wg, options={"box": box})
        rec["box_counts"] = [d.box_counts() for d in doms]
        rec["counts"] = [d.counts() for d in doms]
        parity.run_oracle(wo)
        res = parity.compare_worlds(wg, wo)
        rec["worst"] = max(v for k, v in res.items() if not k.startswith("rupt"))
        rec["rupt"] = {k: v for k, v in res.items() if k.startswith("rupt")}
        try:
            parity.assert_parity(res)
            rec["ok"] = True
        except AssertionError as e:
            rec["ok"] = False
            rec["why"] = str(e)[:400]
        for d in doms:
            d.close()
        wg.close(); wo.close()
    except Exception as e:  # noqa: BLE001
        rec["ok"] = False
        rec["why"] = repr(e)[:400]
    emit(rec)
    return rec.get("ok", False)


def timing_case(case, nstep, boxes=(0, 1, 2), compute_ops=True):
    from eqdyna_b200 import device as dev
    try:
        w = parity.build_world(case, (1, 1, 1), nstep + 12)
        v = w.view(0)
        ref = None
        for box in boxes:
            d = dev.Domain(v, compute_ops=compute_ops)
            d.set_option("box", box)
            d.set_option("timing", 1)
            d.run(1, 10)
            d.set_option("timing", 2)
            d.run(11, 10 + nstep)
            tm = d.timing()
            rec = {"kind": "timing", "case": case, "elements": int(v.Ne), "nstep": nstep, "box": box,
                   "box_counts": d.box_counts(), "ms_per_step": {k: round(x / nstep, 4) for k, x in tm.items()}}
            vel = d.fetch(dev.F_VEL, (3, v.raw.Nn))
            if ref is None:
                ref = vel
            else:
                import numpy as np
                rec["vel_rel_l2_vs_box0"] = float(np.sqrt(((vel - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-300))
            emit(rec)
            d.close()
        w.close()
    except Exception as e:  # noqa: BLE001
        emit({"kind": "timing", "case": case, "ok": False, "why": repr(e)[:400]})


def main():
    emit({"kind": "start", "budget": BUDGET})
    ok = parity_case("test.tpv8", (1, 1, 1), 20, 2)                       # SW + PML, every tile a box tile, one sub-domain
    ok &= parity_case("test.tpv104", (2, 2, 2), 60, 2)                    # RSF (the benchmark's physics), halo, face-first tile order
    if left() > 100:
        timing_case("bench.tpv104_200m", 30)                              # 5.3 M elements
    if left() > 80:
        ok &= parity_case("test.tpv10", (2, 2, 2), 30, 2)                 # warped mesh: box and general tiles mixed, REGX
    if left() > 70:
        ok &= parity_case("test.tpv36", (2, 2, 2), 40, 2)                 # wedges
    if left() > 60:
        ok &= parity_case("test.drv.a6", (2, 2, 1), 8, 2)                 # plastic + body force
    if left() > 50:
        ok &= parity_case("test.tpv8", (2, 2, 1), 20, 2, {"C_Q": 1})      # Q path
    if left() > 50:
        ok &= parity_case("test.tpv8", (1, 1, 1), 20, 1)                  # regular classes only
    if left() > 90:
        timing_case("bench.tpv104_100m", 20, boxes=(0, 2))                # the benchmark mesh
    emit({"kind": "done", "all_parity_ok": bool(ok)})


if __name__ == "__main__":
    main()
