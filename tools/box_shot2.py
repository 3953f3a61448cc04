#!/usr/bin/env python3
"""Second short GPU session for the box operators: parity of the compact-buffer variant
(option "box_compact": three CTAs per SM), then per-kernel timings on the benchmark mesh
for a few tile bricks.  Appends to gpurun_out/box_shot2.jsonl item by item.

  python tools/box_shot2.py [budget_seconds]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import box_shot as bs  # noqa: E402

bs.OUT = os.path.join(ROOT, "gpurun_out", "box_shot2.jsonl")
parity = bs.parity


def parity_compact(case, np_xyz, nstep):
    rec = {"kind": "parity", "case": case, "np": list(np_xyz), "nstep": nstep, "box": 2, "box_compact": 1}
    try:
        wg = parity.build_world(case, np_xyz, nstep)
        wo = parity.build_world(case, np_xyz, nstep)
        doms = parity.run_gpu(wg, options={"box": 2, "box_compact": 1})
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
    bs.emit(rec)


def timing_variants(case, nstep, variants):
    import numpy as np
    from eqdyna_b200 import device as dev
    try:
        w = parity.build_world(case, (1, 1, 1), nstep + 12)
        v = w.view(0)
        ref = None
        for opts_pre, opts_post in variants:
            if bs.left() < 12:
                break
            d = dev.Domain(v, compute_ops=True, options=opts_pre)
            for k, val in opts_post.items():
                d.set_option(k, val)
            d.set_option("timing", 1)
            d.run(1, 10)
            d.set_option("timing", 2)
            d.run(11, 10 + nstep)
            tm = d.timing()
            rec = {"kind": "timing", "case": case, "elements": int(v.Ne), "nstep": nstep, "tiles": opts_pre, "opts": opts_post,
                   "box_counts": d.box_counts(), "ms_per_step": {k: round(x / nstep, 4) for k, x in tm.items()}}
            vel = d.fetch(dev.F_VEL, (3, v.raw.Nn))
            if ref is None:
                ref = vel
            else:
                rec["vel_rel_l2_vs_first"] = float(np.sqrt(((vel - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-300))
            bs.emit(rec)
            d.close()
        w.close()
    except Exception as e:  # noqa: BLE001
        bs.emit({"kind": "timing", "case": case, "ok": False, "why": repr(e)[:400]})


def main():
    bs.emit({"kind": "start", "budget": bs.BUDGET})
    parity_compact("test.tpv8", (1, 1, 1), 20)
    parity_compact("test.tpv104", (2, 2, 2), 60)
    C = {"box": 2, "box_compact": 1}
    timing_variants("bench.tpv104_100m", 20, [
        ({}, {"box": 2}),
        ({}, C),
        ({"reg_bx": 5, "reg_bz": 4, "reg_by": 12}, C),
        ({"reg_bx": 4, "reg_bz": 4, "reg_by": 14}, C),
        ({"reg_bx": 3, "reg_bz": 4, "reg_by": 16}, C),
        ({"reg_bx": 4, "reg_bz": 5, "reg_by": 12}, C),
    ])
    bs.emit({"kind": "done"})


if __name__ == "__main__":
    main()
