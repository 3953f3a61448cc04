"""Shared parity helpers: run one case through the CUDA step library and
through the CPU oracle from the same initial state and compare.

Tolerances (BASELINE.json north_star): fault slip-rate / shear-stress series and
station velocity seismograms within 1e-6 relative L2; rupture time within one dt.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = os.path.join(ROOT, "tests", "golden", "cases")
REF_RESULTS = os.path.join(ROOT, "tests", "golden", "ref_results")

REL_L2_TOL = 1.0e-6

# Physical unit scales used as noise floors: a field whose reference norm is below
# 1e-6 x (unit scale) x sqrt(n) -- slip of 1e-17 m on a fault patch the rupture has
# not reached, say -- is rounding noise in both implementations and is measured
# against that floor instead of against itself.
UNIT = {"slip": 1.0, "sliprate": 1.0, "peakrate": 1.0, "cumslip": 1.0, "traction": 1.0e6, "state": 1.0, "vms": 1.0,
        "finalrate": 1.0}


def rel_l2(a, b, unit=0.0):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = max(np.sqrt(np.sum(b * b)), 1.0e-6 * unit * np.sqrt(max(b.size, 1)))
    num = np.sqrt(np.sum((a - b) ** 2))
    if den == 0.0:
        return 0.0 if num == 0.0 else float("inf")
    return float(num / den)


def series_err(G, O, floor_rel=1.0e-6, unit=0.0):
    """Worst per-station relative L2 of time series G,O shaped (..., nstation).
    A station whose reference series is (numerically) silent -- e.g. slip rate at a
    node the rupture never reaches, normal traction at a free-surface node -- has
    no meaningful relative error; its error is measured against floor_rel times
    the loudest station's norm instead of its own."""
    ns = O.shape[-1]
    norms = [np.sqrt(np.sum(np.asarray(O[..., s], dtype=np.float64) ** 2)) for s in range(ns)]
    top = max(norms) if norms else 0.0
    worst = 0.0
    for s in range(ns):
        num = np.sqrt(np.sum((np.asarray(G[..., s], dtype=np.float64) - O[..., s]) ** 2))
        den = max(norms[s], floor_rel * top, 1.0e-6 * unit * np.sqrt(max(np.asarray(O[..., s]).size, 1)))
        if den == 0.0:
            e = 0.0 if num == 0.0 else float("inf")
        else:
            e = num / den
        worst = max(worst, float(e))
    return worst


def build_world(case, np_xyz=None, nstep=0, switches=None, pre_switches=None):
    """switches are applied after the build (run-time parameters), pre_switches before it
    (those that size arrays: outputGroundMotion, ...)."""
    from eqdyna_b200 import cases
    from eqdyna_b200.host import World
    w = World(cases.materialize(case), np_xyz=np_xyz, nstep=nstep)
    for k, v in (pre_switches or {}).items():
        w.set_switch(k, v)
    w.build()
    for k, v in (switches or {}).items():
        w.set_switch(k, v)
    return w


def run_oracle(world, nstep=None, threads=None):
    import oracle
    n = world.view(0).nstep if nstep is None else nstep
    oracle.run(world, 1, n, threads=threads)
    return world


def run_gpu(world, nstep=None, device=0, chunks=1, options=None, compute_ops=False, pre_options=None):
    """Step every sub-domain of `world` on the GPU and copy the results back into
    the world's host arrays.  Returns the Domain list (caller may read timing).
    pre_options are set before the upload (those that shape it: "march", tile bricks), options after it."""
    from eqdyna_b200 import device as dev
    n = world.view(0).nstep if nstep is None else nstep
    doms = [dev.Domain(world.view(r), device=device, compute_ops=compute_ops, options=pre_options) for r in range(world.size)]
    for d in doms:
        for k, v in (options or {}).items():
            d.set_option(k, v)
    bounds = np.linspace(0, n, chunks + 1).astype(int)
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b <= a:
            continue
        if world.size == 1:
            doms[0].run(a + 1, b)
        else:
            dev.run_group(doms, a + 1, b)
    for d in doms:
        d.fetch_into_view()
    return doms


def compare_worlds(wg, wo, nstep=None, verbose=False):
    """Compare GPU-run world `wg` against oracle-run world `wo`.  Returns a dict of
    error measures; raises AssertionError with a readable message on violation."""
    res = {}
    worst = {}

    def upd(key, val):
        worst[key] = max(worst.get(key, 0.0), val)

    for r in range(wg.size):
        g, o = wg.view(r), wo.view(r)
        n = g.nstep if nstep is None else nstep
        dt = g.params.dt
        # integer mesh arrays are inputs here (same host), nothing to compare.
        upd("disp", rel_l2(g.dispArr, o.dispArr))
        upd("vel", rel_l2(g.velArr, o.velArr))
        upd("v1", rel_l2(g.v1, o.v1))
        upd("accel", rel_l2(g.nodalForceArr, o.nodalForceArr))
        used = g.raw.stressUsed
        upd("stress", rel_l2(g.stressArr[:used], o.stressArr[:used]))
        if g.params.C_elastic == 0:
            upd("pstrain", rel_l2(g.pstrain, o.pstrain))
        npairs = int(np.sum(g.nftnd))
        if npairs:
            for ift in range(g.ntotft):
                k = int(g.nftnd[ift])
                if not k:
                    continue
                fg, fo = g.fric[:, :k, ift], o.fric[:, :k, ift]
                for name, sl in (("slip", slice(70, 73)), ("sliprate", slice(73, 75)), ("peakrate", slice(75, 76)),
                                 ("cumslip", slice(76, 77)), ("traction", slice(77, 80)), ("state", slice(19, 20)),
                                 ("vms", slice(30, 36)), ("finalrate", slice(46, 48))):
                    upd("fric." + name, rel_l2(fg[sl], fo[sl], UNIT[name]))
                tg, to = g.fnft[:k, ift], o.fnft[:k, ift]
                both = (tg < 5000.0) & (to < 5000.0)
                upd("rupt_mismatch", float(np.sum((tg < 5000.0) != (to < 5000.0))))
                if both.any():
                    upd("rupt_time_steps", float(np.max(np.abs(tg[both] - to[both])) / dt))
            if g.nOn:
                hg = g.onFaultQuantHistSCECForm[:, :n, :g.nOn]
                ho = o.onFaultQuantHistSCECForm[:, :n, :g.nOn]
                upd("onfault.sliprate", series_err(hg[1:3], ho[1:3], unit=1.0))
                upd("onfault.shear", series_err(hg[7:9], ho[7:9], unit=1.0e6))
                upd("onfault.normal", series_err(hg[9], ho[9], unit=1.0e6))
                upd("onfault.slip", series_err(hg[4:7], ho[4:7], unit=1.0))
            upd("hypolog", rel_l2(g.hypoLog[:, :n], o.hypoLog[:, :n]))
        if g.nOff:
            sg = g.OffFaultStGramSCEC[:, :n]
            so = o.OffFaultStGramSCEC[:, :n]
            # rows: time, then per station (dof x,y,z) x (disp, vel)
            vg = sg[1:].reshape(g.nOff, 3, 2, -1)
            vo = so[1:].reshape(g.nOff, 3, 2, -1)
            upd("station.vel", series_err(np.moveaxis(vg[:, :, 1, :], 0, -1), np.moveaxis(vo[:, :, 1, :], 0, -1), unit=1.0e-3))
            upd("station.disp", series_err(np.moveaxis(vg[:, :, 0, :], 0, -1), np.moveaxis(vo[:, :, 0, :], 0, -1), unit=1.0e-3))
    res.update(worst)
    if verbose:
        for k in sorted(res):
            print("  %-20s %.3e" % (k, res[k]))
    return res


def assert_parity(res, tol=REL_L2_TOL):
    bad = []
    for k, v in res.items():
        if k == "rupt_time_steps":
            if v > 1.0 + 1e-9:
                bad.append((k, v))
        elif k == "rupt_mismatch":
            if v > 0:
                bad.append((k, v))
        elif not (v <= tol):
            bad.append((k, v))
    assert not bad, "parity violated: " + ", ".join("%s=%.3e" % kv for kv in bad)
