"""Host-side set-up and fetch logic of the step library, exercised WITHOUT a GPU by linking the
library's objects against a stand-in for the CUDA runtime (tools/hoststub/cudart_stub.cpp:
"device" memory is host memory, copies are memcpy, kernels do nothing).  What this can check is
exactly the part of the C ABI that is host code: eqd_set_* conversions (AoS -> SoA, equation
indirection, padding), the tile planner's uploads, and eqd_fetch's way back, including the
staged path large sub-domains take.  It computes no physics -- nothing a kernel would produce
exists in this mode -- and the product never loads the stub (see the header of the .cpp).
Runs in a subprocess so that the stand-in never shares a process with the real library."""
import os
import subprocess
import sys
import textwrap

import parity

ROOT = parity.ROOT

SCRIPT = textwrap.dedent(r"""
    import ctypes as C, os, sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tools", "hoststub"))
    import setup_probe
    lib, stub = setup_probe.build_stub()
    from eqdyna_b200 import build, cases, device as dev
    from eqdyna_b200.host import World
    build.cuda_lib_path = lambda: lib
    S = C.CDLL(stub)
    S.stub_prefault()
    case, nodes_min, device_ops = sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3]))
    w = World(cases.materialize(case), np_xyz=(1, 1, 1), nstep=4)
    w.build(0, sum_shared=False)
    v = w.view(0)
    assert v.Nn >= nodes_min, v.Nn
    rng = np.random.default_rng(7)
    v.dispArr[...] = rng.standard_normal(v.dispArr.shape)
    v.velArr[...] = rng.standard_normal(v.velArr.shape)
    v.v1[...] = rng.standard_normal(v.v1.shape)
    fps = []
    for rep in range(2):
        d = dev.Domain(v, compute_ops=device_ops)
        fp = (C.c_uint64 * 3)()
        S.stub_fingerprint(fp)
        fps.append((fp[0], fp[1], fp[2]))
        disp = d.fetch(dev.F_DISP, (3, v.Nn))
        vel = d.fetch(dev.F_VEL, (3, v.Nn))
        v1 = d.fetch(dev.F_V1, (v.Neq,))
        d.close()
    assert fps[0] == fps[1], fps                      # the uploads are deterministic
    st = v.eqNumStartIndexLoc
    fixed = v.eqNumIndexArr[st] < 0
    free3 = (~fixed) & (v.numOfDofPerNodeArr == 3)
    pml = (~fixed) & (v.numOfDofPerNodeArr == 12)
    # displacement: back as uploaded; fixed boundary nodes are held at rest
    assert np.array_equal(disp[:, ~fixed], v.dispArr[:, ~fixed])
    assert not disp[:, fixed].any() and not vel[:, fixed].any()
    # velocity of a 3-dof node IS its v1 (driver.f90:102-103); a PML node keeps velArr next to its 12 split dofs
    eq3 = v.eqNumIndexArr[st[free3][None, :] + np.arange(3)[:, None]] - 1
    assert np.array_equal(vel[:, free3], v.v1[eq3])
    assert np.array_equal(vel[:, pml], v.velArr[:, pml])
    # v1 round trip through the SoA rows / the 12-row PML block
    assert np.array_equal(v1, v.v1)
    print("OK", case, v.Nn, int(free3.sum()), int(pml.sum()), int(fixed.sum()), "%%016x" %% fps[0][2])
""")


def _run(case, nodes_min, arena_gb, device_ops=0):
    env = dict(os.environ, EQD_STUB_ARENA_GB=str(arena_gb))
    env.pop("EQD_VERBOSE", None)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, case, str(nodes_min), str(device_ops)], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.strip().splitlines()[-1].startswith("OK " + case), r.stdout
    return r.stdout


def test_upload_and_fetch_round_trip_small():
    """tpv8 (260 k nodes): the plain download path of eqd_fetch."""
    _run("test.tpv8", 200000, 1.5)


def test_upload_and_fetch_round_trip_staged():
    """TPV104 at dx = 200 m (5.4 M nodes, 130 MB per field): eqd_fetch's staged path -- pinned
    chunks, three SoA rows per chunk, interleaved by all host threads."""
    _run("bench.tpv104_200m", 3000000, 10, device_ops=1)


RUN_SCRIPT = textwrap.dedent(r"""
    import os, sys
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tools", "hoststub"))
    sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import setup_probe
    lib, stub = setup_probe.build_stub()
    from eqdyna_b200 import build, device as dev
    build.cuda_lib_path = lambda: lib
    import parity
    out = []
    for case, npx, opts in (("test.tpv8", (2, 2, 1), {"box": 2, "box_compact": 1}),
                            ("test.tpv10", (2, 2, 2), {"box": 2}),
                            ("test.tpv36", (1, 1, 1), {"box": 1, "overlap": 0})):
        w = parity.build_world(case, npx, 6)
        if case == "test.tpv36":
            # options that shape the upload go in before eqd_set_mesh: bank-aware element order inside the tiles
            for bo in (1, 2):
                d0 = dev.Domain(w.view(0), options={"bank_order": bo})
                d0.set_option("box", 2)
                d0.run(1, 3)
                assert d0.counts()["launches"] > 0 and d0.box_counts()["regular"] > 0
                d0.close()
        doms = parity.run_gpu(w, options=opts)          # kernels are no-ops here: launch sequence, halo plumbing, fetches
        nbox = sum(d.box_counts()["regular"] for d in doms)
        nreg = sum(d.counts()["regular"] for d in doms)
        npml_box = sum(d.box_counts()["pml"] for d in doms)
        nelem_box = sum(dev.box_check(w.view(r))[0] for r in range(w.size))
        assert all(d.counts()["launches"] > 0 for d in doms)
        assert nbox + npml_box <= nelem_box                # a tile is flagged only if ALL its elements are boxes
        out.append((case, nbox, nreg, npml_box))
        for d in doms:
            d.close()
        w.close()
    # marching class: planner, slot table with the bundles' partial slots, node list, launch sequence with chunked runs
    for case, npx, cops in (("test.tpv104", (2, 2, 1), True), ("test.tpv36", (2, 2, 2), False)):
        w = parity.build_world(case, npx, 6)
        doms = parity.run_gpu(w, options={"box": 2, "box_compact": 1}, pre_options={"march": 2}, chunks=2, compute_ops=cops)
        mc = [d.march_counts() for d in doms]
        assert all(m["elements"] > 0 and m["fused_nodes"] > 0.4 * m["elements"] and m["grid"] == 3 * 148 for m in mc), mc
        assert all(d.box_counts()["regular"] >= m["elements"] for d, m in zip(doms, mc))
        if case == "test.tpv104":
            assert all(m["elements"] == d.counts()["regular"] for d, m in zip(doms, mc))      # rectilinear: all of the regular class
        for d in doms:
            d.close()
        w.close()
    assert out[0][1] == out[0][2] and out[0][3] > 0        # tpv8: rectilinear mesh, every tile of both classes
    assert 0 < out[1][1] < out[1][2]                       # tpv10: box and warped tiles mixed
    assert 0 < out[2][1] < out[2][2] and out[2][3] == 0    # tpv36: wedges excluded; box = 1 leaves the PML class alone
    print("OK", out)
""")


def test_launch_sequence_and_box_flags_under_the_stand_in():
    """eqd_run / eqd_run_group host logic (launch order, in-process halo copies, option handling,
    box-tile flags incl. the rank-face-first permutation) runs to completion; the number of
    elements in flagged tiles is consistent with the exact geometric test."""
    env = dict(os.environ, EQD_STUB_ARENA_GB="4")
    r = subprocess.run([sys.executable, "-c", RUN_SCRIPT % {"root": ROOT}], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.strip().splitlines()[-1].startswith("OK"), r.stdout


BENCH_SCRIPT = textwrap.dedent(r"""
    import os, sys
    sys.path.insert(0, %(root)r)
    import torch
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *_: None
    torch.cuda.synchronize = lambda *_: None
    from eqdyna_b200 import build
    build.cuda_lib_path = lambda: %(lib)r
    import bench
    bench.ClockSampler.run = lambda self: None            # no nvidia-smi here
    sys.argv = ["bench.py", "--case", "test.tpv104", "--steps", "5", "--warmup", "3", "--no-cpu-baseline"]
    bench.main()
""")


def test_bench_native_arm_call_sequence_under_the_stand_in():
    """bench.py's native arm end to end against the real library's host code (kernels no-ops, event
    times made up): every C-ABI call it makes, in its order, and the JSON line's keys.  The numbers
    in the line mean nothing here; the point is that no call fails and no key is missing."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tools", "hoststub"))
    import setup_probe
    lib, stub = setup_probe.build_stub()
    env = dict(os.environ, EQD_STUB_ARENA_GB="3", LD_PRELOAD=stub)   # torch brings the real libcudart into the global scope
    env.pop("EQD_VERBOSE", None)
    r = subprocess.run([sys.executable, "-c", BENCH_SCRIPT % {"root": ROOT, "lib": lib}], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in line, k
    assert line["metric"] == "element-steps/s" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert "workload" in line["config"] and line["config"]["case"] == "test.tpv104"
    tn = line["tuning"]
    assert tn["box"] == 2 and tn["box_compact"] == 1 and tn["march"] == 1 and tn["march_elements"] == 516096
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    rf = line["roofline"]
    assert set(rf) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"} and rf["bound"] == "hbm"
    kernels = {r["kernel"]: r for r in line["rooflines"]}
    assert set(kernels) >= {"k_march", "k_march_pml", "k_node_update"} and "k_tile_reg" not in kernels   # every regular element marches
    assert kernels["k_march"]["units_per_launch"] == 516096
    assert line["gpu_launches"] > 0
