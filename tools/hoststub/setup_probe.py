#!/usr/bin/env python3
"""Profile / fingerprint the HOST side of the set-up path without a GPU.

  python tools/hoststub/setup_probe.py [case] [--device-ops 0|1] [--box N] [--repeat R]

Links the step library's objects (eqdyna_b200/lib/obj/*.o) against tools/hoststub/cudart_stub.cpp
into a scratch directory (kernels do nothing, "device" memory is host memory), runs the calls
bench.py's e2e leg makes before eqd_run -- eqd_create, eqd_set_mesh, eqd_compute_elem_ops |
eqd_set_elem_ops, eqd_set_nodal, eqd_set_fault, eqd_set_halo, eqd_set_stations, finalize (through
eqd_sum_shared's entry, which needs no communicator at 1x1x1) -- with EQD_VERBOSE=1 lap timers,
and prints a fingerprint of every buffer the host uploaded.  Two builds whose fingerprints agree
upload byte-identical data.  Test / profiling infrastructure only: nothing under eqdyna_b200/ knows
about it, and no number a kernel would compute exists in this mode."""
import argparse
import ctypes as C
import glob
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = None   # per-process scratch directory, removed at exit


def build_stub():
    global OUT
    from eqdyna_b200 import build
    build.build_cuda()
    if OUT is None:
        import atexit
        import shutil
        import tempfile
        OUT = tempfile.mkdtemp(prefix="eqd_hoststub_")
        atexit.register(shutil.rmtree, OUT, ignore_errors=True)
    stub = os.path.join(OUT, "libcudart_stub.so")
    lib = os.path.join(OUT, "libeqdyna_b200_hoststub.so")
    src = os.path.join(ROOT, "tools", "hoststub", "cudart_stub.cpp")
    objs = sorted(glob.glob(os.path.join(ROOT, "eqdyna_b200", "lib", "obj", "*.o")))
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-o", stub, src])
    subprocess.check_call(["g++", "-shared", "-o", lib] + objs + [stub, "-ldl", "-lpthread", "-Wl,-rpath," + OUT])
    return lib, stub


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", nargs="?", default="bench.tpv104_200m")
    ap.add_argument("--device-ops", type=int, default=1)
    ap.add_argument("--box", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--np", default="1x1x1", help="decomposition; the probe sets up sub-domain --rank of it (no exchange: eqd_sum_shared is skipped)")
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--march", type=int, default=1)
    args = ap.parse_args()
    decomp = tuple(int(x) for x in args.np.split("x"))
    lib, stub = build_stub()
    os.environ["EQD_VERBOSE"] = "1"
    from eqdyna_b200 import build, cases, device as dev
    from eqdyna_b200.host import World
    build.cuda_lib_path = lambda: lib          # this process only
    S = C.CDLL(stub)
    S.stub_prefault()                          # "device" arena faulted in before anything is timed
    t0 = time.time()
    w = World(cases.materialize(args.case), np_xyz=decomp, nstep=20)
    w.build(args.rank, sum_shared=False)
    v = w.view(args.rank)
    print("[probe] host state built in %.2f s: %d elements, %d nodes" % (time.time() - t0, v.Ne, v.Nn), flush=True)
    for rep in range(args.repeat):
        t0 = time.perf_counter()
        d = dev.Domain(v, compute_ops=bool(args.device_ops), options={"march": args.march})
        t1 = time.perf_counter()
        d.set_option("box", args.box)
        d.set_option("box_compact", 1)
        if decomp == (1, 1, 1):
            d.sum_shared()                     # runs finalize (no neighbours: no communicator needed)
        t2 = time.perf_counter()
        fp = (C.c_uint64 * 3)()
        S.stub_fingerprint(fp)
        print("[probe] rep %d: eqd_create + eqd_set_* %.3f s, finalize %.3f s; uploaded %d buffers, %.3f GB, fingerprint %016x"
              % (rep, t1 - t0, t2 - t1, fp[0], fp[1] / 1e9, fp[2]), flush=True)
        d.close()
    w.close()


if __name__ == "__main__":
    main()
