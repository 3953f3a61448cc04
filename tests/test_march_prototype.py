"""Prototype of the marching kernel planned for the box tiles (tools/proto_march).  The
kernel is written once, as barrier-separated phases (march_kernel.cuh); here g++ compiles it and a
driver runs every phase over all thread ids in turn, so the source the GPU will execute is checked on
the CPU: partial forces and updated stresses equal the element-by-element evaluation from the
reference's stored operators.  nvcc compiles the same header for sm_100a without spills.
Prototype only -- not a product path, no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity


@pytest.fixture(scope="module")
def proto(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("march") / "libmarch_proto.so")
    src = os.path.join(parity.ROOT, "tools", "proto_march", "march_proto.cpp")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", out, src], stderr=subprocess.DEVNULL)
    L = C.CDLL(out)
    L.march_proto.argtypes = [C.c_int32] * 5 + [C.c_void_p] * 11 + [C.c_double, C.c_double] + [C.c_void_p] * 3
    return L


@pytest.mark.parametrize("prefetch", [0, 1, 2], ids=["sync-planes", "prefetched-planes", "prefetched-planes-and-operators"])
@pytest.mark.parametrize("case,min_bundles", [("test.tpv104", 150), ("test.tpv8", 40)])
def test_marching_forces_equal_element_by_element_forces(proto, case, min_bundles, prefetch, monkeypatch):
    if prefetch:
        monkeypatch.setenv("MARCH_PREFETCH", str(prefetch))   # the other schedules of march_kernel.cuh
    else:
        monkeypatch.delenv("MARCH_PREFETCH", raising=False)
    w = parity.build_world(case, (1, 1, 1), 2)
    v = w.view(0)
    r = v.raw
    rng = np.random.default_rng(3)
    vel = np.asfortranarray(rng.standard_normal((3, v.Nn)))
    disp = np.asfortranarray(rng.standard_normal((3, v.Nn)) * 1e-2)
    fm = np.zeros((3, v.Nn), order="F")
    fr = np.zeros((3, v.Nn), order="F")
    st = np.zeros(5, dtype=np.int64)
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    rc = proto.march_proto(v.Nn, v.Ne, r.nx, r.ny, r.nz, P(v.meshCoor), P(v.nodeElemIdRelation), P(v.elemTypeArr),
                           P(v.numOfDofPerNodeArr), P(v.eleshp), P(v.phi), P(v.ss), P(v.eledet), P(v.mat), P(vel), P(disp),
                           v.params.rdampk, v.params.w, P(fm), P(fr), P(st))
    assert rc == 0, "numbering assumption violated at march_proto.cpp:%d" % rc
    assert st[0] >= min_bundles and st[1] % (4 * 16) == 0
    assert st[1] >= 0.7 * st[4]                            # share of the mesh's regular box elements inside bundles
    assert np.abs(fr).max() > 0
    assert np.abs(fm - fr).max() <= 1e-12 * np.abs(fr).max()
    assert st[3] * 1e-18 <= 1e-12                          # updated stresses, relative to the largest
    # nodes outside every bundle receive nothing from either path
    assert np.array_equal(np.abs(fm).sum(axis=0) > 0, np.abs(fr).sum(axis=0) > 0)
    w.close()


def test_prototype_kernel_compiles_for_sm_100a_without_spills(tmp_path):
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    src = os.path.join(parity.ROOT, "tools", "proto_march", "march_kernel.cu")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
                        "-c", src, "-o", str(tmp_path / "march_kernel.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "k_march_reg" in r.stdout and "k_march_reg_pf" in r.stdout and "k_march_reg_pf2" in r.stdout
    assert r.stdout.count("0 bytes spill stores, 0 bytes spill loads") == 3
    # the stand-alone microbenchmark / GPU-vs-host check built on the same header links
    bench = os.path.join(parity.ROOT, "tools", "proto_march", "march_bench.cu")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                        "-o", str(tmp_path / "march_bench"), bench], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
