"""Host-side logic and the C-ABI surface (no GPU): mesh numbering facts, case fixtures,
exported symbols, error behaviour without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parity

ROOT = parity.ROOT


def header_functions(path, prefix):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, txt)))


def test_cuda_library_exports_every_declared_symbol():
    from eqdyna_b200 import build, device
    path = build.cuda_lib_path()
    assert os.path.exists(path), "libeqdyna_b200.so not built (run __graft_entry__.build())"
    lib = C.CDLL(path)
    names = header_functions(os.path.join(ROOT, "include", "eqdyna_b200.h"), "eqd_")
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(device.EXPORTS) == names


def test_host_library_exports_every_declared_symbol():
    from eqdyna_b200 import build
    lib = C.CDLL(build.build_host())
    for n in header_functions(os.path.join(ROOT, "include", "eqdyna_host.h"), "eqh_"):
        assert hasattr(lib, n), "missing export " + n


def test_oracle_is_not_linked_or_imported_by_the_product():
    """The product path must never route through oracle/ (tier rule 3)."""
    for root, _, files in os.walk(os.path.join(ROOT, "eqdyna_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".f90")):
                txt = open(os.path.join(root, f), errors="replace").read()
                if f == "build.py":
                    continue  # builds the checker, does not use it
                assert "import oracle" not in txt and "liboracle" not in txt and "step_oracle" not in txt, f


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from eqdyna_b200 import device
    w = parity.build_world("test.tpv8", (1, 1, 1), 2)
    with pytest.raises(device.StepError):
        device.Domain(w.view(0))
    w.close()


@pytest.mark.parametrize("case,nodes,elems,pairs", [
    ("test.tpv8", (97, 52, 49), 235008, 1891),
    ("test.tpv10", (119, 91, 70), 732780, 1891),
    ("test.tpv104", (141, 76, 71), 735000, 2701),
])
def test_mesh_sizes_and_numbering(case, nodes, elems, pairs):
    """Grid sizes of SURVEY.md section 6 and the numbering facts of section 8(a)."""
    w = parity.build_world(case, (1, 1, 1), 2)
    v = w.view(0)
    nx, ny, nz = v.raw.nx, v.raw.ny, v.raw.nz
    assert (nx, ny, nz) == nodes
    assert v.Ne == elems and int(v.nftnd[0]) == pairs
    assert v.Nn == nx * ny * nz + pairs                      # masters appended after the regular grid
    # node id (ix-1)*ny*nz + (iz-1)*ny + iy: y fastest, then z, then x
    X = v.meshCoor
    assert X[1, 1] > X[1, 0] and X[0, 1] == X[0, 0] and X[2, 1] == X[2, 0]
    assert X[2, ny] > X[2, 0] and X[0, ny * nz] > X[0, 0]
    # masters share their slave's coordinates and are 3-dof
    s, m = v.nsmp[0, :pairs, 0] - 1, v.nsmp[1, :pairs, 0] - 1
    assert np.array_equal(m, nx * ny * nz + np.arange(pairs))
    assert np.array_equal(X[:, s], X[:, m])
    assert np.all(v.numOfDofPerNodeArr[m] == 3)
    # first element: the brick at (ix,iy,iz) = (2,2,2)
    nid = lambda ix, iy, iz: (ix - 1) * ny * nz + (iz - 1) * ny + iy
    assert list(v.nodeElemIdRelation[:, 0]) == [nid(1, 1, 1), nid(2, 1, 1), nid(2, 2, 1), nid(1, 2, 1),
                                                 nid(1, 1, 2), nid(2, 1, 2), nid(2, 2, 2), nid(1, 2, 2)]
    # fixed model boundary: every dof of the node is -1; PML nodes carry 12 dofs
    st = v.eqNumStartIndexLoc
    fixed = v.eqNumIndexArr[st] < 0
    assert fixed.sum() > 0 and np.all(v.numOfDofPerNodeArr[fixed] == 12)
    assert set(np.unique(v.elemTypeArr)) == {1, 2}
    w.close()


def test_wedge_mesh_tpv36():
    """Degenerate wedges (types 11, 12: nodes 3==4 and 7==8 collapsed) and type 13 bricks."""
    w = parity.build_world("test.tpv36", (1, 1, 1), 2)
    v = w.view(0)
    et = v.elemTypeArr
    assert set(np.unique(et)) == {1, 2, 11, 12, 13}
    wed = np.isin(et, (11, 12))
    c = v.nodeElemIdRelation[:, wed]
    assert np.array_equal(c[2], c[3]) and np.array_equal(c[6], c[7])
    idx = np.nonzero(et == 11)[0]
    assert np.all(et[idx + 1] == 12)                          # a wedge cell = two consecutive elements
    assert np.all(v.eledet > 0)
    w.close()


def test_decomposition_partitions_elements_and_shares_one_node_plane():
    w1 = parity.build_world("test.tpv8", (1, 1, 1), 2)
    w4 = parity.build_world("test.tpv8", (2, 2, 1), 2)
    assert sum(w4.view(r).Ne for r in range(4)) == w1.view(0).Ne
    v0, v2 = w4.view(0), w4.view(2)                           # neighbours along x (me = mex*npy*npz + ...)
    assert v0.meshCoor[0].max() == v2.meshCoor[0].min()
    # both sides exchange the same number of dofs on the shared face
    assert v0.numcount[4] == v2.numcount[3]
    # shared-node masses were summed: a face node's mass equals its mass in the undecomposed mesh
    x_face = v0.meshCoor[0].max()
    n0 = int(np.nonzero((v0.meshCoor[0] == x_face) & (v0.numOfDofPerNodeArr == 3))[0][5])
    xyz = v0.meshCoor[:, n0]
    n1 = int(np.nonzero(np.all(w1.view(0).meshCoor[:, :w1.view(0).raw.nx * w1.view(0).raw.ny * w1.view(0).raw.nz] == xyz[:, None], axis=0))[0][0])
    assert v0.fnms[n0] == pytest.approx(w1.view(0).fnms[n1], rel=1e-14)
    w1.close(); w4.close()


def test_bench_fixture_matches_readme_benchmark():
    """TPV104 at dx = 100 m / dt = 0.008 s / 15 s = 1875 steps (README.md:87)."""
    from eqdyna_b200 import cases
    d = cases.materialize("bench.tpv104_100m")
    g = open(os.path.join(d, "bGlobal.txt")).read().split()
    assert "15.0" in g and "0.008" in g and "104" in g
    assert os.path.getsize(os.path.join(d, "on_fault_vars_input.bin")) == 24 + 24 * 361 * 181 * 8


def test_fortran_interface_covers_the_abi():
    """eqdyna_cuda_iface.f90 binds every function of include/eqdyna_b200.h, and the bind(C)
    struct lists the fields of eqd_params in the header's order."""
    f90 = open(os.path.join(ROOT, "eqdyna_b200", "csrc", "fortran", "eqdyna_cuda_iface.f90")).read()
    bound = sorted(set(re.findall(r"bind\(C,\s*name='(eqd_\w+)'\)", f90)))
    assert bound == header_functions(os.path.join(ROOT, "include", "eqdyna_b200.h"), "eqd_")
    hdr = open(os.path.join(ROOT, "include", "eqdyna_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr[hdr.index("typedef struct eqd_params"):hdr.index("} eqd_params;")], flags=re.S)
    cfields = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct eqd_params {", "").strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if decl.split()[0] in ("double", "int32_t") else ""
        cfields += [re.sub(r"\[.*\]", "", n).strip() for n in names.split(",") if n.strip()]
    tbody = f90[f90.index("type, bind(C) :: eqd_params"):f90.index("end type eqd_params")]
    ffields = []
    for line in tbody.splitlines()[1:]:
        if "::" in line:
            ffields += [re.sub(r"\(.*\)", "", n).strip() for n in line.split("::")[1].split(",") if n.strip()]
    assert [c.lower() for c in cfields] == [f.lower() for f in ffields]
    # the ctypes mirror used by the tests has the same order too
    from eqdyna_b200.host import EqdParams
    assert [n for n, _ in EqdParams._fields_] == cfields


@pytest.mark.parametrize("case,frac", [("test.tpv104", 1.0), ("test.tpv10", 0.5), ("test.tpv36", 0.99), ("test.drv.a6", 0.1)])
def test_box_operators(case, frac):
    """Closed-form operators of axis-aligned hexahedra (eqd_box.h, option "box" of the tile
    kernels) against the host's eleshp / phi / ss: strain, B^T t forces and hourglass forces of
    a pseudo-random field agree to rounding on every element that passes the exact box test."""
    from eqdyna_b200 import device
    w = parity.build_world(case, (1, 1, 1), 2)
    v = w.view(0)
    n, dev = device.box_check(v)
    assert n >= frac * v.Ne and n <= v.Ne
    if case == "test.tpv104":
        assert n == v.Ne                                   # planar vertical fault: the whole mesh is rectilinear
    assert np.all(dev < 1e-12), dev
    # wedges never pass
    wed = np.isin(v.elemTypeArr, (11, 12))
    assert n <= v.Ne - int(wed.sum())
    w.close()


def test_standalone_driver_fails_loudly_without_a_gpu(tmp_path):
    """eqdyna_host (eqdyna_b200/csrc/host/eqdyna_host_main.cpp), the compiled counterpart of
    driver_cuda.f90: builds the sub-domains, then must stop with the CUDA error code -- not
    fall back to any CPU path -- when no device is visible."""
    import subprocess
    import torch
    from eqdyna_b200 import build, cases
    exe = build.build_exe()
    assert exe and os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, cases.materialize("test.tpv8"), "-nstep", "2", "-o", str(tmp_path / "out")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CPU path" in r.stderr
    assert "4 sub-domain(s), 235008 elements" in r.stdout
    assert not any(f.startswith("frt") for f in os.listdir(str(tmp_path / "out")))
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 4 and "usage" in r.stderr


@pytest.mark.parametrize("case", ["test.tpv104", "test.tpv36"])
def test_bank_model_and_bank_aware_element_order(case):
    """Modelled shared-memory wavefronts of the tile kernels' corner accesses (eqd_plan_bank_model): an
    ascending element order already is close to conflict-free on the regular class (16-element y runs);
    the PML bricks (6 x 7 x 6) need about two wavefronts per access, and the bank-aware order
    (eqd_set_option "bank_order") lowers that without breaking any planner invariant."""
    from eqdyna_b200 import device
    w = parity.build_world(case, (1, 1, 1), 2)
    v = w.view(0)
    a = device.plan_bank_model(v, 0)
    b = device.plan_bank_model(v, 1)          # raises if the reordered plan violates an invariant
    for cls in ("reg", "pml"):
        ideal, asc, _ = a[cls]
        assert ideal > 0 and asc >= ideal
        assert b[cls][0] == ideal and b[cls][1] == asc and ideal <= b[cls][2] <= asc
    assert a["reg"][1] < 1.35 * a["reg"][0]
    assert a["pml"][1] > 1.8 * a["pml"][0] and b["pml"][2] < 0.9 * a["pml"][1]
    # bank_order = 2: residue numbering of the tile-local nodes of complete bricks (element order kept):
    # the regular class becomes all but conflict free; PML bricks of 6 x 7 x 6 do not fit the 400-node rows
    # (448 needed) and keep ascending ids, smaller ones are renumbered
    c = device.plan_bank_model(v, 2)
    assert c["reg"][2] < 1.12 * c["reg"][0] and c["reg"][2] < a["reg"][1]
    assert c["pml"][2] < a["pml"][1]
    w.close()
