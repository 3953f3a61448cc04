"""mode == 2 (earthquake-cycle hand-off): the host stand-in reads the 12
restart fields of fault.r.nc (netcdf_io.f90:116-185; here the raw dump fault.r.bin) into
fric(7,8,49,47,20,23,31:36) and init_vel (eqdyna3d.f90:191-212) seeds v1 of the split-node pairs
with them.  No GPU: the step library sees nothing but different initial fric / v1."""
import os
import shutil
import struct

import numpy as np
import pytest

import parity

FIELDS = ("shear_strike", "shear_dip", "effective_normal", "slip_rate", "state_variable", "state_normal",
          "vxm", "vym", "vzm", "vxs", "vys", "vzs")
SLOT = (8, 49, 7, 47, 20, 23, 31, 32, 33, 34, 35, 36)


def _restart_case(tmp_path, case="test.tpv104"):
    from eqdyna_b200 import cases
    src = cases.materialize(case)
    dst = str(tmp_path / "case")
    shutil.copytree(src, dst)
    lines = open(os.path.join(dst, "bGlobal.txt")).read().splitlines()
    assert lines[0].split()[0] == "1"
    lines[0] = "2"                                        # mode (readInputFiles.f90:31)
    open(os.path.join(dst, "bGlobal.txt"), "w").write("\n".join(lines) + "\n")
    with open(os.path.join(dst, "on_fault_vars_input.bin"), "rb") as f:
        f.read(8)
        fnx, fnz, nvar, _ = struct.unpack("<4i", f.read(16))
    assert nvar == 24
    ix, iz = np.meshgrid(np.arange(fnx), np.arange(fnz), indexing="ij")
    # field v at grid point (ix, iz): a value that identifies all three
    vals = np.stack([(v + 1) * 1.0e3 + ix * 1.0 + iz * 1.0e-3 for v in range(12)])   # (12,fnx,fnz)
    vals[6:] *= 1.0e-9                                    # velocities: small, so that a few steps stay finite
    with open(os.path.join(dst, "fault.r.bin"), "wb") as f:
        f.write(b"EQDOFV1\0")
        f.write(struct.pack("<4i", fnx, fnz, 12, 0))
        f.write(np.ascontiguousarray(vals.transpose(0, 2, 1)).astype("<f8").tobytes())   # [var][iz][ix]
    return dst, vals


def test_restart_fields_reach_fric_and_v1(tmp_path):
    from eqdyna_b200.host import World
    d, vals = _restart_case(tmp_path)
    w = World(d, np_xyz=(2, 2, 1), nstep=3).build()
    w1 = parity.build_world("test.tpv104", (2, 2, 1), 3)                 # mode 1 twin
    seen = 0
    for r in range(w.size):
        v, u = w.view(r), w1.view(r)
        k = int(v.nftnd[0])
        if not k:
            continue
        seen += k
        p = v.params
        slave = v.nsmp[0, :k, 0] - 1
        master = v.nsmp[1, :k, 0] - 1
        xs, zs = v.meshCoor[0, slave], v.meshCoor[2, slave]
        # grid indices as netcdf_io.f90:155-156 computes them (500 m fault grid of the fixture)
        dx = p.dx
        # the fault of test.tpv104 spans x in [-18 km, 18 km], z in [-18 km, 0]: recover (ii, jj) from the
        # restart value itself instead of re-deriving fxmin / fzmin
        got = v.fric[np.array(SLOT) - 1, :k, 0]                            # (12,k)
        ii = np.rint(got[0] - 1.0e3).astype(int)                           # shear_strike = 1000 + ix + iz/1000
        jj = np.rint((got[0] - 1.0e3 - ii) * 1.0e3).astype(int)
        assert ii.min() >= 0 and jj.min() >= 0
        np.testing.assert_allclose(xs - xs[ii.argmin()], (ii - ii.min()) * dx, atol=1e-6)
        np.testing.assert_allclose(zs - zs[jj.argmin()], (jj - jj.min()) * dx, atol=1e-6)
        for f in range(12):
            np.testing.assert_array_equal(got[f], vals[f][ii, jj])
        # slots the restart does not touch keep the values of the regular on-fault file
        for slot in (1, 2, 3, 9, 10, 11, 12, 13, 46):
            np.testing.assert_array_equal(v.fric[slot - 1, :k, 0], u.fric[slot - 1, :k, 0])
        # init_vel: v1 of the slave / master dofs = fric(34:36) / fric(31:33)
        st = v.eqNumStartIndexLoc
        for j in range(3):
            np.testing.assert_array_equal(v.v1[v.eqNumIndexArr[st[slave] + j] - 1], v.fric[33 + j, :k, 0])
            np.testing.assert_array_equal(v.v1[v.eqNumIndexArr[st[master] + j] - 1], v.fric[30 + j, :k, 0])
    assert seen == 2701 + 37          # 73 x 37 pairs; the column on the shared x face belongs to both ranks
    w.close(); w1.close()


def test_restart_with_the_regular_fields_reproduces_mode_1(tmp_path):
    """A restart file that carries the same tractions / slip rate / state as on_fault_vars_input
    (and zero split-node velocities) must give the mode-1 run bit for bit: the restart only
    overrides initial fric / v1, the loop is the same."""
    from eqdyna_b200.host import World
    d, _ = _restart_case(tmp_path)
    with open(os.path.join(d, "on_fault_vars_input.bin"), "rb") as f:
        f.read(8)
        fnx, fnz, nvar, _ = struct.unpack("<4i", f.read(16))
        ofv = np.frombuffer(f.read(), dtype="<f8").reshape(nvar, fnz, fnx)             # [var][iz][ix], var_id order of netcdf_io.f90:41-64
    zero = np.zeros((fnz, fnx))
    rst = np.stack([ofv[18], ofv[23], ofv[19], ofv[17], ofv[20], np.abs(ofv[19])] + [zero] * 6)
    with open(os.path.join(d, "fault.r.bin"), "wb") as f:
        f.write(b"EQDOFV1\0")
        f.write(struct.pack("<4i", fnx, fnz, 12, 0))
        f.write(np.ascontiguousarray(rst).astype("<f8").tobytes())
    n = 25
    w2 = World(d, np_xyz=(2, 2, 1), nstep=n).build()
    w1 = parity.build_world("test.tpv104", (2, 2, 1), n)
    for r in range(w1.size):
        assert np.array_equal(w1.view(r).fric, w2.view(r).fric) and np.array_equal(w1.view(r).v1, w2.view(r).v1)
    parity.run_oracle(w1)
    parity.run_oracle(w2)
    moved = False
    for r in range(w1.size):
        a, b = w1.view(r), w2.view(r)
        for name in ("dispArr", "velArr", "fric", "fnft", "onFaultQuantHistSCECForm"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        moved = moved or float(np.abs(b.velArr).max()) > 0
    assert moved
    w1.close(); w2.close()


def test_mode2_without_restart_file_is_an_error(tmp_path):
    from eqdyna_b200.host import World
    d, _ = _restart_case(tmp_path)
    os.remove(os.path.join(d, "fault.r.bin"))
    with pytest.raises(RuntimeError) as e:
        World(d, np_xyz=(1, 1, 1), nstep=2)
    assert "fault.r.bin" in str(e.value)


def test_cycle_hand_off_frt_to_restart(tmp_path):
    """One EQdyna run hands its final on-fault state to the next one: frt.txt<me> ->
    tools/restart_from_frt.py (the mapping of scripts/plotRuptureDynamics:71-82) -> fault.r.bin ->
    a mode == 2 world whose fric / v1 start from the first run's final tractions, slip rate,
    state and split-node velocities (to the 7 digits frt.txt carries)."""
    import subprocess
    import sys
    from eqdyna_b200.host import World
    n = 60
    w1 = parity.build_world("test.tpv104", (2, 2, 1), n)
    parity.run_oracle(w1)
    run_dir = str(tmp_path / "run1")
    for r in range(w1.size):
        w1.write_outputs(r, run_dir)
    d, _ = _restart_case(tmp_path)
    os.remove(os.path.join(d, "fault.r.bin"))
    tool = os.path.join(parity.ROOT, "tools", "restart_from_frt.py")
    r = subprocess.run([sys.executable, tool, run_dir, d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "73 x 37" in r.stdout
    w2 = World(d, np_xyz=(2, 2, 1), nstep=2).build()
    checked = 0
    for rk in range(w1.size):
        a, b = w1.view(rk), w2.view(rk)
        k = int(a.nftnd[0])
        if not k:
            continue
        # final state of run 1 (fric slots as frt.txt prints them) == initial state of run 2
        for src, dst in ((79, 8), (80, 49), (78, 7), (47, 47), (20, 20), (23, 23),
                         (31, 31), (32, 32), (33, 33), (34, 34), (35, 35), (36, 36)):
            x, y = a.fric[src - 1, :k, 0], b.fric[dst - 1, :k, 0]
            np.testing.assert_allclose(y, x, rtol=6e-7, atol=1e-30)
        st = b.eqNumStartIndexLoc
        slave = b.nsmp[0, :k, 0] - 1
        np.testing.assert_array_equal(b.v1[b.eqNumIndexArr[st[slave]] - 1], b.fric[33, :k, 0])
        checked += k
        assert float(np.abs(a.fric[46, :k, 0]).max()) > 1e-3          # the rupture was under way: not a trivial state
    assert checked > 2701
    w1.close(); w2.close()


def test_nc_to_bin_with_a_stand_in_for_netcdf4(tmp_path, monkeypatch):
    """tools/nc_to_bin.py converts the reference workflow's netCDF inputs where netCDF4 exists; here a
    minimal stand-in for the module serves the arrays, and the output must equal the fixture that
    tools/gen_case_fixtures.py captured from scripts/case.setup."""
    import importlib.util
    import sys
    import types
    from eqdyna_b200 import cases
    src = cases.materialize("test.tpv8")
    with open(os.path.join(src, "on_fault_vars_input.bin"), "rb") as f:
        raw = f.read()
    fnx, fnz, nvar, _ = struct.unpack("<4i", raw[8:24])
    ofv = np.frombuffer(raw[24:], dtype="<f8").reshape(nvar, fnz, fnx)
    spec = importlib.util.spec_from_file_location("nc_to_bin", os.path.join(parity.ROOT, "tools", "nc_to_bin.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class FakeDataset:
        def __init__(self, path):
            assert path.endswith("on_fault_vars_input.nc")
            self.variables = {n: ofv[i] for i, n in enumerate(mod.ON_FAULT)}
    monkeypatch.setitem(sys.modules, "netCDF4", types.SimpleNamespace(Dataset=FakeDataset))
    d = str(tmp_path)
    mod.main(["nc_to_bin.py", d])
    assert open(os.path.join(d, "on_fault_vars_input.bin"), "rb").read() == raw
    assert len(mod.RESTART) == 12 and mod.RESTART == list(FIELDS)


@pytest.mark.parametrize("case", ["test.tpv104", "test.meng2023a"])
def test_restart_tool_reproduces_the_reference_restart_file(case, tmp_path):
    """Golden vector: the reference ships, next to its golden frt.txt*, the fault.dyna.r.nc that
    scripts/plotRuptureDynamics (generateNcRestart) derived from them.  The file is NETCDF4 (HDF5, no
    reader in this image) but its twelve (dip, strike) float64 arrays are stored contiguously, in the
    order netcdf_io.f90:139-150 reads them: tools/restart_from_frt.py, fed the golden frt.txt*, must
    produce exactly those bytes."""
    import gzip
    import importlib.util
    import golden_io
    from eqdyna_b200 import cases
    spec = importlib.util.spec_from_file_location("restart_from_frt", os.path.join(parity.ROOT, "tools", "restart_from_frt.py"))
    rf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rf)
    run = str(tmp_path / "golden_run")
    os.makedirs(run)
    for f in ("frt.txt0", "frt.txt2"):
        open(os.path.join(run, f), "w").write(golden_io.read_text(golden_io.golden_path(case, f)))
    fields, fnx, fnz = rf.gather(run, cases.materialize(case))
    mine = np.ascontiguousarray(fields).astype("<f8").tobytes()           # [12][fnz][fnx]
    raw = gzip.open(golden_io.golden_path(case, "fault.dyna.r.nc"), "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    first = raw.find(mine[:fnx * fnz * 8])
    assert first > 0, "shear_strike of the tool is not in the reference's file"
    assert raw[first:first + len(mine)] == mine, "the twelve restart arrays differ from the reference's"
    assert np.abs(fields[0]).max() > 1.0e6                                # tractions in Pa, not a trivial block of zeros
