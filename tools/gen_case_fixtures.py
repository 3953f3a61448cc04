#!/usr/bin/env python3
"""Generate case-input fixtures by running the UNMODIFIED reference case workflow.

For every listed case this script does, in a scratch directory, what
`scripts/create.newcase` + `./case.setup` do in the reference
(/root/reference/scripts/create.newcase:12-31, scripts/case.setup:319-328):
copy `case_input/<case>/user_defined_params.py` next to the reference's
`defaultParameters.py`, `lib.py`, `case.setup`, `generateFaultInterface`, then
execute `case.setup` and (for insertFaultType>0) `generateFaultInterface`.

The container has neither netCDF4 nor matplotlib, so both modules are replaced
by recording stubs:  the netCDF4 stub captures every variable written to
`on_fault_vars_input.nc` (case.setup:78-185) and we dump those 24 fields to
`on_fault_vars_input.bin` (raw float64; layout documented in
eqdyna_b200/csrc/host/eqh_io.cpp).  The five `b*.txt` files and
`bFault_Rough_Geometry.txt` are written by the reference code itself and are
therefore byte-identical to what the reference workflow produces.

Outputs go to tests/golden/cases/<case>/.  Only runs where /root/reference
exists (this container); the fixtures are committed so the GPU box needs nothing.
"""
import os
import runpy
import shutil
import struct
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("EQDYNAROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_ROOT = os.path.join(os.path.dirname(HERE), "tests", "golden", "cases")

# order of the 24 fields = order of var_id(1..24) in src/netcdf_io.f90:41-64
NC_NAMES = [
    "sw_fs", "sw_fd", "sw_D0", "rsf_a", "rsf_b", "rsf_Dc", "rsf_v0", "rsf_r0",
    "rsf_fw", "rsf_vw", "tp_a_hy", "tp_a_th", "tp_rouc", "tp_lambda", "tp_h",
    "tp_Tini", "tp_pini", "init_slip_rate", "init_strike_shear",
    "init_normal_stress", "init_state", "tw_t0", "cohesion", "init_dip_shear",
]

CASES = [
    "test.tpv8", "test.tpv10", "test.tpv104", "test.tpv36", "test.drv.a6",
    "test.tpv1053d", "test.meng2023a", "test.meng2023cb", "test.tpv37",
]


class _Var:
    def __init__(self, store, name, shape):
        self._store, self._name = store, name
        store[name] = np.zeros(shape, dtype=np.float64)

    def __setitem__(self, key, val):
        self._store[self._name][key] = val

    def __setattr__(self, k, v):
        if k.startswith("_"):
            object.__setattr__(self, k, v)


class _Dataset:
    captured = {}

    def __init__(self, fname, mode="r", format=None):
        self.dims = {}
        _Dataset.captured = {}

    def createDimension(self, name, n):
        self.dims[name] = n
        return n

    def createVariable(self, name, dtype, dims):
        shape = tuple(self.dims[d] for d in dims)
        return _Var(_Dataset.captured, name, shape)

    def close(self):
        pass


def _install_stubs():
    nc = types.ModuleType("netCDF4")
    nc.Dataset = _Dataset
    sys.modules["netCDF4"] = nc
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")

    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    for fn in ("figure", "contourf", "gca", "colorbar", "title", "savefig",
               "rc", "contour", "subplot", "close"):
        setattr(plt, fn, _Any())
    mpl.pyplot = plt
    anim = types.ModuleType("matplotlib.animation")
    mpl.animation = anim
    mpl.rc = _Any()
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["matplotlib.animation"] = anim


def gen_case(case, out_root=OUT_ROOT, subst=None, out_name=None, gz=False):
    """Run the reference workflow for `case`.  `subst` is an optional list of
    (old, new) text replacements applied to the scratch copy of
    user_defined_params.py (used for the benchmark-resolution variants, e.g.
    TPV104 at dx = 100 m: README.md:87 / misc/model_path_initial.m:5-8);
    `out_name` names the fixture directory; gz compresses the on-fault dump."""
    _install_stubs()
    work = tempfile.mkdtemp(prefix="eqd_case_")
    try:
        for f in os.listdir(os.path.join(REF, "case_input", case)):
            shutil.copy(os.path.join(REF, "case_input", case, f), work)
        for f in ("defaultParameters.py", "lib.py", "case.setup",
                  "generateFaultInterface"):
            shutil.copy(os.path.join(REF, "scripts", f), work)
        if subst:
            up = os.path.join(work, "user_defined_params.py")
            txt = open(up).read()
            for old, new in subst:
                if old not in txt:
                    raise RuntimeError("substitution source %r not found in %s" % (old, case))
                txt = txt.replace(old, new)
            open(up, "w").write(txt)
        cwd = os.getcwd()
        os.chdir(work)
        sys.path.insert(0, work)
        for m in ("user_defined_params", "defaultParameters", "lib"):
            sys.modules.pop(m, None)
        try:
            # case.setup -> create_model_input_file() etc.  os.system('./generateFaultInterface')
            # inside it fails harmlessly (no interpreter deps); we run it below instead.
            g = runpy.run_path(os.path.join(work, "case.setup"), run_name="case_setup")
            g["create_model_input_file"]()
            g["create_station_input_file"]()
            g["netcdf_write_on_fault_vars"]()
            par = g["par"]
            fields = dict(_Dataset.captured)
            if par.insertFaultType > 0:
                runpy.run_path(os.path.join(work, "generateFaultInterface"),
                               run_name="__main__")
        finally:
            os.chdir(cwd)
            sys.path.remove(work)
        out = os.path.join(out_root, out_name or case)
        os.makedirs(out, exist_ok=True)
        names = ["bGlobal.txt", "bModelGeometry.txt", "bFaultGeometry.txt",
                 "bMaterial.txt", "bStations.txt"]
        if par.insertFaultType > 0:
            names.append("bFault_Rough_Geometry.txt")
        for n in names:
            shutil.copy(os.path.join(work, n), os.path.join(out, n))
        nfz, nfx = fields["sw_fs"].shape
        import gzip
        opener = (lambda p: gzip.open(p + ".gz", "wb", compresslevel=9)) if gz else (lambda p: open(p, "wb"))
        with opener(os.path.join(out, "on_fault_vars_input.bin")) as f:
            f.write(b"EQDOFV1\0")
            f.write(struct.pack("<iii", nfx, nfz, len(NC_NAMES)))
            f.write(struct.pack("<i", 0))
            for n in NC_NAMES:
                a = np.ascontiguousarray(fields[n], dtype="<f8")
                assert a.shape == (nfz, nfx)
                f.write(a.tobytes())
        return par
    finally:
        shutil.rmtree(work, ignore_errors=True)


# benchmark-resolution variants (BASELINE.json configs)
VARIANTS = {
    "bench.tpv104_100m": ("test.tpv104", [("par.term = 5.", "par.term = 15."), ("par.dx = 500.", "par.dx = 100."),
                                          ("par.dt = 0.5*par.dx/par.vp", "par.dt = 0.008")]),
    "bench.tpv104_200m": ("test.tpv104", [("par.term = 5.", "par.term = 15."), ("par.dx = 500.", "par.dx = 200."),
                                          ("par.dt = 0.5*par.dx/par.vp", "par.dt = 0.016")]),
    # TPV36 (15-degree thrust, wedges) at finer resolutions than the shipped 500 m (README.md:5,86 quotes 50 m on 512 cores)
    "bench.tpv36_100m": ("test.tpv36", [("par.dx   = 500.", "par.dx   = 100.")]),
    "bench.tpv36_200m": ("test.tpv36", [("par.dx   = 500.", "par.dx   = 200.")]),
}

if __name__ == "__main__":
    cases = sys.argv[1:] or CASES
    for c in list(cases):
        if c in VARIANTS:
            base, subst = VARIANTS[c]
            par = gen_case(base, subst=subst, out_name=c, gz=True)
            print("fixture written:", c, "nfx,nfz =", par.nfx, par.nfz, "dx =", par.dx, "dt =", par.dt, "term =", par.term)
            cases.remove(c)
    for c in cases:
        par = gen_case(c)
        print("fixture written:", c, "nfx,nfz =", par.nfx, par.nfz,
              "np =", par.nx, par.ny, par.nz)
