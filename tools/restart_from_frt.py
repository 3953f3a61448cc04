#!/usr/bin/env python3
"""frt.txt<me> of a finished run -> fault.r.bin, the restart input of a mode == 2 run.

The reference hands an earthquake cycle on through netCDF: scripts/plotRuptureDynamics:29-82 gathers
12 fields from the frt.txt* files on the (strike, dip) fault grid and generateNcRestart (:201-262)
writes them to fault.dyna.r.nc; EQquasi returns fault.r.nc, which netcdf_read_on_fault_eqdyna_restart
(src/netcdf_io.f90:116-185) loads.  This tool does the first half without netCDF4 and writes the raw
container the stand-in host reads (eqdyna_b200/csrc/host/eqh_io.cpp): same columns, same grid indices
(round((x - fxmin)/dx), round((z - fzmin)/dz)), same 12 fields in the same order.

  python tools/restart_from_frt.py <run_dir with frt.txt*> <case_dir> [out=<case_dir>/fault.r.bin]
"""
import glob
import os
import struct
import sys

import numpy as np

# 0-based columns of frt.txt (library_output.f90:160-205) for shear_strike, shear_dip, effective_normal,
# slip_rate, state_variable, state_normal, vxm, vym, vzm, vxs, vys, vzs (plotRuptureDynamics:71-82)
COLS = (12, 13, 11, 10, 20, 21, 14, 15, 16, 17, 18, 19)


def fault_grid(case_dir):
    """fxmin, fzmin, dx, dz, fnx, fnz of fault 1 from bFaultGeometry.txt / bModelGeometry.txt and the
    header of on_fault_vars_input.bin."""
    # readInputFiles.f90:124-131: a title line, then "fxmin fxmax", "fymin fymax", "fzmin fzmax" per fault
    g = [l.split() for l in open(os.path.join(case_dir, "bFaultGeometry.txt")) if l.strip()]
    fxmin, fxmax = float(g[1][0]), float(g[1][1])
    fzmin, fzmax = float(g[3][0]), float(g[3][1])
    with open(os.path.join(case_dir, "on_fault_vars_input.bin"), "rb") as f:
        assert f.read(8)[:7] == b"EQDOFV1"
        fnx, fnz, _, _ = struct.unpack("<4i", f.read(16))
    dx = (fxmax - fxmin) / (fnx - 1)
    dz = (fzmax - fzmin) / (fnz - 1)
    return fxmin, fzmin, dx, dz, fnx, fnz


def gather(run_dir, case_dir):
    fxmin, fzmin, dx, dz, fnx, fnz = fault_grid(case_dir)
    out = np.zeros((12, fnz, fnx))
    seen = np.zeros((fnz, fnx), dtype=bool)
    files = sorted(glob.glob(os.path.join(run_dir, "frt.txt*")))
    if not files:
        raise SystemExit("no frt.txt* in " + run_dir)
    for fn in files:
        a = np.atleast_2d(np.loadtxt(fn))
        ii = np.rint((a[:, 0] - fxmin) / dx).astype(int)
        jj = np.rint((a[:, 2] - fzmin) / dz).astype(int)
        for k, c in enumerate(COLS):
            out[k, jj, ii] = a[:, c]
        seen[jj, ii] = True
    if not seen.all():
        raise SystemExit("frt.txt* do not cover the fault grid (%d of %d nodes)" % (seen.sum(), seen.size))
    return out, fnx, fnz


def write_bin(path, fields, fnx, fnz):
    with open(path, "wb") as f:
        f.write(b"EQDOFV1\0")
        f.write(struct.pack("<4i", fnx, fnz, fields.shape[0], 0))
        f.write(np.ascontiguousarray(fields).astype("<f8").tobytes())     # [var][iz][ix]


def main(argv):
    if len(argv) < 3:
        raise SystemExit(__doc__)
    run_dir, case_dir = argv[1], argv[2]
    out = argv[3] if len(argv) > 3 else os.path.join(case_dir, "fault.r.bin")
    fields, fnx, fnz = gather(run_dir, case_dir)
    write_bin(out, fields, fnx, fnz)
    print("wrote %s: 12 fields on %d x %d fault nodes" % (out, fnx, fnz))


if __name__ == "__main__":
    main(sys.argv)
