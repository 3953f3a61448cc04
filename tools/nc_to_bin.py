#!/usr/bin/env python3
"""on_fault_vars_input.nc / fault.r.nc -> the raw containers the stand-in host reads.

For users of the reference's own case workflow (scripts/case.setup:78-185 writes
on_fault_vars_input.nc with netCDF4; EQquasi writes fault.r.nc): run this where netCDF4 is installed.

  python tools/nc_to_bin.py <case_dir>

Writes <case_dir>/on_fault_vars_input.bin (24 fields, order of var_id(1..24) in
src/netcdf_io.f90:41-64) and, if fault.r.nc exists, <case_dir>/fault.r.bin (12 fields, order of
src/netcdf_io.f90:139-150).  Layout: char[8] "EQDOFV1\\0", int32 nfx, nfz, nvar, 0, float64 [nvar][nfz][nfx]."""
import os
import struct
import sys

import numpy as np

ON_FAULT = ["sw_fs", "sw_fd", "sw_D0", "rsf_a", "rsf_b", "rsf_Dc", "rsf_v0", "rsf_r0", "rsf_fw", "rsf_vw", "tp_a_hy", "tp_a_th",
            "tp_rouc", "tp_lambda", "tp_h", "tp_Tini", "tp_pini", "init_slip_rate", "init_strike_shear", "init_normal_stress",
            "init_state", "tw_t0", "cohesion", "init_dip_shear"]
RESTART = ["shear_strike", "shear_dip", "effective_normal", "slip_rate", "state_variable", "state_normal",
           "vxm", "vym", "vzm", "vxs", "vys", "vzs"]


def convert(nc_path, names, out_path):
    import netCDF4
    ds = netCDF4.Dataset(nc_path)
    fields = [np.asarray(ds.variables[n][:], dtype=np.float64) for n in names]      # each (dip, strike) = (nfz, nfx)
    nfz, nfx = fields[0].shape
    with open(out_path, "wb") as f:
        f.write(b"EQDOFV1\0")
        f.write(struct.pack("<4i", nfx, nfz, len(names), 0))
        f.write(np.ascontiguousarray(np.stack(fields)).astype("<f8").tobytes())
    return nfx, nfz


def main(argv):
    if len(argv) != 2:
        raise SystemExit(__doc__)
    d = argv[1]
    print("on_fault_vars_input.bin: %d x %d" % convert(os.path.join(d, "on_fault_vars_input.nc"), ON_FAULT,
                                                       os.path.join(d, "on_fault_vars_input.bin")))
    if os.path.exists(os.path.join(d, "fault.r.nc")):
        print("fault.r.bin: %d x %d" % convert(os.path.join(d, "fault.r.nc"), RESTART, os.path.join(d, "fault.r.bin")))


if __name__ == "__main__":
    main(sys.argv)
