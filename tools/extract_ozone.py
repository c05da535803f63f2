#!/usr/bin/env python3
"""Extract one month (January) of the reference's zonal-mean ozone climatology
(input/INPUT/ozone_1990.nc, netCDF-3: ozone_1990(time=12, pfull=59, lat=64, lon=1) kg/kg) into the
small fixture mima_b200/data/ozone_1990_jan.npz used by the synthetic-column generator for
BASELINE.json config 4 (4xCO2 + stratospheric ozone input).  Input synthesis only."""
import os
import numpy as np
from scipy.io import netcdf_file

REF = os.environ.get("MIMA_REFERENCE", "/root/reference")
f = netcdf_file(os.path.join(REF, "input/INPUT/ozone_1990.nc"), "r", mmap=False)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mima_b200", "data", "ozone_1990_jan.npz")
np.savez_compressed(out, lat=f.variables["lat"][:].astype(np.float64),
                    pfull=f.variables["pfull"][:].astype(np.float64),
                    ozone=f.variables["ozone_1990"][0, :, :, 0].astype(np.float64))
print("wrote", out, os.path.getsize(out))
