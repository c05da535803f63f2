#!/usr/bin/env python3
"""Turn the reference's LW known-answer cases into a committed fixture (tests/golden/std_atm.npz).

Source (authoring container only): /root/reference/doc_rrtm/runs_std_atm/input_rrtm_{MLS,MLW,SAW,TROP}-clr and
output_rrtm_*-clr -- AER's standard-atmosphere runs of RRTMG_LW v4.85 (rtrnmr + taumol Rev 1.8 + setcoef Rev 1.6,
the routines MiMA links), clear sky, emissivity 1.  The inputs are TAPE5 records 2.1.1-2.1.2 (layer pressure and
temperature, level pressure/temperature, H2O CO2 O3 N2O CO CH4 O2 as volume mixing ratios or column densities --
normalised here to mixing ratios relative to dry air -- and the dry-air column); the outputs are the broadband (10-3250 cm-1) level fluxes and heating rates, 4-5 printed decimals.

These vectors pin the LW *coefficients*: they can only be reproduced with the real rrtmg_lw_k_g.f90 tables, which
are stripped from the reference checkout (tests/test_std_atm.py skips the comparison until
mima_b200/data/rrtmg_lw_kg.bin exists; tools/build_tables.py builds it from the file).
"""
import os
import sys

import numpy as np

SRC = os.environ.get("STD_ATM_DIR", "/root/reference/doc_rrtm/runs_std_atm")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "std_atm.npz")


def fnum(s):
    s = s.strip()
    return float(s) if s else np.nan


def parse_input(path):
    lines = open(path).read().split("\n")
    i = next(k for k, l in enumerate(lines) if l.startswith("$"))
    tbound = float(lines[i + 2].split()[0])
    hdr = lines[i + 3]
    nlay = int(hdr[2:5])
    nmol = int(hdr[5:10])
    assert nmol == 7, nmol
    k = i + 4
    pavel, tavel, pz, tz, vmr, broad = [], [], [], [], [], []
    for lay in range(nlay):
        l = lines[k].ljust(85)
        pavel.append(float(l[0:15])); tavel.append(float(l[15:25]))
        if lay == 0:
            pz.append(fnum(l[48:56])); tz.append(fnum(l[56:63]))
        pz.append(fnum(l[70:78])); tz.append(fnum(l[78:85]))
        k += 1
        if lines[k].strip().isdigit():          # optional cross-section flag line
            k += 1
        vals = [float(x) for x in lines[k].split()]
        assert len(vals) == 8, (path, lay, lines[k])
        v = np.array(vals[:7])
        if v[0] > 1.0:
            # column densities [molecules/cm2] (RRTM: values > 1): dry column = broadening gases + listed gases except H2O
            coldry = vals[7] + v[1:].sum()
            v = v / coldry
        else:
            # volume mixing ratios relative to dry air
            coldry = vals[7] / (1.0 - v[1:].sum())
        vmr.append(v); broad.append(coldry)
        k += 1
    return dict(tbound=tbound, pavel=np.array(pavel), tavel=np.array(tavel), pz=np.array(pz), tz=np.array(tz),
                vmr=np.array(vmr).T, coldry=np.array(broad))


def parse_output(path, nlay):
    lines = open(path).read().split("\n")
    i = next(k for k, l in enumerate(lines) if "Wavenumbers:   10.0 - 3250.0" in l)
    rows = []
    for l in lines[i + 3: i + 3 + nlay + 1]:
        f = l.split()
        rows.append([int(f[0]), float(f[1]), float(f[2]), float(f[3]), float(f[4]), float(f[5])])
    r = np.array(rows)[::-1]                    # level 0 (surface) first
    assert (r[:, 0] == np.arange(nlay + 1)).all()
    return dict(plev=r[:, 1], uflx=r[:, 2], dflx=r[:, 3], fnet=r[:, 4], hr=r[:, 5])


def main():
    out = {}
    for case in ("MLS", "MLW", "SAW", "TROP"):
        a = parse_input(os.path.join(SRC, f"input_rrtm_{case}-clr"))
        b = parse_output(os.path.join(SRC, f"output_rrtm_{case}-clr"), len(a["pavel"]))
        for k, v in {**a, **b}.items():
            out[f"{case}.{k}"] = np.asarray(v)
        print(case, len(a["pavel"]), "layers; surface up", b["uflx"][0], "TOA up", b["uflx"][-1])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
