#!/usr/bin/env python3
"""Build the coefficient blobs the oracle and the CUDA library load at init.

Run in the authoring container (needs /root/reference); the outputs are committed because the
GPU box has no copy of the reference:

  mima_b200/data/rrtmg_sw_kg.bin        SW k-distribution data, 16 g-points per band, parsed from
                                        SW/src/rrtmg_sw_k_g.f90 (subroutines sw_kgb16..29), plus the
                                        reference atmosphere of SW/src/rrtmg_sw_setcoef.f90:289-343.
  mima_b200/data/rrtmg_lw_ref.bin       LW data that IS present in the reference: pref/preflog/tref,
                                        chi_mls (LW/src/rrtmg_lw_setcoef.f90:418-576) and the Planck
                                        tables totplnk/totplk16 (+derivatives) (:586-1990).
  mima_b200/data/rrtmg_lw_kg_synth.bin  LW k-distribution data with the exact shapes declared in
                                        LW/modules/rrlw_kg01..16.f90 but SYNTHETIC values: the real
                                        file LW/src/rrtmg_lw_k_g.f90 is stripped from the reference
                                        checkout (.MISSING_LARGE_BLOBS).  Smooth, positive, seeded.

Blob format (little endian): 8-byte magic "RRTMGTB1", uint32 n, then n records
{char name[32]; uint32 ndim; uint32 dims[4]; uint64 offset_in_doubles}, then the doubles.
Every array is stored in Fortran (column-major) element order with the dimensions the Fortran
module declares, g-point dimension still 16 wide ("o" = original arrays).  The 16 -> ngc reduction
(cmbgbNN) is done at init by the consumer, not here.
"""
import os
import re
import struct
import sys

import numpy as np

REF = os.environ.get("MIMA_REFERENCE", "/root/reference")
RR = os.path.join(REF, "src/atmos_param/rrtm_radiation")
LW = os.path.join(RR, "rrtmg_lw/gcm_model")
SW = os.path.join(RR, "rrtmg_sw/gcm_model")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mima_b200", "data")

MAGIC = b"RRTMGTB1"


def write_blob(path, arrays):
    """arrays: dict name -> np.ndarray (any order); stored column-major."""
    names = list(arrays)
    recs = []
    off = 0
    payload = []
    for n in names:
        a = np.asarray(arrays[n], dtype=np.float64)
        dims = list(a.shape) + [1] * (4 - a.ndim)
        assert a.ndim <= 4 and len(n) < 32, n
        recs.append(struct.pack("<32sI4IQ", n.encode(), a.ndim, *dims, off))
        flat = np.asfortranarray(a).ravel(order="F")
        payload.append(flat.tobytes())
        off += flat.size
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(names)))
        for r in recs:
            f.write(r)
        for p in payload:
            f.write(p)
    print(f"wrote {path}: {len(names)} arrays, {off} doubles, {os.path.getsize(path)} bytes")


def read_blob(path):
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:8] == MAGIC
    (n,) = struct.unpack_from("<I", raw, 8)
    pos = 12
    recs = []
    for _ in range(n):
        name, ndim, d0, d1, d2, d3, off = struct.unpack_from("<32sI4IQ", raw, pos)
        pos += struct.calcsize("<32sI4IQ")
        recs.append((name.rstrip(b"\0").decode(), ndim, (d0, d1, d2, d3)[:ndim], off))
    data = np.frombuffer(raw, dtype="<f8", offset=pos)
    out = {}
    for name, ndim, dims, off in recs:
        cnt = int(np.prod(dims))
        out[name] = data[off:off + cnt].reshape(dims, order="F")
    return out


# ----------------------------------------------------------------------------------------------
# Minimal Fortran reader: module array declarations and "name(slice) = (/ ... /)" assignments.
# ----------------------------------------------------------------------------------------------

def strip_comment(line):
    i = line.find("!")
    return line if i < 0 else line[:i]


def parse_decls(path, params):
    """Return {name: (shape tuple, lower-bounds tuple)} for real arrays declared in a module."""
    txt = [strip_comment(l) for l in open(path)]
    # join continuation lines
    joined = []
    cur = ""
    for l in txt:
        s = l.rstrip()
        if not s.strip():
            continue
        if cur:
            s2 = s.strip()
            if s2.startswith("&"):
                s2 = s2[1:]
            cur += s2
        else:
            cur = s
        if cur.rstrip().endswith("&"):
            cur = cur.rstrip()[:-1]
            continue
        joined.append(cur)
        cur = ""
    for l in joined:
        m = re.match(r"\s*integer\(kind=im\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(\d+)", l)
        if m:
            params[m.group(1)] = int(m.group(2))
    decls = {}

    def dimlist(s):
        shape, lows = [], []
        for d in s.split(","):
            d = d.strip()
            if ":" in d:
                lo, hi = d.split(":")
                lo, hi = int(params.get(lo.strip(), lo)), int(params.get(hi.strip(), hi))
            else:
                lo, hi = 1, int(params.get(d, d))
            shape.append(hi - lo + 1)
            lows.append(lo)
        return tuple(shape), tuple(lows)

    for l in joined:
        m = re.match(r"\s*real\(kind=rb\)\s*,\s*dimension\(([^)]*)\)\s*::\s*(.*)", l)
        if m:
            sh = dimlist(m.group(1))
            for n in m.group(2).split(","):
                decls[n.strip()] = sh
            continue
        m = re.match(r"\s*real\(kind=rb\)\s*::\s*(.*)", l)
        if m:
            for n, d in re.findall(r"(\w+)\s*\(([^)]*)\)", m.group(1)):
                decls[n] = dimlist(d)
            # scalars
            rest = re.sub(r"\w+\s*\([^)]*\)", "", m.group(1))
            for n in rest.split(","):
                if n.strip():
                    decls[n.strip()] = ((), ())
    return decls


NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][-+]?\d+)?")


def parse_numbers(s):
    s = s.replace("_rb", "")
    vals = []
    for tok in NUM.findall(s):
        vals.append(float(tok.replace("d", "e").replace("D", "e")))
    return vals


def parse_assignments(lines, decls):
    """Execute 'name(slice) = (/ ... /)' and 'name = scalar' statements on numpy arrays."""
    arrays = {}
    for n, (shape, lows) in decls.items():
        arrays[n] = np.full(shape, np.nan, order="F") if shape else np.nan
    i = 0
    nl = len(lines)
    while i < nl:
        l = strip_comment(lines[i]).rstrip()
        m = re.match(r"\s*(\w+)\s*(\(([^)]*)\))?\s*=\s*(.*)$", l)
        if not m or m.group(1) not in decls:
            i += 1
            continue
        name, sl, rhs = m.group(1), m.group(3), m.group(4)
        # gather continuation
        stmt = rhs
        while stmt.rstrip().endswith("&"):
            i += 1
            nxt = strip_comment(lines[i]).strip()
            if not nxt:                      # comment line inside a continued statement
                continue
            if nxt.startswith("&"):
                nxt = nxt[1:]
            stmt = stmt.rstrip()[:-1] + " " + nxt
        i += 1
        shape, lows = decls[name]
        if "(/" in stmt:
            body = stmt[stmt.index("(/") + 2: stmt.rindex("/)")]
            vals = parse_numbers(body)
        else:
            vals = parse_numbers(stmt)
            if not vals:
                continue
        if not shape:
            arrays[name] = vals[0]
            continue
        idx = []
        for k, d in enumerate(sl.split(",")):
            d = d.strip()
            if d == ":":
                idx.append(slice(None))
            elif ":" in d:
                lo, hi = d.split(":")
                idx.append(slice(int(lo) - lows[k], int(hi) - lows[k] + 1))
            else:
                idx.append(int(d) - lows[k])
        tgt = arrays[name][tuple(idx)]
        v = np.array(vals, dtype=np.float64)
        assert tgt.size == v.size, (name, sl, tgt.shape, v.size)
        arrays[name][tuple(idx)] = v.reshape(tgt.shape, order="F")
    return arrays


def split_subroutines(path, pat):
    out = {}
    cur = None
    for l in open(path):
        m = re.match(r"\s*subroutine\s+" + pat, l)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if re.match(r"\s*end subroutine", l):
            cur = None
            continue
        if cur is not None:
            out[cur].append(l)
    return out


# ----------------------------------------------------------------------------------------------
def build_sw():
    params = {}
    for l in open(os.path.join(SW, "modules/parrrsw.f90")):
        m = re.match(r"\s*integer\(kind=im\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(\d+)", strip_comment(l))
        if m:
            params[m.group(1)] = int(m.group(2))
    subs = split_subroutines(os.path.join(SW, "src/rrtmg_sw_k_g.f90"), r"sw_kgb(\d+)")
    out = {}
    for b in range(16, 30):
        decls = parse_decls(os.path.join(SW, f"modules/rrsw_kg{b}.f90"), dict(params))
        # original (16-g) arrays end in 'o'; scalar rayl is shared
        want = {n: d for n, d in decls.items()
                if (n.endswith("o") and d[0] and d[0][-1] == 16) or (n.endswith("o") and d[0] and d[0][0] == 16)
                or (n == "rayl" and not d[0])}
        arrs = parse_assignments(subs[str(b)], want)
        for n, a in arrs.items():
            if isinstance(a, float):
                assert not np.isnan(a), (b, n)
                out[f"sw{b}.{n}"] = np.array([a])
            else:
                assert not np.isnan(a).any(), (b, n, int(np.isnan(a).sum()))
                out[f"sw{b}.{n}"] = a
    # reference atmosphere (SW copy)
    subs = split_subroutines(os.path.join(SW, "src/rrtmg_sw_setcoef.f90"), r"(swatmref)")
    decls = {"pref": ((59,), (1,)), "preflog": ((59,), (1,)), "tref": ((59,), (1,))}
    arrs = parse_assignments(subs["swatmref"], decls)
    for n, a in arrs.items():
        assert not np.isnan(a).any()
        out[f"swref.{n}"] = a
    # ECMWF aerosol optical properties per band and aerosol type (iaer = 6): rsrtaua/rsrpiza/rsrasya(nbndsw, naerec),
    # written row by row as `name(ib, :) = (/ six values /)` in swaerpr (rrtmg_sw_init.f90:370-470)
    subs = split_subroutines(os.path.join(SW, "src/rrtmg_sw_init.f90"), r"(swaerpr)")
    text = " ".join(strip_comment(l).replace("&", " ") for l in subs["swaerpr"])
    for name in ("rsrtaua", "rsrpiza", "rsrasya"):
        a = np.full((14, 6), np.nan, order="F")
        for m in re.finditer(name + r"\(\s*(\d+)\s*,\s*:\s*\)\s*=\s*\(/(.*?)/\)", text, re.S):
            a[int(m.group(1)) - 1, :] = parse_numbers(m.group(2))
        assert not np.isnan(a).any(), name
        out[f"swaer.{name}"] = a
    # cloud optical properties of cldprop_sw's parameterisations (swcldpr, rrtmg_sw_init.f90:1519-3341)
    subs = split_subroutines(os.path.join(SW, "src/rrtmg_sw_init.f90"), r"(swcldpr)")
    decls = {n: ((58, 14), (1, 16)) for n in ("extliq1", "ssaliq1", "asyliq1")}
    decls.update({n: ((43, 14), (1, 16)) for n in ("extice2", "ssaice2", "asyice2")})
    decls.update({n: ((46, 14), (1, 16)) for n in ("extice3", "ssaice3", "asyice3", "fdlice3")})
    decls.update({n: ((5,), (1,)) for n in ("abari", "bbari", "cbari", "dbari", "ebari", "fbari")})
    arrs = parse_assignments(subs["swcldpr"], decls)
    assert set(arrs) == set(decls), sorted(set(decls) - set(arrs))
    for n, a in arrs.items():
        assert not np.isnan(a).any(), n
        out[f"swcld.{n}"] = a
    return out


def build_lw_ref():
    out = {}
    subs = split_subroutines(os.path.join(LW, "src/rrtmg_lw_setcoef.f90"),
                             r"(lwatmref|lwavplank|lwavplankderiv)\b")
    decls = {"pref": ((59,), (1,)), "preflog": ((59,), (1,)), "tref": ((59,), (1,)),
             "chi_mls": ((7, 59), (1, 1))}
    arrs = parse_assignments(subs["lwatmref"], decls)
    for n, a in arrs.items():
        assert not np.isnan(a).any(), n
        out[f"lwref.{n}"] = a
    decls = {"totplnk": ((181, 16), (1, 1)), "totplk16": ((181,), (1,))}
    arrs = parse_assignments(subs["lwavplank"], decls)
    for n, a in arrs.items():
        assert not np.isnan(a).any(), n
        out[f"lwref.{n}"] = a
    decls = {"totplnkderiv": ((181, 16), (1, 1)), "totplk16deriv": ((181,), (1,))}
    arrs = parse_assignments(subs["lwavplankderiv"], decls)
    for n, a in arrs.items():
        assert not np.isnan(a).any(), n
        out[f"lwref.{n}"] = a
    # cloud absorption coefficients of cldprop's parameterisations (lwcldpr, rrtmg_lw_init.f90:2018-2656)
    subs = split_subroutines(os.path.join(LW, "src/rrtmg_lw_init.f90"), r"(lwcldpr)")
    decls = {"absice0": ((2,), (1,)), "absice2": ((43, 16), (1, 1)), "absice3": ((46, 16), (1, 1)), "absliq1": ((58, 16), (1, 1))}
    arrs = parse_assignments(subs["lwcldpr"], decls)
    for n, a in arrs.items():
        assert not np.isnan(a).any(), n
        out[f"lwcld.{n}"] = a
    text = " ".join(strip_comment(l).replace("&", " ") for l in subs["lwcldpr"])
    a1 = np.full((2, 5), np.nan, order="F")
    for m in re.finditer(r"absice1\(\s*(\d+)\s*,\s*:\s*\)\s*=\s*\(/(.*?)/\)", text, re.S):
        a1[int(m.group(1)) - 1, :] = parse_numbers(m.group(2))
    assert not np.isnan(a1).any()
    out["lwcld.absice1"] = a1
    for name in ("abscld1", "absliq0"):
        m = re.search(name + r"\s*=\s*([-+0-9.eEdD]+)_rb", text)
        out[f"lwcld.{name}"] = np.array([float(m.group(1).replace("d", "e").replace("D", "e"))])
    return out


def build_lw_synth(seed=20240917):
    """Synthetic LW k-distribution with the declared shapes (rrlw_kgNN.f90).

    Design: absorption rises ~5 decades across the 16 g-points (as real k-distributions do), varies
    smoothly with the binary-species index, temperature index and reference pressure level, and is
    perturbed by 5 % seeded log-normal noise so that no two table rows are proportional.  Planck
    fractions are positive and sum to one over the 16 g-points for every eta column.
    """
    rng = np.random.default_rng(seed)
    params = {}
    for l in open(os.path.join(LW, "modules/parrrtm.f90")):
        m = re.match(r"\s*integer\(kind=im\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(\d+)", strip_comment(l))
        if m:
            params[m.group(1)] = int(m.group(2))
    out = {}
    g = np.arange(16)
    gshape = 10.0 ** (-4.0 + 5.2 * (g / 15.0) ** 1.5)          # 1e-4 .. ~16
    # per-band strength of the key species (arbitrary but fixed)
    kscale = {1: 3.0, 2: 1.0, 3: 0.6, 4: 2.0, 5: 0.8, 6: 0.02, 7: 0.3, 8: 0.05, 9: 0.4, 10: 1.5,
              11: 2.5, 12: 0.2, 13: 0.05, 14: 5.0, 15: 0.5, 16: 0.1}
    for b in range(1, 17):
        decls = parse_decls(os.path.join(LW, f"modules/rrlw_kg{b:02d}.f90"), dict(params))
        for n, (shape, lows) in decls.items():
            original = n.startswith(("kao", "kbo")) or n in (
                "selfrefo", "forrefo", "fracrefao", "fracrefbo", "ccl4o", "cfc11adjo", "cfc12o", "cfc22adjo")
            if not original or not shape:
                continue
            assert 16 in (shape[-1], shape[0]), (b, n, shape)
            if n in ("fracrefao", "fracrefbo"):
                # g first: (16) or (16, neta)
                neta = shape[1] if len(shape) == 2 else 1
                base = np.exp(-0.5 * ((g[:, None] - (5.0 + 0.6 * np.arange(neta)[None, :] + 0.3 * b)) / 4.5) ** 2)
                base = base * (1.0 + 0.05 * rng.standard_normal(base.shape)).clip(0.5)
                base /= base.sum(axis=0, keepdims=True)
                a = base.reshape(shape, order="F")
            elif n in ("kao", "kbo"):
                # (neta?, 5, npress, 16)
                a = np.empty(shape, order="F")
                npress = shape[-2]
                plev = np.arange(npress) + (0 if n == "kao" else 12)
                pfac = np.exp(-0.045 * plev)                       # weaker lines aloft
                tfac = 1.0 + 0.12 * (np.arange(5) - 2)              # temperature dependence
                core = tfac[:, None, None] * pfac[None, :, None] * gshape[None, None, :]
                if len(shape) == 4:
                    neta = shape[0]
                    eta = np.linspace(0.0, 1.0, neta)
                    efac = 0.35 + 0.65 * eta ** 0.7 + 0.25 * (1 - eta) ** 2
                    core = efac[:, None, None, None] * core[None]
                a[...] = 8.0 * kscale[b] * core * np.exp(0.05 * rng.standard_normal(shape))
            elif n == "selfrefo":
                tf = np.exp(-0.06 * np.arange(10))
                a = (2.0e-2 * kscale[b] * tf[:, None] * gshape[None, :] ** 0.8
                     * np.exp(0.05 * rng.standard_normal(shape)))
            elif n == "forrefo":
                tf = 1.0 + 0.1 * np.arange(shape[0])
                a = (2.0e-4 * kscale[b] * tf[:, None] * gshape[None, :] ** 0.6
                     * np.exp(0.05 * rng.standard_normal(shape)))
            elif n.startswith("kao_m") or n.startswith("kbo_m"):
                # minor species: (19,16) or (neta,19,16)
                tf = 1.0 + 0.03 * np.arange(19)
                core = tf[:, None] * gshape[None, :] ** 0.9
                if len(shape) == 3:
                    eta = np.linspace(0.0, 1.0, shape[0])
                    core = (0.5 + eta ** 1.3)[:, None, None] * core[None]
                minor_scale = {"mn2": 2e-7, "mn2o": 3.0, "mo3": 0.5, "mco2": 5e-3, "mco": 2.0, "mo2": 2e-7}
                key = n.split("_")[1]
                a = minor_scale[key] * core * np.exp(0.05 * rng.standard_normal(shape))
            else:
                # halocarbon cross sections (ccl4o, cfc11adjo, cfc12o, cfc22adjo): (16,)
                a = 40.0 * (1.0 + 0.4 * np.sin(0.7 * g + b)) * np.exp(0.05 * rng.standard_normal(shape))
            assert a.shape == shape and (a > 0).all(), (b, n)
            out[f"lw{b:02d}.{n}"] = np.asfortranarray(a)
    out["lwmeta.synthetic"] = np.array([1.0])      # what rrtmg_b200_tables_info() reports: these are not AER coefficients
    return out


LW_ORIGINALS = ("selfrefo", "forrefo", "fracrefao", "fracrefbo", "ccl4o", "cfc11adjo", "cfc12o", "cfc22adjo")


def lw_decls(b, params):
    """The unreduced (16 g-point) arrays of LW band b as LW/modules/rrlw_kgNN.f90 declares them."""
    decls = parse_decls(os.path.join(LW, f"modules/rrlw_kg{b:02d}.f90"), dict(params))
    return {n: d for n, d in decls.items() if d[0] and (n.startswith(("kao", "kbo")) or n in LW_ORIGINALS)}


def lw_params():
    params = {}
    for l in open(os.path.join(LW, "modules/parrrtm.f90")):
        m = re.match(r"\s*integer\(kind=im\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(\d+)", strip_comment(l))
        if m:
            params[m.group(1)] = int(m.group(2))
    return params


def build_lw_real(kg_path):
    """Parse the real LW k-distribution file LW/src/rrtmg_lw_k_g.f90 (AER RRTMG_LW v4.85; subroutines lw_kgb01..16,
    the same `name(:, j, k, ig) = (/ ... /)` literal format as rrtmg_sw_k_g.f90) into the lwNN.* arrays of the blob.
    The file is stripped from the reference checkout (.MISSING_LARGE_BLOBS): this runs wherever it is available
    (MIMA_LW_KG=/path/to/rrtmg_lw_k_g.f90) and is exercised here by a round trip through write_kg_fortran()."""
    params = lw_params()
    subs = split_subroutines(kg_path, r"lw_kgb(\d+)")
    out = {}
    for b in range(1, 17):
        want = lw_decls(b, params)
        arrs = parse_assignments(subs[f"{b:02d}"], want)
        for n, a in arrs.items():
            assert not np.isnan(a).any(), (b, n, int(np.isnan(a).sum()))
            out[f"lw{b:02d}.{n}"] = a
    return out


def write_kg_fortran(path, arrays, prefix="lw", digits=17):
    """Write arrays {"lwNN.name": ndarray} as Fortran source in the k_g literal format (one assignment per leading
    column, continuation lines of five values, `_rb` kind suffix, subscripts with the declared lower bounds, e.g.
    kbo(:, jt, 13:59, ig)).  Test helper for build_lw_real()."""
    params = lw_params()
    lows_of = {}
    for b in range(1, 17):
        for n, (shape, lows) in lw_decls(b, params).items():
            lows_of[f"lw{b:02d}.{n}"] = lows
    bands = sorted({k.split(".")[0] for k in arrays if re.match(r"^[a-z]+\d\d\.", k)})
    with open(path, "w") as f:
        for bk in bands:
            f.write(f"      subroutine {prefix}_kgb{bk[len(prefix):]}\n      use parkind, only : im => kind_im, rb => kind_rb\n"
                    f"      use rr{prefix}_kg{bk[len(prefix):]}\n      implicit none\n      save\n\n")
            for key in sorted(k for k in arrays if k.startswith(bk + ".")):
                name = key.split(".")[1]
                a = np.asfortranarray(arrays[key])
                lead = a.shape[0]
                rest = a.shape[1:]
                for idx in np.ndindex(*rest[::-1]) if rest else [()]:
                    idx = idx[::-1]
                    col = a[(slice(None),) + idx]
                    lows = lows_of.get(key, (1,) * a.ndim)
                    sl = ",".join([":"] + [str(i + lows[d + 1]) for d, i in enumerate(idx)])
                    f.write(f"      {name}({sl}) = (/ &\n")
                    vals = [f"{v:.{digits}e}_rb" for v in col]
                    for i in range(0, lead, 5):
                        tail = ", &" if i + 5 < lead else " /)"
                        f.write("        & " + ",".join(vals[i:i + 5]) + tail + "\n")
            f.write("\n      end subroutine " + f"{prefix}_kgb{bk[len(prefix):]}\n\n")


# ---- rrtmg_lw.nc: the netCDF form of the same coefficients (LW/src/rrtmg_lw_read_nc.f90, LW/modules/rrlw_ncpar.f90)
# Eight variables hold all bands; every module array is one hyperslab of one variable, read with start = 1 everywhere
# except (absorber,) band, g-point set 1, and count = the array's own extents padded with ones on the left
# (read_nc.f90:30-110 for band 1, the other bands alike).  Dimension order below is the Fortran one (first fastest);
# the file stores the reverse.
NC_DIMS = {"keylower": 9, "keyupper": 5, "Tdiff": 5, "plower": 13, "pupper": 47, "Tself": 10, "Tforeign": 4, "T": 19,
           "band": 16, "GPoint": 16, "GPointSet": 2, "Absorber": 12}
NC_ABSORBERS = ["N2", "CCL4", "CFC11", "CFC12", "CFC22", "H2O", "CO2", "O3", "N2O", "CO", "CH4", "O2"]      # rrlw_ncpar.f90
NC_VARS = {
    "PlanckFractionLowerAtmos": ("GPoint", "keylower", "band", "GPointSet"),
    "PlanckFractionUpperAtmos": ("GPoint", "keyupper", "band", "GPointSet"),
    "KeySpeciesAbsorptionCoefficientsLowerAtmos": ("keylower", "Tdiff", "plower", "GPoint", "band", "GPointSet"),
    "KeySpeciesAbsorptionCoefficientsUpperAtmos": ("keyupper", "Tdiff", "pupper", "GPoint", "band", "GPointSet"),
    "H20SelfAbsorptionCoefficients": ("Tself", "GPoint", "band", "GPointSet"),
    "H20ForeignAbsorptionCoefficients": ("Tforeign", "GPoint", "band", "GPointSet"),
    "AbsorptionCoefficientsLowerAtmos": ("keylower", "T", "GPoint", "Absorber", "band", "GPointSet"),
    "AbsorptionCoefficientsUpperAtmos": ("keyupper", "T", "GPoint", "Absorber", "band", "GPointSet"),
}
NC_MINOR = {"mn2": "N2", "mn2o": "N2O", "mo3": "O3", "mco2": "CO2", "mco": "CO", "mo2": "O2"}
NC_XSEC = {"ccl4o": "CCL4", "cfc11adjo": "CFC11", "cfc12o": "CFC12", "cfc22adjo": "CFC22"}


def nc_slab(name, shape, band):
    """(variable, index tuple in Fortran dimension order) of module array `name` (extents `shape`) of LW band `band`."""
    b, gs = band - 1, 0
    if name in ("fracrefao", "fracrefbo"):
        var = "PlanckFractionLowerAtmos" if name == "fracrefao" else "PlanckFractionUpperAtmos"
        k = slice(0, shape[1]) if len(shape) == 2 else 0
        return var, (slice(0, shape[0]), k, b, gs)
    if name in ("kao", "kbo"):
        var = "KeySpeciesAbsorptionCoefficients" + ("LowerAtmos" if name == "kao" else "UpperAtmos")
        lead = (slice(0, shape[0]),) if len(shape) == 4 else (0,)
        t, pr, g = shape[-3:]
        return var, lead + (slice(0, t), slice(0, pr), slice(0, g), b, gs)
    if name == "selfrefo":
        return "H20SelfAbsorptionCoefficients", (slice(0, shape[0]), slice(0, shape[1]), b, gs)
    if name == "forrefo":
        return "H20ForeignAbsorptionCoefficients", (slice(0, shape[0]), slice(0, shape[1]), b, gs)
    if name in NC_XSEC:
        return "AbsorptionCoefficientsLowerAtmos", (0, 0, slice(0, shape[0]), NC_ABSORBERS.index(NC_XSEC[name]), b, gs)
    m = re.match(r"k([ab])o_(m\w+)$", name)
    if m and m.group(2) in NC_MINOR:
        var = "AbsorptionCoefficients" + ("LowerAtmos" if m.group(1) == "a" else "UpperAtmos")
        lead = (slice(0, shape[0]),) if len(shape) == 3 else (0,)
        t, g = shape[-2:]
        return var, lead + (slice(0, t), slice(0, g), NC_ABSORBERS.index(NC_MINOR[m.group(2)]), b, gs)
    raise KeyError(name)


def lw_template():
    """Names and declared extents of every unreduced LW array, from the synthetic blob (built from rrlw_kgNN.f90)."""
    return {k: v.shape for k, v in read_blob(os.path.join(OUT, "rrtmg_lw_kg_synth.bin")).items() if not k.startswith("lwmeta.")}


def build_lw_from_nc(nc_path, template=None):
    """Read rrtmg_lw.nc (netCDF classic / 64-bit offset, through scipy.io) into the lwNN.* arrays of the blob, as
    lw_kgb01..16 of rrtmg_lw_read_nc.f90 fill the rrlw_kgNN modules.  SURVEY.md section 8f rank 3.  The file is not
    part of the reference checkout; exercised by a round trip through write_lw_nc()."""
    from scipy.io import netcdf_file
    template = template or lw_template()
    try:
        f = netcdf_file(nc_path, "r", mmap=False)
    except (TypeError, ValueError) as e:
        raise SystemExit(f"{nc_path}: not a netCDF classic file ({e}); convert with `nccopy -k classic`")
    out = {}
    try:
        cache = {}
        for key, shape in template.items():
            band, name = int(key[2:4]), key.split(".")[1]
            var, idx = nc_slab(name, shape, band)
            if var not in cache:
                v = f.variables[var]
                want = tuple(NC_DIMS[d] for d in NC_VARS[var])[::-1]
                if tuple(v.shape) != want:
                    raise SystemExit(f"{nc_path}: {var} has shape {tuple(v.shape)}, expected {want}")
                cache[var] = np.array(v[:], dtype=np.float64).T          # Fortran dimension order
            a = np.asfortranarray(cache[var][idx])
            assert a.shape == tuple(shape), (key, a.shape, shape)
            out[key] = a
    finally:
        f.close()
    return out


def write_lw_nc(nc_path, arrays):
    """Write {"lwNN.name": ndarray} in the rrtmg_lw.nc layout (test helper for build_lw_from_nc)."""
    from scipy.io import netcdf_file
    full = {v: np.zeros(tuple(NC_DIMS[d] for d in dims), order="F") for v, dims in NC_VARS.items()}
    for key, a in arrays.items():
        var, idx = nc_slab(key.split(".")[1], a.shape, int(key[2:4]))
        full[var][idx] = a
    f = netcdf_file(nc_path, "w")
    try:
        for d, n in NC_DIMS.items():
            f.createDimension(d, n)
        for v, dims in NC_VARS.items():
            nv = f.createVariable(v, "d", dims[::-1])
            nv[:] = full[v].T
    finally:
        f.close()


def main():
    os.makedirs(OUT, exist_ok=True)
    sw = build_sw()
    write_blob(os.path.join(OUT, "rrtmg_sw_kg.bin"), sw)
    lwref = build_lw_ref()
    write_blob(os.path.join(OUT, "rrtmg_lw_ref.bin"), lwref)
    lw = build_lw_synth()
    write_blob(os.path.join(OUT, "rrtmg_lw_kg_synth.bin"), lw)
    # the real LW coefficients, wherever the file exists (it is stripped from the reference checkout)
    kg = os.environ.get("MIMA_LW_KG", os.path.join(LW, "src/rrtmg_lw_k_g.f90"))
    if os.path.exists(kg):
        write_blob(os.path.join(OUT, "rrtmg_lw_kg.bin"), build_lw_real(kg))
        print("wrote rrtmg_lw_kg.bin from", kg)
    elif os.environ.get("MIMA_LW_NC") and os.path.exists(os.environ["MIMA_LW_NC"]):
        write_blob(os.path.join(OUT, "rrtmg_lw_kg.bin"), build_lw_from_nc(os.environ["MIMA_LW_NC"]))
        print("wrote rrtmg_lw_kg.bin from", os.environ["MIMA_LW_NC"])
    else:
        print("no rrtmg_lw_k_g.f90 at", kg, "-> LW stays on the synthetic blob")
    # round-trip check
    for fn, src in (("rrtmg_sw_kg.bin", sw), ("rrtmg_lw_ref.bin", lwref), ("rrtmg_lw_kg_synth.bin", lw)):
        back = read_blob(os.path.join(OUT, fn))
        for n, a in src.items():
            assert np.array_equal(np.asarray(a).reshape(back[n].shape, order="F"), back[n]), n
    print("round-trip ok")


def emit_lw_kg(path):
    """Write the LW coefficients the oracle and the CUDA library load (the real blob if there is one, else the synthetic
    one) as Fortran source in the layout of the reference's stripped LW/src/rrtmg_lw_k_g.f90 (subroutines lw_kgb01..16
    assigning the rrlw_kgNN module arrays).  oracle/Makefile translates it next to the reference sources, so that the
    translated rrtmg_lw_ini finds the routines it calls (rrtmg_lw_init.f90:80-95)."""
    real = os.path.join(OUT, "rrtmg_lw_kg.bin")
    arrays = read_blob(real if os.path.exists(real) else os.path.join(OUT, "rrtmg_lw_kg_synth.bin"))
    write_kg_fortran(path, arrays)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--emit-lw-kg":
        emit_lw_kg(sys.argv[2])
        sys.exit(0)
    sys.exit(main())
