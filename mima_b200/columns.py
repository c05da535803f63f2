"""Seeded synthetic atmospheric columns in the layout `run_rrtmg` hands to rrtmg_sw / rrtmg_lw.

Follows the generator specified in SURVEY.md section 8(d).  The marshaling conventions are the caller's
(src/atmos_param/rrtm_radiation/rrtm_radiation.f90:652-677): pressures in hPa, level index 1 = surface,
column index = longitude fastest then latitude, H2O as specific humidity, O3 as mass mixing ratio,
other gases as volume mixing ratio, h2o >= 2e-7, 100 <= T <= 370.

All arrays are Fortran-ordered float64 with shape (ncol, nlay) / (ncol, nlay+1) / (ncol,), i.e. the
column index is contiguous exactly as in the Fortran interface.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

RESOLUTIONS = {
    # name: (nlon, nlat, nlay)
    "T42L40": (128, 64, 40),
    "T85L40": (256, 128, 40),
    "T170L60": (512, 256, 60),
    "T341L80": (1024, 512, 80),
}


@dataclasses.dataclass
class Columns:
    ncol: int
    nlay: int
    nlon: int
    nlat: int
    play: np.ndarray
    plev: np.ndarray
    tlay: np.ndarray
    tlev: np.ndarray
    tsfc: np.ndarray
    h2o: np.ndarray       # specific humidity
    o3: np.ndarray        # mass mixing ratio
    co2: np.ndarray       # vmr
    ch4: np.ndarray
    n2o: np.ndarray
    o2: np.ndarray
    cfc11: np.ndarray
    cfc12: np.ndarray
    cfc22: np.ndarray
    ccl4: np.ndarray
    emis: np.ndarray      # (ncol, 16)
    albedo: np.ndarray    # (ncol,)
    coszen: np.ndarray    # (ncol,)
    scon: float = 1370.0
    adjes: float = 1.0
    dyofyr: int = 0

    def rows(self, j0: int, j1: int) -> "Columns":
        """Contiguous block of latitude rows [j0, j1): a pointer offset in column space
        (rrtm_radiation.f90:652 -- longitude fastest)."""
        c0, c1 = j0 * self.nlon, j1 * self.nlon
        kw = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if isinstance(v, np.ndarray):
                kw[f.name] = np.asfortranarray(v[c0:c1])
            else:
                kw[f.name] = v
        kw["ncol"] = c1 - c0
        kw["nlat"] = j1 - j0
        return Columns(**kw)

    def take(self, idx) -> "Columns":
        kw = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            kw[f.name] = np.asfortranarray(v[idx]) if isinstance(v, np.ndarray) else v
        kw["ncol"] = len(kw["tsfc"])
        kw["nlon"] = kw["ncol"]
        kw["nlat"] = 1
        return Columns(**kw)


def sigma_half_levels(nlay: int) -> np.ndarray:
    """MiMA 'uneven_sigma' half levels, TOA first (vert_coordinate.f90:217-242, input.nml:33-36):
    b_k = exp(-7.9*(0.1*zeta + 0.9*zeta**1.4)), zeta = 1-(k-1)/L, b_1 = 0."""
    k = np.arange(1, nlay + 2)
    zeta = 1.0 - (k - 1.0) / nlay
    b = np.exp(-7.9 * (0.1 * zeta + 0.9 * zeta ** 1.4))
    b[0] = 0.0
    return b


def _qsat(T, p_hpa):
    es = 6.112 * np.exp(17.67 * (T - 273.15) / (T - 29.65))      # hPa (Bolton)
    es = np.minimum(es, 0.5 * p_hpa)
    return 0.622 * es / (p_hpa - 0.378 * es)


def _ozone_file_profile(lat_deg, p_hpa):
    """Config C4: zonal-mean ozone (kg/kg) from input/INPUT/ozone_1990.nc -- one month, linear in
    latitude and log-pressure.  The values are read from the small fixture
    mima_b200/data/ozone_1990_jan.npz extracted once by tools/extract_ozone.py."""
    path = os.path.join(os.path.dirname(__file__), "data", "ozone_1990_jan.npz")
    d = np.load(path)
    flat, fp, fo3 = d["lat"], d["pfull"], d["ozone"]       # (nlat), (nlev) hPa increasing, (nlev, nlat)
    lp = np.log(fp)
    out = np.empty(p_hpa.shape)
    li = np.clip(np.searchsorted(flat, lat_deg) - 1, 0, len(flat) - 2)
    wl = np.clip((lat_deg - flat[li]) / (flat[li + 1] - flat[li]), 0.0, 1.0)
    prof = fo3[:, li] * (1 - wl) + fo3[:, li + 1] * wl     # (nlev, ncol)
    lq = np.log(p_hpa)
    for c in range(p_hpa.shape[0]):
        out[c] = np.interp(lq[c], lp, prof[:, c])
    return np.maximum(out, 0.0)


def make_columns(resolution: str = "T42L40", *, seed: int = 20240917, co2_ppmv: float = 390.0,
                 ozone: str = "analytic", night: bool = False, secondary_gases: bool = False,
                 nlon: int | None = None, nlat: int | None = None, nlay: int | None = None,
                 lat_rows: tuple[int, int] | None = None) -> Columns:
    """Build one synthetic batch.

    night=False: cos(zenith) floored at 0.02 so every column is sunlit (the headline workload);
    night=True : roughly half the columns get coszen=0 (realistic instantaneous insolation).
    ozone='file' uses the reference's ozone_1990.nc climatology (config C4).
    lat_rows=(j0, j1) generates only those latitude rows of the full grid (rank-local block) with
    exactly the values the full-grid call would produce for them.
    """
    if resolution in RESOLUTIONS:
        rl, rt, ry = RESOLUTIONS[resolution]
    else:
        rl, rt, ry = 128, 64, 40
    nlon = nlon or rl
    nlat = nlat or rt
    nlay = nlay or ry
    j0, j1 = lat_rows if lat_rows is not None else (0, nlat)
    L = nlay

    lat_all = np.arcsin(np.linspace(-1.0 + 1.0 / nlat, 1.0 - 1.0 / nlat, nlat))
    cols = []
    for j in range(j0, j1):
        # one independent stream per latitude row => row-sharded generation is reproducible
        rng = np.random.default_rng([seed, j])
        phi = lat_all[j]
        s2 = np.sin(phi) ** 2
        n = nlon
        ps = np.clip(1000.0 * (1.0 + 0.03 * rng.standard_normal(n)), 850.0, 1050.0)
        b = sigma_half_levels(L)                                  # TOA -> surface
        ph = ps[:, None] * b[None, :]                             # (n, L+1) TOA first
        # Simmons-Burridge full levels (press_and_geopot.f90:146-173)
        pf = np.empty((n, L))
        lnph = np.log(np.maximum(ph[:, 1:], 1e-300))
        with np.errstate(divide="ignore", invalid="ignore"):
            dp = ph[:, 1:] - ph[:, :-1]
            lnpf = (ph[:, 1:] * lnph - ph[:, :-1] * np.log(np.maximum(ph[:, :-1], 1e-300))) / dp - 1.0
        lnpf[:, 0] = lnph[:, 0] - 1.0
        pf = np.exp(lnpf)
        ph[:, 0] = 0.5 * pf[:, 0]                                 # rrtm_radiation.f90:655-656
        Ts = 300.0 - 45.0 * s2 + 2.0 * rng.standard_normal(n)
        Ttrop = 200.0 + 15.0 * s2
        T = np.maximum(Ts[:, None] * (pf / ps[:, None]) ** 0.19, Ttrop)
        lp = np.log(pf)
        strat = pf < 30.0
        Tstrat = Ttrop + (270.0 - Ttrop) * (np.log(30.0) - lp) / (np.log(30.0) - np.log(1.0))
        above1 = pf < 1.0
        Tmeso = 270.0 - 10.0 * (np.log(1.0) - lp)
        T = np.where(strat, np.where(above1, Tmeso, Tstrat), T)
        T = T + 1.5 * rng.standard_normal((n, L))
        T = np.clip(T, 100.0, 370.0)
        # interface temperatures (interp_temp, rrtm_radiation.f90:422-461)
        Th = np.empty((n, L + 1))
        lph = np.log(ph)
        w = (lph[:, 1:-1] - lp[:, :-1]) / (lp[:, 1:] - lp[:, :-1])
        Th[:, 1:-1] = T[:, :-1] * (1 - w) + T[:, 1:] * w
        Th[:, -1] = Ts
        Th[:, 0] = 0.5 * (3.0 * T[:, 0] - T[:, 1]) if L > 1 else T[:, 0]      # a single layer has nothing to extrapolate from
        Th = np.clip(Th, 100.0, 370.0)
        q = 0.8 * _qsat(T, pf) * (pf / ps[:, None]) ** 3 * np.exp(0.2 * rng.standard_normal((n, L)))
        qstrat = (2.0 + 2.0 * rng.random((n, L))) * 1e-6
        q = np.where(pf < 100.0, qstrat, np.maximum(q, qstrat))
        q = np.maximum(q, 2e-7)
        if ozone == "file":
            o3 = _ozone_file_profile(np.full(n, np.degrees(phi)), pf)
        else:
            o3 = 1.5e-5 * np.exp(-0.5 * ((lp - np.log(10.0)) / 1.1) ** 2) + 4e-8
            o3 = o3 * (1.0 + 0.1 * rng.standard_normal((n, L))).clip(0.5, 1.5)
        alb = rng.uniform(0.05, 0.8, n)
        cz = np.maximum(0.0, np.cos(phi) / np.pi) + 0.05 * rng.standard_normal(n)
        if night:
            lon = np.arange(n) * (2 * np.pi / n)
            cz = np.cos(phi) * np.cos(lon - np.pi) + 0.02 * rng.standard_normal(n)
            cz = np.where(cz < 0.0, 0.0, cz)
        else:
            cz = np.maximum(cz, 0.02)
        cols.append(dict(pf=pf, ph=ph, T=T, Th=Th, Ts=Ts, q=q, o3=o3, alb=alb, cz=cz))

    def cat(key, flip):
        a = np.concatenate([c[key] for c in cols], axis=0)
        if flip:
            a = a[:, ::-1]                                        # surface first (rrtm_radiation.f90:652)
        return np.asfortranarray(a)

    ncol = (j1 - j0) * nlon
    ones = np.ones((ncol, L), order="F")
    zeros = np.zeros((ncol, L), order="F")
    if secondary_gases:
        ch4, n2o, o2 = 1.8e-6 * ones, 3.2e-7 * ones, 0.209 * ones
        cfc11, cfc12, cfc22, ccl4 = 2.5e-10 * ones, 5.3e-10 * ones, 2.0e-10 * ones, 9.0e-11 * ones
    else:
        ch4 = n2o = o2 = cfc11 = cfc12 = cfc22 = ccl4 = zeros
    return Columns(
        ncol=ncol, nlay=L, nlon=nlon, nlat=j1 - j0,
        play=cat("pf", True), plev=cat("ph", True), tlay=cat("T", True), tlev=cat("Th", True),
        tsfc=np.concatenate([c["Ts"] for c in cols]),
        h2o=cat("q", True), o3=cat("o3", True), co2=np.asfortranarray(co2_ppmv * 1e-6 * ones),
        ch4=ch4, n2o=n2o, o2=o2, cfc11=cfc11, cfc12=cfc12, cfc22=cfc22, ccl4=ccl4,
        emis=np.ones((ncol, 16), order="F"),
        albedo=np.concatenate([c["alb"] for c in cols]),
        coszen=np.concatenate([c["cz"] for c in cols]),
    )


def wild_columns(seed: int, nlay: int = 40, nlon: int = 64, nlat: int = 8) -> Columns:
    """Columns far outside the bench generator's climate, for branch coverage in the parity tests: surface pressures from
    300 to 1080 hPa, temperature profiles shifted by up to +-35 K, humidity over four orders of magnitude, ozone x 0..5,
    CO2 x 0.2..30, every secondary gas and CFC x 0..5, emissivities 0.6..1, albedo 0..1, a fifth of the columns at night,
    earth-sun distance, day of year and solar constant varied."""
    rng = np.random.default_rng(seed)
    c = make_columns("T42L40", nlon=nlon, nlat=nlat, nlay=nlay, secondary_gases=True, seed=seed)
    n, L = c.ncol, c.nlay
    f = rng.uniform(0.3, 1.08, n)
    c.play = np.asfortranarray(c.play * f[:, None])
    c.plev = np.asfortranarray(c.plev * f[:, None])
    dT = rng.uniform(-35, 35, n)[:, None] + rng.normal(0, 4, (n, L))
    c.tlay = np.asfortranarray(np.clip(c.tlay + dT, 100, 370))
    c.tlev = np.asfortranarray(np.clip(c.tlev + np.concatenate([dT, dT[:, -1:]], 1), 100, 370))
    c.tsfc = np.clip(c.tsfc + dT[:, 0], 100, 370)
    c.h2o = np.asfortranarray(np.clip(c.h2o * np.exp(rng.normal(0, 2.0, (n, L))), 2e-7, 0.06))
    c.o3 = np.asfortranarray(c.o3 * rng.uniform(0, 5, (n, L)))
    c.co2 = np.asfortranarray(c.co2 * np.exp(rng.uniform(np.log(0.2), np.log(30), (n, 1))) * np.ones((1, L)))
    for k, nom in (("ch4", 1.8e-6), ("n2o", 3.2e-7), ("o2", 0.209), ("cfc11", 2.5e-10), ("cfc12", 5.3e-10),
                   ("cfc22", 2e-10), ("ccl4", 9e-11)):
        setattr(c, k, np.asfortranarray(nom * rng.uniform(0, 5, (n, L))))
    c.o2 = np.asfortranarray(np.clip(c.o2, 0, 0.5))
    c.emis = np.asfortranarray(rng.uniform(0.6, 1.0, (n, 16)))
    c.albedo = rng.uniform(0, 1, n)
    c.coszen = np.where(rng.random(n) < 0.2, 0.0, rng.uniform(0, 1, n))
    c.adjes = float(rng.uniform(0.9, 1.1))
    c.dyofyr = int(rng.integers(0, 366))
    c.scon = float(rng.uniform(1300, 1400))
    return c


def make_gcm_state(resolution: str = "T42L40", **kw) -> dict:
    """The same synthetic atmosphere as make_columns(), in the layout MiMA's run_rrtmg receives
    (see gcm_state_from_columns)."""
    c = make_columns(resolution, **kw)
    nlat_full = kw["nlat"] if "nlat" in kw else RESOLUTIONS.get(resolution, (c.nlon, c.nlat, c.nlay))[1]
    return gcm_state_from_columns(c, nlat_full=nlat_full, j0=kw.get("lat_rows", (0, c.nlat))[0])


def gcm_state_from_columns(c: Columns, nlat_full: int | None = None, j0: int = 0) -> dict:
    """A Columns batch rearranged into the dummy arguments of run_rrtmg (rrtm_radiation.f90:471-503): fields
    (lon, lat, lev) with level 1 = top, pressures in Pa, p_half(top) = 0, lat/lon in radians, plus geopotential
    heights for interp_temp (hydrostatic, z_half(k=1) = 0 as in the model, rrtm_radiation.f90:433) and a zero
    tendency array."""
    si, sj, sk = c.nlon, c.nlat, c.nlay
    nlat_full = nlat_full or sj

    def fms(a):          # (ncol, n) surface-first -> (lon, lat, n) top-first
        return np.asfortranarray(a.reshape((si, sj, a.shape[1]), order="F")[:, :, ::-1])

    p_full = fms(c.play) * 100.0
    p_half = fms(c.plev) * 100.0
    p_half[:, :, 0] = 0.0
    t = fms(c.tlay)
    q = fms(c.h2o)
    o3f = fms(c.o3)
    t_surf = np.asfortranarray(c.tsfc.reshape((si, sj), order="F"))
    albedo = np.asfortranarray(c.albedo.reshape((si, sj), order="F"))
    lat1 = np.arcsin(np.linspace(-1.0 + 1.0 / nlat_full, 1.0 - 1.0 / nlat_full, nlat_full))
    lat = np.asfortranarray(np.broadcast_to(lat1[j0:j0 + sj][None, :], (si, sj)))
    lon = np.asfortranarray(np.broadcast_to((np.arange(si) * (2.0 * np.pi / si))[:, None], (si, sj)))
    rd_g = 287.04 / 9.80
    z_half = np.zeros((si, sj, sk + 1), order="F")
    z_full = np.zeros((si, sj, sk), order="F")
    for k in range(sk - 1, -1, -1):          # integrate upwards from the surface
        pb = p_half[:, :, k + 1]
        z_full[:, :, k] = z_half[:, :, k + 1] + rd_g * t[:, :, k] * np.log(pb / p_full[:, :, k])
        if k > 0:
            z_half[:, :, k] = z_half[:, :, k + 1] + rd_g * t[:, :, k] * np.log(pb / p_half[:, :, k])
    z_half[:, :, 0] = 0.0
    return dict(si=si, sj=sj, sk=sk, lat=lat, lon=lon, p_full=p_full, p_half=p_half, t=t, q=q, o3f=o3f,
                t_surf=t_surf, albedo=albedo, z_full=z_full, z_half=z_half, tdt=np.zeros((si, sj, sk), order="F"))
