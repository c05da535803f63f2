import numpy as np, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from mima_b200 import rrtmg
from mima_b200.columns import make_columns
from test_oracle_lw_clouds import cloud_field
rrtmg.set_device(0); rrtmg.rrtmg_lw_ini(allow_synthetic_lw=True); rrtmg.rrtmg_sw_ini()
c = make_columns("T42L40", nlon=32, nlat=4, night=True)
rng = np.random.default_rng(1)
cl = cloud_field(c, rng)
rrtmg.lw_from_columns(c); rrtmg.sw_from_columns(c)
# column kernels: a ragged last tile, less than one tile, the aerosol instantiation of lw_column, both block shapes
for n in (77, 5):
    c2 = c.take(np.arange(n))
    rrtmg.lw_from_columns(c2); rrtmg.sw_from_columns(c2)
rrtmg.lw_from_columns(c, tauaer=np.asfortranarray(rng.uniform(0.0, 0.05, (c.ncol, c.nlay, 16))))
for cw in (8, 16, 0):
    rrtmg.set_option("col_warps", cw)
    rrtmg.lw_from_columns(c2); rrtmg.sw_from_columns(c2)
# the staged clear-sky kernels
rrtmg.set_option("lw_fused", 0); rrtmg.set_option("sw_fused", 0)
rrtmg.lw_from_columns(c); rrtmg.sw_from_columns(c)
rrtmg.set_option("lw_fused", 1); rrtmg.set_option("sw_fused", 1)
import os
if not os.environ.get("SAN_NO_TMA"):
    rrtmg.lw_from_columns(c, idrv=1)
for icld in (1, 2):
    rrtmg.lw_from_columns(c, icld=icld, clouds=cl, idrv=1)
shp = (14, c.ncol, c.nlay)
cld = (rng.uniform(size=(c.ncol, c.nlay)) < 0.3).astype(float)
asm = rng.uniform(0.7, 0.9, shp)
swcl = dict(cldfr=np.asfortranarray(cld), taucld=np.asfortranarray(rng.uniform(0, 20, shp) * cld[None]),
            ssacld=np.asfortranarray(rng.uniform(0.5, 0.99999, shp)), asmcld=np.asfortranarray(asm), fsfcld=np.asfortranarray(asm * asm))
a3 = (c.ncol, c.nlay, 14)
aer = dict(tauaer=np.asfortranarray(rng.uniform(0, 0.3, a3)), ssaaer=np.asfortranarray(rng.uniform(0.6, 0.999, a3)), asmaer=np.asfortranarray(rng.uniform(0.3, 0.8, a3)))
rrtmg.sw_from_columns(c, icld=2, iaer=10, clouds=swcl, aerosols=aer)
rrtmg.sw_from_columns(c, iaer=10, aerosols=aer)
shp2 = (c.ncol, c.nlay)
wcl = dict(cldfr=np.asfortranarray((rng.uniform(size=shp2) < 0.3).astype(float)),
           cicewp=np.asfortranarray(rng.uniform(0, 30, shp2)), cliqwp=np.asfortranarray(rng.uniform(0, 60, shp2)),
           reice=np.asfortranarray(rng.uniform(14, 120, shp2)), reliq=np.asfortranarray(rng.uniform(3, 50, shp2)))
for ice in (1, 2, 3):
    rrtmg.sw_from_columns(c, icld=2, inflgsw=2, iceflgsw=ice, liqflgsw=1, clouds=wcl)
for infl, ice, liq in ((1, 0, 0), (2, 0, 0), (2, 1, 1), (2, 2, 1), (2, 3, 1)):
    rrtmg.lw_from_columns(c, icld=2, inflglw=infl, iceflglw=ice, liqflglw=liq, clouds=wcl)
aer6 = dict(ecaer=np.asfortranarray(rng.uniform(0, 0.05, (c.ncol, c.nlay, 6))))
rrtmg.sw_from_columns(c, iaer=6, aerosols=aer6)
c80 = make_columns("T341L80", nlon=32, nlat=2, night=True)
rrtmg.lw_from_columns(c80); rrtmg.sw_from_columns(c80)
from mima_b200 import rrtm_radiation as rr
print("sanitizer workload done")
