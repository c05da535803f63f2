"""The two-stream flux recurrences of the default SW solver, checked on the CPU against a literal restatement of
vrtqdr_sw (SW/src/rrtmg_sw_vrtqdr.f90:103-150).

mima_b200/csrc/sw_solver.cu evaluates reftra once per (g, layer) cell and replaces one of the reference's two adding
recurrences by a flux propagation (variants 3/4: top-down first; variant 2: bottom-up first).  Both are algebraic
identities of the reference's formulas; this test pins the algebra itself with numpy, for random layer
properties, independent of the CUDA code (the GPU tests compare the kernels with the C oracle).
"""
import numpy as np
import pytest


def random_layers(rng, nlay, n):
    """Random layer properties: direct transmittance dbt, total transmittance for the direct beam tra >= dbt,
    reflectances with ref + tra <= 1 and refd + trad <= 1 (some layers almost opaque, some almost empty).  The direct
    and diffuse sets are drawn independently -- more general than reftra's output, which the identities do not need."""
    tau = 10.0 ** rng.uniform(-6, 1.5, (n, nlay))
    dbt = np.exp(-tau / rng.uniform(0.05, 1.0, (n, 1)))
    scat = rng.uniform(0.0, 1.0, (n, nlay)) * (1.0 - dbt)          # scattered part of the direct beam
    f = rng.uniform(0.0, 1.0, (n, nlay))
    ref, tra = scat * f, dbt + scat * (1.0 - f)
    absd = np.exp(-1.66 * tau)
    sd = rng.uniform(0.0, 1.0, (n, nlay)) * (1.0 - absd)
    fd = rng.uniform(0.0, 1.0, (n, nlay))
    refd, trad = sd * fd, absd + sd * (1.0 - fd)
    return ref, refd, tra, trad, dbt


def vrtqdr_reference(ref, refd, tra, trad, dbt, albp, albd):
    """rrtmg_sw_vrtqdr.f90 with the level index counted from the surface (level s below layer s): returns pfu, pfd."""
    n, L = ref.shape
    rup, rupd = np.empty((n, L + 1)), np.empty((n, L + 1))
    rup[:, 0], rupd[:, 0] = albp, albd                              # :103-108 (surface)
    for l in range(L):                                              # :112-121, bottom -> top
        zr = 1.0 / (1.0 - refd[:, l] * rupd[:, l])
        rup[:, l + 1] = ref[:, l] + trad[:, l] * ((tra[:, l] - dbt[:, l]) * rupd[:, l] + dbt[:, l] * rup[:, l]) * zr
        rupd[:, l + 1] = refd[:, l] + trad[:, l] * trad[:, l] * rupd[:, l] * zr
    tdn, rdnd, tdbt = np.empty((n, L + 1)), np.empty((n, L + 1)), np.empty((n, L + 1))
    tdn[:, L], rdnd[:, L], tdbt[:, L] = 1.0, 0.0, 1.0               # :125-128 (top)
    for l in range(L - 1, -1, -1):                                  # :130-140, top -> bottom
        zr = 1.0 / (1.0 - refd[:, l] * rdnd[:, l + 1])
        tdn[:, l] = tdbt[:, l + 1] * tra[:, l] + trad[:, l] * ((tdn[:, l + 1] - tdbt[:, l + 1])
                                                              + tdbt[:, l + 1] * ref[:, l] * rdnd[:, l + 1]) * zr
        rdnd[:, l] = refd[:, l] + trad[:, l] * trad[:, l] * rdnd[:, l + 1] * zr
        tdbt[:, l] = dbt[:, l] * tdbt[:, l + 1]
    zr = 1.0 / (1.0 - rdnd * rupd)                                  # :144-150
    pfu = (tdbt * rup + (tdn - tdbt) * rupd) * zr
    pfd = tdbt + (tdn - tdbt + tdbt * rup * rdnd) * zr
    return pfu, pfd


def top_down_first(ref, refd, tra, trad, dbt, albp, albd):
    """sw_solver_warp_kernel / sw_solver_kernel OPT bit 4: the top-down recurrence first, then U_a = zp U_b + zq."""
    n, L = ref.shape
    zp, zq = np.empty((n, L)), np.empty((n, L))
    rdnd_l, tdn_l = np.empty((n, L + 1)), np.empty((n, L + 1))
    tdn, rdnd, tdbt = np.ones(n), np.zeros(n), np.ones(n)
    for s in range(L, -1, -1):
        tdn_l[:, s], rdnd_l[:, s] = tdn, rdnd
        if s == 0:
            break
        l = s - 1
        dif = tdn - tdbt
        zr = 1.0 / (1.0 - refd[:, l] * rdnd)
        zp[:, l] = trad[:, l] * zr
        zq[:, l] = (ref[:, l] * tdbt + refd[:, l] * dif) * zr
        tdn_n = tdbt * tra[:, l] + trad[:, l] * (dif + tdbt * ref[:, l] * rdnd) * zr
        rdnd = refd[:, l] + trad[:, l] * trad[:, l] * rdnd * zr
        tdbt = dbt[:, l] * tdbt
        tdn = tdn_n
    pfu, pfd = np.empty((n, L + 1)), np.empty((n, L + 1))
    u = (albp * tdbt + albd * (tdn - tdbt)) / (1.0 - albd * rdnd)
    for s in range(L + 1):
        if s > 0:
            u = zp[:, s - 1] * u + zq[:, s - 1]
        pfu[:, s] = u
        pfd[:, s] = tdn_l[:, s] + rdnd_l[:, s] * u
    return pfu, pfd


def bottom_up_first(ref, refd, tra, trad, dbt, albp, albd):
    """sw_solver_kernel OPT bit 3: the bottom-up recurrence first, then D_below = fa D + fb S, S_below = dbt S."""
    n, L = ref.shape
    rup, rupd = np.empty((n, L + 1)), np.empty((n, L + 1))
    fa, fb = np.empty((n, L)), np.empty((n, L))
    rup[:, 0], rupd[:, 0] = albp, albd
    for l in range(L):
        zr = 1.0 / (1.0 - refd[:, l] * rupd[:, l])
        fa[:, l] = trad[:, l] * zr
        fb[:, l] = ((tra[:, l] - dbt[:, l]) + refd[:, l] * (rup[:, l] * dbt[:, l])) * zr
        rup[:, l + 1] = ref[:, l] + trad[:, l] * ((tra[:, l] - dbt[:, l]) * rupd[:, l] + dbt[:, l] * rup[:, l]) * zr
        rupd[:, l + 1] = refd[:, l] + trad[:, l] * trad[:, l] * rupd[:, l] * zr
    pfu, pfd = np.empty((n, L + 1)), np.empty((n, L + 1))
    D, S = np.zeros(n), np.ones(n)
    for s in range(L, -1, -1):
        pfu[:, s] = rupd[:, s] * D + rup[:, s] * S
        pfd[:, s] = D + S
        if s > 0:
            D = fa[:, s - 1] * D + fb[:, s - 1] * S
            S = dbt[:, s - 1] * S
    return pfu, pfd


@pytest.mark.parametrize("nlay", [1, 2, 40, 60, 128])
def test_flux_propagation_equals_vrtqdr(nlay):
    rng = np.random.default_rng(100 + nlay)
    n = 4000
    props = random_layers(rng, nlay, n)
    albp, albd = rng.uniform(0.0, 0.95, n), rng.uniform(0.0, 0.95, n)
    pfu, pfd = vrtqdr_reference(*props, albp, albd)
    assert (pfd[:, -1] == 1.0).all() and (pfu >= 0).all()
    for scheme in (top_down_first, bottom_up_first):
        u, d = scheme(*props, albp, albd)
        # relative to the incident flux (= 1 at the top): what the g-point sums and the 1e-6 flux tolerance see
        assert np.max(np.abs(u - pfu) / np.maximum(pfu, 1.0)) < 2e-13, scheme.__name__
        assert np.max(np.abs(d - pfd) / np.maximum(pfd, 1.0)) < 2e-13, scheme.__name__
        # and relative to the value itself wherever the flux is not negligible
        big = pfd > 1e-6
        assert np.max(np.abs(d - pfd)[big] / pfd[big]) < 1e-11, scheme.__name__


def test_black_surface_and_empty_atmosphere():
    """Limits: no atmosphere -> pfd = 1, pfu = direct albedo at every level; black surface under a non-scattering
    atmosphere -> no upward flux."""
    n, L = 16, 40
    one, zero = np.ones((n, L)), np.zeros((n, L))
    albp, albd = np.linspace(0, 0.9, n), np.linspace(0.9, 0, n)
    for scheme in (vrtqdr_reference, top_down_first, bottom_up_first):
        u, d = scheme(zero, zero, one, one, one, albp, albd)
        assert np.allclose(d, 1.0, rtol=0, atol=1e-15) and np.allclose(u, albp[:, None], rtol=0, atol=1e-15)
        dbt = np.full((n, L), 0.9)
        u, d = scheme(zero, zero, dbt, dbt, dbt, np.zeros(n), np.zeros(n))
        assert (u == 0).all() and np.allclose(d[:, 0], 0.9 ** L, rtol=1e-13)
