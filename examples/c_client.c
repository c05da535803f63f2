/* c_client.c -- the drop-in boundary used from plain C: what a compiled host (MiMA's Fortran through the ISO_C_BINDING
 * shim, shim/rrtmg_b200_shim.f90) does, without Python in between.
 *
 *   c_client <data_dir> <columns.bin> <fluxes.bin> [cp_air]
 *
 * columns.bin (written by tests/test_c_client.py): int32 ncol, nlay; then, as float64 column-major arrays in the order of
 * the rrtmg_sw / rrtmg_lw dummy lists: play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, albedo, coszen, and the
 * scalars adjes, scon.  fluxes.bin receives swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc, uflx, dflx, hr, uflxc, dflxc, hrc.
 * The calls are the ones of rrtm_radiation.f90:686-748 with MiMA's fixed switches (icld = iaer = idrv = 0, zero secondary
 * gases, emissivity 1, dyofyr = 0).
 *
 * Build: gcc -std=c99 -O2 -I include examples/c_client.c -L mima_b200/lib -lrrtmg_b200 -Wl,-rpath,$PWD/mima_b200/lib
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rrtmg_b200.h"

static void die(const char *what, int rc)
{
    fprintf(stderr, "c_client: %s failed with %d: %s\n", what, rc, rrtmg_b200_last_error());
    exit(1);
}

static double *read_arr(FILE *f, size_t n)
{
    double *a = (double *)malloc(n * sizeof(double));
    if (!a || fread(a, sizeof(double), n, f) != n) { fprintf(stderr, "c_client: short read\n"); exit(2); }
    return a;
}

int main(int argc, char **argv)
{
    if (argc != 4 && argc != 5) { fprintf(stderr, "usage: c_client <data_dir> <columns.bin> <fluxes.bin> [cp_air]\n"); return 2; }
    char path[1024];
    int rc;
    if ((rc = rrtmg_b200_set_device(0))) die("set_device", rc);
    const char *blobs[3] = {"rrtmg_lw_ref.bin", "rrtmg_lw_kg_synth.bin", "rrtmg_sw_kg.bin"};
    for (int i = 0; i < 3; ++i) {
        snprintf(path, sizeof path, "%s/%s", argv[1], blobs[i]);
        if ((rc = rrtmg_b200_load_tables(path))) die(path, rc);
    }
    const double cp_air = argc == 5 ? strtod(argv[4], NULL) : 1004.64;   /* constants_mod: rdgas / kappa */
    if ((rc = rrtmg_b200_lw_init(cp_air))) die("lw_init", rc);       /* physics_driver.f90:577-578 */
    if ((rc = rrtmg_b200_sw_init(cp_air))) die("sw_init", rc);

    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 2; }
    int32_t dims[2];
    if (fread(dims, sizeof(int32_t), 2, f) != 2) return 2;
    const int ncol = dims[0], nlay = dims[1];
    const size_t L = (size_t)ncol * nlay, V = (size_t)ncol * (nlay + 1);
    double *play = read_arr(f, L), *plev = read_arr(f, V), *tlay = read_arr(f, L), *tlev = read_arr(f, V);
    double *tsfc = read_arr(f, ncol), *h2o = read_arr(f, L), *o3 = read_arr(f, L), *co2 = read_arr(f, L);
    double *albedo = read_arr(f, ncol), *coszen = read_arr(f, ncol), *sc = read_arr(f, 2);
    fclose(f);

    double *out[12];
    for (int i = 0; i < 12; ++i) out[i] = (double *)calloc((i % 3 == 2) ? L : V, sizeof(double));
    int icld = 0, iaer = 0;
    rc = rrtmg_b200_sw(ncol, nlay, &icld, &iaer, play, plev, tlay, tlev, tsfc, h2o, o3, co2, NULL, NULL, NULL,
                       albedo, albedo, albedo, albedo, coszen, sc[0], 0, sc[1],
                       0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL,
                       out[0], out[1], out[2], out[3], out[4], out[5]);
    if (rc) die("rrtmg_b200_sw", rc);
    rc = rrtmg_b200_lw(ncol, nlay, &icld, 0, play, plev, tlay, tlev, tsfc, h2o, o3, co2, NULL, NULL, NULL,
                       NULL, NULL, NULL, NULL, NULL,
                       0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL,
                       out[6], out[7], out[8], out[9], out[10], out[11], NULL, NULL);
    if (rc) die("rrtmg_b200_lw", rc);

    /* error path: a call the library must refuse (water-path cloud optics without the water-path arrays) */
    icld = 2;
    rc = rrtmg_b200_lw(ncol, nlay, &icld, 0, play, plev, tlay, tlev, tsfc, h2o, o3, co2, NULL, NULL, NULL,
                       NULL, NULL, NULL, NULL, NULL,
                       2, 0, 0, tlay, tlay, NULL, NULL, NULL, NULL, NULL,
                       out[6], out[7], out[8], out[9], out[10], out[11], NULL, NULL);
    if (rc != RRTMG_B200_ERR_BAD_ARGUMENT) { fprintf(stderr, "c_client: expected ERR_BAD_ARGUMENT, got %d\n", rc); return 3; }

    f = fopen(argv[3], "wb");
    if (!f) { perror(argv[3]); return 2; }
    for (int i = 0; i < 12; ++i) fwrite(out[i], sizeof(double), (i % 3 == 2) ? L : V, f);
    fclose(f);
    printf("c_client: %d columns x %d layers; column 0: SW down at the surface %.6f W/m2, OLR %.6f W/m2; launches %ld\n", ncol,
           nlay, out[1][0], out[6][(size_t)nlay * ncol], rrtmg_b200_launch_count());
    if ((rc = rrtmg_b200_finalize())) die("finalize", rc);
    return 0;
}
