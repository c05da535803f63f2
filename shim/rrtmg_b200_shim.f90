! rrtmg_b200_shim.f90 -- ISO_C_BINDING shim that puts librrtmg_b200.so behind MiMA's existing RRTMG module
! and procedure names, so that physics_driver.f90:132-135,577-578 and rrtm_radiation.f90:473-474,686-748
! compile and run unchanged.  Build it INSTEAD of
!     rrtmg_lw/gcm_model/src/rrtmg_lw_rad.nomcica.f90, rrtmg_lw_init.f90 (keep rrtmg_lw_k_g.f90 + modules/)
!     rrtmg_sw/gcm_model/src/rrtmg_sw_rad.nomcica.f90, rrtmg_sw_init.f90 (keep rrtmg_sw_k_g.f90 + modules/)
! and link with -lrrtmg_b200 -lcudart.  (Not compiled in the authoring container: no Fortran compiler.)
!
! Assumed-shape dummies may be non-contiguous or expression temporaries (ch4_val*ones, 10*ones); the
! CONTIGUOUS attribute makes the compiler pass a packed copy when needed.  Arrays MiMA never varies
! (zeros for the secondary gases, tauaer = 0, emis = 1) are detected and passed as C_NULL_PTR so that they
! never cross PCIe (extension documented in include/rrtmg_b200.h).

module rrtmg_b200_c
  use iso_c_binding
  implicit none
  interface
     integer(c_int) function rrtmg_b200_set_device(local_rank) bind(c)
       import; integer(c_int), value :: local_rank
     end function
     integer(c_int) function rrtmg_b200_set_table(name, data, ndim, dims) bind(c)
       import; character(kind=c_char) :: name(*); real(c_double) :: data(*)
       integer(c_int), value :: ndim; integer(c_int) :: dims(*)
     end function
     integer(c_int) function rrtmg_b200_load_tables(path) bind(c)
       import; character(kind=c_char) :: path(*)
     end function
     integer(c_int) function rrtmg_b200_lw_init(cpdair) bind(c)
       import; real(c_double), value :: cpdair
     end function
     integer(c_int) function rrtmg_b200_sw_init(cpdair) bind(c)
       import; real(c_double), value :: cpdair
     end function
     integer(c_int) function rrtmg_b200_lw(ncol, nlay, icld, idrv, play, plev, tlay, tlev, tsfc, &
          h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, &
          inflglw, iceflglw, liqflglw, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer, &
          uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt) bind(c)
       import
       integer(c_int), value :: ncol, nlay, idrv, inflglw, iceflglw, liqflglw
       integer(c_int) :: icld
       type(c_ptr), value :: play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, &
            cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer, &
            uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt
     end function
     integer(c_int) function rrtmg_b200_sw(ncol, nlay, icld, iaer, play, plev, tlay, tlev, tsfc, &
          h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, asdir, asdif, aldir, aldif, coszen, adjes, dyofyr, scon, &
          inflgsw, iceflgsw, liqflgsw, cldfr, taucld, ssacld, asmcld, fsfcld, cicewp, cliqwp, reice, reliq, &
          tauaer, ssaaer, asmaer, ecaer, swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc) bind(c)
       import
       integer(c_int), value :: ncol, nlay, dyofyr, inflgsw, iceflgsw, liqflgsw
       integer(c_int) :: icld, iaer
       real(c_double), value :: adjes, scon
       type(c_ptr), value :: play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, &
            asdir, asdif, aldir, aldif, coszen, cldfr, taucld, ssacld, asmcld, fsfcld, cicewp, cliqwp, reice, &
            reliq, tauaer, ssaaer, asmaer, ecaer, swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc
     end function
     function rrtmg_b200_last_error() bind(c) result(p)
       import; type(c_ptr) :: p
     end function
  end interface
  ! MiMA never reads the clear-sky outputs of either routine (rrtm_radiation.f90:716, 752, 790-791 use swuflx,
  ! swdflx, swhr, uflx, dflx, hr only).  Setting this to .false. passes C_NULL_PTR for uflxc, dflxc, hrc / swuflxc,
  ! swdflxc, swhrc: the library then leaves those dummies untouched and does not copy them back over PCIe.
  logical, save :: b200_clear_sky_outputs = .true.
contains
  function optout(a) result(p)      ! clear-sky output wanted?
    real(c_double), contiguous, target, intent(inout) :: a(:,:)
    type(c_ptr) :: p
    p = c_null_ptr
    if (b200_clear_sky_outputs) p = c_loc(a)
  end function
  subroutine b200_check(rc, where)
    ! C status -> the model's fatal-error convention (cf. error_mesg(...,FATAL), rrtm_radiation.f90:527-528)
    use fms_mod, only: error_mesg, FATAL
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(len=32) :: code
    if (rc /= 0) then
       write(code, '(i0)') rc
       call error_mesg(where, 'librrtmg_b200 returned error '//trim(code), FATAL)
    end if
  end subroutine
  function opt2(a) result(p)        ! all-zero (ncol,nlay) input -> NULL: "absent" in the C ABI
    real(c_double), contiguous, target, intent(in) :: a(:,:)
    type(c_ptr) :: p
    p = c_null_ptr
    if (any(a /= 0._c_double)) p = c_loc(a)
  end function
end module rrtmg_b200_c

module rrtmg_lw_init
  use iso_c_binding
  use rrtmg_b200_c
  implicit none
contains
  subroutine rrtmg_lw_ini(cpdair)                         ! replaces LW/src/rrtmg_lw_init.f90:28
    use rrlw_kg01, only: kao1 => kao, kbo1 => kbo           ! ... likewise rrlw_kg02..16, rrlw_wvn, rrlw_ref
    real(c_double), intent(in) :: cpdair
    ! 1. fill the unreduced module arrays exactly as the stock code does: lw_kgb01..16, lwatmref, lwavplank
    !    (these data routines stay in the build: rrtmg_lw_k_g.f90, rrtmg_lw_setcoef.f90 data section);
    ! 2. register each with its blob name, e.g.
    !       call reg('lw01.kao', kao1)   ->  rrtmg_b200_set_table('lw01.kao'//c_null_char, kao1, 3, [5,13,16])
    !    (names: tools/build_tables.py; one line per array of LW/modules/rrlw_kg01..16.f90, plus
    !     lwref.pref/preflog/tref/chi_mls/totplnk);
    ! 3. reduce 16 -> ngc g-points, build exp/tfn tables, upload:
    call b200_check(rrtmg_b200_lw_init(cpdair), 'rrtmg_lw_ini')
  end subroutine
end module rrtmg_lw_init

module rrtmg_lw_rad
  use iso_c_binding
  use rrtmg_b200_c
  implicit none
contains
  subroutine rrtmg_lw(ncol, nlay, icld, idrv, play, plev, tlay, tlev, tsfc, &
       h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, &
       inflglw, iceflglw, liqflglw, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer, &
       uflx, dflx, hr, uflxc, dflxc, hrc, duflx_dt, duflxc_dt)   ! LW/src/rrtmg_lw_rad.nomcica.f90:80-89
    integer, intent(in) :: ncol, nlay, idrv, inflglw, iceflglw, liqflglw
    integer, intent(inout) :: icld
    real(c_double), contiguous, target, intent(in) :: play(:,:), plev(:,:), tlay(:,:), tlev(:,:), tsfc(:)
    real(c_double), contiguous, target, intent(in) :: h2ovmr(:,:), o3vmr(:,:), co2vmr(:,:), ch4vmr(:,:), &
         n2ovmr(:,:), o2vmr(:,:), cfc11vmr(:,:), cfc12vmr(:,:), cfc22vmr(:,:), ccl4vmr(:,:), emis(:,:)
    real(c_double), contiguous, target, intent(in) :: cldfr(:,:), cicewp(:,:), cliqwp(:,:), reice(:,:), reliq(:,:)
    real(c_double), contiguous, target, intent(in) :: taucld(:,:,:), tauaer(:,:,:)
    real(c_double), contiguous, target, intent(inout) :: uflx(:,:), dflx(:,:), hr(:,:), uflxc(:,:), dflxc(:,:), hrc(:,:)
    real(c_double), contiguous, target, intent(out), optional :: duflx_dt(:,:), duflxc_dt(:,:)
    integer(c_int) :: icld_c
    type(c_ptr) :: pemis, paer, pcldfr, ptaucld, pdu, pduc, pwp(4)
    icld_c = icld
    pemis = c_null_ptr; if (any(emis /= 1._c_double)) pemis = c_loc(emis)
    paer = c_null_ptr;  if (any(tauaer /= 0._c_double)) paer = c_loc(tauaer)
    ! cloud arrays are never dereferenced for icld = 0 (as in the reference); pass NULL.  With icld > 0 the library
    ! takes the cloud fraction and the band optical depths (inflglw = 0); water-path inputs return error 2.
    pcldfr = c_null_ptr; ptaucld = c_null_ptr; pwp = c_null_ptr
    if (icld /= 0) then
       pcldfr = c_loc(cldfr); ptaucld = c_loc(taucld)
       if (inflglw >= 1) then
          pwp(1) = c_loc(cicewp); pwp(2) = c_loc(cliqwp); pwp(3) = c_loc(reice); pwp(4) = c_loc(reliq)
       endif
    endif
    pdu = c_null_ptr; pduc = c_null_ptr
    if (idrv == 1 .and. present(duflx_dt) .and. present(duflxc_dt)) then
       pdu = c_loc(duflx_dt); pduc = c_loc(duflxc_dt)
    endif
    call b200_check(rrtmg_b200_lw(ncol, nlay, icld_c, idrv, c_loc(play), c_loc(plev), c_loc(tlay), c_loc(tlev), &
         c_loc(tsfc), c_loc(h2ovmr), c_loc(o3vmr), c_loc(co2vmr), opt2(ch4vmr), opt2(n2ovmr), opt2(o2vmr), &
         opt2(cfc11vmr), opt2(cfc12vmr), opt2(cfc22vmr), opt2(ccl4vmr), pemis, inflglw, iceflglw, liqflglw, &
         pcldfr, ptaucld, pwp(1), pwp(2), pwp(3), pwp(4), paer, &
         c_loc(uflx), c_loc(dflx), c_loc(hr), optout(uflxc), optout(dflxc), optout(hrc), pdu, pduc), 'rrtmg_lw')
    icld = icld_c
  end subroutine
end module rrtmg_lw_rad

module rrtmg_sw_init
  use iso_c_binding
  use rrtmg_b200_c
  implicit none
contains
  subroutine rrtmg_sw_ini(cpdair)                         ! replaces SW/src/rrtmg_sw_init.f90:28
    real(c_double), intent(in) :: cpdair
    ! sw_kgb16..29 (rrtmg_sw_k_g.f90) fill rrsw_kg16..29; register 'sw16.kao' ... as for LW, or simply
    !   call b200_check(rrtmg_b200_load_tables('INPUT/rrtmg_sw_kg.bin'//c_null_char), 'rrtmg_sw_ini')
    call b200_check(rrtmg_b200_sw_init(cpdair), 'rrtmg_sw_ini')
  end subroutine
end module rrtmg_sw_init

module rrtmg_sw_rad
  use iso_c_binding
  use rrtmg_b200_c
  implicit none
contains
  subroutine rrtmg_sw(ncol, nlay, icld, iaer, play, plev, tlay, tlev, tsfc, &
       h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, asdir, asdif, aldir, aldif, &
       coszen, adjes, dyofyr, scon, inflgsw, iceflgsw, liqflgsw, cldfr, &
       taucld, ssacld, asmcld, fsfcld, cicewp, cliqwp, reice, reliq, tauaer, ssaaer, asmaer, ecaer, &
       swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc)           ! SW/src/rrtmg_sw_rad.nomcica.f90:78-88
    integer, intent(in) :: ncol, nlay, dyofyr, inflgsw, iceflgsw, liqflgsw
    integer, intent(inout) :: icld, iaer
    real(c_double), intent(in) :: adjes, scon
    real(c_double), contiguous, target, intent(in) :: play(:,:), plev(:,:), tlay(:,:), tlev(:,:), tsfc(:)
    real(c_double), contiguous, target, intent(in) :: h2ovmr(:,:), o3vmr(:,:), co2vmr(:,:), ch4vmr(:,:), n2ovmr(:,:), o2vmr(:,:)
    real(c_double), contiguous, target, intent(in) :: asdir(:), asdif(:), aldir(:), aldif(:), coszen(:)
    ! cloud / aerosol dummies: MiMA passes LW-shaped arrays here (rrtm_radiation.f90:692-708); they are never
    ! read for icld = 0 / iaer = 0, so they are declared assumed-size and forwarded only for icld > 0 (cldfr and the
    ! four band arrays, inflgsw = 0) and iaer = 10 (the three band arrays)
    real(c_double), target, intent(in) :: cldfr(*), taucld(*), ssacld(*), asmcld(*), fsfcld(*), cicewp(*), cliqwp(*), &
         reice(*), reliq(*), tauaer(*), ssaaer(*), asmaer(*), ecaer(*)
    real(c_double), contiguous, target, intent(inout) :: swuflx(:,:), swdflx(:,:), swhr(:,:), swuflxc(:,:), swdflxc(:,:), swhrc(:,:)
    integer(c_int) :: icld_c, iaer_c
    type(c_ptr) :: pc(5), pa(3), pec, pw(4)
    icld_c = icld; iaer_c = iaer
    pc = c_null_ptr; pa = c_null_ptr; pec = c_null_ptr; pw = c_null_ptr
    if (icld /= 0 .and. inflgsw == 2) then
       pw(1) = c_loc(cicewp); pw(2) = c_loc(cliqwp); pw(3) = c_loc(reice); pw(4) = c_loc(reliq)
    endif
    if (iaer == 6) pec = c_loc(ecaer)
    if (icld /= 0) then
       pc(1) = c_loc(cldfr); pc(2) = c_loc(taucld); pc(3) = c_loc(ssacld); pc(4) = c_loc(asmcld); pc(5) = c_loc(fsfcld)
    endif
    if (iaer == 10) then
       pa(1) = c_loc(tauaer); pa(2) = c_loc(ssaaer); pa(3) = c_loc(asmaer)
    endif
    call b200_check(rrtmg_b200_sw(ncol, nlay, icld_c, iaer_c, c_loc(play), c_loc(plev), c_loc(tlay), c_loc(tlev), &
         c_loc(tsfc), c_loc(h2ovmr), c_loc(o3vmr), c_loc(co2vmr), opt2(ch4vmr), opt2(n2ovmr), opt2(o2vmr), &
         c_loc(asdir), c_loc(asdif), c_loc(aldir), c_loc(aldif), c_loc(coszen), adjes, dyofyr, scon, &
         inflgsw, iceflgsw, liqflgsw, pc(1), pc(2), pc(3), pc(4), pc(5), pw(1), &
         pw(2), pw(3), pw(4), pa(1), pa(2), pa(3), pec, &
         c_loc(swuflx), c_loc(swdflx), c_loc(swhr), optout(swuflxc), optout(swdflxc), optout(swhrc)), 'rrtmg_sw')
    icld = icld_c; iaer = iaer_c
  end subroutine
end module rrtmg_sw_rad

! ---------------------------------------------------------------------------------------------------------
! Optional second stage (SURVEY.md section 8f, ranks 1-2): the radiation step of run_rrtmg on the device.
! With this module the body of run_rrtmg between "we know now that we want to run radiation"
! (rrtm_radiation.f90:585) and write_diag_rrtm (:829) becomes ONE call: lon sub-sampling, the vertical flip,
! Pa -> hPa, the clamps, interp_temp, compute_zenith, both RRTMG calls, K/day -> K/s, the lon
! re-interpolation and the surface fluxes all run on the GPU, and only the GCM state crosses PCIe.
! The alarm (dt_rad), the interpolator_mod reads (ozone, h2o, radiation files) and the diag manager stay as
! they are.  type(b200_rad_config) mirrors struct rrtmg_b200_rad_config (include/rrtmg_b200.h) field by field.
module rrtm_radiation_b200
  use iso_c_binding
  use rrtmg_b200_c, only: b200_check
  implicit none
  type, bind(c) :: b200_rad_config
     integer(c_int) :: include_secondary_gases, do_fixed_water, do_zm_tracers, do_rad_time_avg, dt_rad_avg, &
          lonstep, do_zm_rad, use_dyofyr, solday, days_per_year
     real(c_double) :: scale_ozone, o3_val, ch4_val, n2o_val, o2_val, cfc11_val, cfc12_val, cfc22_val, ccl4_val, &
          h2o_lower_limit, temp_lower_limit, temp_upper_limit, co2ppmv, fixed_water, fixed_water_pres, &
          fixed_water_lat, slowdown_rad, obliq, solr_cnst, solrad, equinox_day
  end type
  interface
     integer(c_int) function rrtmg_b200_run_rrtmg(cfg, si, sj, sk, seconds, days, lat, lon, p_full, p_half, albedo, &
          q, t, t_surf_rad, z_full, z_half, t_half_in, o3f, tdt, coszen, flux_sw, flux_lw, tdt_rad, tdt_sw, tdt_lw, &
          olr, isr, t_half_out) bind(c)
       import
       type(b200_rad_config), intent(in) :: cfg
       integer(c_int), value :: si, sj, sk, seconds, days
       type(c_ptr), value :: lat, lon, p_full, p_half, albedo, q, t, t_surf_rad, z_full, z_half, t_half_in, o3f, &
            tdt, coszen, flux_sw, flux_lw, tdt_rad, tdt_sw, tdt_lw, olr, isr, t_half_out
     end function
  end interface
contains
  ! Called from run_rrtmg in place of rrtm_radiation.f90:585-808.  `cfg` is filled once in rrtm_radiation_init
  ! from the namelist variables of rrtm_vars / rrtm_astro (logicals -> 0/1; dt_rad_avg after :336-340;
  ! days_per_year from get_time(length_of_year())).
  subroutine run_rrtmg_b200(cfg, seconds, days, lat, lon, p_full, p_half, albedo, q, t, t_surf_rad, z_full, z_half, &
       o3f, have_o3, tdt, coszen, flux_sw, flux_lw, tdt_rad, t_half)
    type(b200_rad_config), intent(in) :: cfg
    integer, intent(in) :: seconds, days
    real(c_double), contiguous, target, intent(in) :: lat(:,:), lon(:,:), albedo(:,:), t_surf_rad(:,:)
    real(c_double), contiguous, target, intent(in) :: p_full(:,:,:), p_half(:,:,:), q(:,:,:), t(:,:,:)
    real(c_double), contiguous, target, intent(in) :: z_full(:,:,:), z_half(:,:,:), o3f(:,:,:)
    logical, intent(in) :: have_o3                        ! do_read_ozone
    real(c_double), contiguous, target, intent(inout) :: tdt(:,:,:)
    real(c_double), contiguous, target, intent(out) :: coszen(:,:), flux_sw(:,:), flux_lw(:,:), tdt_rad(:,:,:), t_half(:,:,:)
    type(c_ptr) :: po3
    po3 = c_null_ptr; if (have_o3) po3 = c_loc(o3f)
    call b200_check(rrtmg_b200_run_rrtmg(cfg, size(t,1), size(t,2), size(t,3), seconds, days, &
         c_loc(lat), c_loc(lon), c_loc(p_full), c_loc(p_half), c_loc(albedo), c_loc(q), c_loc(t), c_loc(t_surf_rad), &
         c_loc(z_full), c_loc(z_half), c_null_ptr, po3, c_loc(tdt), c_loc(coszen), c_loc(flux_sw), c_loc(flux_lw), &
         c_loc(tdt_rad), c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_loc(t_half)), 'run_rrtmg')
  end subroutine
end module rrtm_radiation_b200
