! rl_capi_mod.f90 -- iso_c_binding interfaces to libradlite_b200.so (include/radlite_b200.h).
! Compiled together with the reference sources (see fortran/Makefile.gpu and INTEGRATION.md).
! NOTE: no Fortran compiler exists in the build image of this repository, so this file has been
! checked by reading only; the C++ host program (radlite_b200/csrc/host_main.cpp) exercises the
! identical ABI.
module rl_capi_mod
  use iso_c_binding
  implicit none
  interface
     integer(c_int) function rl_create(ctx, device) bind(c, name='rl_create')
       import :: c_ptr, c_int
       type(c_ptr), intent(out) :: ctx
       integer(c_int), value :: device
     end function
     subroutine rl_destroy(ctx) bind(c, name='rl_destroy')
       import :: c_ptr
       type(c_ptr), value :: ctx
     end subroutine
     type(c_ptr) function rl_last_error(ctx) bind(c, name='rl_last_error')
       import :: c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function rl_set_grid_ghosted(ctx, nr, nt, rc_m1, tc_m1) &
          bind(c, name='rl_set_grid_ghosted')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: nr, nt
       real(c_double), intent(in) :: rc_m1(*), tc_m1(*)
     end function
     integer(c_int) function rl_set_medium(ctx, rho, abund, vel, linewidth, umass_av) &
          bind(c, name='rl_set_medium')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(in) :: rho(*), abund(*), vel(*), linewidth(*)
       real(c_double), value :: umass_av
     end function
     integer(c_int) function rl_set_lines(ctx, nlines, nlevels, lev_up, lev_down, linefreq, aud, &
          gdeg, popul) bind(c, name='rl_set_lines')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: nlines, nlevels
       integer(c_int), intent(in) :: lev_up(*), lev_down(*)
       real(c_double), intent(in) :: linefreq(*), aud(*), gdeg(*), popul(*)
     end function
     integer(c_int) function rl_set_line_dust(ctx, src, alp) bind(c, name='rl_set_line_dust')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(in) :: src(*), alp(*)
     end function
     integer(c_int) function rl_set_camera(ctx, anginf, nphi, nrext, dbdr, rstar, imethod, nrref) &
          bind(c, name='rl_set_camera')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: anginf, rstar
       integer(c_int), value :: nphi, nrext, dbdr, imethod, nrref
     end function
     integer(c_int) function rl_set_bc(ctx, in_itype, out_itype, ncf, cont_freq_nu, starspec, isrf) &
          bind(c, name='rl_set_bc')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: in_itype, out_itype, ncf
       real(c_double), intent(in) :: cont_freq_nu(*), starspec(*), isrf(*)
     end function
     integer(c_int) function rl_set_options(ctx, subgrid, nonredundant, levthres, aksmax) &
          bind(c, name='rl_set_options')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: subgrid, nonredundant
       real(c_double), value :: levthres, aksmax
     end function
     ! opaque-wall start (include/radlite_b200.h): 0 integrates every segment like the reference; default 64
     integer(c_int) function rl_set_wall_tau(ctx, tau) bind(c, name='rl_set_wall_tau')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: tau
     end function
     real(c_double) function rl_get_executed(ctx) bind(c, name='rl_get_executed')
       import :: c_ptr, c_double
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function rl_get_camera_dims(ctx, nrr, nphi, nray) bind(c, name='rl_get_camera_dims')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
       integer(c_int), intent(out) :: nrr, nphi, nray
     end function
     integer(c_int) function rl_get_rings(ctx, rays_r, imcir_ri) bind(c, name='rl_get_rings')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: rays_r(*), imcir_ri(*)
     end function
     integer(c_int) function rl_render(ctx, iline0, nl, nfr, vmax_kms, dist_cm, flux, imcir, cmask, &
          tau_center, maserflag, velo) bind(c, name='rl_render')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: iline0, nl, nfr
       real(c_double), value :: vmax_kms, dist_cm
       real(c_double), intent(out) :: flux(*)
       type(c_ptr), value :: imcir, cmask        ! c_null_ptr when the cube is not wanted
       real(c_double), intent(out) :: tau_center(*), velo(*)
       integer(c_int), intent(out) :: maserflag(*)
     end function
     ! multi-GPU by camera-ring block (one process per GPU): rings ring_lo..ring_hi, 0 = central beam
     integer(c_int) function rl_render_rings(ctx, iline0, nl, nfr, vmax_kms, dist_cm, ring_lo, ring_hi, &
          ringsum, imcir) bind(c, name='rl_render_rings')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: iline0, nl, nfr, ring_lo, ring_hi
       real(c_double), value :: vmax_kms, dist_cm
       real(c_double), intent(out) :: ringsum(*)   ! (nfr, 0:nrr, nl)
       type(c_ptr), value :: imcir                 ! c_null_ptr when the cube is not wanted
     end function
     integer(c_int) function rl_flux_from_rings(ctx, nl, nfr, dist_cm, ringsum, flux) &
          bind(c, name='rl_flux_from_rings')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: nl, nfr
       real(c_double), value :: dist_cm
       real(c_double), intent(in) :: ringsum(*)
       real(c_double), intent(out) :: flux(*)
     end function
  end interface
end module rl_capi_mod
