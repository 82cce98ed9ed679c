!! fedem_b200_mod.f90 -- ISO_C_BINDING interface to libfedem_b200.so (include/fedem_b200.h).
!!
!! This is the thin layer the reference's Fortran host code (src/vpmStress/stress.f90, gage.f90,
!! src/vpmSolver/stressRecoveryModule.f90) uses to call the B200 kernels: every C entry point is
!! bound by its exact name; derived types mirror the C structs member by member.  Index arrays
!! are passed exactly as SamType holds them (1-based, src/vpmCommon/samModule.f90:27-66); real
!! arrays are real(c_double) = real(dp).  See fortran/stress_b200_driver.f90 for the call
!! sequence that replaces the time loop of stress.f90:357-435 and INTEGRATION.md for the build.
!!
!! No Fortran compiler exists in the build image, so this file is delivered as source;
!! tests/test_abi_cpu.py checks that it binds every symbol the header declares, and
!! tests/c_abi_driver.c exercises the identical symbols from plain C.

module fedem_b200_mod

  use, intrinsic :: iso_c_binding

  implicit none

  integer(c_int), parameter :: FSR_OK = 0, FSR_ERR_ARG = -1, FSR_ERR_CUDA = -2
  integer(c_int), parameter :: FSR_ERR_ALLOC = -3, FSR_ERR_STATE = -4, FSR_ERR_LIMIT = -5
  integer(c_int), parameter :: FSR_NBEAM = 32
  integer(c_int), parameter :: FSR_HIST_GAGE_MAJOR = 0, FSR_HIST_STEP_MAJOR = 1
  integer(c_int), parameter :: FSR_GAGE_NVAL = 24

  !> struct fsr_sam: the SamType subset read by initiateSAM (samStressModule.f90:273-316)
  type, bind(C) :: fsr_sam
     integer(c_int) :: nnod, nel, ndof, ndof1, ndof2, ngen, neq, nceq, nmmnpc, nmmceq
     type(c_ptr)    :: madof, msc, mpmnpc, mmnpc, melcon, mpmceq, mmceq, ttcc, meqn, meqn1, meqn2
  end type fsr_sam

  !> struct fsr_elmdata: what ffl_getcoor/getmat/getthick/getbeamsection/getpinflags/getelmid return
  type, bind(C) :: fsr_elmdata
     type(c_ptr) :: xyz, emod, rny, thk, elmid, beam
  end type fsr_elmdata

  !> struct fsr_options
  type, bind(C) :: fsr_options
     integer(c_int) :: device, stressForm, step_tile
     integer(c_int) :: reserved(5)
  end type fsr_options

  !> struct fsr_rosette: one strain rosette (strainGageModule.f90:184-237, &STRAIN_ROSETTE)
  type, bind(C) :: fsr_rosette
     integer(c_int) :: id, numnod, ngage, zero_init
     integer(c_int) :: nodes(4)
     real(c_double) :: rpos(12), zpos, emod, nu, alpha_gages, gate
     real(c_double) :: sncurve(4)
  end type fsr_rosette

  !> struct fsr_strain_coat: one strain coat element as ffl_getstraincoat delivers it
  type, bind(C) :: fsr_strain_coat
     integer(c_int) :: id, nnod, npts, elm_id
     integer(c_int) :: nodes(8)
     integer(c_int) :: mat_id(3), res_set(3), sn_curve(2,3)
     real(c_double) :: emod(3), nu(3), zpos(3), scf(3)
  end type fsr_strain_coat

  !> struct fsr_rdb_options: what writeStressHeader takes from the command line and from sup%id
  type, bind(C) :: fsr_rdb_options
     integer(c_int) :: out_mask, double_precision, rdbinc, part_base_id, part_user_id
     type(c_ptr)    :: part_descr, model_file, link_file   !< NUL-terminated C strings (c_loc of a c_char array)
     type(c_ptr)    :: elmid                               !< integer(c_int) elmid(nel)
     type(c_ptr)    :: module_name
     type(c_ptr)    :: minex                               !< integer(c_int) minex(nnod)
     type(c_ptr)    :: sup_tr_init                         !< real(c_double) supTrInit(3,4)
  end type fsr_rdb_options

  interface

     ! ---- life cycle ----------------------------------------------------------------------
     function fsr_part_create (part, sam, elm, opt) bind(C,name="fsr_part_create") result(ierr)
       import :: c_ptr, c_int, fsr_sam, fsr_elmdata, fsr_options
       type(c_ptr)      , intent(out) :: part
       type(fsr_sam)    , intent(in)  :: sam
       type(fsr_elmdata), intent(in)  :: elm
       type(fsr_options), intent(in)  :: opt
       integer(c_int) :: ierr
     end function fsr_part_create

     function fsr_set_recovery (part, B, ldB, E, ldE) bind(C,name="fsr_set_recovery") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: part
       real(c_double), intent(in) :: B(*), E(*)
       integer(c_int), value      :: ldB, ldE
       integer(c_int) :: ierr
     end function fsr_set_recovery

     subroutine fsr_part_destroy (part) bind(C,name="fsr_part_destroy")
       import :: c_ptr
       type(c_ptr), value :: part
     end subroutine fsr_part_destroy

     function fsr_set_stream (part, stream) bind(C,name="fsr_set_stream") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part, stream
       integer(c_int) :: ierr
     end function fsr_set_stream

     ! ---- sizes -----------------------------------------------------------------------------
     function fsr_num_result_points (part) bind(C,name="fsr_num_result_points") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: n
     end function fsr_num_result_points

     function fsr_result_point_offsets (part, off) bind(C,name="fsr_result_point_offsets") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr)   , value       :: part
       integer(c_int), intent(out) :: off(*)
       integer(c_int) :: ierr
     end function fsr_result_point_offsets

     function fsr_ndim (part) bind(C,name="fsr_ndim") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: n
     end function fsr_ndim

     function fsr_vms_size (part) bind(C,name="fsr_vms_size") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: n
     end function fsr_vms_size

     ! ---- the hot path ----------------------------------------------------------------------
     function fsr_recover (part, Q, ldq, nsteps, vm_hist) bind(C,name="fsr_recover") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: part
       real(c_double), intent(in) :: Q(ldq,*)
       integer(c_int), value      :: ldq, nsteps
       type(c_ptr)   , value      :: vm_hist   !< c_loc of a real(dp) array, or c_null_ptr
       integer(c_int) :: ierr
     end function fsr_recover

     function fsr_recover_dev (part, Q_dev, ldq, nsteps, vm_hist_dev, ld_vm, stream) &
          &                   bind(C,name="fsr_recover_dev") result(ierr)
       import :: c_ptr, c_int, c_size_t
       type(c_ptr)      , value :: part, Q_dev, vm_hist_dev, stream
       integer(c_int)   , value :: ldq, nsteps
       integer(c_size_t), value :: ld_vm
       integer(c_int) :: ierr
     end function fsr_recover_dev

     function fsr_reset_envelope (part) bind(C,name="fsr_reset_envelope") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: ierr
     end function fsr_reset_envelope

     function fsr_get_envelope (part, vm_max, vm_min) bind(C,name="fsr_get_envelope") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: part
       real(c_double), intent(out) :: vm_max(*), vm_min(*)
       integer(c_int) :: ierr
     end function fsr_get_envelope

     function fsr_envelope_dev (part, vm_max_dev, vm_min_dev) bind(C,name="fsr_envelope_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value       :: part
       type(c_ptr), intent(out) :: vm_max_dev, vm_min_dev
       integer(c_int) :: ierr
     end function fsr_envelope_dev

     function fsr_copy_envelope_dev (part, vm_max_dst_dev, vm_min_dst_dev, stream) &
          &                         bind(C,name="fsr_copy_envelope_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part, vm_max_dst_dev, vm_min_dst_dev, stream
       integer(c_int) :: ierr
     end function fsr_copy_envelope_dev

     function fsr_recover_step_full (part, q, resmat, stress, strain, sres, sv) &
          &                         bind(C,name="fsr_recover_step_full") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: part
       real(c_double), intent(in) :: q(*)
       type(c_ptr)   , value      :: resmat, stress, strain, sres, sv  !< c_loc(array) or c_null_ptr
       integer(c_int) :: ierr
     end function fsr_recover_step_full

     function fsr_get_vms (part, q, vms, nvms) bind(C,name="fsr_get_vms") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: part
       real(c_double), intent(in)  :: q(*)
       real(c_double), intent(out) :: vms(*)
       integer(c_int), value       :: nvms
       integer(c_int) :: ierr
     end function fsr_get_vms

     function fsr_expand (part, Q, ldq, nsteps, U_host) bind(C,name="fsr_expand") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: part
       real(c_double), intent(in)  :: Q(ldq,*)
       integer(c_int), value       :: ldq, nsteps
       real(c_double), intent(out) :: U_host(*)
       integer(c_int) :: ierr
     end function fsr_expand

     function fsr_expand_rows (part, Q, ldq, nsteps, rows, nrows, out) bind(C,name="fsr_expand_rows") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: part
       integer(c_int), value       :: ldq, nsteps, nrows
       real(c_double), intent(in)  :: Q(ldq,*)
       integer(c_int), intent(in)  :: rows(*)       ! 0-based nodal DOF indices
       real(c_double), intent(out) :: out(nrows,*)  ! (nrows, nsteps)
       integer(c_int) :: ierr
     end function fsr_expand_rows

     ! ---- strain gages (fedem_gage path) ----------------------------------------------------
     function fsr_gage_create (gages, part, ros, nros) bind(C,name="fsr_gage_create") result(ierr)
       import :: c_ptr, c_int, fsr_rosette
       type(c_ptr)      , intent(out) :: gages
       type(c_ptr)      , value       :: part
       type(fsr_rosette), intent(in)  :: ros(*)
       integer(c_int)   , value       :: nros
       integer(c_int) :: ierr
     end function fsr_gage_create

     function fsr_gage_num_series (gages) bind(C,name="fsr_gage_num_series") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: gages
       integer(c_int) :: n
     end function fsr_gage_num_series

     function fsr_gage_get_bcart (gages, bcart) bind(C,name="fsr_gage_get_bcart") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: gages
       real(c_double), intent(out) :: bcart(*)
       integer(c_int) :: ierr
     end function fsr_gage_get_bcart

     function fsr_gage_recover (gages, Q, ldq, nsteps, values) bind(C,name="fsr_gage_recover") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: gages
       real(c_double), intent(in) :: Q(ldq,*)
       integer(c_int), value      :: ldq, nsteps
       type(c_ptr)   , value      :: values    !< c_loc of real(dp) (FSR_GAGE_NVAL,nros,nsteps) or c_null_ptr
       integer(c_int) :: ierr
     end function fsr_gage_recover

     function fsr_gage_recover_dev (gages, Q_dev, ldq, nsteps, values_dev, stream) &
          &                        bind(C,name="fsr_gage_recover_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr)   , value :: gages, Q_dev, values_dev, stream
       integer(c_int), value :: ldq, nsteps
       integer(c_int) :: ierr
     end function fsr_gage_recover_dev

     function fsr_gage_fatigue (gages, Q, ldq, nsteps, to_mpa, gate, curve, bin_size, nbins, damage, &
          &                    ncycles, bins, status) bind(C,name="fsr_gage_fatigue") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: gages
       real(c_double), intent(in) :: Q(ldq,*), curve(4)
       integer(c_int), value      :: ldq, nsteps, nbins
       real(c_double), value      :: to_mpa, gate, bin_size
       type(c_ptr)   , value      :: damage, ncycles, bins, status
       integer(c_int) :: ierr
     end function fsr_gage_fatigue

     function fsr_gage_fatigue_begin (gages, to_mpa, default_gate, default_curve, bin_size, nbins, &
          &                          stack_cap) bind(C,name="fsr_gage_fatigue_begin") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: gages
       real(c_double), value      :: to_mpa, default_gate, bin_size
       real(c_double), intent(in) :: default_curve(4)
       integer(c_int), value      :: nbins, stack_cap
       integer(c_int) :: ierr
     end function fsr_gage_fatigue_begin

     function fsr_gage_fatigue_feed_dev (gages, Q_dev, ldq, step0, nsteps, mode, n_pending, stream) &
          &                             bind(C,name="fsr_gage_fatigue_feed_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr)   , value :: gages, Q_dev, n_pending, stream
       integer(c_int), value :: ldq, step0, nsteps, mode
       integer(c_int) :: ierr
     end function fsr_gage_fatigue_feed_dev

     function fsr_gage_fatigue_end (gages, damage, ncycles, bins, status) &
          &                        bind(C,name="fsr_gage_fatigue_end") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: gages, damage, ncycles, bins, status
       integer(c_int) :: ierr
     end function fsr_gage_fatigue_end

     subroutine fsr_gage_destroy (gages) bind(C,name="fsr_gage_destroy")
       import :: c_ptr
       type(c_ptr), value :: gages
     end subroutine fsr_gage_destroy

     ! ---- strain coat summary: replaces calcStrainCoatData / calcAngleData / BiAxMean / BiAxStdDev (strainCoatModule.f90) ----
     function fsr_coat_begin (gages, angle_bins, biaxial_gate) bind(C,name="fsr_coat_begin") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value :: gages
       integer(c_int), value :: angle_bins
       real(c_double), value :: biaxial_gate
       integer(c_int) :: ierr
     end function fsr_coat_begin

     function fsr_coat_feed (gages, Q, ldq, nsteps) bind(C,name="fsr_coat_feed") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: gages
       integer(c_int), value      :: ldq, nsteps
       real(c_double), intent(in) :: Q(ldq,*)
       integer(c_int) :: ierr
     end function fsr_coat_feed

     function fsr_coat_feed_dev (gages, Q_dev, ldq, nsteps, stream) bind(C,name="fsr_coat_feed_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr)   , value :: gages, Q_dev, stream
       integer(c_int), value :: ldq, nsteps
       integer(c_int) :: ierr
     end function fsr_coat_feed_dev

     function fsr_coat_end (gages, env, summary, nbiax) bind(C,name="fsr_coat_end") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: gages
       type(c_ptr), value :: env, summary, nbiax   ! real(c_double) env(nros,8), summary(nros,6), integer(c_int) nbiax(nros), or c_null_ptr
       integer(c_int) :: ierr
     end function fsr_coat_end

     function fsr_gage_set_coat_fatigue (gages, scf) bind(C,name="fsr_gage_set_coat_fatigue") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: gages
       type(c_ptr), value :: scf   ! real(c_double) scf(nros), or c_null_ptr
       integer(c_int) :: ierr
     end function fsr_gage_set_coat_fatigue

     function fsr_sn_read (lib, path) bind(C,name="fsr_sn_read") result(ierr)
       import :: c_ptr, c_int, c_char
       type(c_ptr), intent(out) :: lib
       character(kind=c_char), intent(in) :: path(*)
       integer(c_int) :: ierr
     end function fsr_sn_read

     subroutine fsr_sn_free (lib) bind(C,name="fsr_sn_free")
       import :: c_ptr
       type(c_ptr), value :: lib
     end subroutine fsr_sn_free

     function fsr_sn_num_standards (lib) bind(C,name="fsr_sn_num_standards") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: lib
       integer(c_int) :: n
     end function fsr_sn_num_standards

     function fsr_sn_num_curves (lib, std_index) bind(C,name="fsr_sn_num_curves") result(n)
       import :: c_ptr, c_int
       type(c_ptr)   , value :: lib
       integer(c_int), value :: std_index
       integer(c_int) :: n
     end function fsr_sn_num_curves

     function fsr_sn_get (lib, std_index, curve_index, std_id, loga, m, logN0, cap) bind(C,name="fsr_sn_get") result(nseg)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: lib
       integer(c_int), value       :: std_index, curve_index, cap
       integer(c_int), intent(out) :: std_id
       real(c_double), intent(out) :: loga(*), m(*), logN0(*)
       integer(c_int) :: nseg
     end function fsr_sn_get

     function fsr_sn_value (lib, std_index, curve_index, s) bind(C,name="fsr_sn_value") result(n)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value :: lib
       integer(c_int), value :: std_index, curve_index
       real(c_double), value :: s
       real(c_double) :: n
     end function fsr_sn_value

     ! ---- fatigue (ffp_addpoint / ffp_getdamage / ffp_getnumcycles) ---------------------------
     function fsr_fatigue (device, hist, ngage, nsteps, gate, curve, bin_size, nbins, damage, &
          &               ncycles, bins) bind(C,name="fsr_fatigue") result(ierr)
       import :: c_ptr, c_int, c_double
       integer(c_int), value       :: device, ngage, nsteps, nbins
       real(c_double), intent(in)  :: hist(nsteps,*), curve(4)
       real(c_double), value       :: gate, bin_size
       real(c_double), intent(out) :: damage(*)
       integer(c_int), intent(out) :: ncycles(*)
       type(c_ptr)   , value       :: bins
       integer(c_int) :: ierr
     end function fsr_fatigue

     function fsr_fatigue_dev (device, hist_dev, ld_hist, ngage, nsteps, gate, curve, bin_size, nbins, &
          &                   damage_dev, ncycles_dev, bins_dev, stream) &
          &                   bind(C,name="fsr_fatigue_dev") result(ierr)
       import :: c_ptr, c_int, c_double, c_size_t
       integer(c_int)   , value      :: device, ngage, nsteps, nbins
       type(c_ptr)      , value      :: hist_dev, damage_dev, ncycles_dev, bins_dev, stream
       integer(c_size_t), value      :: ld_hist
       real(c_double)   , value      :: gate, bin_size
       real(c_double)   , intent(in) :: curve(4)
       integer(c_int) :: ierr
     end function fsr_fatigue_dev

     function fsr_fatigue_create (f, device, ngage, gate, curve, bin_size, nbins, stack_cap) &
          &                      bind(C,name="fsr_fatigue_create") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , intent(out) :: f
       integer(c_int), value       :: device, ngage, nbins, stack_cap
       real(c_double), value       :: gate, bin_size
       real(c_double), intent(in)  :: curve(4)
       integer(c_int) :: ierr
     end function fsr_fatigue_create

     function fsr_fatigue_set_gage_params (f, gate, curve) bind(C,name="fsr_fatigue_set_gage_params") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: f, gate, curve
       integer(c_int) :: ierr
     end function fsr_fatigue_set_gage_params

     function fsr_fatigue_reset (f) bind(C,name="fsr_fatigue_reset") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: f
       integer(c_int) :: ierr
     end function fsr_fatigue_reset

     function fsr_fatigue_locate_dev (f, hist_dev, ld, layout, step0, nsteps, n_pending, stream) &
          &                          bind(C,name="fsr_fatigue_locate_dev") result(ierr)
       import :: c_ptr, c_int, c_size_t
       type(c_ptr)      , value :: f, hist_dev, stream
       integer(c_size_t), value :: ld
       integer(c_int)   , value :: layout, step0, nsteps
       integer(c_int)   , intent(out) :: n_pending
       integer(c_int) :: ierr
     end function fsr_fatigue_locate_dev

     function fsr_fatigue_feed_dev (f, hist_dev, ld, layout, step0, nsteps, stream) &
          &                        bind(C,name="fsr_fatigue_feed_dev") result(ierr)
       import :: c_ptr, c_int, c_size_t
       type(c_ptr)      , value :: f, hist_dev, stream
       integer(c_size_t), value :: ld
       integer(c_int)   , value :: layout, step0, nsteps
       integer(c_int) :: ierr
     end function fsr_fatigue_feed_dev

     function fsr_fatigue_finish (f, damage, ncycles, bins, status) bind(C,name="fsr_fatigue_finish") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: f, damage, ncycles, bins, status
       integer(c_int) :: ierr
     end function fsr_fatigue_finish

     function fsr_fatigue_finish_dev (f, stream) bind(C,name="fsr_fatigue_finish_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: f, stream
       integer(c_int) :: ierr
     end function fsr_fatigue_finish_dev

     function fsr_fatigue_results_dev (f, damage_dev, ncycles_dev, bins_dev, status_dev) &
          &                           bind(C,name="fsr_fatigue_results_dev") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value       :: f
       type(c_ptr), intent(out) :: damage_dev, ncycles_dev, bins_dev, status_dev
       integer(c_int) :: ierr
     end function fsr_fatigue_results_dev

     subroutine fsr_fatigue_destroy (f) bind(C,name="fsr_fatigue_destroy")
       import :: c_ptr
       type(c_ptr), value :: f
     end subroutine fsr_fatigue_destroy

     ! ---- file formats and history assembly (host only) ---------------------------------------
     function fsr_fmx_write (path, tag, checksum, A, n, single_precision) bind(C,name="fsr_fmx_write") result(ierr)
       import :: c_char, c_int, c_double, c_long_long
       character(kind=c_char), intent(in) :: path(*), tag(*)   !< NUL-terminated
       integer(c_int)        , value      :: checksum, single_precision
       real(c_double)        , intent(in) :: A(*)
       integer(c_long_long)  , value      :: n
       integer(c_int) :: ierr
     end function fsr_fmx_write

     function fsr_fmx_read (path, tag_out, tag_cap, checksum, is_single, A, n) bind(C,name="fsr_fmx_read") result(ierr)
       import :: c_char, c_int, c_double, c_long_long
       character(kind=c_char), intent(in)  :: path(*)
       character(kind=c_char), intent(out) :: tag_out(*)
       integer(c_int)        , value       :: tag_cap
       integer(c_int)        , intent(out) :: checksum, is_single
       real(c_double)        , intent(out) :: A(*)
       integer(c_long_long)  , value       :: n
       integer(c_int) :: ierr
     end function fsr_fmx_read

     function fsr_fsm_read_mpar (path, checksum, mpar, cap) bind(C,name="fsr_fsm_read_mpar") result(npar)
       import :: c_char, c_int
       character(kind=c_char), intent(in)  :: path(*)
       integer(c_int)        , intent(out) :: checksum, mpar(*)
       integer(c_int)        , value       :: cap
       integer(c_int) :: npar
     end function fsr_fsm_read_mpar

     function fsr_fsm_read (path, madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon, mpmceq, mmceq, ttcc, &
          &                meqn, meqn1, meqn2) bind(C,name="fsr_fsm_read") result(ierr)
       import :: c_char, c_ptr, c_int
       character(kind=c_char), intent(in) :: path(*)
       type(c_ptr), value :: madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon, mpmceq, mmceq, ttcc, meqn, meqn1, meqn2
       integer(c_int) :: ierr
     end function fsr_fsm_read

     function fsr_fsm_write (path, checksum, npar, mpar, madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon, &
          &                 mpmceq, mmceq, ttcc, meqn, meqn1, meqn2) bind(C,name="fsr_fsm_write") result(ierr)
       import :: c_char, c_ptr, c_int
       character(kind=c_char), intent(in) :: path(*)
       integer(c_int)        , value      :: checksum, npar
       integer(c_int)        , intent(in) :: mpar(*)
       type(c_ptr), value :: madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon, mpmceq, mmceq, ttcc, meqn, meqn1, meqn2
       integer(c_int) :: ierr
     end function fsr_fsm_write

     function fsr_build_finit (nsteps, ntriads, sup_tr, triad_ur, tr_undef, ndofs, first_dof, ngen, gen_ur, &
          &                   gen_first_dof, Q, ldq) bind(C,name="fsr_build_finit") result(ierr)
       import :: c_int, c_double
       integer(c_int), value       :: nsteps, ntriads, ngen, gen_first_dof, ldq
       real(c_double), intent(in)  :: sup_tr(3,4,*), triad_ur(3,4,ntriads,*), tr_undef(3,4,*), gen_ur(ngen,*)
       integer(c_int), intent(in)  :: ndofs(*), first_dof(*)
       real(c_double), intent(out) :: Q(ldq,*)
       integer(c_int) :: ierr
     end function fsr_build_finit

     ! readSupElModes (modesRoutines.f90:121-203): eigenvector components -> columns of Q for a mode-shape expansion
     function fsr_build_mode_finit (ntriads, supTr, ndofs, first_dof, triad_eig, ngen, gen_first_dof, gen_eig, ncomp, Q, ldq) &
          &                        bind(C,name="fsr_build_mode_finit") result(ierr)
       import :: c_int, c_double
       integer(c_int), value       :: ntriads, ngen, gen_first_dof, ncomp, ldq
       real(c_double), intent(in)  :: supTr(3,4), triad_eig(*), gen_eig(*)
       integer(c_int), intent(in)  :: ndofs(*), first_dof(*)
       real(c_double), intent(out) :: Q(ldq,*)
       integer(c_int) :: ierr
     end function fsr_build_mode_finit

     ! ---- .frs results database (replaces ffr_init/ffr_findptr/ffr_getdata for the recovery path) ----
     function fsr_frs_open (db, paths, nfiles) bind(C,name="fsr_frs_open") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr)   , intent(out) :: db
       type(c_ptr)   , intent(in)  :: paths(*)   !< c_loc of NUL-terminated file names
       integer(c_int), value       :: nfiles
       integer(c_int) :: ierr
     end function fsr_frs_open

     subroutine fsr_frs_close (db) bind(C,name="fsr_frs_close")
       import :: c_ptr
       type(c_ptr), value :: db
     end subroutine fsr_frs_close

     function fsr_frs_num_steps (db) bind(C,name="fsr_frs_num_steps") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: db
       integer(c_int) :: n
     end function fsr_frs_num_steps

     function fsr_frs_get_steps (db, stepno, time, cap) bind(C,name="fsr_frs_get_steps") result(n)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: db
       integer(c_int), intent(out) :: stepno(*)
       real(c_double), intent(out) :: time(*)
       integer(c_int), value       :: cap
       integer(c_int) :: n
     end function fsr_frs_get_steps

     function fsr_frs_find (db, var_path, og_type, base_id) bind(C,name="fsr_frs_find") result(handle)
       import :: c_ptr, c_char, c_int
       type(c_ptr)           , value      :: db
       character(kind=c_char), intent(in) :: var_path(*), og_type(*)
       integer(c_int)        , value      :: base_id
       integer(c_int) :: handle
     end function fsr_frs_find

     function fsr_frs_var_size (db, handle) bind(C,name="fsr_frs_var_size") result(n)
       import :: c_ptr, c_int
       type(c_ptr)   , value :: db
       integer(c_int), value :: handle
       integer(c_int) :: n
     end function fsr_frs_var_size

     function fsr_frs_read (db, handle, step0, nsteps, data, nw, ld) bind(C,name="fsr_frs_read") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: db
       integer(c_int), value       :: handle, step0, nsteps, nw, ld
       real(c_double), intent(out) :: data(ld,*)
       integer(c_int) :: ierr
     end function fsr_frs_read

     function fsr_frs_reduced_history (db, sup_base_id, ntriads, triad_base_id, ndofs, first_dof, tr_undef, ngen, &
          &                           gen_first_dof, step0, nsteps, Q, ldq) &
          &                           bind(C,name="fsr_frs_reduced_history") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: db
       integer(c_int), value       :: sup_base_id, ntriads, ngen, gen_first_dof, step0, nsteps, ldq
       integer(c_int), intent(in)  :: triad_base_id(*), ndofs(*), first_dof(*)
       real(c_double), intent(in)  :: tr_undef(3,4,*)
       real(c_double), intent(out) :: Q(ldq,*)
       integer(c_int) :: ierr
     end function fsr_frs_reduced_history

     function fsr_frs_create (w, path, checksum, header_text, payload_bytes) bind(C,name="fsr_frs_create") result(ierr)
       import :: c_ptr, c_char, c_int, c_long_long
       type(c_ptr)           , intent(out) :: w
       character(kind=c_char), intent(in)  :: path(*), header_text(*)
       integer(c_int)        , value       :: checksum
       integer(c_long_long)  , value       :: payload_bytes
       integer(c_int) :: ierr
     end function fsr_frs_create

     function fsr_frs_create_tagged (w, path, tag, checksum, header_text, payload_bytes) &
          &                         bind(C,name="fsr_frs_create_tagged") result(ierr)
       import :: c_ptr, c_char, c_int, c_long_long
       type(c_ptr)           , intent(out) :: w
       character(kind=c_char), intent(in)  :: path(*), tag(*), header_text(*)
       integer(c_int)        , value       :: checksum
       integer(c_long_long)  , value       :: payload_bytes
       integer(c_int) :: ierr
     end function fsr_frs_create_tagged

     function fsr_frs_write_step (w, stepno, time, payload) bind(C,name="fsr_frs_write_step") result(n)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value :: w, payload
       integer(c_int), value :: stepno
       real(c_double), value :: time
       integer(c_int) :: n
     end function fsr_frs_write_step

     function fsr_frs_finish (w) bind(C,name="fsr_frs_finish") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: w
       integer(c_int) :: ierr
     end function fsr_frs_finish

     ! ---- FE part file (.ftl): replaces ffl_init + the per-element ffl_get* calls -------------
     function fsr_ftl_open (ftl, path) bind(C,name="fsr_ftl_open") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr)           , intent(out) :: ftl
       character(kind=c_char), intent(in)  :: path(*)
       integer(c_int) :: ierr
     end function fsr_ftl_open

     subroutine fsr_ftl_close (ftl) bind(C,name="fsr_ftl_close")
       import :: c_ptr
       type(c_ptr), value :: ftl
     end subroutine fsr_ftl_close

     function fsr_ftl_version (ftl) bind(C,name="fsr_ftl_version") result(iver)
       import :: c_ptr, c_int
       type(c_ptr), value :: ftl
       integer(c_int) :: iver
     end function fsr_ftl_version

     function fsr_ftl_activate_groups (ftl, groups) bind(C,name="fsr_ftl_activate_groups") result(nignored)
       import :: c_ptr, c_char, c_int
       type(c_ptr)           , value      :: ftl
       character(kind=c_char), intent(in) :: groups(*)
       integer(c_int) :: nignored
     end function fsr_ftl_activate_groups

     function fsr_ftl_sizes (ftl, sz) bind(C,name="fsr_ftl_sizes") result(nael)
       import :: c_ptr, c_int
       type(c_ptr)   , value       :: ftl
       integer(c_int), intent(out) :: sz(12)
       integer(c_int) :: nael
     end function fsr_ftl_sizes

     function fsr_ftl_get_nodes (ftl, madof, minex, mnode, msc, xyz) bind(C,name="fsr_ftl_get_nodes") result(nnod)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: ftl
       integer(c_int), intent(out) :: madof(*), minex(*), mnode(*), msc(*)
       real(c_double), intent(out) :: xyz(3,*)
       integer(c_int) :: nnod
     end function fsr_ftl_get_nodes

     function fsr_ftl_get_topology (ftl, use_andes, melcon, mpmnpc, mmnpc) bind(C,name="fsr_ftl_get_topology") result(nel)
       import :: c_ptr, c_int
       type(c_ptr)   , value       :: ftl
       integer(c_int), value       :: use_andes
       integer(c_int), intent(out) :: melcon(*), mpmnpc(*), mmnpc(*)
       integer(c_int) :: nel
     end function fsr_ftl_get_topology

     function fsr_ftl_get_elmdata (ftl, emod, rny, rho, thk, elmid, beam, status) &
          &                       bind(C,name="fsr_ftl_get_elmdata") result(nbad)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: ftl
       real(c_double), intent(out) :: emod(*), rny(*), rho(*), thk(*), beam(32,*)
       integer(c_int), intent(out) :: elmid(*), status(*)
       integer(c_int) :: nbad
     end function fsr_ftl_get_elmdata

     function fsr_ftl_ext2int (ftl, is_node, id) bind(C,name="fsr_ftl_ext2int") result(intid)
       import :: c_ptr, c_int
       type(c_ptr)   , value :: ftl
       integer(c_int), value :: is_node, id
       integer(c_int) :: intid
     end function fsr_ftl_ext2int

     function fsr_ftl_num_strain_coats (ftl) bind(C,name="fsr_ftl_num_strain_coats") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: ftl
       integer(c_int) :: n
     end function fsr_ftl_num_strain_coats

     function fsr_ftl_get_strain_coats (ftl, coats, cap) bind(C,name="fsr_ftl_get_strain_coats") result(n)
       import :: c_ptr, c_int, fsr_strain_coat
       type(c_ptr)          , value       :: ftl
       type(fsr_strain_coat), intent(out) :: coats(*)
       integer(c_int)       , value       :: cap
       integer(c_int) :: n
     end function fsr_ftl_get_strain_coats

     ! ---- solver input file (.fsi): replaces readSolverData ----------------------------------
     function fsr_fsi_open (fsi, path, part_base_id) bind(C,name="fsr_fsi_open") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr)           , intent(out) :: fsi
       character(kind=c_char), intent(in)  :: path(*)
       integer(c_int)        , value       :: part_base_id
       integer(c_int) :: ierr
     end function fsr_fsi_open

     subroutine fsr_fsi_close (fsi) bind(C,name="fsr_fsi_close")
       import :: c_ptr
       type(c_ptr), value :: fsi
     end subroutine fsr_fsi_close

     function fsr_fsi_part (fsi, user_id, descr, dcap, ntriads, ngen, supPos, gravity, model_file, mcap) &
          &                bind(C,name="fsr_fsi_part") result(base_id)
       import :: c_ptr, c_char, c_int, c_double
       type(c_ptr)           , value       :: fsi
       integer(c_int)        , intent(out) :: user_id, ntriads, ngen
       character(kind=c_char), intent(out) :: descr(*), model_file(*)
       integer(c_int)        , value       :: dcap, mcap
       real(c_double)        , intent(out) :: supPos(3,4), gravity(3)
       integer(c_int) :: base_id
     end function fsr_fsi_part

     function fsr_fsi_triads (fsi, base_id, user_id, ndofs, first_dof, trUndeformed, ur) &
          &                  bind(C,name="fsr_fsi_triads") result(gen_first_dof)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: fsi
       integer(c_int), intent(out) :: base_id(*), user_id(*), ndofs(*), first_dof(*)
       real(c_double), intent(out) :: trUndeformed(3,4,*), ur(3,4,*)
       integer(c_int) :: gen_first_dof
     end function fsr_fsi_triads

     ! ReadStrainGages (strainGageModule.f90:78-237): the &STRAIN_ROSETTE records of a rosette input file
     function fsr_fsi_read_rosettes (path, link_base_id, ros, user_id, descr, descr_stride, cap) &
          &                         bind(C,name="fsr_fsi_read_rosettes") result(nros)
       import :: c_ptr, c_char, c_int
       character(kind=c_char), intent(in) :: path(*)
       integer(c_int), value :: link_base_id, descr_stride, cap
       type(c_ptr)   , value :: ros      ! type(fsr_rosette) array, or c_null_ptr to count
       type(c_ptr)   , value :: user_id  ! integer(c_int) array or c_null_ptr
       type(c_ptr)   , value :: descr    ! character buffer cap*descr_stride or c_null_ptr
       integer(c_int) :: nros
     end function fsr_fsi_read_rosettes

     ! ---- stress results database: replaces writeStressHeader + writeStressDB/writeStrMeasureDB ----
     function fsr_rdb_create (rdb, part, path, opt) bind(C,name="fsr_rdb_create") result(ierr)
       import :: c_ptr, c_char, c_int, fsr_rdb_options
       type(c_ptr)           , intent(out) :: rdb
       type(c_ptr)           , value       :: part
       character(kind=c_char), intent(in)  :: path(*)
       type(fsr_rdb_options) , intent(in)  :: opt
       integer(c_int) :: ierr
     end function fsr_rdb_create

     function fsr_rdb_build_header (nnod, madof, nel, melcon, opt, header, cap, step_bytes) &
          &                        bind(C,name="fsr_rdb_build_header") result(nchar)
       import :: c_ptr, c_char, c_int, c_long_long, fsr_rdb_options
       integer(c_int)        , value       :: nnod, nel, cap
       integer(c_int)        , intent(in)  :: madof(*), melcon(*)
       type(fsr_rdb_options) , intent(in)  :: opt
       character(kind=c_char), intent(out) :: header(*)
       integer(c_long_long)  , intent(out) :: step_bytes
       integer(c_int) :: nchar
     end function fsr_rdb_build_header

     function fsr_rdb_step_bytes (rdb) bind(C,name="fsr_rdb_step_bytes") result(nbytes)
       import :: c_ptr, c_long_long
       type(c_ptr), value :: rdb
       integer(c_long_long) :: nbytes
     end function fsr_rdb_step_bytes

     function fsr_rdb_header (rdb, buf, cap) bind(C,name="fsr_rdb_header") result(nchar)
       import :: c_ptr, c_char, c_int
       type(c_ptr)           , value       :: rdb
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int)        , value       :: cap
       integer(c_int) :: nchar
     end function fsr_rdb_header

     function fsr_rdb_path (rdb, buf, cap) bind(C,name="fsr_rdb_path") result(nchar)
       import :: c_ptr, c_char, c_int
       type(c_ptr)           , value       :: rdb
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int)        , value       :: cap
       integer(c_int) :: nchar
     end function fsr_rdb_path

     !> Q(ldq,nsteps) = [finit; vg] of every step of the window, supTr(3,4,nsteps) (only read when the
     !> total displacements are written)
     function fsr_rdb_write_steps (rdb, Q, ldq, nsteps, stepno, time, supTr) &
          &                       bind(C,name="fsr_rdb_write_steps") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value      :: rdb
       real(c_double), intent(in) :: Q(ldq,*), time(*), supTr(3,4,*)
       integer(c_int), value      :: ldq, nsteps
       integer(c_int), intent(in) :: stepno(*)
       integer(c_int) :: ierr
     end function fsr_rdb_write_steps

     subroutine fsr_total_nodal_displacement (x0, u, nd, T, T0, utot) bind(C,name="fsr_total_nodal_displacement")
       import :: c_int, c_double
       real(c_double), intent(in)  :: x0(3), u(*), T(3,4), T0(3,4)
       integer(c_int), value       :: nd
       real(c_double), intent(out) :: utot(*)
     end subroutine fsr_total_nodal_displacement

     function fsr_rdb_create_group (rdb, group, path, opt) bind(C,name="fsr_rdb_create_group") result(ierr)
       import :: c_ptr, c_int, c_char, fsr_rdb_options
       type(c_ptr), intent(out) :: rdb
       type(c_ptr), value       :: group
       character(kind=c_char), intent(in) :: path(*)
       type(fsr_rdb_options), intent(in)  :: opt
       integer(c_int) :: ierr
     end function fsr_rdb_create_group

     function fsr_rdb_flush (rdb, t, n) bind(C,name="fsr_rdb_flush") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value     :: rdb
       real(c_double), intent(out) :: t(*)
       integer(c_int), value  :: n
       integer(c_int) :: ierr
     end function fsr_rdb_flush

     function fsr_rdb_close (rdb) bind(C,name="fsr_rdb_close") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: rdb
       integer(c_int) :: ierr
     end function fsr_rdb_close

     function fsr_family_counts (part, counts, cap) bind(C,name="fsr_family_counts") result(nfam)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int), intent(out) :: counts(*)
       integer(c_int), value :: cap
       integer(c_int) :: nfam
     end function fsr_family_counts

     function fsr_vm_path_info (part, info, cap) bind(C,name="fsr_vm_path_info") result(nfilled)
       import :: c_ptr, c_int, c_long_long
       type(c_ptr), value :: part
       integer(c_long_long), intent(out) :: info(*)
       integer(c_int), value :: cap
       integer(c_int) :: nfilled
     end function fsr_vm_path_info

     function fsr_recover_displacements (part, sv_hist, nsteps, vm_hist) bind(C,name="fsr_recover_displacements") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: part
       real(c_double), intent(in) :: sv_hist(*)
       integer(c_int), value      :: nsteps
       type(c_ptr), value         :: vm_hist
       integer(c_int) :: ierr
     end function fsr_recover_displacements

     function fsr_rdb_write_steps_displacements (rdb, sv_hist, nsteps, stepno, time, sup_tr) &
          &   bind(C,name="fsr_rdb_write_steps_displacements") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: rdb
       real(c_double), intent(in) :: sv_hist(*), time(*), sup_tr(*)
       integer(c_int), value      :: nsteps
       integer(c_int), intent(in) :: stepno(*)
       integer(c_int) :: ierr
     end function fsr_rdb_write_steps_displacements

     ! ---- streaming form of the hot path ---------------------------------------------------------
     function fsr_recover_async (part, Q, ldq, nsteps) bind(C,name="fsr_recover_async") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: part
       real(c_double), intent(in) :: Q(*)
       integer(c_int), value      :: ldq, nsteps
       integer(c_int) :: ierr
     end function fsr_recover_async

     function fsr_get_envelope_async (part, vm_max, vm_min) bind(C,name="fsr_get_envelope_async") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: part
       real(c_double), intent(out) :: vm_max(*), vm_min(*)
       integer(c_int) :: ierr
     end function fsr_get_envelope_async

     function fsr_envelope_wait (part) bind(C,name="fsr_envelope_wait") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: ierr
     end function fsr_envelope_wait

     function fsr_synchronize (part) bind(C,name="fsr_synchronize") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: ierr
     end function fsr_synchronize

     ! ---- element blocks and multi-GPU (sharded.cu) ------------------------------------------
     function fsr_split_elements (sam, elm, nblocks, e_cut) bind(C,name="fsr_split_elements") result(ierr)
       import :: c_int, fsr_sam, fsr_elmdata
       type(fsr_sam), intent(in)     :: sam
       type(fsr_elmdata), intent(in) :: elm
       integer(c_int), value         :: nblocks
       integer(c_int), intent(out)   :: e_cut(*)
       integer(c_int) :: ierr
     end function fsr_split_elements

     function fsr_blockdef_create (def, sam, elm, opt, e0, e1) bind(C,name="fsr_blockdef_create") result(ierr)
       import :: c_ptr, c_int, fsr_sam, fsr_elmdata, fsr_options
       type(c_ptr), intent(out)      :: def
       type(fsr_sam), intent(in)     :: sam
       type(fsr_elmdata), intent(in) :: elm
       type(fsr_options), intent(in) :: opt
       integer(c_int), value         :: e0, e1
       integer(c_int) :: ierr
     end function fsr_blockdef_create

     function fsr_blockdef_sam (def) bind(C,name="fsr_blockdef_sam") result(p)
       import :: c_ptr
       type(c_ptr), value :: def
       type(c_ptr) :: p   !< pointer to a fsr_sam, use c_f_pointer
     end function fsr_blockdef_sam

     function fsr_blockdef_elm (def) bind(C,name="fsr_blockdef_elm") result(p)
       import :: c_ptr
       type(c_ptr), value :: def
       type(c_ptr) :: p   !< pointer to a fsr_elmdata
     end function fsr_blockdef_elm

     function fsr_blockdef_info (def, info, rows1, nodes) bind(C,name="fsr_blockdef_info") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: def
       integer(c_int), intent(out) :: info(10), rows1(*), nodes(*)
       integer(c_int) :: ierr
     end function fsr_blockdef_info

     subroutine fsr_blockdef_destroy (def) bind(C,name="fsr_blockdef_destroy")
       import :: c_ptr
       type(c_ptr), value :: def
     end subroutine fsr_blockdef_destroy

     function fsr_part_create_block (part, sam, elm, opt, e0, e1) bind(C,name="fsr_part_create_block") result(ierr)
       import :: c_ptr, c_int, fsr_sam, fsr_elmdata, fsr_options
       type(c_ptr), intent(out)      :: part
       type(fsr_sam), intent(in)     :: sam
       type(fsr_elmdata), intent(in) :: elm
       type(fsr_options), intent(in) :: opt
       integer(c_int), value         :: e0, e1   !< 0-based element range [e0, e1)
       integer(c_int) :: ierr
     end function fsr_part_create_block

     function fsr_block_info (part, info) bind(C,name="fsr_block_info") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int), intent(out) :: info(10)
       integer(c_int) :: ierr
     end function fsr_block_info

     function fsr_block_rows (part, rows1, nodes) bind(C,name="fsr_block_rows") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int), intent(out) :: rows1(*), nodes(*)
       integer(c_int) :: ierr
     end function fsr_block_rows

     function fsr_set_recovery_parent (part, B, ldB, E, ldE) bind(C,name="fsr_set_recovery_parent") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: part
       real(c_double), intent(in) :: B(*), E(*)
       integer(c_int), value      :: ldB, ldE
       integer(c_int) :: ierr
     end function fsr_set_recovery_parent

     function fsr_group_create (group, sam, elm, opt, devices, ndev) bind(C,name="fsr_group_create") result(ierr)
       import :: c_ptr, c_int, fsr_sam, fsr_elmdata, fsr_options
       type(c_ptr), intent(out)      :: group
       type(fsr_sam), intent(in)     :: sam
       type(fsr_elmdata), intent(in) :: elm
       type(fsr_options), intent(in) :: opt
       integer(c_int), intent(in)    :: devices(*)
       integer(c_int), value         :: ndev
       integer(c_int) :: ierr
     end function fsr_group_create

     function fsr_group_set_recovery (group, B, ldB, E, ldE) bind(C,name="fsr_group_set_recovery") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: group
       real(c_double), intent(in) :: B(*), E(*)
       integer(c_int), value      :: ldB, ldE
       integer(c_int) :: ierr
     end function fsr_group_set_recovery

     function fsr_group_num_blocks (group) bind(C,name="fsr_group_num_blocks") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: group
       integer(c_int) :: n
     end function fsr_group_num_blocks

     function fsr_group_num_result_points (group) bind(C,name="fsr_group_num_result_points") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: group
       integer(c_int) :: n
     end function fsr_group_num_result_points

     function fsr_group_ndim (group) bind(C,name="fsr_group_ndim") result(n)
       import :: c_ptr, c_int
       type(c_ptr), value :: group
       integer(c_int) :: n
     end function fsr_group_ndim

     function fsr_group_block (group, b) bind(C,name="fsr_group_block") result(part)
       import :: c_ptr, c_int
       type(c_ptr), value    :: group
       integer(c_int), value :: b
       type(c_ptr) :: part
     end function fsr_group_block

     function fsr_group_recover (group, Q, ldq, nsteps, vm_hist) bind(C,name="fsr_group_recover") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: group
       real(c_double), intent(in) :: Q(*)
       integer(c_int), value      :: ldq, nsteps
       type(c_ptr), value         :: vm_hist   !< c_loc of a real(c_double) array [npts,nsteps], or c_null_ptr
       integer(c_int) :: ierr
     end function fsr_group_recover

     function fsr_group_synchronize (group) bind(C,name="fsr_group_synchronize") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: group
       integer(c_int) :: ierr
     end function fsr_group_synchronize

     function fsr_group_reset_envelope (group) bind(C,name="fsr_group_reset_envelope") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: group
       integer(c_int) :: ierr
     end function fsr_group_reset_envelope

     function fsr_group_get_envelope (group, vm_max, vm_min) bind(C,name="fsr_group_get_envelope") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: group
       real(c_double), intent(out) :: vm_max(*), vm_min(*)
       integer(c_int) :: ierr
     end function fsr_group_get_envelope

     function fsr_group_last_timing (group, t, n) bind(C,name="fsr_group_last_timing") result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: group
       real(c_double), intent(out) :: t(*)
       integer(c_int), value :: n
       integer(c_int) :: ierr
     end function fsr_group_last_timing

     function fsr_group_timing_reset (group) bind(C,name="fsr_group_timing_reset") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: group
       integer(c_int) :: ierr
     end function fsr_group_timing_reset

     subroutine fsr_group_destroy (group) bind(C,name="fsr_group_destroy")
       import :: c_ptr
       type(c_ptr), value :: group
     end subroutine fsr_group_destroy

     function fsr_comm_unique_id (id, cap) bind(C,name="fsr_comm_unique_id") result(ierr)
       import :: c_char, c_int
       character(kind=c_char), intent(out) :: id(*)   !< 128 bytes
       integer(c_int), value :: cap
       integer(c_int) :: ierr
     end function fsr_comm_unique_id

     function fsr_comm_init_rank (comm, id, rank, world, device) bind(C,name="fsr_comm_init_rank") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr), intent(out) :: comm
       character(kind=c_char), intent(in) :: id(*)
       integer(c_int), value :: rank, world, device
       integer(c_int) :: ierr
     end function fsr_comm_init_rank

     function fsr_comm_broadcast (comm, buf_dev, count, root, stream) bind(C,name="fsr_comm_broadcast") result(ierr)
       import :: c_ptr, c_int, c_long_long
       type(c_ptr), value :: comm, buf_dev, stream
       integer(c_long_long), value :: count
       integer(c_int), value :: root
       integer(c_int) :: ierr
     end function fsr_comm_broadcast

     function fsr_comm_gather_envelope (comm, block, pt0, npts, vm_max_root_dev, vm_min_root_dev, root, stream) &
          &   bind(C,name="fsr_comm_gather_envelope") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: comm, block, vm_max_root_dev, vm_min_root_dev, stream
       integer(c_int), intent(in) :: pt0(*), npts(*)
       integer(c_int), value :: root
       integer(c_int) :: ierr
     end function fsr_comm_gather_envelope

     subroutine fsr_comm_destroy (comm) bind(C,name="fsr_comm_destroy")
       import :: c_ptr
       type(c_ptr), value :: comm
     end subroutine fsr_comm_destroy

     function fsr_nccl_version () bind(C,name="fsr_nccl_version") result(v)
       import :: c_int
       integer(c_int) :: v
     end function fsr_nccl_version

     ! ---- the fedem_stress program (same exported names as the reference's stressInterface.C) --
     subroutine initSolverArgs (argc, argv) bind(C,name="initSolverArgs")
       import :: c_int, c_ptr
       integer(c_int), value :: argc
       type(c_ptr)           :: argv(*)   !< C strings
     end subroutine initSolverArgs

     function solveStress () bind(C,name="solveStress") result(ierr)
       import :: c_int
       integer(c_int) :: ierr
     end function solveStress

     !> ffr_getNextStep loop over a sorted key list: indices (0-based) of the time steps to process
     function fsr_select_steps (times, n, start, stop, tinc, out, cap) bind(C,name="fsr_select_steps") result(nsel)
       import :: c_int, c_double
       real(c_double), intent(in)  :: times(*)
       integer(c_int), value       :: n, cap
       real(c_double), value       :: start, stop, tinc
       integer(c_int), intent(out) :: out(*)
       integer(c_int) :: nsel
     end function fsr_select_steps

     subroutine fsr_cmdline_reset () bind(C,name="fsr_cmdline_reset")
     end subroutine fsr_cmdline_reset

     subroutine fsr_stress_define_options () bind(C,name="fsr_stress_define_options")
     end subroutine fsr_stress_define_options

     subroutine fsr_gage_define_options () bind(C,name="fsr_gage_define_options")
     end subroutine fsr_gage_define_options

     function solveGage () bind(C,name="solveGage") result(ierr)
       import :: c_int
       integer(c_int) :: ierr
     end function solveGage

     subroutine fsr_modes_define_options () bind(C,name="fsr_modes_define_options")
     end subroutine fsr_modes_define_options

     function solveModes () bind(C,name="solveModes") result(ierr)
       import :: c_int
       integer(c_int) :: ierr
     end function solveModes

     subroutine fsr_fpp_define_options () bind(C,name="fsr_fpp_define_options")
     end subroutine fsr_fpp_define_options

     function solveFpp () bind(C,name="solveFpp") result(ierr)
       import :: c_int
       integer(c_int) :: ierr
     end function solveFpp

     subroutine fsr_cmdline_add_bool (name, value) bind(C,name="fsr_cmdline_add_bool")
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int), value :: value
     end subroutine fsr_cmdline_add_bool

     subroutine fsr_cmdline_add_int (name, value) bind(C,name="fsr_cmdline_add_int")
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int), value :: value
     end subroutine fsr_cmdline_add_int

     subroutine fsr_cmdline_add_double (name, value) bind(C,name="fsr_cmdline_add_double")
       import :: c_char, c_double
       character(kind=c_char), intent(in) :: name(*)
       real(c_double), value :: value
     end subroutine fsr_cmdline_add_double

     subroutine fsr_cmdline_add_string (name, value) bind(C,name="fsr_cmdline_add_string")
       import :: c_char
       character(kind=c_char), intent(in) :: name(*), value(*)
     end subroutine fsr_cmdline_add_string

     subroutine fsr_cmdline_init (argc, argv) bind(C,name="fsr_cmdline_init")
       import :: c_int, c_ptr
       integer(c_int), value :: argc
       type(c_ptr)           :: argv(*)
     end subroutine fsr_cmdline_init

     function fsr_cmdline_read_file (path) bind(C,name="fsr_cmdline_read_file") result(ok)
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: path(*)
       integer(c_int) :: ok
     end function fsr_cmdline_read_file

     !> ffa_cmdlinearg_getbool / getint / getdouble / getstring / isSet
     function fsr_cmdline_get_bool (name) bind(C,name="fsr_cmdline_get_bool") result(v)
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int) :: v
     end function fsr_cmdline_get_bool

     function fsr_cmdline_get_int (name) bind(C,name="fsr_cmdline_get_int") result(v)
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int) :: v
     end function fsr_cmdline_get_int

     function fsr_cmdline_get_double (name) bind(C,name="fsr_cmdline_get_double") result(v)
       import :: c_char, c_double
       character(kind=c_char), intent(in) :: name(*)
       real(c_double) :: v
     end function fsr_cmdline_get_double

     function fsr_cmdline_get_string (name, out, cap) bind(C,name="fsr_cmdline_get_string") result(nchar)
       import :: c_char, c_int
       character(kind=c_char), intent(in)  :: name(*)
       character(kind=c_char), intent(out) :: out(*)
       integer(c_int), value :: cap
       integer(c_int) :: nchar
     end function fsr_cmdline_get_string

     function fsr_cmdline_is_set (name) bind(C,name="fsr_cmdline_is_set") result(v)
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int) :: v
     end function fsr_cmdline_is_set

     ! ---- in-core part state (the solver's recovery list, polled by fedempy) ------------------
     function fsr_recovery_register (base_id, part, minex) bind(C,name="fsr_recovery_register") result(ierr)
       import :: c_ptr, c_int
       integer(c_int), value      :: base_id
       type(c_ptr)   , value      :: part
       integer(c_int), intent(in) :: minex(*)
       integer(c_int) :: ierr
     end function fsr_recovery_register

     function fsr_recovery_unregister (base_id) bind(C,name="fsr_recovery_unregister") result(ierr)
       import :: c_int
       integer(c_int), value :: base_id
       integer(c_int) :: ierr
     end function fsr_recovery_unregister

     function fsr_recovery_update_parts (nparts, base_ids, istep, time, timeStep, q) &
          &   bind(C,name="fsr_recovery_update_parts") result(ierr)
       import :: c_int, c_double, c_ptr
       integer(c_int), value      :: nparts, istep
       integer(c_int), intent(in) :: base_ids(*)
       real(c_double), value      :: time, timeStep
       type(c_ptr)   , intent(in) :: q(*)      !< c_loc of every part's [finit; vg]
       integer(c_int) :: ierr
     end function fsr_recovery_update_parts

     function fsr_recovery_options (args) bind(C,name="fsr_recovery_options") result(ierr)
       import :: c_char, c_int
       character(kind=c_char), intent(in) :: args(*)   !< "-recovery 1 -partVMStress 3 ..." null-terminated
       integer(c_int) :: ierr
     end function fsr_recovery_options

     function fsr_recovery_register_part (base_id, user_id, descr, part, minex, supTrInit) &
          &   bind(C,name="fsr_recovery_register_part") result(ierr)
       import :: c_char, c_int, c_double, c_ptr
       integer(c_int), value      :: base_id, user_id
       character(kind=c_char), intent(in) :: descr(*)
       type(c_ptr), value         :: part
       integer(c_int), intent(in) :: minex(*)
       real(c_double), intent(in) :: supTrInit(*)   !< sup%supTrInit, 3x4 column-major
       integer(c_int) :: ierr
     end function fsr_recovery_register_part

     function fsr_recovery_update_parts_save (nparts, base_ids, istep, time, timeStep, q, supTr, doSave) &
          &   bind(C,name="fsr_recovery_update_parts_save") result(ierr)
       import :: c_int, c_double, c_ptr
       integer(c_int), value      :: nparts, istep
       integer(c_int), intent(in) :: base_ids(*)
       real(c_double), value      :: time, timeStep
       type(c_ptr)   , intent(in) :: q(*)      !< c_loc of every part's [finit; vg]
       type(c_ptr)   , intent(in) :: supTr(*)  !< c_loc of every part's sup%supTr (3x4)
       integer(c_int), value      :: doSave
       integer(c_int) :: ierr
     end function fsr_recovery_update_parts_save

     function fsr_recovery_close () bind(C,name="fsr_recovery_close") result(ierr)
       import :: c_int
       integer(c_int) :: ierr
     end function fsr_recovery_close

     function fsr_recovery_file (base_id, buf, cap) bind(C,name="fsr_recovery_file") result(length)
       import :: c_char, c_int
       integer(c_int), value :: base_id, cap
       character(kind=c_char), intent(out) :: buf(*)
       integer(c_int) :: length
     end function fsr_recovery_file

     function fsr_recovery_update (base_id, istep, time, timeStep, q) bind(C,name="fsr_recovery_update") result(ierr)
       import :: c_int, c_double
       integer(c_int), value      :: base_id, istep
       real(c_double), value      :: time, timeStep
       real(c_double), intent(in) :: q(*)
       integer(c_int) :: ierr
     end function fsr_recovery_update

     function getPartDeformationStateSize (bid) bind(C,name="getPartDeformationStateSize") result(ndat)
       import :: c_int
       integer(c_int), value :: bid
       integer(c_int) :: ndat
     end function getPartDeformationStateSize

     function getPartStressStateSize (bid) bind(C,name="getPartStressStateSize") result(ndat)
       import :: c_int
       integer(c_int), value :: bid
       integer(c_int) :: ndat
     end function getPartStressStateSize

     function savePartDeformationState (bid, data, ndat) bind(C,name="savePartDeformationState") result(ok)
       import :: c_int, c_double, c_bool
       integer(c_int), value       :: bid, ndat
       real(c_double), intent(out) :: data(*)
       logical(c_bool) :: ok
     end function savePartDeformationState

     function savePartStressState (bid, data, ndat) bind(C,name="savePartStressState") result(ok)
       import :: c_int, c_double, c_bool
       integer(c_int), value       :: bid, ndat
       real(c_double), intent(out) :: data(*)
       logical(c_bool) :: ok
     end function savePartStressState

     ! ---- diagnostics ------------------------------------------------------------------------
     function fsr_last_error () bind(C,name="fsr_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg    !< NUL-terminated C string, see fsr_error_message below
     end function fsr_last_error

     function fsr_kernel_launches (reset) bind(C,name="fsr_kernel_launches") result(n)
       import :: c_int, c_long_long
       integer(c_int), value :: reset
       integer(c_long_long) :: n
     end function fsr_kernel_launches

     function fsr_last_timing (part, t_ms, n) bind(C,name="fsr_last_timing") result(m)
       import :: c_ptr, c_int, c_double
       type(c_ptr)   , value       :: part
       real(c_double), intent(out) :: t_ms(*)
       integer(c_int), value       :: n
       integer(c_int) :: m
     end function fsr_last_timing

     function fsr_timing_reset (part) bind(C,name="fsr_timing_reset") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: part
       integer(c_int) :: ierr
     end function fsr_timing_reset

  end interface

contains

  !> Copies the library's last error text into a Fortran string (for reportError).
  subroutine fsr_error_message (text)
    character(len=*), intent(out) :: text
    character(kind=c_char), pointer :: p(:)
    type(c_ptr) :: cmsg
    integer     :: i
    text = ' '
    cmsg = fsr_last_error()
    if (.not. c_associated(cmsg)) return
    call c_f_pointer (cmsg, p, [len(text)])
    do i = 1, len(text)
       if (p(i) == c_null_char) exit
       text(i:i) = p(i)
    end do
  end subroutine fsr_error_message

end module fedem_b200_mod
