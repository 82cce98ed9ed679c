!! stress_b200_driver.f90 -- how the reference's Fortran host code calls the B200 library.
!!
!! This is the replacement for the time loop of src/vpmStress/stress.f90:357-435: instead of
!!   do while (ffr_getnextstep(...)); readSupElDisplacements; calcIntDisplacements; calcStresses
!! per time step, the driver collects the reduced displacements of a window of steps into
!! Q(ndim,nWin) and makes ONE call.  Everything before the loop (ffl_init, initiateSAM,
!! readSolverData, openBandEmatrices, ffr_init, readResponsePointers) stays as it is; the arrays handed
!! over are the reference's own (SamType members, ffl_* results, the in-core B and E of
!! DiskMatrixType%vala).  The results database is written by the library: fsr_rdb_create replaces
!! writeStressHeader (same header grammar), fsr_rdb_write_steps replaces writeTimeStepDB + the per-element
!! writeStressDB / writeStrMeasureDB / writeDisplacementDB calls of a whole window of steps -- the records
!! are formed on the device.  (A host that prefers to keep its own writer calls fsr_recover / fsr_recover_step_full
!! instead and gets the arrays back; the library-only program is bin/fedem_stress, csrc/stress_driver.cu.)
!!
!! Delivered as source: the build image has no Fortran compiler (see DESIGN.md section 1).

subroutine stress_b200 (sam, xyz, emod, rny, thk, elmid, beam, Bmat, Emat, ngen, &
     &                  triads, sup, rPointers, startTime, stopTime, tInc, &
     &                  rdbFile, outMask, lDouble, lpu, ierr)

  use, intrinsic :: iso_c_binding
  use fedem_b200_mod
  use SamModule                 , only : SamType
  use TriadTypeModule           , only : TriadType
  use SupElTypeModule           , only : SupElType
  use DisplacementModule        , only : readSupElDisplacements
  use FFrExtractorInterface     , only : ffr_getnextstep
  use KindModule                , only : dp, i8
  use ReportErrorModule         , only : reportError, error_p, debugFileOnly_p

  implicit none

  type(SamType)  , intent(in), target :: sam
  real(dp)       , intent(in), target :: xyz(:,:), emod(:), rny(:), thk(:), beam(:,:)
  integer        , intent(in), target :: elmid(:)
  real(dp)       , intent(in)         :: Bmat(:,:), Emat(:,:)
  integer        , intent(in)         :: ngen, lpu
  type(TriadType), intent(inout)      :: triads(:)
  type(SupElType), intent(inout)      :: sup
  integer        , intent(in)         :: rPointers(:)
  real(dp)       , intent(in)         :: startTime, stopTime, tInc
  character(len=*), intent(in)        :: rdbFile   !< -rdbfile
  integer        , intent(in)         :: outMask   !< FSR_OUT_* bits from -SR -stress -strain -vmStress ...
  logical        , intent(in)         :: lDouble   !< -double
  integer        , intent(out)        :: ierr

  integer, parameter  :: nWin = 512        !< time steps per device batch
  type(fsr_sam)       :: csam
  type(fsr_elmdata)   :: celm
  type(fsr_options)   :: copt
  type(fsr_rdb_options) :: ropt
  type(c_ptr)         :: part, rdb
  character(kind=c_char,len=:), allocatable, target :: cDescr
  real(dp), allocatable, target :: Q(:,:), supTr(:,:,:), vmMax(:), vmMin(:), time(:)
  integer , allocatable :: stepNo(:)
  real(dp), target    :: supTrInit(3,4)
  real(dp)            :: currTime
  integer(i8)         :: iStep
  integer             :: ndim, npts, n
  character(len=512)  :: msg

  !! --- model hand-over (replaces initiateSAM's index work and the per-step ffl_* lookups)
  csam%nnod = sam%nnod;  csam%nel = sam%nel;  csam%ndof = sam%ndof
  csam%ndof1 = sam%ndof1; csam%ndof2 = sam%ndof2; csam%ngen = ngen
  csam%neq = sam%neq;  csam%nceq = sam%nceq
  csam%nmmnpc = size(sam%mmnpc);  csam%nmmceq = size(sam%mmceq)
  csam%madof  = c_loc(sam%madof);  csam%msc   = c_loc(sam%msc)
  csam%mpmnpc = c_loc(sam%mpmnpc); csam%mmnpc = c_loc(sam%mmnpc)
  csam%melcon = c_loc(sam%melcon); csam%mpmceq = c_loc(sam%mpmceq)
  csam%mmceq  = c_loc(sam%mmceq);  csam%ttcc  = c_loc(sam%ttcc)
  csam%meqn   = c_loc(sam%meqn);   csam%meqn1 = c_loc(sam%meqn1)
  csam%meqn2  = c_loc(sam%meqn2)
  celm%xyz = c_loc(xyz);  celm%emod = c_loc(emod);  celm%rny = c_loc(rny)
  celm%thk = c_loc(thk);  celm%elmid = c_loc(elmid); celm%beam = c_loc(beam)
  copt%device = 0;  copt%stressForm = 0;  copt%step_tile = nWin;  copt%reserved = 0

  ierr = fsr_part_create(part,csam,celm,copt)
  if (ierr < 0) goto 900
  if (ierr > 0) write(lpu,"('  ** ',I8,' elements failed; they get hugeVal results')") ierr

  !! --- replaces openBandEmatrices: B and E go to the GPU once
  ierr = fsr_set_recovery(part,Bmat,size(Bmat,1),Emat,size(Emat,1))
  if (ierr < 0) goto 900

  ndim = fsr_ndim(part)
  npts = fsr_num_result_points(part)
  allocate(Q(ndim,nWin),supTr(3,4,nWin),time(nWin),stepNo(nWin),vmMax(npts),vmMin(npts))

  !! --- replaces writeStressHeader (saveStressModule.f90:120-247)
  cDescr = trim(sup%id%descr)//c_null_char
  supTrInit = sup%supTrInit
  ropt%out_mask = outMask;  ropt%double_precision = merge(1,0,lDouble);  ropt%rdbinc = 1
  ropt%part_base_id = sup%id%baseId;  ropt%part_user_id = sup%id%userId
  ropt%part_descr = c_loc(cDescr);  ropt%model_file = c_null_ptr;  ropt%link_file = c_null_ptr
  ropt%elmid = c_loc(elmid);  ropt%module_name = c_null_ptr
  ropt%minex = c_loc(sam%minex);  ropt%sup_tr_init = c_loc(supTrInit)
  ierr = fsr_rdb_create(rdb,part,trim(rdbFile)//c_null_char,ropt)
  if (ierr < 0) goto 900

  !! --- the time loop, batched
  n = 0
  currTime = startTime - 1.0_dp
  do while (ffr_getnextstep(startTime,stopTime,tInc,currTime,iStep))
     call readSupElDisplacements (triads,sup,rPointers,0,lpu,ierr)   ! fills sup%finit, sup%genDOFs%ur
     if (ierr /= 0) goto 900
     n = n + 1
     Q(1:sam%ndof2,n) = sup%finit(1:sam%ndof2)
     if (ngen > 0) Q(sam%ndof2+1:ndim,n) = sup%genDOFs%ur(1:ngen)
     supTr(:,:,n) = sup%supTr;  time(n) = currTime;  stepNo(n) = int(iStep)
     if (n == nWin) then
        !! K1 expansion + record kernels for n steps, records appended to the .frs file
        ierr = fsr_rdb_write_steps(rdb,Q,ndim,n,stepNo,time,supTr)
        if (ierr < 0) goto 900
        n = 0
     end if
  end do
  if (n > 0) then
     ierr = fsr_rdb_write_steps(rdb,Q,ndim,n,stepNo,time,supTr)
     if (ierr < 0) goto 900
  end if

  ierr = fsr_rdb_close(rdb)
  !! (the running von Mises envelopes of fsr_get_envelope belong to the fsr_recover path, not to the record path)
  call fsr_part_destroy (part)
  return

900 call fsr_error_message (msg)
  call reportError (error_p,trim(msg),addString='stress_b200')
  call reportError (debugFileOnly_p,'stress_b200')
  call fsr_part_destroy (part)

end subroutine stress_b200
