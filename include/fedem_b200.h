/*
 * fedem_b200.h -- C ABI of the B200-native stress-recovery library (libfedem_b200.so).
 *
 * This is the drop-in boundary for the fedem_stress / fedem_gage hot path of
 * SAP-archive/fedem-solvers: plain pointers and sizes, no C++/torch types, callable from
 * Fortran through ISO_C_BINDING (fortran/fedem_b200_mod.f90), from C/C++ and from Python/ctypes.
 * Each entry point names the reference routine(s) it replaces (paths relative to the reference
 * checkout).  The reference calls those routines once per time step; this library is batched:
 * the caller collects the reduced history Q for a window of steps and makes one call.
 *
 * Conventions (identical to what the reference stores, so a Fortran caller passes its arrays
 * untouched): all index arrays are 1-based as in the .fsm file; matrices are column-major;
 * host pointers unless the name says _dev.  Every function returns 0 on success, <0 on a fatal
 * error (message via fsr_last_error), >0 as a warning count (e.g. number of failed elements,
 * which get hugeVal results like stressRoutines.f90:237-241,264-268).
 *
 * There is NO CPU fallback: every call fails with FSR_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef FEDEM_B200_H
#define FEDEM_B200_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSR_OK 0
#define FSR_ERR_ARG (-1)
#define FSR_ERR_CUDA (-2)
#define FSR_ERR_ALLOC (-3)
#define FSR_ERR_STATE (-4)
#define FSR_ERR_LIMIT (-5)

#define FSR_NBEAM 32 /* doubles of beam data per element, see fsr_elmdata.beam */

typedef struct fsr_part fsr_part; /* opaque per-superelement handle (one FE part on one GPU) */

/* SamType subset (src/vpmCommon/samModule.f90:27-66) exactly as initiateSAM reads it from the
 * .fsm file (src/vpmStress/samStressModule.f90:273-316), BEFORE the msc remap of :245-256. */
typedef struct fsr_sam {
  int nnod, nel, ndof, ndof1, ndof2, ngen, neq, nceq, nmmnpc, nmmceq;
  const int *madof;  /* [nnod+1] */
  const int *msc;    /* [ndof] status codes as stored: 2 = external, 1 = free, 0 = fixed */
  const int *mpmnpc; /* [nel+1]  */
  const int *mmnpc;  /* [nmmnpc] */
  const int *melcon; /* [nel]    */
  const int *mpmceq; /* [nceq+1] (may be NULL if nceq == 0) */
  const int *mmceq;  /* [nmmceq] */
  const double *ttcc;/* [nmmceq] */
  const int *meqn;   /* [ndof]   */
  const int *meqn1;  /* [ndof1]  */
  const int *meqn2;  /* [ndof2]  */
} fsr_sam;

/* What the reference pulls per element, per step, from the FE model through
 * ffl_getcoor / ffl_getmat / ffl_getthick / ffl_getbeamsection / ffl_getpinflags / ffl_getelmid
 * (fedem-foundation/src/FFlLib/FFlLinkHandler_F.C:699-1193); here handed over once. */
typedef struct fsr_elmdata {
  const double *xyz;  /* [3*nnod] x,y,z of every internal node (node order of madof)          */
  const double *emod; /* [nel] Young's modulus                                               */
  const double *rny;  /* [nel] Poisson's ratio                                               */
  const double *thk;  /* [nel] shell thickness (ffl_getthick: uniform over the element)      */
  const int *elmid;   /* [nel] external element id; < 1 = not in the -group selection; NULL = all */
  const double *beam; /* [nel*FSR_NBEAM] for type-11 elements, else ignored; may be NULL.
                         [0:5) X(1:5), [5:10) Y(1:5), [10:15) Z(1:5) as ffl_getcoor returns them
                         for beams (ends incl. eccentricity, Z-direction point, the two nodes);
                         [15:29) BSEC(1:14) of ffl_getbeamsection; [29] IPA, [30] IPB pin flags */
} fsr_elmdata;

/* Arithmetic-changing options of fedem_stress (src/vpmStress/stressmain.C:68-79) + run control */
typedef struct fsr_options {
  int device;     /* CUDA device ordinal                                                    */
  int stressForm; /* -stressForm (solids): 0 = nodal evaluation (default), else Gauss extrap. */
  int step_tile;  /* time steps per device batch (0 = automatic from free HBM)              */
  int reserved[5];/* [0]: element processing order, 0 = Morton order of the centroids (default,
                     L2 reuse of shared nodes), 1 = SAM order; results are in SAM order either way
                     [1]: -ffqStressForm + 1 (legacy FFQ4 shells, type 22), 0 = the default formulation 2
                     [2]: -fftStressForm + 1 (legacy FFT3 shells, type 21), 0 = the default formulation 1
                     Type 21 is recovered for the default formulation (STR21 = STR23 statement by statement), type 22
                     for formulations 2 (STR22 = STR24) and 1 (one Gauss point; parts without ANDES quads),
                     elStressModule.f90:521-733; with another value those elements get no results */
} fsr_options;

/* Output selection bits = the -vmStress ... switches of stressmain.C:46-60 */
#define FSR_OUT_VMSTRESS 0x001
#define FSR_OUT_MAXPSTRESS 0x002
#define FSR_OUT_MINPSTRESS 0x004
#define FSR_OUT_MAXSSTRESS 0x008
#define FSR_OUT_VMSTRAIN 0x010
#define FSR_OUT_MAXPSTRAIN 0x020
#define FSR_OUT_MINPSTRAIN 0x040
#define FSR_OUT_MAXSSTRAIN 0x080
#define FSR_OUT_STRESS 0x100
#define FSR_OUT_STRAIN 0x200
#define FSR_OUT_SR 0x400
#define FSR_OUT_DEFORMATION 0x800

/* ---- life cycle ------------------------------------------------------------------------ */

/* Replaces initiateSAM (src/vpmStress/samStressModule.f90:39-263: dofPosIn2, index maps) and
 * the per-step ffl_* lookups + element-matrix rebuilds of ElStress
 * (src/vpmStress/elStressModule.f90:129-229): uploads the model and builds every element's
 * stress operator once on the GPU. */
int fsr_part_create(fsr_part **part, const fsr_sam *sam, const fsr_elmdata *elm,
                    const fsr_options *opt);

/* Replaces openBandEmatrices (src/vpmStress/displacementModule.f90:645-790): takes the
 * reducer's B (ndof1 x ndof2, leading dimension ldB) and E (ndof1 x ngen, ldE) as dmOpen
 * holds them in core and folds dofPosIn2, meqn1/meqn2 scatter and disExpand (constraint
 * equations) into one row operator R[ndof x (ndof2+ngen)] in nodal DOF order on the GPU. */
int fsr_set_recovery(fsr_part *part, const double *B, int ldB, const double *E, int ldE);

void fsr_part_destroy(fsr_part *part);

/* Makes every later call on this handle run on the caller's stream (cudaStream_t as void*;
 * NULL = the legacy default stream) instead of the handle's private stream. */
int fsr_set_stream(fsr_part *part, void *stream);

/* ---- sizes ------------------------------------------------------------------------------- */
int fsr_num_result_points(const fsr_part *part);           /* sum of nstrp over active elements */
int fsr_result_point_offsets(const fsr_part *part, int *off /* [nel+1] */);
int fsr_ndim(const fsr_part *part);                        /* ndof2 + ngen (mpar(24))            */

/* ---- the hot path --------------------------------------------------------------------------
 * Replaces the time loop body of stress.f90:361-435: calcIntDisplacements
 * (displacementModule.f90:931-1024) + calcStresses (stressRoutines.f90:48-342) for nsteps
 * steps at once.  Q is ndim x nsteps column-major, column s = [finit(1:ndof2); vg(1:ngen)] of
 * step s (what readSupElDisplacements/BuildFinit deliver, supElTypeModule.f90:1067-1114).
 *  vm_hist : optional [nsteps x npts] step-major von Mises stress of every result point
 *            (row s = the reference's resMat(1,:) of step s for all elements in SAM order);
 *            NULL = keep only the envelopes.
 * The running envelopes (strainCoatModule.f90:159-166,410-420 semantics: max starts at 0,
 * min at hugeVal) accumulate across calls until fsr_reset_envelope. */
int fsr_recover(fsr_part *part, const double *Q, int ldq, int nsteps, double *vm_hist);

/* Same, with Q already resident on the device (ldq-strided) and the von Mises history left on
 * the device: vm_hist_dev is [nsteps x ld_vm] step-major or NULL.  Asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream). */
int fsr_recover_dev(fsr_part *part, const double *Q_dev, int ldq, int nsteps,
                    double *vm_hist_dev, size_t ld_vm, void *stream);

/* The streaming form for hosts that feed window after window: fsr_recover_async queues a window (H2D copy of Q, K1, K2 +
 * envelope) and returns; fsr_get_envelope_async delivers the envelopes as they are after the work queued so far -- a
 * device-side snapshot in stream order, then the PCIe copy on a second stream while the next windows compute;
 * fsr_envelope_wait waits for the pending read-backs, fsr_synchronize for everything.  Q and the two result arrays
 * should be page-locked (cudaHostAlloc / cudaHostRegister): with pageable memory the calls still work but block. */
int fsr_recover_async(fsr_part *part, const double *Q, int ldq, int nsteps);
int fsr_get_envelope_async(fsr_part *part, double *vm_max, double *vm_min);
int fsr_envelope_wait(fsr_part *part);
int fsr_synchronize(fsr_part *part);

/* calcStresses on nodal displacements that are already there: what fedem_stress does without -fsifile, when the results
 * files hold a direct solution ("Vectors|Dynamic response|Displacement" of the part, stress.f90:397 -> readIntDisplacements,
 * displacementModule.f90:865-904).  No B / E matrices and no expansion: sv_hist is [nsteps x ndof] step-major in nodal DOF
 * order (madof).  vm_hist as fsr_recover; the envelopes accumulate. */
int fsr_recover_displacements(fsr_part *part, const double *sv_hist, int nsteps, double *vm_hist);

int fsr_reset_envelope(fsr_part *part);
int fsr_get_envelope(fsr_part *part, double *vm_max, double *vm_min); /* host [npts] each */
int fsr_envelope_dev(fsr_part *part, double **vm_max_dev, double **vm_min_dev);
/* device-to-device copy of the envelopes into caller-owned device buffers ([npts] each, either
 * may be NULL), asynchronous on `stream` -- the hand-over point to an NCCL gather. */
int fsr_copy_envelope_dev(fsr_part *part, double *vm_max_dst_dev, double *vm_min_dst_dev, void *stream);

/* Full result set for ONE step (what fedem_stress writes per step when all of -SR -stress
 * -strain -vmStress ... are on, stressRoutines.f90:234-331).  q = [finit; vg] (ndim).
 *  resmat [8 x npts] col-major per point: vmStress,maxP,minP,maxShear, then the same for strain
 *  stress/strain [6 x npts] (first ncmp rows used; tensorial shear strain as ElStress :244-253)
 *  sres [24 x nel] SR(1:6, node 1:4) shell stress resultants / SF(1:6, 1:2) beam section forces
 *  sv [ndof] expanded nodal displacements (calcIntDisplacements output).  Any may be NULL. */
int fsr_recover_step_full(fsr_part *part, const double *q, double *resmat, double *stress,
                          double *strain, double *sres, double *sv);

/* The in-core von Mises state that fedempy reads through getPartStressStateSize /
 * savePartStressState (src/vpmSolver/solverInterface.C:946,993 -> getGroupVMSsize / getStress,
 * src/vpmSolver/stressRecoveryModule.f90:203-280,718-747): for every active element that has
 * stress points, in SAM order, [iel, nenod, nstrp, vm(1..nstrp)] as doubles
 * (stressRoutines.f90:324-331).  fsr_vms_size returns the array length, fsr_get_vms fills it
 * for the step whose reduced displacements are q = [finit; vg]. */
int fsr_vms_size(const fsr_part *part);
int fsr_get_vms(fsr_part *part, const double *q, double *vms, int nvms);

/* Expansion only (calcIntDisplacements for a batch): U_host [nsteps x ndof] step-major. */
int fsr_expand(fsr_part *part, const double *Q, int ldq, int nsteps, double *U_host);
/* The same for a few DOFs only (0-based indices into the nodal DOF vector): out [nsteps x nrows] step-major; what
 * CalcRosetteDisplacements (strainRosetteModule.f90:846) needs of the H_el rows of the rosette nodes. */
int fsr_expand_rows(fsr_part *part, const double *Q, int ldq, int nsteps, const int *rows, int nrows, double *out);

/* ---- element blocks and multi-GPU (one box, NCCL over NVLink / NVSwitch) ------------------------------------
 * The reference recovers a part in one serial element loop, one process per part (stress.f90:126-128; in the solver
 * stressRecoveryModule.f90:1021-1061 loops over the parts).  Elements are independent given the nodal displacements and
 * result points belong to elements, so a part is cut into contiguous (SAM order) element blocks of equal cost; a block is
 * a self-contained part handle: its own nodes, every external node (ndof2 and the reduced history Q stay the parent's),
 * the masters of the constraint equations it depends on, the B / E rows of exactly those nodes.  Results of a block are
 * bit-identical to the parent's for its elements.
 *  fsr_split_elements      : e_cut[nblocks+1], 0-based element indices, block b = [e_cut[b], e_cut[b+1])
 *  fsr_part_create_block   : like fsr_part_create for the elements [e0, e1) of the part given by sam / elm
 *  fsr_block_info          : info[10] = e0, e1, first result point in the parent's order, result points, nodes, B/E rows of
 *                            the block, B/E rows of the parent (ndof1), result points of the parent, nodal DOFs of the
 *                            block, elements of the parent
 *  fsr_block_rows          : rows1[ndof1 of the block] = rows (0-based) of the parent's B / E the block keeps, nodes[nnod of
 *                            the block] = parent node numbers (1-based); either may be NULL
 *  fsr_set_recovery_parent : fsr_set_recovery with the PARENT's B / E (ldB, ldE >= the parent's ndof1); the block picks its
 *                            rows.  (fsr_set_recovery on a block handle expects the rows already gathered.) */
int fsr_split_elements(const fsr_sam *sam, const fsr_elmdata *elm, int nblocks, int *e_cut);
/* host only (no device needed): the SAM / element arrays of the block [e0, e1) for hosts that keep their own copy.  The
 * pointers of fsr_blockdef_sam / _elm stay valid until fsr_blockdef_destroy; info[10], rows1, nodes as fsr_block_info /
 * fsr_block_rows (any may be NULL). */
typedef struct fsr_blockdef fsr_blockdef;
int fsr_blockdef_create(fsr_blockdef **def, const fsr_sam *sam, const fsr_elmdata *elm, const fsr_options *opt,
                        int e0, int e1);
const fsr_sam *fsr_blockdef_sam(const fsr_blockdef *def);
const fsr_elmdata *fsr_blockdef_elm(const fsr_blockdef *def);
int fsr_blockdef_info(const fsr_blockdef *def, int *info, int *rows1, int *nodes);
void fsr_blockdef_destroy(fsr_blockdef *def);
int fsr_part_create_block(fsr_part **part, const fsr_sam *sam, const fsr_elmdata *elm, const fsr_options *opt,
                          int e0, int e1);
int fsr_block_info(const fsr_part *part, int *info);
int fsr_block_rows(const fsr_part *part, int *rows1, int *nodes);
int fsr_set_recovery_parent(fsr_part *part, const double *B, int ldB, const double *E, int ldE);

/* One process, several GPUs: the blocks of one part on `devices` (NULL / ndev <= 0: all visible devices; ndev > 0 with
 * devices == NULL: the first ndev).  fsr_group_recover copies Q to the first device, ncclBroadcast's it to the others and
 * runs K1 + K2 of every block concurrently; vm_hist (optional) is [nsteps x npts] step-major in the PARENT's result-point
 * order.  fsr_group_get_envelope gathers the per-block envelopes on the first device with ncclSend / ncclRecv (block b
 * lands at its first result point: concatenation is the parent's order) and copies them to the host.  opt->device is
 * ignored.  fsr_group_create returns the number of failed elements like fsr_part_create. */
typedef struct fsr_group fsr_group;
int fsr_group_create(fsr_group **group, const fsr_sam *sam, const fsr_elmdata *elm, const fsr_options *opt,
                     const int *devices, int ndev);
int fsr_group_set_recovery(fsr_group *group, const double *B, int ldB, const double *E, int ldE);
int fsr_group_num_blocks(const fsr_group *group);
int fsr_group_num_result_points(const fsr_group *group);
int fsr_group_ndim(const fsr_group *group);
fsr_part *fsr_group_block(fsr_group *group, int b); /* borrowed handle of block b */
int fsr_group_recover(fsr_group *group, const double *Q, int ldq, int nsteps, double *vm_hist);
int fsr_group_synchronize(fsr_group *group);
int fsr_group_reset_envelope(fsr_group *group);
int fsr_group_get_envelope(fsr_group *group, double *vm_max, double *vm_min); /* host [npts of the parent] each */
/* t[0] = K1, t[1] = K2 device time (ms) of the slowest block since fsr_group_timing_reset, t[2] = tiles,
 * t[3] = fastest / slowest block (load balance) */
int fsr_group_last_timing(fsr_group *group, double *t, int n);
int fsr_group_timing_reset(fsr_group *group);
void fsr_group_destroy(fsr_group *group);

/* One process per GPU (MPI or torchrun style hosts): a communicator over the ranks' devices.  Rank 0 calls
 * fsr_comm_unique_id (128 bytes), the host passes the id to the other ranks by its own means, every rank calls
 * fsr_comm_init_rank.  fsr_comm_broadcast sends the reduced history window from the root's device buffer to the same
 * buffer on every rank; fsr_comm_gather_envelope sends this rank's block envelopes to the root, which receives block r at
 * pt0[r] of its two [parent npts] device buffers (pt0 / npts [world]: fsr_block_info of every rank's block).  Both are
 * asynchronous on `stream` (cudaStream_t as void*) -- pass the stream the block runs on. */
typedef struct fsr_comm fsr_comm;
int fsr_comm_unique_id(char *id, int cap);
int fsr_comm_init_rank(fsr_comm **comm, const char *id, int rank, int world, int device);
int fsr_comm_broadcast(fsr_comm *comm, double *buf_dev, long long count, int root, void *stream);
int fsr_comm_gather_envelope(fsr_comm *comm, fsr_part *block, const int *pt0, const int *npts,
                             double *vm_max_root_dev, double *vm_min_root_dev, int root, void *stream);
void fsr_comm_destroy(fsr_comm *comm);
int fsr_nccl_version(void); /* version code of the NCCL library bound at run time, < 0 if none */

/* ---- strain gages + fatigue (fedem_gage path) ------------------------------------------------
 * Replaces ffp_addpoint / ffp_getdamage / ffp_getnumcycles
 * (fedem-foundation/src/FFpLib/FFpFatigue/FFpFatigue_F.C:37-141) for ngage independent scalar
 * histories at once: PVX peak-valley extraction, rainflow counting, Miner sum on a two-slope
 * NorSok S-N curve, cycle histogram with the bin edges of reportDamage
 * (src/vpmStress/strainGageModule.f90:827-848: s0 = 0, s1 = s0 + binSize, ...).
 *  hist   [ngage x nsteps] gage-major (one contiguous history per gage), host
 *  curve  {loga1, loga2, m1, m2}
 *  damage [ngage]; ncycles [ngage] counted cycles; bins [ngage x nbins] (may be NULL) */
int fsr_fatigue(int device, const double *hist, int ngage, int nsteps, double gate,
                const double *curve, double bin_size, int nbins, double *damage, int *ncycles,
                int *bins);
int fsr_fatigue_dev(int device, const double *hist_dev, size_t ld_hist, int ngage, int nsteps,
                    double gate, const double *curve, double bin_size, int nbins,
                    double *damage_dev, int *ncycles_dev, int *bins_dev, void *stream);

/* Streaming form of the same (what ffp_addpoint does sample by sample, FFpFatigue_F.C:37-44, but
 * without ever holding a whole history): a handle owns ngage per-gage states on one GPU and is fed
 * tiles of time steps in order.  Because the reference's PVX starts at the FIRST turning point,
 * which it finds by a look-ahead scan (FFpPVXprocessor::locateFirstTP, FFpFatigue.C:129-163), the
 * caller runs fsr_fatigue_locate_dev over the leading tiles until *n_pending == 0 (normally the
 * first tile) and then fsr_fatigue_feed_dev over ALL tiles from step 0.
 *  layout FSR_HIST_GAGE_MAJOR: hist_dev[g*ld + t]; FSR_HIST_STEP_MAJOR: hist_dev[t*ld + g];
 *  step0 = global index of the tile's first step.
 * fsr_fatigue_finish closes the residue (processFinish, FFpFatigue.C:274-320) and returns
 *  damage[ngage], ncycles[ngage], bins[ngage x nbins] (-1 like ffp_getnumcycles when there are no
 *  cycles or the bin starts above the largest range), status[ngage]: 0 ok, 1 = the reference's
 *  closure failure (cycles counted so far are kept, as ffp_getdamage ignores it), 2 = residue
 *  stack capacity exceeded (results -1).  Return value: number of gages with status != 0. */
typedef struct fsr_fatigue_state fsr_fatigue_state;
#define FSR_HIST_GAGE_MAJOR 0
#define FSR_HIST_STEP_MAJOR 1
int fsr_fatigue_create(fsr_fatigue_state **f, int device, int ngage, double gate, const double *curve,
                       double bin_size, int nbins, int stack_cap /* 0 = 1024 points per gage */);
/* per-gage gate values [ngage] and S-N curves [ngage x 4] (rosette%gateValue / %snCurve,
 * strainRosetteModule.f90:36-37); either may be NULL to keep the common value */
int fsr_fatigue_set_gage_params(fsr_fatigue_state *f, const double *gate, const double *curve);
int fsr_fatigue_reset(fsr_fatigue_state *f);
int fsr_fatigue_locate_dev(fsr_fatigue_state *f, const double *hist_dev, size_t ld, int layout, int step0,
                           int nsteps, int *n_pending, void *stream);
int fsr_fatigue_feed_dev(fsr_fatigue_state *f, const double *hist_dev, size_t ld, int layout, int step0,
                         int nsteps, void *stream);
int fsr_fatigue_finish(fsr_fatigue_state *f, double *damage, int *ncycles, int *bins, int *status);
int fsr_fatigue_finish_dev(fsr_fatigue_state *f, void *stream); /* asynchronous; results stay on device */
int fsr_fatigue_results_dev(fsr_fatigue_state *f, double **damage_dev, int **ncycles_dev, int **bins_dev,
                            int **status_dev);
void fsr_fatigue_destroy(fsr_fatigue_state *f);

/* ---- strain rosettes (fedem_gage) ------------------------------------------------------------
 * One &STRAIN_ROSETTE record as readStrainGageData delivers it
 * (src/vpmStress/strainGageModule.f90:107-237): nodes are INTERNAL node numbers (1-based, as in
 * mmnpc), rpos = posInGl(3,4) column-major (rosette X, Y, Z axes and position; the .fsi format
 * gives it explicitly), ngage / alpha_gages from the rosette type (SINGLE_GAGE 1/0,
 * DOUBLE_GAGE_90 2/pi/2, TRIPLE_GAGE_60 3/pi/3, TRIPLE_GAGE_45 3/pi/4), gate and sncurve
 * <= 0 = use the run's defaults (reportDamage, strainGageModule.f90:806-812). */
typedef struct fsr_rosette {
  int id, numnod, ngage, zero_init;
  int nodes[4];
  double rpos[12];
  double zpos, emod, nu, alpha_gages, gate;
  double sncurve[4]; /* loga1, loga2, m1, m2 */
} fsr_rosette;
typedef struct fsr_gages fsr_gages;
/* values per rosette and step written by fsr_gage_recover: [0:3) epsC, [3:6) epsP (max, min,
 * signed abs max), 6 gammaMax, 7 epsVM, 8 alpha1, 9 alphaGamma, [10:13) sigmaC, [13:16) sigmaP,
 * 16 tauMax, 17 sigmaVM, [18:21) gage strains, [21:24) gage stresses
 * (StrainRosetteType, strainRosetteModule.f90:29-58; calcRosetteStrains :251-324) */
#define FSR_GAGE_NVAL 24

/* Replaces ElDispFromSupElDisp + InitStrainRosette + InitStrainGages (gage.f90:169-232,
 * displacementModule.f90:1096-1202, strainRosetteModule.f90:587-812, strainGageModule.f90:604-661):
 * Bcart(3 x ndim) of every rosette, on the GPU, from the row operator of `part` (which must have
 * had fsr_set_recovery).  The part may be destroyed afterwards (like closeBandEmatrices, gage.f90:256). */
int fsr_gage_create(fsr_gages **gages, fsr_part *part, const fsr_rosette *ros, int nros);
int fsr_gage_num_series(const fsr_gages *gages); /* 4 per rosette: max principal stress, legs 1-3 */
int fsr_gage_get_bcart(fsr_gages *gages, double *bcart /* [nros][3 x ndim] column-major */);
/* Replaces the time loop body of gage.f90:293-355 (CalcZeroStartRosetteStrains,
 * CalcRosetteStrains) for nsteps steps: values [nsteps][nros][FSR_GAGE_NVAL] (may be NULL). */
int fsr_gage_recover(fsr_gages *gages, const double *Q, int ldq, int nsteps, double *values);
int fsr_gage_recover_dev(fsr_gages *gages, const double *Q_dev, int ldq, int nsteps,
                         double *values_dev, void *stream);
/* Replaces AddFatiguePoints + reportDamage (strainGageModule.f90:691-716,778-862): rainflow and
 * damage of sigmaP(1)*toMPa and every leg stress*toMPa; series 4*r = rosette r max principal,
 * 4*r+k = leg k (unused legs: no cycles).  damage/ncycles/status [4*nros], bins [4*nros x nbins]. */
int fsr_gage_fatigue(fsr_gages *gages, const double *Q, int ldq, int nsteps, double to_mpa,
                     double gate, const double *curve, double bin_size, int nbins, double *damage,
                     int *ncycles, int *bins, int *status);
/* streaming form for a device-resident history (see fsr_fatigue_locate_dev for the protocol):
 * mode 0 = locate pass (stops at the first tile after which *n_pending == 0), 1 = counting pass */
int fsr_gage_fatigue_begin(fsr_gages *gages, double to_mpa, double default_gate,
                           const double *default_curve, double bin_size, int nbins, int stack_cap);
int fsr_gage_fatigue_feed_dev(fsr_gages *gages, const double *Q_dev, int ldq, int step0, int nsteps,
                              int mode, int *n_pending, void *stream);
int fsr_gage_fatigue_end(fsr_gages *gages, double *damage, int *ncycles, int *bins, int *status);
void fsr_gage_destroy(fsr_gages *gages);

/* ---- strain coat summary (fedem_fpp's recovery-summary loop) -------------------------------------------
 * calcStrainCoatData (src/vpmStress/strainCoatModule.f90:315-480) with every rosette of `gages` as one coat result
 * point: running envelopes of the principal strains / stresses, max shear and von Mises (updateMax / updateMin, max
 * from 0, min from hugeVal), the angle bins of the principal directions (updateAngBin: -angleBins - 1 bins over 180
 * degrees) and the biaxiality sums gated on the signed abs-max principal stress (-biAxialGate); fsr_coat_end applies
 * calcAngleData (:481-547, useOldRange = false) and BiAxMean / BiAxStdDev (:672-704).
 *  env     [8][nros]  epsMax, epsMin, sigMax, sigMin, gammaMax, tauMax, vmeMax, vmsMax
 *  summary [6][nros]  stress range, strain range (largest range inside one angle bin), most popular angle [deg],
 *                     angle spread [deg], biaxiality mean, biaxiality standard deviation
 *  nbiax   [nros]     number of steps above the biaxiality gate.            Any output may be NULL. */
int fsr_coat_begin(fsr_gages *gages, int angle_bins /* -angleBins, 541 */, double biaxial_gate /* -biAxialGate, 10 */);
int fsr_coat_feed(fsr_gages *gages, const double *Q, int ldq, int nsteps);
int fsr_coat_feed_dev(fsr_gages *gages, const double *Q_dev, int ldq, int nsteps, void *stream);
int fsr_coat_end(fsr_gages *gages, double *env, double *summary, int *nbiax);
/* Fatigue of strain coat result points (fatigueAddPoint / fatigueDamage, src/vpmStress/fatigueModule.f90:40-118, with
 * calcStrainCoatData's fatValue = sigmaP(3) * toMPaScale * SCF, strainCoatModule.f90:385-400): after this call the first fatigue
 * series of every rosette (fsr_gage_fatigue*: series 4 r) is the SIGNED ABS-MAX principal stress times to_mpa times scf[r]
 * instead of the largest principal stress.  scf [nros]; NULL switches back. */
int fsr_gage_set_coat_fatigue(fsr_gages *gages, const double *scf);

/* S-N curve library file of fedem_fpp -SNfile (FFpSNCurveLib::readSNCurves / read, fedem-foundation/src/FFpLib/FFpFatigue/
 * FFpSNCurveLib.C:132-277): entries <standard name, standard id (0 NorSok, 1 British), <curve name, <loga1, m1, loga2, m2, ..>
 * [, thickness exponent]>, ..>, '#' comment lines.  Curves failing the reference's validity checks are dropped, standards
 * without a valid curve too, so (std_index, curve_index) address the same curves as FFpSNCurveLib::getCurve.
 *  fsr_sn_get : nseg line segments (loga[k], m[k], k < nseg <= cap) and for NorSok the nseg - 1 intersections logN0; returns
 *               nseg, or < 0 when the indices are out of range.  std_id may be NULL.
 *  fsr_sn_value : FFpSNCurve::getValue, cycles to failure at stress range s (FFpSNCurve.C:22-47); < 0 = no such curve. */
typedef struct fsr_sn_lib fsr_sn_lib;
int fsr_sn_read(fsr_sn_lib **lib, const char *path);
void fsr_sn_free(fsr_sn_lib *lib);
int fsr_sn_num_standards(const fsr_sn_lib *lib);
int fsr_sn_num_curves(const fsr_sn_lib *lib, int std_index);
int fsr_sn_get(const fsr_sn_lib *lib, int std_index, int curve_index, int *std_id, double *loga, double *m, double *logN0, int cap);
double fsr_sn_value(const fsr_sn_lib *lib, int std_index, int curve_index, double s);

/* ---- file formats and history assembly on the drop-in surface (host only) -------------------
 * Tagged binary files as written by writeTagDB / read by readTagDB (src/vpmUtilities/binaryDB.c:
 * 643-733; header layout FFaTag.C:192-297): 30-char tag, 0x1234 endian mark, 8-byte checksum field,
 * ";1.0;\n" (46 bytes), raw arrays.  Files of the other endianness are swapped on read.
 *
 * .fmx (dmOpen, src/vpmUtilities/diskMatrixModule.f90:263-301): column-major values; the dimensions
 * come from the .fsm.  tag NULL = "#FEDEM disk matrix" (the E-matrix uses "#FEDEM generalized
 * modes", displacementModule.f90:667); single_precision appends " SP" and stores floats. */
int fsr_fmx_write(const char *path, const char *tag, int checksum, const double *A, long long n,
                  int single_precision);
int fsr_fmx_read(const char *path, char *tag_out, int tag_cap, int *checksum, int *is_single,
                 double *A, long long n);
/* .fsm (saveSAM, src/vpmReducer/samReducerModule.f90:586-672; readSAMarrays,
 * src/vpmStress/samStressModule.f90:273-316).  fsr_fsm_read_mpar returns npar and the first `cap`
 * entries of mpar (mpar(1)=nnod, (2)=nel, (3)=ndof, (4)=ndof1, (5)=ndof2, (7)=nceq, (11)=neq,
 * (15)=nmmnpc, (16)=nmmceq, (18)=part base id, (22)=ngen, (24)=ndim); the caller sizes the arrays
 * and calls fsr_fsm_read (NULL = skip that array). */
int fsr_fsm_read_mpar(const char *path, int *checksum, int *mpar, int cap);
int fsr_fsm_read(const char *path, int *madof, int *minex, int *mnnn, int *msc, int *mpmnpc,
                 int *mmnpc, int *melcon, int *mpmceq, int *mmceq, double *ttcc, int *meqn,
                 int *meqn1, int *meqn2);
int fsr_fsm_write(const char *path, int checksum, int npar, const int *mpar, const int *madof,
                  const int *minex, const int *mnnn, const int *msc, const int *mpmnpc,
                  const int *mmnpc, const int *melcon, const int *mpmceq, const int *mmceq,
                  const double *ttcc, const int *meqn, const int *meqn1, const int *meqn2);
/* BuildFinit (src/vpmCommon/supElTypeModule.f90:1067-1114) for nsteps steps at once: co-rotated
 * deformational displacements of the triads + the component-mode amplitudes = the columns of Q.
 * 3x4 position matrices column-major: sup_tr [nsteps][12], triad_ur [nsteps][ntriads][12],
 * tr_undef [ntriads][12]; ndofs/first_dof [ntriads] (first_dof 1-based); gen_ur [nsteps][ngen]. */
int fsr_build_finit(int nsteps, int ntriads, const double *sup_tr, const double *triad_ur,
                    const double *tr_undef, const int *ndofs, const int *first_dof, int ngen,
                    const double *gen_ur, int gen_first_dof, double *Q, int ldq);
/* readSupElModes (src/vpmStress/modesRoutines.f90:121-203): the same for a mode shape -- fedem_modes is the K1 expansion
 * (fsr_expand / fsr_recover) with eigenvectors as Q.  triad_eig = the "Eigenvectors|Mode n" variables of the part's triads as
 * read from the solver results (global directions, nDOFs x ncomp values each, concatenated in triad order), gen_eig the
 * Part's [ngen x ncomp]; ncomp = 1 (2 for damped modes: real and imaginary part).  Q [ldq x ncomp] column-major. */
int fsr_build_mode_finit(int ntriads, const double *sup_tr, const int *ndofs, const int *first_dof,
                         const double *triad_eig, int ngen, int gen_first_dof, const double *gen_eig, int ncomp,
                         double *Q, int ldq);

/* .frs results database, reader side = what the recovery path uses of FFrExtractor through
 * ffr_init / ffr_findptr / ffr_getdata / ffr_setposition / ffr_increment
 * (fedem-foundation/src/FFrLib/FFrExtractor_F.C:33-263; header grammar FFrResultContainer.C:234-528,
 * record layout :534-585, time keys :714-860).  One handle owns any number of files (the solver
 * splits its output over th_p_*.frs / th_s_*.frs); the time steps are the sorted union of the
 * physical-time keys of all files.
 *  fsr_frs_find  : var_path = item-group names and the variable name joined by '|', og_type =
 *                  "Triad", "Part", ... (NULL/"" = a top-level variable), base_id = the object's base
 *                  id.  Returns a variable handle >= 0, or -1 when no file holds it.
 *  fsr_frs_read  : nw values of that variable for steps [step0, step0+nsteps) into data[s*ld + i]
 *                  (FLOAT 32/64 and INT 8..64 on file -> double); error if a step lacks the variable
 *                  or it is shorter than nw (ffr_getdata's ierr). */
typedef struct fsr_frs fsr_frs;
int fsr_frs_open(fsr_frs **db, const char *const *paths, int nfiles);
void fsr_frs_close(fsr_frs *db);
int fsr_frs_num_steps(const fsr_frs *db);
int fsr_frs_get_steps(const fsr_frs *db, int *stepno, double *time, int cap); /* returns the step count */
int fsr_frs_find(fsr_frs *db, const char *var_path, const char *og_type, int base_id);
int fsr_frs_var_size(const fsr_frs *db, int handle);
int fsr_frs_read(fsr_frs *db, int handle, int step0, int nsteps, double *data, int nw, int ld);
/* readSupElDisplacements (src/vpmStress/displacementModule.f90:434-524) for a window of steps:
 * "Position matrix" of every triad (6 DOFs; "Position" for 3-DOF triads) and of the part,
 * "Generalized displacement" of the part, then BuildFinit -> Q[ldq x nsteps].  Arguments as
 * fsr_build_finit; *_base_id = the base ids the solver wrote the objects with. */
int fsr_frs_reduced_history(fsr_frs *db, int sup_base_id, int ntriads, const int *triad_base_id,
                            const int *ndofs, const int *first_dof, const double *tr_undef, int ngen,
                            int gen_first_dof, int step0, int nsteps, double *Q, int ldq);
/* .frs writer core (src/vpmCommon/rdbModule.f90: openRDBfile :268-403, writeTimeStepDB :669-736):
 * tag line, the caller's text header (heading lines, VARIABLES:, DATABLOCKS: sections), "DATA:",
 * then per step int32 step number + double time + payload_bytes of the caller's record. */
typedef struct fsr_frs_writer fsr_frs_writer;
int fsr_frs_create(fsr_frs_writer **w, const char *path, int checksum, const char *header_text,
                   long long payload_bytes);
/* the same with another file tag (openRDBfile's 4th argument): "#FEDEM modal data" for the fedem_modes results */
int fsr_frs_create_tagged(fsr_frs_writer **w, const char *path, const char *tag, int checksum, const char *header_text,
                          long long payload_bytes);
int fsr_frs_write_step(fsr_frs_writer *w, int stepno, double time, const void *payload); /* returns steps written */
int fsr_frs_finish(fsr_frs_writer *w);

/* ---- FE part file (.ftl) ----------------------------------------------------------------------
 * Replaces what fedem_stress pulls from the FE-model singleton through ffl_init (stress.f90:111 ->
 * fedem-foundation/src/FFlLib/FFlLinkHandler_F.C:56-160) and, per element per step, through ffl_getcoor /
 * ffl_getmat / ffl_getthick / ffl_getbeamsection / ffl_getpinflags / ffl_getelmid (:699-1193): reads the
 * .ftl text file (grammar FFlIOAdaptors/FFlFedemReader.C:456-763) once and returns the same numbers as
 * flat arrays in SAM order, i.e. the members of fsr_elmdata.
 *  fsr_ftl_sizes        : sz[12] = nnod, nel, ndof, nmnpc, nmat, nxnod, npbeam, nrgd, nrbar, nwavgm, nprop,
 *                         ncons exactly as ffl_getsize (:367-431); returns the number of elements whose
 *                         calculation flag is on (its ierr).  initiateSAM compares these with the .fsm
 *                         (samStressModule.f90:118-135).
 *  fsr_ftl_activate_groups : the -group option, "55", "<33,22,44>", "<PMAT 33, PTHICK 55>" (FFlUtils.C:18-61);
 *                         returns the number of non-existing groups that were ignored.
 *  fsr_ftl_get_nodes    : ffl_getnodes (:459-556): madof [nnod+1], minex [nnod], mnode [nnod], msc [ndof],
 *                         xyz [nnod][3] (extra nodes of pinned beam ends last, minex < 0); returns nnod.
 *  fsr_ftl_get_topology : ffl_gettopol (:558-672): melcon [nel] (21/22 -> 23/24 when use_andes), mpmnpc
 *                         [nel+1], mmnpc [nmnpc]; returns nel.
 *  fsr_ftl_get_elmdata  : emod, rny, rho, thk [nel]; elmid [nel] (negative: outside the -group selection);
 *                         beam [nel][FSR_NBEAM]; status [nel] (0 ok, -2 no material, -3 invalid Poisson's
 *                         ratio / no beam section, -4 no shell thickness); returns the number of elements
 *                         with status != 0 (those get hugeVal results, stressRoutines.f90:237-241). */
typedef struct fsr_ftl fsr_ftl;
int fsr_ftl_open(fsr_ftl **ftl, const char *path);
void fsr_ftl_close(fsr_ftl *ftl);
int fsr_ftl_version(const fsr_ftl *ftl);
int fsr_ftl_activate_groups(fsr_ftl *ftl, const char *groups);
int fsr_ftl_sizes(const fsr_ftl *ftl, int *sz);
int fsr_ftl_get_nodes(const fsr_ftl *ftl, int *madof, int *minex, int *mnode, int *msc, double *xyz);
int fsr_ftl_get_topology(const fsr_ftl *ftl, int use_andes, int *melcon, int *mpmnpc, int *mmnpc);
int fsr_ftl_get_elmdata(const fsr_ftl *ftl, double *emod, double *rny, double *rho, double *thk, int *elmid,
                        double *beam, int *status);
int fsr_ftl_ext2int(const fsr_ftl *ftl, int is_node, int id); /* ffl_ext2int (:716-737) */
/* Strain coat elements of the FE part (STRCT3 / STRCQ4 / STRCT6 / STRCQ8 with their PSTRC result sets, PFATIGUE data and the
 * {FE id} reference to the underlying finite element) as ffl_getstraincoat delivers them one by one (FFlLinkHandler_F.C:1587-1701):
 * nodes = internal node numbers (1-based; every second node of the 6- and 8-noded elements is skipped, < 0 = non-existing node),
 * per result set k < npts: res_set 1 / 2 / 3 = "Bottom" / "Mid" / "Top" (0 = other), material id, E, nu of the PSTRC's PMAT, zpos =
 * PHEIGHT height or PTHICKREF factor * PTHICK thickness, sn_curve = {snCurveStd, snCurveIndex} of the PFATIGUE (-1, -1 without
 * one) and its stress concentration factor.  fsr_ftl_num_strain_coats = ffl_getnostrc (:1753-1762, calculation flag on);
 * fsr_ftl_get_strain_coats fills at most cap entries in element order and returns the total count. */
typedef struct fsr_strain_coat {
  int id, nnod, npts, elm_id;
  int nodes[8];
  int mat_id[3], res_set[3], sn_curve[3][2];
  double emod[3], nu[3], zpos[3], scf[3];
} fsr_strain_coat;
int fsr_ftl_num_strain_coats(const fsr_ftl *ftl);
int fsr_ftl_get_strain_coats(const fsr_ftl *ftl, fsr_strain_coat *out, int cap);

/* ---- solver input file (.fsi) -----------------------------------------------------------------------
 * Replaces readSolverData (src/vpmStress/displacementModule.f90:138-229 -> InitiateSupEls1/InitiateTriads/
 * InitiateSupEls2, src/vpmStress/initiateTriadAndSupElTypeModule.f90:34-319): from the Fortran namelist file of
 * the dynamics run, the &SUP_EL record with id = part_base_id, its &TRIAD_UNDPOS and &TRIAD records, &HEADING
 * modelFile and &ENVIRONMENT gravity.
 *  fsr_fsi_part   : returns the part's base id; user id, description, numTriads, numGenDOFs, supPos as a
 *                   column-major 3x4 matrix (= sup%supTr = sup%supTrInit), gravity[3], model file name.
 *  fsr_fsi_triads : per triad in the order of triadIds (= the order of the reduced DOFs in finit): base id,
 *                   user id, nDOFs, first reduced DOF (1-based), TrUndeformed and initial position ur
 *                   (column-major 3x4 each); returns sup%genDOFs%firstDOF.  Arguments may be NULL. */
typedef struct fsr_fsi fsr_fsi;
int fsr_fsi_open(fsr_fsi **fsi, const char *path, int part_base_id);
void fsr_fsi_close(fsr_fsi *fsi);
int fsr_fsi_part(const fsr_fsi *fsi, int *user_id, char *descr, int dcap, int *ntriads, int *ngen,
                 double *sup_pos, double *gravity, char *model_file, int mcap);
int fsr_fsi_triads(const fsr_fsi *fsi, int *base_id, int *user_id, int *ndofs, int *first_dof,
                   double *tr_undef, double *ur);
/* ReadStrainGages (src/vpmStress/strainGageModule.f90:78-237): the &STRAIN_ROSETTE records of a rosette input file in
 * .fsi format (gage.f90:137-139); ros[k].nodes are the EXTERNAL node numbers of the file (map them with
 * fsr_ftl_ext2int; solveGage also applies checkRosette's orientation swap).  A record of another part or with an
 * unknown type is an input error, as in the reference.  ros may be NULL to count.  Returns the record count. */
int fsr_fsi_read_rosettes(const char *path, int link_base_id, fsr_rosette *ros, int *user_id, char *descr,
                          int descr_stride, int cap);

/* ---- stress results database (.frs), written from the GPU ------------------------------------------
 * Replaces writeStressHeader (src/vpmStress/saveStressModule.f90:120-247, header grammar :625-1430) and the
 * per-element writeStressDB / writeStrMeasureDB calls of calcStresses (stressRoutines.f90:234-310,
 * saveStressModule.f90:1527-1633): the file has the reference's text header (meta data, VARIABLES:, item
 * group definitions, DATABLOCKS: with the Part's "Elements" list) and, per time step, int32 step number +
 * float64 time + [per node the 3 or 6 deformational (and total) displacements if FSR_OUT_DEFORMATION,
 * writeDisplacementDB :1437-1515] + for every element in SAM order [SR(6,nenod) if FSR_OUT_SR] then per result point
 * [stress][strain][selected measures], float32 unless double_precision.  fsr_rdb_write_steps recovers a
 * whole window of steps on the device (K1 + record kernels) and appends their records.
 * The file name gets the reference's "_<rdbinc>" suffix before the extension when rdbinc > 0
 * (openRDBfile, src/vpmCommon/rdbModule.f90:305-322). */
typedef struct fsr_rdb_options {
  unsigned out_mask;      /* FSR_OUT_* bits: -SR -stress -strain -vmStress ... (stressmain.C:46-60)       */
  int double_precision;   /* -double                                                                     */
  int rdbinc;             /* -rdbinc                                                                     */
  int part_base_id, part_user_id;
  const char *part_descr; /* sup%id%descr (the link file name when there is no solver input file)        */
  const char *model_file; /* AssociatedModelFileName, may be NULL                                        */
  const char *link_file;  /* ModelName (-linkfile), may be NULL                                          */
  const int *elmid;       /* [nel] external element ids as ffl_getelmid (< 1: skipped); NULL = 1..nel    */
  const char *module_name;/* NULL = "fedem_stress"                                                       */
  const int *minex;       /* [nnod] external node ids (FSR_OUT_DEFORMATION); NULL = 1..nnod              */
  const double *sup_tr_init; /* [12] column-major 3x4 initial position of the part (sup%supTrInit): with
                             FSR_OUT_DEFORMATION also the total displacements are written, as the current
                             reference does (iDef = 3, stress.f90:292; calcTotalNodalDisplacement,
                             displacementModule.f90:1694-1745); NULL = deformational displacements only   */
} fsr_rdb_options;
typedef struct fsr_rdb fsr_rdb;
int fsr_rdb_create(fsr_rdb **rdb, fsr_part *part, const char *path, const fsr_rdb_options *opt);
/* the same for a part recovered by a group of element blocks on several GPUs: every device fills the record slots of its
 * elements and a strided device-to-host copy drops them into the full step records (nodal deformation output is written
 * by a single device only: use fsr_rdb_create for FSR_OUT_DEFORMATION) */
int fsr_rdb_create_group(fsr_rdb **rdb, fsr_group *group, const char *path, const fsr_rdb_options *opt);
/* host only: the header text and the bytes per step for a part given by its SAM element type codes (melcon)
 * and opt->elmid; header may be NULL to query the length, which is returned */
int fsr_rdb_build_header(int nnod, const int *madof, int nel, const int *melcon, const fsr_rdb_options *opt,
                         char *header, int cap, long long *step_bytes);
long long fsr_rdb_step_bytes(const fsr_rdb *rdb);            /* bytes per time step incl. the 12-byte key */
int fsr_rdb_header(const fsr_rdb *rdb, char *buf, int cap);  /* the text header; returns its length      */
int fsr_rdb_path(const fsr_rdb *rdb, char *buf, int cap);    /* the actual file name                     */
/* sup_tr [nsteps][12]: column-major 3x4 position matrix of the part at every step; only read when the
 * total displacements are written (opt->sup_tr_init given), else may be NULL */
int fsr_rdb_write_steps(fsr_rdb *rdb, const double *Q, int ldq, int nsteps, const int *stepno,
                        const double *time, const double *sup_tr);
/* the same from nodal displacements that are already there (fsr_recover_displacements): sv_hist [nsteps x ndof] */
int fsr_rdb_write_steps_displacements(fsr_rdb *rdb, const double *sv_hist, int nsteps, const int *stepno,
                                      const double *time, const double *sup_tr);
/* fsr_rdb_write_steps is pipelined: it returns once the window is queued; while the device recovers tile n, tile n-1
 * crosses PCIe and tile n-2 is appended to the file by a writer thread.  fsr_rdb_flush waits until every record handed
 * over so far is on file and reports where the time went: t[0] = device time of the tiles (H2D of Q, K1, record
 * kernels), t[1] = device-to-host copies, t[2] = file writes -- milliseconds, each summed over the tiles (the three
 * overlap in wall time) -- t[3] = bytes written, t[4] = tiles, t[5] = the part of t[0] spent on the H2D copy of Q and the
 * expansion (K1).  t may be NULL.  Returns the entries written. */
int fsr_rdb_flush(fsr_rdb *rdb, double *t, int n);
/* calcTotalNodalDisplacement for one node on the host (the same code the record kernel runs): x0[3], u[nd],
 * nd = 3 or 6, T / T0 = current / initial 3x4 position matrices; utot[nd] */
void fsr_total_nodal_displacement(const double *x0, const double *u, int nd, const double *T, const double *T0,
                                  double *utot);
int fsr_rdb_close(fsr_rdb *rdb);

/* ---- the fedem_stress program ------------------------------------------------------------------------
 * The reference's launcher entry points, same names (src/vpmStress/stressInterface.C:86-116):
 * initSolverArgs (re)starts the option parser over the given arguments and defines the option table of stressmain.C:22-79 (+ -fao/-fco/-fop/-cwd/-help/-debug ... of
 * cmdLineArgInitStd.C / cmdLineArgInit.C) over the given arguments, solveStress runs subroutine stress
 * (src/vpmStress/stress.f90): -linkfile .ftl, -samfile .fsm, -fsifile .fsi, -Bmatfile/-eigfile/-dispfile .fmx,
 * -frsfile solver results, -statm/-stotm/-tinc, -group, -SR -stress -strain -vmStress ... -deformation -double,
 * -rdbfile/-rdbinc stress results database.  Returns 0 on success.  bin/fedem_stress is main() over these. */
void initSolverArgs(int argc, char **argv);
int solveStress(void);
/* solveGage (src/vpmStress/stressInterface.C:118-123) runs subroutine gage (src/vpmStress/gage.f90) with the option
 * table of gagemain.C:22-61: the same part / history inputs as solveStress, -rosfile with the &STRAIN_ROSETTE records
 * (.fsi format), rosette strains and stresses of every selected step to -rdbfile (saveStrainGageModule.f90 grammar),
 * -fatigue > 0: rainflow + damage report of sigmaP(1) and the gage legs in the -resfile (reportDamage).
 * bin/fedem_gage is main() over initSolverArgs + solveGage. */
int solveGage(void);
/* solveModes (src/vpmStress/stressInterface.C:125-130, modesmain.C:15-56) runs subroutine modes (src/vpmStress/modes.f90):
 * for every time listed in -recover_modes <t1 m1 m2 ..> <t2 ..> the dynamic response and the listed eigenmodes of the
 * solver's modal results ("Eigenvectors|Mode n" of the part's triads and of the part) are expanded to all nodes (K1) and
 * written as vector data to one modal results file (writeModesHeader / writeDisplacementDB "Vectors" grammar, file tag
 * "#FEDEM modal data"); -damped = complex modes (Re / Im).  Different mode lists per time, -write_nodes or -energy_density
 * give one file for the dynamic response and one per mode instead (writeModeHeader :259-343, file increments in creation
 * order), with the nodal form (writeNodesHeader :459-537) and the scaled strain energy density per result point
 * (calcStrainEnergyDensity, modesRoutines.f90:219-305: one full K1 + K2 pass per mode).  VTF export is not part of this build. */
int solveModes(void);
/* solveFpp (src/vpmStress/stressInterface.C:117-124, fppmain.C:15-72) runs subroutine fpp (src/vpmStress/fpp.f90): the strain coat
 * elements of the FE part (-surface selects the result sets) become rosettes in their element coordinate systems
 * (initiateStrainCoats, strainCoatModule.f90:172-312; InitStrainRosette with useElCoordSys); over the selected time steps the
 * running envelopes, angle bins and biaxiality sums (calcStrainCoatData) and, with -HistDataType 1 and S-N curves assigned
 * (PFATIGUE + -SNfile), rainflow damage of the signed abs-max principal stress; ONE summary record per run on the strain coat
 * results database (saveStrainCoatModule.f90: file tag "#FEDEM strain coat data", step 1 at the stop time), the tables of
 * printStrainCoatInput / printStrainCoatData in the -resfile with -debug > 0.  -writeHistory gives the per-step history form
 * instead (writeHistoryHeader / writeHistoryDB).  Not part of this build: the nCode FPP plug-in (-HistDataType < 0, -fppfile;
 * FT_HAS_FPPINTERFACE) and residual stress import.  bin/fedem_fpp is main() over initSolverArgs + solveFpp. */
int solveFpp(void);
void fsr_fpp_define_options(void);     /* the option table of fedem_fpp (fppmain.C:21-68) */
void fsr_modes_define_options(void);   /* the option table of fedem_modes (modesmain.C:22-52) */
/* ffr_getnextstep (fedem-foundation/src/FFrLib/FFrExtractorInterface.f90:134-170) over a sorted key list:
 * indices of the time steps the stress loop visits for -statm start -stotm stop -tinc tinc; returns their
 * number (out may be NULL) */
int fsr_select_steps(const double *times, int n, double start, double stop, double tinc, int *out, int cap);
/* the option parser (FFaCmdLineArg semantics) behind a C face, one global instance like the reference's */
void fsr_cmdline_reset(void);
void fsr_stress_define_options(void);
void fsr_gage_define_options(void);   /* the option table of fedem_gage (gagemain.C:22-61) */
void fsr_cmdline_add_bool(const char *name, int value);
void fsr_cmdline_add_int(const char *name, int value);
void fsr_cmdline_add_double(const char *name, double value);
void fsr_cmdline_add_string(const char *name, const char *value);
void fsr_cmdline_init(int argc, char **argv);
int fsr_cmdline_read_file(const char *path);
int fsr_cmdline_get_bool(const char *name);
int fsr_cmdline_get_int(const char *name);
double fsr_cmdline_get_double(const char *name);
int fsr_cmdline_get_string(const char *name, char *out, int cap);
int fsr_cmdline_is_set(const char *name);

/* ---- in-core part state for fedempy ----------------------------------------------------------------
 * The four functions fedempy's FedemSolver.save_part_state / get_part_*_state_size call
 * (PythonAPI/src/fedempy/solver.py:524-629), with the reference's names, argument lists and data layout
 * (src/vpmSolver/solverInterface.C:940-1001 -> solverModule.f90:2183-2263 -> stressRecoveryModule.f90:115-267):
 *  data[0:4] = step number, time, time step size, part base id; then
 *  deformation: 3 translational deformations per node in SAM node order (0 for minex <= 0), size 3*nnod + 4;
 *  stress     : the vms array of fsr_get_vms, size fsr_vms_size + 4.
 * Sizes: -1 = no such part, -999 = no part registered at all (the reference's "not allocated").
 * fsr_recovery_register makes a part known under its base id (minex [nnod] may be NULL);
 * fsr_recovery_update is the solver's per-step recovery: it expands q = [finit; vg] of the converged step and
 * evaluates the von Mises stresses on the device, keeping both in core for the save calls. */
int fsr_recovery_register(int base_id, fsr_part *part, const int *minex);
int fsr_recovery_unregister(int base_id);
int fsr_recovery_update(int base_id, int step, double time, double time_step, const double *q);
/* the same for several parts at once -- the loop over the parts of stressRecoveryModule.f90:1021-1061: the device work of
 * all parts (each on its own stream / device) is queued first and waited for afterwards, so the parts of a mechanism
 * overlap; q[k] = [finit; vg] of part base_ids[k] */
int fsr_recovery_update_parts(int nparts, const int *base_ids, int step, double time, double time_step,
                              const double *const *q);
/* The recovery switches of the dynamics solver (src/vpmSolver/solverInterface.C:447-452), given like on its command line:
 *   -recovery N         1 = stress, 2 = gages, 3 = both; stress recovery runs for mod(N, 2) == 1 (stressRecoveryModule.f90:1025)
 *   -partVMStress N     0 off, 1 von Mises to the frs file, 2 through the state array (savePartStressState), 3 both (:583,655)
 *   -partDeformation N  0 off, 1 deformational displacements to the frs file, 2 / 3 total displacements as well (:1017,1183)
 *   -frs3file NAME      one file, or <"a.frs","b.frs">: one per recovered part in registration order (:771-815,860-884)
 *   -double, -modelfile NAME; other options are skipped.  Switches not named take the solver's defaults (0, 1, 1).
 * The switches apply to the parts registered afterwards.  Without this call the registry keeps stress recovery and the
 * state arrays on and writes no file.  fsr_recovery_register_part also names the part in the frs header (user_id, descr)
 * and takes its initial position sup_tr_init [12] (column-major 3x4; NULL = no total displacements).
 * fsr_recovery_update_parts_save is the per-step call with the solver's doSave flag (0 = recoverNotSave) and the parts'
 * positions sup_tr[k] [12] at the step (NULL when no total displacements are written): the step is appended to the
 * parts' frs files through the pipelined results-database writer.  fsr_recovery_close = closeRecovery: flushes and closes
 * the files, empties the registry, resets the switches.  fsr_recovery_file: the frs file of a part (length, 0 = none). */
int fsr_recovery_options(const char *args);
int fsr_recovery_register_part(int base_id, int user_id, const char *descr, fsr_part *part, const int *minex,
                               const double *sup_tr_init);
int fsr_recovery_update_parts_save(int nparts, const int *base_ids, int step, double time, double time_step,
                                   const double *const *q, const double *const *sup_tr, int do_save);
int fsr_recovery_close(void);
int fsr_recovery_file(int base_id, char *buf, int cap);
int getPartDeformationStateSize(int bid);
int getPartStressStateSize(int bid);
bool savePartDeformationState(int bid, double *data, int ndat);
bool savePartStressState(int bid, double *data, int ndat);

/* ---- diagnostics --------------------------------------------------------------------------- */
const char *fsr_last_error(void);
/* Number of kernels this library launched since the counter was last reset (bench evidence). */
long long fsr_kernel_launches(int reset);
/* Device time (ms), summed over the step tiles launched since the last fsr_timing_reset (or the
 * start of the last fsr_recover), split per kernel family: t[0] = Q packing + K1 expansion,
 * t[1] = K2 element kernels (+fused envelope), t[2] = number of tiles timed.  Measured with CUDA
 * events recorded on the stream the kernels were launched on; synchronises on the last event.
 * Returns the number of entries written. */
int fsr_last_timing(fsr_part *part, double *t_ms, int n);
/* counts[3 f + 0..2] = elements of family f (0 ANDES quad, 1 ANDES triangle, 2 TET10, 3 beam, 4 HEX20, 5 HEX8, 6 TET4,
 * 7 WEDG6, 8 WEDG15, 9 TRI6, 10 QUAD8), of which on the geometry fast path (flat quads: membrane / bending split;
 * straight-sided TET10: corner gradients), of which on the general kernel.  cap = entries of counts.  Returns the
 * number of families. */
int fsr_family_counts(const fsr_part *part, int *counts, int cap);
/* What the von Mises path of the part expands (K1) and where its quadrilaterals go.  Nodes whose flat quadrilaterals all
 * lie in one plane get four rows (u, v, theta1, theta2) in the axes of that plane instead of six global ones (the rotation
 * is folded into the recovery operator once, fsr_set_recovery); FSR_QUAD_PLANAR=0 in the environment switches that off.
 * info[0] = rows K1 expands per step tile, [1] = nodal DOFs of the part, [2] = in-plane rows, [3] = 128-row tiles of the
 * global operator still expanded (read by other elements), [4] = quadrilaterals in the in-plane form, [5] = flat
 * quadrilaterals on global rows, [6] = quadrilaterals on the dense operator.  Returns the number of entries written. */
int fsr_vm_path_info(const fsr_part *part, long long *info, int cap);
int fsr_timing_reset(fsr_part *part);

#ifdef __cplusplus
}
#endif
#endif /* FEDEM_B200_H */
