"""Host-side mirror of the reference's file layer for the recovery path: `.fmx` disk matrices
(dmOpen, src/vpmUtilities/diskMatrixModule.f90:190-409), `.fsm` SAM files (saveSAM /
initiateSAM, src/vpmReducer/samReducerModule.f90:586-672, src/vpmStress/samStressModule.f90:39-316)
and the assembly of the reduced history from position matrices (BuildFinit,
src/vpmCommon/supElTypeModule.f90:1067-1114).  The byte-level work is in libfedem_b200.so
(csrc/io_files.cu); file naming follows the reducer (<part>_SAM.fsm, <part>_B.fmx, <part>_E.fmx,
reducer.f90:367-405)."""
import ctypes as C
import os
import numpy as np

from . import _lib
from ._lib import check
from .model import SamData, ElementData, PartModel

F64 = np.float64
I32 = np.int32
DM_TAG = "#FEDEM disk matrix"
GM_TAG = "#FEDEM generalized modes"
NPAR = 50   # size of mpar (samModule.f90:261)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def write_fmx(path, A, tag=DM_TAG, checksum=0, single_precision=False):
    lib = _lib.load_library()
    A = np.asfortranarray(A, F64)
    check(lib.fsr_fmx_write(os.fsencode(path), tag.encode(), int(checksum), _dp(A), A.size, int(single_precision)),
          "fsr_fmx_write")


def read_fmx(path, nrows, ncols, want_tag=None):
    """Returns (A [nrows, ncols] Fortran order, tag, checksum, stored_as_float)."""
    lib = _lib.load_library()
    A = np.zeros((nrows, ncols), F64, order="F")
    tag = C.create_string_buffer(40)
    cs, sp = C.c_int(), C.c_int()
    check(lib.fsr_fmx_read(os.fsencode(path), tag, 40, C.byref(cs), C.byref(sp), _dp(A), A.size), "fsr_fmx_read")
    t = tag.value.decode()
    if want_tag is not None and t != want_tag:   # dmOpen's wantTag check (diskMatrixModule.f90:291-295)
        raise _lib.FsrError(f"Invalid disk matrix file {path}: file tag '{t}', expected '{want_tag}'")
    return A, t, cs.value, bool(sp.value)


def sam_mpar(sam, part_id=0):
    """mpar(50) as the reducer fills it (samModule.f90:267-290, reducer.f90:373): only the entries the
    recovery path reads are set."""
    mpar = np.zeros(NPAR, I32)
    mpar[0], mpar[1], mpar[2], mpar[3], mpar[4] = sam.nnod, sam.nel, sam.ndof, sam.ndof1, sam.ndof2
    mpar[6], mpar[10], mpar[14], mpar[15] = sam.nceq, sam.neq, len(sam.mmnpc), len(sam.mmceq)
    mpar[17], mpar[21], mpar[23] = part_id, sam.ngen, sam.ndof2 + sam.ngen
    return mpar


def write_fsm(path, sam, checksum=0, part_id=0):
    lib = _lib.load_library()
    c = lambda a: np.ascontiguousarray(a, I32)
    mpar = sam_mpar(sam, part_id)
    minex = c(sam.minex if sam.minex is not None else np.arange(1, sam.nnod + 1))
    mnnn = np.ones(sam.nnod, I32)
    arrs = [c(sam.madof), minex, mnnn, c(sam.msc), c(sam.mpmnpc), c(sam.mmnpc), c(sam.melcon), c(sam.mpmceq),
            c(sam.mmceq if len(sam.mmceq) else np.zeros(1, I32))]
    ttcc = np.ascontiguousarray(sam.ttcc if len(sam.ttcc) else np.zeros(1), F64)
    tail = [c(sam.meqn), c(sam.meqn1 if sam.ndof1 else np.zeros(1, I32)), c(sam.meqn2 if sam.ndof2 else np.zeros(1, I32))]
    check(lib.fsr_fsm_write(os.fsencode(path), int(checksum), NPAR, _ip(mpar), *[_ip(a) for a in arrs], _dp(ttcc),
                            *[_ip(a) for a in tail]), "fsr_fsm_write")


def read_fsm(path):
    """initiateSAM's file part: returns (SamData, mpar, checksum).  msc is returned as stored
    (2 = external, 1 = free, 0 = fixed/dependent), i.e. BEFORE the remap of samStressModule.f90:245-256."""
    lib = _lib.load_library()
    mpar = np.zeros(64, I32)
    cs = C.c_int()
    npar = check(lib.fsr_fsm_read_mpar(os.fsencode(path), C.byref(cs), _ip(mpar), 64), "fsr_fsm_read_mpar")
    nnod, nel, ndof, ndof1, ndof2, nceq, neq, nmmnpc, nmmceq = (int(mpar[i]) for i in (0, 1, 2, 3, 4, 6, 10, 14, 15))
    z = lambda n: np.zeros(max(n, 1), I32)
    madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon = z(nnod + 1), z(nnod), z(nnod), z(ndof), z(nel + 1), z(nmmnpc), z(nel)
    mpmceq, mmceq, ttcc = z(nceq + 1), z(nmmceq), np.zeros(max(nmmceq, 1), F64)
    meqn, meqn1, meqn2 = z(ndof), z(ndof1), z(ndof2)
    check(lib.fsr_fsm_read(os.fsencode(path), _ip(madof), _ip(minex), _ip(mnnn), _ip(msc), _ip(mpmnpc), _ip(mmnpc),
                           _ip(melcon), _ip(mpmceq), _ip(mmceq), _dp(ttcc), _ip(meqn), _ip(meqn1), _ip(meqn2)),
          "fsr_fsm_read")
    sam = SamData(nnod=nnod, nel=nel, ndof=ndof, ndof1=ndof1, ndof2=ndof2, ngen=int(mpar[21]), neq=neq, nceq=nceq,
                  madof=madof[:nnod + 1], msc=msc[:ndof], mpmnpc=mpmnpc[:nel + 1], mmnpc=mmnpc[:nmmnpc],
                  melcon=melcon[:nel], meqn=meqn[:ndof], meqn1=meqn1[:ndof1], meqn2=meqn2[:ndof2],
                  mpmceq=mpmceq[:nceq + 1], mmceq=mmceq[:nmmceq], ttcc=ttcc[:nmmceq], minex=minex[:nnod])
    return sam, mpar[:npar].copy(), cs.value


def save_part(prefix, part, checksum=0, part_id=0, b_single_precision=False):
    """What the reducer leaves for one FE part: <prefix>_SAM.fsm, <prefix>_B.fmx, <prefix>_E.fmx."""
    write_fsm(prefix + "_SAM.fsm", part.sam, checksum, part_id)
    if part.B is not None and part.sam.ndof2 > 0:
        write_fmx(prefix + "_B.fmx", part.B, DM_TAG, checksum, b_single_precision)
    if part.E is not None and part.sam.ngen > 0:
        write_fmx(prefix + "_E.fmx", part.E, GM_TAG, checksum)


def load_part(prefix, elm: ElementData):
    """initiateSAM + openBandEmatrices from files; the element data (the .ftl side) is handed in."""
    sam, mpar, cs = read_fsm(prefix + "_SAM.fsm")
    B = E = None
    if sam.ndof2 > 0 and sam.ndof1 > 0:
        B, _, csb, _ = read_fmx(prefix + "_B.fmx", sam.ndof1, sam.ndof2, DM_TAG)
        if csb != cs:   # the reference compares the checksums of the FE data (dmOpen fileChkSum)
            raise _lib.FsrError(f"{prefix}_B.fmx: checksum {csb} does not match the SAM file ({cs})")
    if sam.ngen > 0 and sam.ndof1 > 0:
        E, _, _, _ = read_fmx(prefix + "_E.fmx", sam.ndof1, sam.ngen, GM_TAG)
    return PartModel(sam=sam, elm=elm, B=B, E=E, name=os.path.basename(prefix))


def build_finit(sup_tr, triad_ur, tr_undef, ndofs, first_dof, gen_ur=None, gen_first_dof=0, ndim=None):
    """BuildFinit for a window of steps.  sup_tr [nsteps, 3, 4], triad_ur [nsteps, ntriads, 3, 4],
    tr_undef [ntriads, 3, 4]; returns Q [ndim, nsteps] (Fortran order)."""
    lib = _lib.load_library()
    sup_tr = np.asarray(sup_tr, F64); triad_ur = np.asarray(triad_ur, F64); tr_undef = np.asarray(tr_undef, F64)
    ns, nt = sup_tr.shape[0], tr_undef.shape[0]
    cm = lambda a: np.ascontiguousarray(np.swapaxes(a, -1, -2))     # [.., 3, 4] -> column-major 12 doubles
    ndofs = np.ascontiguousarray(ndofs, I32); first_dof = np.ascontiguousarray(first_dof, I32)
    ngen = 0 if gen_ur is None else np.asarray(gen_ur).shape[1]
    g = np.ascontiguousarray(gen_ur, F64) if ngen else None
    if ndim is None:
        ndim = int(max([f + min(n, 6) - 1 for f, n in zip(first_dof, ndofs)] + [gen_first_dof + ngen - 1]))
    Q = np.zeros((ndim, ns), F64, order="F")
    check(lib.fsr_build_finit(ns, nt, _dp(cm(sup_tr)), _dp(cm(triad_ur)), _dp(cm(tr_undef)), _ip(ndofs), _ip(first_dof),
                              ngen, _dp(g), int(gen_first_dof), _dp(Q), ndim), "fsr_build_finit")
    return Q
