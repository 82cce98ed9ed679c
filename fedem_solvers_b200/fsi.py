"""Solver input files (.fsi): reader binding (csrc/io_fsi.cu, replaces readSolverData of
src/vpmStress/displacementModule.f90:138-229) and a writer for the records the stress recovery reads
(&HEADING, &ENVIRONMENT, &TRIAD, &SUP_EL, &TRIAD_UNDPOS), in the layout the FEDEM GUI writes them."""
import ctypes as C
import os
from dataclasses import dataclass
import numpy as np

from . import _lib
from ._lib import check

F64 = np.float64
I32 = np.int32


@dataclass
class SolverPart:
    """What fedem_stress keeps of the solver model for one part (SupElType + its TriadTypes)."""
    base_id: int
    user_id: int
    descr: str
    ngen: int
    sup_pos: np.ndarray         # [3, 4] sup%supTr = sup%supTrInit
    gravity: np.ndarray         # [3]
    model_file: str
    triad_base_id: np.ndarray   # [ntriads], order of triadIds = order of the reduced DOFs
    triad_user_id: np.ndarray
    ndofs: np.ndarray
    first_dof: np.ndarray       # 1-based position in finit
    tr_undef: np.ndarray        # [ntriads, 3, 4] sup%TrUndeformed
    triad_ur: np.ndarray        # [ntriads, 3, 4] initial triad positions
    gen_first_dof: int

    @property
    def ndim(self):
        return self.gen_first_dof - 1 + self.ngen


def read_fsi(path, part_base_id):
    lib = _lib.load_library()
    h = C.c_void_p()
    check(lib.fsr_fsi_open(C.byref(h), os.fsencode(path), int(part_base_id)), "fsr_fsi_open")
    try:
        user, nt, ng = C.c_int(), C.c_int(), C.c_int()
        descr, mfile = C.create_string_buffer(256), C.create_string_buffer(1024)
        sp, g = np.zeros(12, F64), np.zeros(3, F64)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        base = check(lib.fsr_fsi_part(h, C.byref(user), descr, 256, C.byref(nt), C.byref(ng), dp(sp), dp(g), mfile, 1024),
                     "fsr_fsi_part")
        n = nt.value
        tb, tu, nd, fd = (np.zeros(max(n, 1), I32) for _ in range(4))
        tr, ur = np.zeros((max(n, 1), 12), F64), np.zeros((max(n, 1), 12), F64)
        gfd = check(lib.fsr_fsi_triads(h, ip(tb), ip(tu), ip(nd), ip(fd), dp(tr), dp(ur)), "fsr_fsi_triads")
        cm = lambda a: np.swapaxes(a[:n].reshape(n, 4, 3), 1, 2).copy()   # column-major 12 -> [3, 4]
        return SolverPart(base_id=base, user_id=user.value, descr=descr.value.decode("latin1"), ngen=ng.value,
                          sup_pos=sp.reshape(4, 3).T.copy(), gravity=g, model_file=mfile.value.decode("latin1"),
                          triad_base_id=tb[:n], triad_user_id=tu[:n], ndofs=nd[:n], first_dof=fd[:n], tr_undef=cm(tr),
                          triad_ur=cm(ur), gen_first_dof=gfd)
    finally:
        lib.fsr_fsi_close(h)


def _mat(name, m, indent):
    rows = ["  ".join(f"{v: .9e}" for v in r) for r in np.asarray(m, F64)]
    pad = " " * (indent + len(name) + 3)
    return f"{' ' * indent}{name} = " + ("\n" + pad).join(rows)


def write_fsi(path, parts, gravity=(0.0, 0.0, 0.0), model_file="model.fmm"):
    """parts: list of SolverPart (first_dof / gen_first_dof are derived by the reader and not written)."""
    out = ["&HEADING", f"  modelFile = '{model_file}'", "  version = 3.0", "/", "",
           "&ENVIRONMENT", "  gravity = " + " ".join(f"{v: .9e}" for v in gravity), "/", ""]
    seen = set()
    for p in parts:
        for j, b in enumerate(p.triad_base_id):
            if int(b) in seen:
                continue
            seen.add(int(b))
            out += ["&TRIAD", f"  id = {int(b)}", f"  extId = {int(p.triad_user_id[j])}", f"  extDescr = 'Triad {int(b)}'",
                    f"  nDOFs = {int(p.ndofs[j])}", _mat("ur ", p.triad_ur[j], 2), "/", ""]
    for p in parts:
        out += ["&SUP_EL", f"  id = {p.base_id}", f"  extId = {p.user_id}", f"  extDescr = '{p.descr}'",
                f"  numGenDOFs = {p.ngen}", f"  numTriads = {len(p.triad_base_id)}",
                "  triadIds = " + " ".join(str(int(b)) for b in p.triad_base_id), _mat("supPos", p.sup_pos, 2), "/"]
        for j, b in enumerate(p.triad_base_id):
            out += ["&TRIAD_UNDPOS", f"  supElId = {p.base_id}", f"  triadId = {int(b)}",
                    _mat("undPosInSupElSystem", p.tr_undef[j], 2), "/"]
        out.append("")
    with open(path, "w") as f:
        f.write("\n".join(out))
