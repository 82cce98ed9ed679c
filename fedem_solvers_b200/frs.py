"""Host-side mirror of the reference's results-database layer for the recovery path: the `.frs` reader
the stress modules drive through ffr_init / ffr_findPtr / ffr_getData
(fedem-foundation/src/FFrLib/FFrExtractor_F.C:33-263, FFrExtractorInterface.f90), the assembly of the
reduced history from it (readSupElDisplacements, src/vpmStress/displacementModule.f90:434-524) and the
`.frs` writer core of src/vpmCommon/rdbModule.f90.  The byte-level work is in libfedem_b200.so
(csrc/io_frs.cu)."""
import ctypes as C
import os
import numpy as np

from . import _lib
from ._lib import check

F64 = np.float64
I32 = np.int32


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class FrsReader:
    """One FFrExtractor-like handle over any number of .frs files (ffr_init)."""

    def __init__(self, paths):
        if isinstance(paths, (str, bytes, os.PathLike)):
            paths = [paths]
        self.lib = _lib.load_library()
        self.paths = [os.fspath(p) for p in paths]
        arr = (C.c_char_p * len(self.paths))(*[os.fsencode(p) for p in self.paths])
        self._h = C.c_void_p()
        check(self.lib.fsr_frs_open(C.byref(self._h), arr, len(self.paths)), "fsr_frs_open")
        n = check(self.lib.fsr_frs_num_steps(self._h), "fsr_frs_num_steps")
        self.step_numbers = np.zeros(max(n, 1), I32)
        self.times = np.zeros(max(n, 1), F64)
        check(self.lib.fsr_frs_get_steps(self._h, _ip(self.step_numbers), _dp(self.times), n), "fsr_frs_get_steps")
        self.step_numbers, self.times = self.step_numbers[:n], self.times[:n]

    @property
    def nsteps(self):
        return len(self.times)

    def find(self, var_path, og_type="", base_id=0):
        """ffr_findPtr: returns a variable handle, or None when no file holds the variable."""
        h = self.lib.fsr_frs_find(self._h, var_path.encode(), (og_type or "").encode(), int(base_id))
        if h < -1:
            check(h, "fsr_frs_find")
        return None if h < 0 else h

    def var_size(self, handle):
        return check(self.lib.fsr_frs_var_size(self._h, handle), "fsr_frs_var_size")

    def read(self, handle, step0=0, nsteps=None, nw=None):
        """ffr_getData over a window of steps: [nsteps, nw] doubles."""
        nsteps = self.nsteps - step0 if nsteps is None else nsteps
        nw = self.var_size(handle) if nw is None else nw
        out = np.zeros((nsteps, nw), F64)
        check(self.lib.fsr_frs_read(self._h, handle, step0, nsteps, _dp(out), nw, nw), "fsr_frs_read")
        return out

    def reduced_history(self, sup_base_id, triad_base_ids, ndofs, first_dof, tr_undef, ngen=0, gen_first_dof=0,
                        step0=0, nsteps=None, ndim=None):
        """readSupElDisplacements + BuildFinit for a window of steps -> Q [ndim, nsteps] (Fortran order).
        tr_undef [ntriads, 3, 4] = sup%TrUndeformed."""
        nsteps = self.nsteps - step0 if nsteps is None else nsteps
        tb = np.ascontiguousarray(triad_base_ids, I32)
        nd = np.ascontiguousarray(ndofs, I32)
        fd = np.ascontiguousarray(first_dof, I32)
        tu = np.ascontiguousarray(np.swapaxes(np.asarray(tr_undef, F64), -1, -2))  # column-major 3x4
        if ndim is None:
            ndim = int(max([f + min(n, 6) - 1 for f, n in zip(fd, nd)] + [gen_first_dof + ngen - 1]))
        Q = np.zeros((ndim, nsteps), F64, order="F")
        check(self.lib.fsr_frs_reduced_history(self._h, int(sup_base_id), len(tb), _ip(tb), _ip(nd), _ip(fd), _dp(tu),
                                               int(ngen), int(gen_first_dof), step0, nsteps, _dp(Q), ndim),
              "fsr_frs_reduced_history")
        return Q

    def close(self):
        if self._h:
            self.lib.fsr_frs_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrsWriter:
    """openRDBfile + writeTimeStepDB: tag line, text header, 'DATA:', then one record per step
    (int32 step number, float64 time, payload)."""

    def __init__(self, path, header_text, payload_bytes, checksum=0):
        self.lib = _lib.load_library()
        self._h = C.c_void_p()
        self.payload_bytes = int(payload_bytes)
        check(self.lib.fsr_frs_create(C.byref(self._h), os.fsencode(path), int(checksum), header_text.encode("latin1"),
                                      self.payload_bytes), "fsr_frs_create")

    def write_step(self, stepno, time, payload):
        buf = np.ascontiguousarray(payload)
        if buf.nbytes != self.payload_bytes:
            raise _lib.FsrError(f"record payload is {buf.nbytes} bytes, the header declares {self.payload_bytes}")
        return check(self.lib.fsr_frs_write_step(self._h, int(stepno), float(time), buf.ctypes.data_as(C.c_void_p)),
                     "fsr_frs_write_step")

    def close(self):
        if self._h:
            check(self.lib.fsr_frs_finish(self._h), "fsr_frs_finish")
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def solver_header(triads, parts, module="fedem_solver"):
    """Text header of a primary-results file as the dynamics solver writes it for the recovery path
    (layout of fedem-foundation/src/FFrLib/FFrTests/response_0001/timehist_prim_0001/th_p_1.frs):
    triads = [(base_id, user_id, descr)], parts = [(base_id, user_id, descr, ngen)].
    Returns (text, payload_bytes_per_step)."""
    lines = [" InformationText         = response data base file;",
             f" Module                  = {module};",
             "VARIABLES:",
             '<1;"Time step number";NONE;INT;32;NUMBER>',
             '<2;"Physical time";TIME;FLOAT;64;SCALAR>',
             '<3;"Position matrix";NONE;FLOAT;64;TMAT34;(3,4);(("x","y","z"),("i1","i2","i3","position"))>']
    gen_ids = {}
    for p in parts:
        ng = p[3]
        if ng > 0 and ng not in gen_ids:
            gen_ids[ng] = 4 + len(gen_ids)
            lines.append(f'<{gen_ids[ng]};"Generalized displacement";LENGTH;FLOAT;64;VECTOR;({ng})>')
    lines += ["DATABLOCKS:", "<1><2>"]
    nbytes = 0
    for b, u, d in triads:
        lines.append(f'{{"Triad";{b};{u};"{d}";<3>}}')
        nbytes += 96
    for b, u, d, ng in parts:
        g = f"<{gen_ids[ng]}>" if ng > 0 else ""
        lines.append(f'{{"Part";{b};{u};"{d}";<3>{g}}}')
        nbytes += 96 + 8 * ng
    return "\n".join(lines) + "\n", nbytes
