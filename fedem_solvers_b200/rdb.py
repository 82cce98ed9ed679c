"""Stress results database (.frs) writer: host-side mirror of writeStressHeader + the writeStressDB /
writeStrMeasureDB / writeDisplacementDB calls of the fedem_stress time loop
(src/vpmStress/saveStressModule.f90:120-247,1437-1633; stress.f90:292-432) on top of csrc/io_rdb.cu, which
forms the records of a whole window of time steps on the GPU."""
import ctypes as C
import os
import numpy as np

from . import _lib
from ._lib import FsrRdbOptions, check

F64 = np.float64
I32 = np.int32

# the -vmStress ... switches (include/fedem_b200.h FSR_OUT_*); order = resMat rows of stressRoutines.f90:273-287
OUT = dict(vmStress=0x001, maxPStress=0x002, minPStress=0x004, maxSStress=0x008, vmStrain=0x010, maxPStrain=0x020,
           minPStrain=0x040, maxSStrain=0x080, stress=0x100, strain=0x200, SR=0x400, deformation=0x800)


def out_mask(**flags):
    m = 0
    for k, v in flags.items():
        if v:
            m |= OUT[k]
    return m


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _options(mask, double, rdbinc, base_id, user_id, descr, model_file, link_file, elmid, minex, sup_tr_init, keep):
    o = FsrRdbOptions(out_mask=mask, double_precision=int(double), rdbinc=int(rdbinc), part_base_id=int(base_id),
                      part_user_id=int(user_id))
    o.part_descr = descr.encode() if descr else None
    o.model_file = model_file.encode() if model_file else None
    o.link_file = link_file.encode() if link_file else None
    if elmid is not None:
        keep.append(np.ascontiguousarray(elmid, I32)); o.elmid = _ip(keep[-1])
    if minex is not None:
        keep.append(np.ascontiguousarray(minex, I32)); o.minex = _ip(keep[-1])
    if sup_tr_init is not None:   # [3, 4] -> column-major 12 doubles
        keep.append(np.ascontiguousarray(np.asarray(sup_tr_init, F64).T.ravel())); o.sup_tr_init = _dp(keep[-1])
    return o


def build_header(madof, melcon, mask, double=False, base_id=1, user_id=1, descr="", model_file=None, link_file=None,
                 elmid=None, minex=None, sup_tr_init=None):
    """Header text and bytes per time step of the file fedem_stress would write for this part (host only)."""
    lib = _lib.load_library()
    keep = []
    o = _options(mask, double, 0, base_id, user_id, descr, model_file, link_file, elmid, minex, sup_tr_init, keep)
    madof = np.ascontiguousarray(madof, I32)
    melcon = np.ascontiguousarray(melcon, I32)
    sb = C.c_longlong()
    n = check(lib.fsr_rdb_build_header(len(madof) - 1, _ip(madof), len(melcon), _ip(melcon), C.byref(o), None, 0, C.byref(sb)),
              "fsr_rdb_build_header")
    buf = C.create_string_buffer(n + 1)
    check(lib.fsr_rdb_build_header(len(madof) - 1, _ip(madof), len(melcon), _ip(melcon), C.byref(o), buf, n + 1, C.byref(sb)),
          "fsr_rdb_build_header")
    return buf.value.decode("latin1"), sb.value


class StressRdb:
    """writeStressHeader on creation, writeTimeStepDB + the per-element writes per step in write_steps."""

    def __init__(self, recovery, path, mask, double=False, rdbinc=0, base_id=1, user_id=1, descr="", model_file=None,
                 link_file=None, elmid=None, minex=None, sup_tr_init=None):
        self.lib = _lib.load_library()
        self._h = C.c_void_p()
        self._keep = []
        o = _options(mask, double, rdbinc, base_id, user_id, descr, model_file, link_file, elmid, minex, sup_tr_init,
                     self._keep)
        if hasattr(recovery, "nblocks"):   # GroupRecovery: element blocks on several GPUs fill their slots of every record
            check(self.lib.fsr_rdb_create_group(C.byref(self._h), recovery._h, os.fsencode(path), C.byref(o)), "fsr_rdb_create_group")
        else:
            check(self.lib.fsr_rdb_create(C.byref(self._h), recovery._h, os.fsencode(path), C.byref(o)), "fsr_rdb_create")
        self.step_bytes = self.lib.fsr_rdb_step_bytes(self._h)
        buf = C.create_string_buffer(4096)
        self.lib.fsr_rdb_path(self._h, buf, 4096)
        self.path = buf.value.decode()
        self.ndim = recovery.ndim

    def header(self):
        n = self.lib.fsr_rdb_header(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.lib.fsr_rdb_header(self._h, buf, n + 1)
        return buf.value.decode("latin1")

    def write_steps(self, Q, stepno, time, sup_tr=None):
        """Q [ndim, nsteps]; stepno / time [nsteps]; sup_tr [nsteps, 3, 4] when total displacements are written."""
        Q = np.asfortranarray(Q, F64)
        assert Q.shape[0] == self.ndim
        stepno = np.ascontiguousarray(stepno, I32)
        time = np.ascontiguousarray(time, F64)
        st = None
        if sup_tr is not None:
            st = np.ascontiguousarray(np.swapaxes(np.asarray(sup_tr, F64), -1, -2))
        check(self.lib.fsr_rdb_write_steps(self._h, _dp(Q), Q.shape[0], Q.shape[1], _ip(stepno), _dp(time), _dp(st)),
              "fsr_rdb_write_steps")

    def flush(self):
        """Waits for the pipeline (device -> PCIe -> file) to drain; returns where the time went."""
        t = np.zeros(6, F64)
        check(self.lib.fsr_rdb_flush(self._h, _dp(t), 6), "fsr_rdb_flush")
        return dict(compute_ms=t[0], d2h_ms=t[1], disk_ms=t[2], bytes=int(t[3]), tiles=int(t[4]), k1_ms=t[5])

    def close(self):
        if self._h:
            check(self.lib.fsr_rdb_close(self._h), "fsr_rdb_close")
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
