"""Host-side mirror of the reference's recovery driver interface, on top of the C ABI.

``StressRecovery`` plays the role of one FE part inside ``fedem_stress``
(src/vpmStress/stress.f90:112-435) or of one entry of the solver's recovery list
(src/vpmSolver/stressRecoveryModule.f90:517-768,991-1225): it is created from the SAM data and
the element data (initiateSAM + ffl_*), receives the B and E matrices (openBandEmatrices) and is
then driven with a window of the reduced history.  Method names follow the reference routines
they replace; argument meaning and error behaviour (negative = fatal, positive = number of
failed elements which get hugeVal results) follow ``ierr`` of the Fortran."""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import FsrSam, FsrElmData, FsrOptions, check
from .model import PartModel

F64 = np.float64
I32 = np.int32


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def c_part_structs(part: PartModel, keep):
    """fsr_sam / fsr_elmdata views of a PartModel; `keep` collects the arrays the pointers refer to."""
    s, e = part.sam, part.elm

    def ci(a):
        a = np.ascontiguousarray(a, I32)
        keep.append(a)
        return a

    def cd(a):
        a = np.ascontiguousarray(a, F64)
        keep.append(a)
        return a

    sam = FsrSam(nnod=s.nnod, nel=s.nel, ndof=s.ndof, ndof1=s.ndof1, ndof2=s.ndof2, ngen=s.ngen,
                 neq=s.neq, nceq=s.nceq, nmmnpc=len(s.mmnpc), nmmceq=len(s.mmceq))
    sam.madof = _ip(ci(s.madof)); sam.msc = _ip(ci(s.msc)); sam.mpmnpc = _ip(ci(s.mpmnpc))
    sam.mmnpc = _ip(ci(s.mmnpc)); sam.melcon = _ip(ci(s.melcon)); sam.mpmceq = _ip(ci(s.mpmceq))
    sam.mmceq = _ip(ci(s.mmceq if len(s.mmceq) else np.zeros(1, I32)))
    sam.ttcc = _dp(cd(s.ttcc if len(s.ttcc) else np.zeros(1)))
    sam.meqn = _ip(ci(s.meqn)); sam.meqn1 = _ip(ci(s.meqn1 if s.ndof1 else np.zeros(1, I32)))
    sam.meqn2 = _ip(ci(s.meqn2 if s.ndof2 else np.zeros(1, I32)))
    elm = FsrElmData()
    elm.xyz = _dp(cd(e.xyz)); elm.emod = _dp(cd(e.emod)); elm.rny = _dp(cd(e.rny)); elm.thk = _dp(cd(e.thk))
    elm.elmid = _ip(ci(e.elmid)) if e.elmid is not None else None
    elm.beam = _dp(cd(e.beam)) if e.beam is not None else None
    return sam, elm


def c_options(device=0, stress_form=0, step_tile=0, elem_order=0, ffq_stress_form=2, fft_stress_form=1):
    opt = FsrOptions(device=device, stressForm=stress_form, step_tile=step_tile)
    opt.reserved[0] = elem_order  # 0 = Morton order of element centroids, 1 = SAM order
    opt.reserved[1] = ffq_stress_form + 1   # -ffqStressForm / -fftStressForm of the legacy shells (types 22 / 21)
    opt.reserved[2] = fft_stress_form + 1
    return opt


def split_elements(part: PartModel, nblocks):
    """fsr_split_elements: contiguous element ranges [(e0, e1), ...] of equal cost (native; host only)."""
    lib = _lib.load_library()
    keep = []
    sam, elm = c_part_structs(part, keep)
    cut = np.zeros(nblocks + 1, I32)
    check(lib.fsr_split_elements(C.byref(sam), C.byref(elm), nblocks, _ip(cut)), "fsr_split_elements")
    return [(int(cut[b]), int(cut[b + 1])) for b in range(nblocks)]


class StressRecovery:
    """One FE part -- or, with `block=(e0, e1)`, the element block [e0, e1) of it (fsr_part_create_block: own nodes, all
    external DOFs, the B / E rows of its nodes; results bit-identical to the parent's for its elements)."""

    def __init__(self, part: PartModel, device=0, stress_form=0, step_tile=0, elem_order=0, ffq_stress_form=2, fft_stress_form=1,
                 block=None):
        self._lib = _lib.load_library()
        self._h = C.c_void_p()
        s = part.sam
        keep = []
        sam, elm = c_part_structs(part, keep)
        opt = c_options(device, stress_form, step_tile, elem_order, ffq_stress_form, fft_stress_form)
        if block is None:
            rc = self._lib.fsr_part_create(C.byref(self._h), C.byref(sam), C.byref(elm), C.byref(opt))
            self.n_failed = check(rc, "fsr_part_create")
        else:
            rc = self._lib.fsr_part_create_block(C.byref(self._h), C.byref(sam), C.byref(elm), C.byref(opt), int(block[0]), int(block[1]))
            self.n_failed = check(rc, "fsr_part_create_block")
        del keep
        self.ndim = self._lib.fsr_ndim(self._h)
        self.npts = self._lib.fsr_num_result_points(self._h)
        info = np.zeros(10, I32)
        check(self._lib.fsr_block_info(self._h, _ip(info)), "fsr_block_info")
        self.e0, self.e1, self.pt0 = int(info[0]), int(info[1]), int(info[2])
        self.nnod, self.ndof1, self.parent_ndof1, self.parent_npts = int(info[4]), int(info[5]), int(info[6]), int(info[7])
        self.is_block = block is not None
        self.ndof = int(info[8])
        self.nel = self.e1 - self.e0
        if part.B is not None or part.E is not None:
            self.open_B_and_E_matrices(part.B, part.E)

    def block_rows(self):
        """(rows of the parent's B / E kept by the block, 0-based; parent node numbers of its nodes, 1-based)"""
        rows, nodes = np.zeros(max(self.ndof1, 1), I32), np.zeros(max(self.nnod, 1), I32)
        check(self._lib.fsr_block_rows(self._h, _ip(rows), _ip(nodes)), "fsr_block_rows")
        return rows[:self.ndof1], nodes[:self.nnod]

    # ---- openBandEmatrices (displacementModule.f90:645-790) --------------------------------
    def open_B_and_E_matrices(self, B, E):
        B = np.asfortranarray(B, F64) if B is not None and B.size else None
        E = np.asfortranarray(E, F64) if E is not None and E.size else None
        ldB = B.shape[0] if B is not None else 0
        ldE = E.shape[0] if E is not None else 0
        if self.is_block and max(ldB, ldE) == self.parent_ndof1 and self.parent_ndof1 != self.ndof1:
            check(self._lib.fsr_set_recovery_parent(self._h, _dp(B), ldB, _dp(E), ldE), "fsr_set_recovery_parent")   # the parent's matrices
        else:
            check(self._lib.fsr_set_recovery(self._h, _dp(B), ldB, _dp(E), ldE), "fsr_set_recovery")

    # ---- stress.f90:361-435 time loop, batched ----------------------------------------------
    def recover(self, Q, want_history=True):
        """Q: [ndim, nsteps] (column = [finit; vg] of a step).  Returns the von Mises history
        [nsteps, npts] (row s = resMat(1,:) of step s, stressRoutines.f90:273-276) or None."""
        Q = np.asfortranarray(Q, F64)
        assert Q.shape[0] == self.ndim, (Q.shape, self.ndim)
        nsteps = Q.shape[1]
        vm = np.empty((nsteps, self.npts), F64) if want_history else None
        check(self._lib.fsr_recover(self._h, _dp(Q), Q.shape[0], nsteps, _dp(vm)), "fsr_recover")
        return vm

    def recover_displacements(self, sv, want_history=True):
        """calcStresses on nodal displacements that are already there (stress.f90:397 readIntDisplacements, the direct
        solution on the results files): sv [nsteps, ndof].  Returns the von Mises history like recover."""
        sv = np.ascontiguousarray(sv, F64)
        assert sv.ndim == 2 and sv.shape[1] == self.ndof, (sv.shape, self.ndof)
        vm = np.empty((sv.shape[0], self.npts), F64) if want_history else None
        check(self._lib.fsr_recover_displacements(self._h, _dp(sv), sv.shape[0], _dp(vm)), "fsr_recover_displacements")
        return vm

    def recover_dev(self, q_ptr, ldq, nsteps, vm_ptr=None, ld_vm=0, stream=None):
        """Device-pointer variant (torch tensors: pass .data_ptr()); asynchronous."""
        check(self._lib.fsr_recover_dev(self._h, C.c_void_p(q_ptr), ldq, nsteps,
                                        C.c_void_p(vm_ptr) if vm_ptr else None, ld_vm,
                                        C.c_void_p(stream) if stream else None), "fsr_recover_dev")

    def recover_async(self, Q):
        """Queues a window (Q [ndim, nsteps], Fortran order, ideally page-locked) and returns; see envelope_async / synchronize."""
        assert Q.flags.f_contiguous and Q.dtype == F64 and Q.shape[0] == self.ndim
        check(self._lib.fsr_recover_async(self._h, _dp(Q), Q.shape[0], Q.shape[1]), "fsr_recover_async")

    def envelope_async(self, out_max, out_min):
        """Envelopes after the work queued so far -> out_max / out_min (page-locked arrays), copied while later windows compute."""
        check(self._lib.fsr_get_envelope_async(self._h, _dp(out_max), _dp(out_min)), "fsr_get_envelope_async")

    def envelope_wait(self):
        check(self._lib.fsr_envelope_wait(self._h), "fsr_envelope_wait")

    def synchronize(self):
        check(self._lib.fsr_synchronize(self._h), "fsr_synchronize")

    def reset_envelope(self):
        check(self._lib.fsr_reset_envelope(self._h), "fsr_reset_envelope")

    def set_stream(self, stream_ptr):
        """Run all later calls on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        check(self._lib.fsr_set_stream(self._h, C.c_void_p(stream_ptr) if stream_ptr else None), "fsr_set_stream")

    def envelope(self, out_max=None, out_min=None):
        """Running (max, min) of von Mises per result point (strainCoatModule.f90:159-166,410-420)."""
        mx = out_max if out_max is not None else np.empty(self.npts, F64)
        mn = out_min if out_min is not None else np.empty(self.npts, F64)
        check(self._lib.fsr_get_envelope(self._h, _dp(mx), _dp(mn)), "fsr_get_envelope")
        return mx, mn

    def copy_envelope_dev(self, max_ptr, min_ptr, stream=None):
        """Envelopes into caller-owned device buffers (torch tensors: .data_ptr()), asynchronous."""
        check(self._lib.fsr_copy_envelope_dev(self._h, C.c_void_p(max_ptr) if max_ptr else None,
                                              C.c_void_p(min_ptr) if min_ptr else None,
                                              C.c_void_p(stream) if stream else None), "fsr_copy_envelope_dev")

    def envelope_dev_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        check(self._lib.fsr_envelope_dev(self._h, C.byref(a), C.byref(b)), "fsr_envelope_dev")
        return a.value, b.value

    # ---- calcIntDisplacements for a batch ------------------------------------------------------
    def calc_int_displacements(self, Q):
        Q = np.asfortranarray(Q, F64)
        nsteps = Q.shape[1]
        U = np.empty((nsteps, self.ndof), F64)
        check(self._lib.fsr_expand(self._h, _dp(Q), Q.shape[0], nsteps, _dp(U)), "fsr_expand")
        return U

    # ---- calcStresses with every output switch on, one step -------------------------------------
    def calc_stresses(self, q):
        q = np.ascontiguousarray(q, F64)
        res = dict(resmat=np.zeros((self.npts, 8), F64), stress=np.zeros((self.npts, 6), F64),
                   strain=np.zeros((self.npts, 6), F64), sres=np.zeros((self.nel, 24), F64),
                   sv=np.zeros(self.ndof, F64))
        rc = self._lib.fsr_recover_step_full(self._h, _dp(q), _dp(res["resmat"]), _dp(res["stress"]),
                                             _dp(res["strain"]), _dp(res["sres"]), _dp(res["sv"]))
        check(rc, "fsr_recover_step_full")
        return res

    def result_point_offsets(self):
        off = np.zeros(self.nel + 1, I32)
        check(self._lib.fsr_result_point_offsets(self._h, _ip(off)), "fsr_result_point_offsets")
        return off

    def last_timing(self):
        t = np.zeros(3, F64)
        self._lib.fsr_last_timing(self._h, _dp(t), 3)
        return dict(k1_ms=t[0], k2_ms=t[1], tiles=int(t[2]))

    def family_counts(self):
        """{family: (elements, on the geometry fast path, on the general kernel)}"""
        names = ["quad", "tri", "tet10", "beam", "hex20", "hex8", "tet4", "wedg6", "wedg15", "tri6", "quad8"]
        c = np.zeros(3 * len(names), I32)
        check(self._lib.fsr_family_counts(self._h, _ip(c), len(c)), "fsr_family_counts")
        return {n: tuple(int(v) for v in c[3 * i:3 * i + 3]) for i, n in enumerate(names) if c[3 * i] > 0}

    def vm_path_info(self):
        """Rows the von Mises path expands per step tile and the split of the quadrilaterals (fsr_vm_path_info)."""
        v = (C.c_longlong * 7)()
        check(self._lib.fsr_vm_path_info(self._h, v, 7), "fsr_vm_path_info")
        keys = ["k1_rows", "ndof", "inplane_rows", "global_row_tiles", "quads_inplane", "quads_flat", "quads_dense"]
        return {k: int(x) for k, x in zip(keys, v)}

    def timing_reset(self):
        self._lib.fsr_timing_reset(self._h)

    def close(self):
        if self._h:
            self._lib.fsr_part_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupRecovery:
    """One part on several GPUs of this process (fsr_group_*): element blocks, Q broadcast and envelope gather with NCCL
    inside the library.  devices: list of CUDA ordinals, or None for all visible ones."""

    def __init__(self, part: PartModel, devices=None, stress_form=0, step_tile=0, elem_order=0, ffq_stress_form=2, fft_stress_form=1):
        self._lib = _lib.load_library()
        self._h = C.c_void_p()
        keep = []
        sam, elm = c_part_structs(part, keep)
        opt = c_options(0, stress_form, step_tile, elem_order, ffq_stress_form, fft_stress_form)
        dv = np.ascontiguousarray(devices, I32) if devices is not None else None
        rc = self._lib.fsr_group_create(C.byref(self._h), C.byref(sam), C.byref(elm), C.byref(opt), _ip(dv), len(dv) if dv is not None else 0)
        self.n_failed = check(rc, "fsr_group_create")
        self.nblocks = self._lib.fsr_group_num_blocks(self._h)
        self.npts = self._lib.fsr_group_num_result_points(self._h)
        self.ndim = self._lib.fsr_group_ndim(self._h)
        if part.B is not None or part.E is not None:
            B = np.asfortranarray(part.B, F64) if part.B is not None and part.B.size else None
            E = np.asfortranarray(part.E, F64) if part.E is not None and part.E.size else None
            check(self._lib.fsr_group_set_recovery(self._h, _dp(B), B.shape[0] if B is not None else 0, _dp(E),
                                                   E.shape[0] if E is not None else 0), "fsr_group_set_recovery")

    def recover(self, Q, want_history=True):
        Q = np.asfortranarray(Q, F64)
        assert Q.shape[0] == self.ndim
        vm = np.empty((Q.shape[1], self.npts), F64) if want_history else None
        check(self._lib.fsr_group_recover(self._h, _dp(Q), Q.shape[0], Q.shape[1], _dp(vm)), "fsr_group_recover")
        return vm

    def synchronize(self):
        check(self._lib.fsr_group_synchronize(self._h), "fsr_group_synchronize")

    def reset_envelope(self):
        check(self._lib.fsr_group_reset_envelope(self._h), "fsr_group_reset_envelope")

    def envelope(self, out_max=None, out_min=None):
        mx = out_max if out_max is not None else np.empty(self.npts, F64)
        mn = out_min if out_min is not None else np.empty(self.npts, F64)
        check(self._lib.fsr_group_get_envelope(self._h, _dp(mx), _dp(mn)), "fsr_group_get_envelope")
        return mx, mn

    def last_timing(self):
        t = np.zeros(4, F64)
        self._lib.fsr_group_last_timing(self._h, _dp(t), 4)
        return dict(k1_ms=t[0], k2_ms=t[1], tiles=int(t[2]), balance=t[3])

    def timing_reset(self):
        self._lib.fsr_group_timing_reset(self._h)

    def close(self):
        if self._h:
            self._lib.fsr_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """One process per GPU: the library's own NCCL communicator (fsr_comm_*).  `exchange(id_bytes_or_None) -> id_bytes` is the
    host's way of passing rank 0's 128-byte id to the other ranks (MPI_Bcast, torch.distributed.broadcast_object_list, ...)."""

    def __init__(self, rank, world, device, exchange):
        self._lib = _lib.load_library()
        self._h = C.c_void_p()
        self.rank, self.world = rank, world
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(self._lib.fsr_comm_unique_id(buf, 128), "fsr_comm_unique_id")
        ident = exchange(bytes(buf.raw) if rank == 0 else None)
        check(self._lib.fsr_comm_init_rank(C.byref(self._h), ident, rank, world, device), "fsr_comm_init_rank")

    def broadcast(self, buf_ptr, count, root=0, stream=None):
        check(self._lib.fsr_comm_broadcast(self._h, C.c_void_p(buf_ptr), int(count), root, C.c_void_p(stream) if stream else None),
              "fsr_comm_broadcast")

    def gather_envelope(self, rec, pt0, npts, max_ptr=None, min_ptr=None, root=0, stream=None):
        pt0 = np.ascontiguousarray(pt0, I32); npts = np.ascontiguousarray(npts, I32)
        check(self._lib.fsr_comm_gather_envelope(self._h, rec._h, _ip(pt0), _ip(npts), C.c_void_p(max_ptr) if max_ptr else None,
                                                 C.c_void_p(min_ptr) if min_ptr else None, root, C.c_void_p(stream) if stream else None),
              "fsr_comm_gather_envelope")

    def close(self):
        if self._h:
            self._lib.fsr_comm_destroy(self._h)
            self._h = C.c_void_p()


def fatigue(hist, gate, curve, bin_size=0.0, nbins=0, device=0):
    """ffp_getdamage / ffp_getnumcycles for many gages at once
    (fedem-foundation/src/FFpLib/FFpFatigue/FFpFatigue_F.C:81-141).
    hist: [ngage, nsteps].  Returns (damage[ngage], ncycles[ngage], bins[ngage, nbins] or None)."""
    lib = _lib.load_library()
    hist = np.ascontiguousarray(hist, F64)
    ng, ns = hist.shape
    curve = np.ascontiguousarray(curve, F64)
    damage = np.zeros(ng, F64); ncyc = np.zeros(ng, I32)
    bins = np.zeros((ng, nbins), I32) if nbins > 0 else None
    check(lib.fsr_fatigue(device, _dp(hist), ng, ns, float(gate), _dp(curve), float(bin_size), nbins,
                          _dp(damage), _ip(ncyc), _ip(bins)), "fsr_fatigue")
    return damage, ncyc, bins


class FatigueCounter:
    """Streaming PVX + rainflow + damage for ngage histories resident on one GPU: the batched,
    device-side form of ffp_addpoint ... ffp_getdamage (FFpFatigue_F.C:37-124).  Tiles of time steps
    are device arrays (torch tensors: pass .data_ptr()); see include/fedem_b200.h for the locate /
    feed protocol."""
    GAGE_MAJOR, STEP_MAJOR = 0, 1

    def __init__(self, ngage, gate, curve, bin_size=0.0, nbins=0, stack_cap=0, device=0):
        self._lib = _lib.load_library()
        self._h = C.c_void_p()
        self.ngage, self.nbins = ngage, nbins
        curve = np.ascontiguousarray(curve, F64)
        check(self._lib.fsr_fatigue_create(C.byref(self._h), device, ngage, float(gate), _dp(curve),
                                           float(bin_size), nbins, stack_cap), "fsr_fatigue_create")

    def set_gage_params(self, gate=None, curve=None):
        g = np.ascontiguousarray(gate, F64) if gate is not None else None
        c = np.ascontiguousarray(curve, F64) if curve is not None else None
        check(self._lib.fsr_fatigue_set_gage_params(self._h, _dp(g), _dp(c)), "fsr_fatigue_set_gage_params")

    def reset(self):
        check(self._lib.fsr_fatigue_reset(self._h), "fsr_fatigue_reset")

    def locate(self, hist_ptr, ld, layout, step0, nsteps, stream=None, want_pending=True):
        n = C.c_int(-1)
        check(self._lib.fsr_fatigue_locate_dev(self._h, C.c_void_p(hist_ptr), ld, layout, step0, nsteps,
                                               C.byref(n) if want_pending else None,
                                               C.c_void_p(stream) if stream else None), "fsr_fatigue_locate_dev")
        return n.value

    def feed(self, hist_ptr, ld, layout, step0, nsteps, stream=None):
        check(self._lib.fsr_fatigue_feed_dev(self._h, C.c_void_p(hist_ptr), ld, layout, step0, nsteps,
                                             C.c_void_p(stream) if stream else None), "fsr_fatigue_feed_dev")

    def finish(self):
        """Returns dict(damage, ncycles, bins, status) as host arrays."""
        damage = np.zeros(self.ngage, F64); ncyc = np.zeros(self.ngage, I32); status = np.zeros(self.ngage, I32)
        bins = np.zeros((self.ngage, self.nbins), I32) if self.nbins > 0 else None
        nwarn = check(self._lib.fsr_fatigue_finish(self._h, _dp(damage), _ip(ncyc), _ip(bins), _ip(status)),
                      "fsr_fatigue_finish")
        return dict(damage=damage, ncycles=ncyc, bins=bins, status=status, nwarn=nwarn)

    def close(self):
        if self._h:
            self._lib.fsr_fatigue_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
