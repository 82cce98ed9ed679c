"""ctypes access to the CPU checker under oracle/ (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it).

Oracle     -> oracle/liboracle.so        (C restatement of the reference's Fortran/C++ path)
Reference  -> oracle/_ref/libfedem_ref.so (the reference's own C++, compiled unmodified)"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
F64 = np.float64
I32 = np.int32
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)


def _dp(a):
    return a.ctypes.data_as(_D) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_I) if a is not None else None


class OrcSam(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nnod", "nel", "ndof", "ndof1", "ndof2", "neq", "nceq", "ngen")] + [
        ("madof", _I), ("mpmnpc", _I), ("mmnpc", _I), ("melcon", _I), ("meqn", _I), ("meqn1", _I),
        ("meqn2", _I), ("dofPosIn2", _I), ("mpmceq", _I), ("mmceq", _I), ("ttcc", _D)]


class OrcElm(C.Structure):
    _fields_ = [("xyz", _D), ("emod", _D), ("rny", _D), ("thk", _D), ("elmid", _I), ("beam", _D)]


class OrcRosette(C.Structure):
    _fields_ = [("id", C.c_int), ("numnod", C.c_int), ("ngage", C.c_int), ("zero_init", C.c_int),
                ("nodes", C.c_int * 4), ("rpos", C.c_double * 12), ("zpos", C.c_double), ("emod", C.c_double),
                ("nu", C.c_double), ("alpha_gages", C.c_double), ("gate", C.c_double), ("sncurve", C.c_double * 4)]


def ensure_built():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])
    return so


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(ensure_built())
        L = self.lib
        L.orc_von_mises.restype = C.c_double
        L.orc_max_shear_value.restype = C.c_double
        L.orc_sn_norsok.restype = C.c_double
        L.orc_damage.restype = C.c_double
        L.orc_sn_norsok.argtypes = [C.c_double] * 5
        L.orc_damage.argtypes = [_D, _D, C.c_int] + [C.c_double] * 4
        L.orc_pvx.argtypes = [_D, C.c_int, C.c_double, _D]
        L.orc_rainflow.argtypes = [_D, C.c_int, C.c_double, _D, _D]
        L.orc_cubic_solve.argtypes = [C.c_double] * 4 + [_D]
        L.orc_str24.argtypes = [_D, _D, _D, C.c_double, C.c_double, _D, _D, _D, _D, _D, _D]
        L.orc_str23.argtypes = [_D, _D, _D, C.c_double, C.c_double, _D, _D, _D, _D, _D, _D]
        L.orc_str41.argtypes = [_D, _D, _D, C.c_double, C.c_double, C.c_int, _D, _D, _D]
        L.orc_mat_times_vec.argtypes = [C.c_int, C.c_int, _D, _D, _D, C.c_int]
        L.orc_recover_history.argtypes = [C.POINTER(OrcSam), C.POINTER(OrcElm), _D, _D, _D, C.c_int, _I,
                                          _D, _D, _D, C.c_int]
        L.orc_calc_stresses.argtypes = [C.POINTER(OrcSam), C.POINTER(OrcElm), _D, _I, _D, _D, _D, _D, C.c_int]
        L.orc_calc_int_displacements.argtypes = [C.POINTER(OrcSam), _D, _D, _D, _D, _D, _D]
        L.orc_result_point_offsets.argtypes = [C.POINTER(OrcSam), _I, _I]
        L.orc_dof_pos_in2.argtypes = [C.c_int, C.c_int, _I, _I, _I, _I]

    # ---- model plumbing -------------------------------------------------------------------
    def bind_part(self, part):
        s, e = part.sam, part.elm
        keep = {}

        def ci(k, a):
            keep[k] = np.ascontiguousarray(a, I32)
            return _ip(keep[k])

        def cd(k, a):
            keep[k] = np.ascontiguousarray(a, F64)
            return _dp(keep[k])

        pos = np.zeros(max(s.ndof2, 1), I32)
        msc = np.ascontiguousarray(s.msc, I32); meqn = np.ascontiguousarray(s.meqn, I32)
        meqn2 = np.ascontiguousarray(s.meqn2 if s.ndof2 else np.zeros(1, I32), I32)
        self.lib.orc_dof_pos_in2(s.ndof, s.ndof2, _ip(msc), _ip(meqn), _ip(meqn2), _ip(pos))
        sam = OrcSam(nnod=s.nnod, nel=s.nel, ndof=s.ndof, ndof1=s.ndof1, ndof2=s.ndof2, neq=s.neq,
                     nceq=s.nceq, ngen=s.ngen)
        sam.madof = ci("madof", s.madof); sam.mpmnpc = ci("mpmnpc", s.mpmnpc); sam.mmnpc = ci("mmnpc", s.mmnpc)
        sam.melcon = ci("melcon", s.melcon); sam.meqn = ci("meqn", s.meqn)
        sam.meqn1 = ci("meqn1", s.meqn1 if s.ndof1 else np.zeros(1, I32))
        sam.meqn2 = ci("meqn2", meqn2); sam.dofPosIn2 = ci("pos", pos)
        sam.mpmceq = ci("mpmceq", s.mpmceq)
        sam.mmceq = ci("mmceq", s.mmceq if len(s.mmceq) else np.zeros(1, I32))
        sam.ttcc = cd("ttcc", s.ttcc if len(s.ttcc) else np.zeros(1))
        elm = OrcElm()
        elm.xyz = cd("xyz", e.xyz); elm.emod = cd("emod", e.emod); elm.rny = cd("rny", e.rny)
        elm.thk = cd("thk", e.thk)
        elm.elmid = ci("elmid", e.elmid) if e.elmid is not None else None
        elm.beam = cd("beam", e.beam) if e.beam is not None else None
        ptoff = np.zeros(s.nel + 1, I32)
        npts = self.lib.orc_result_point_offsets(C.byref(sam), elm.elmid, _ip(ptoff))
        keep["dofPosIn2"] = pos
        return dict(sam=sam, elm=elm, keep=keep, ptoff=ptoff, npts=npts, part=part)

    def expand(self, b, q):
        """calcIntDisplacements for one step: q = [finit; vg] -> sv[ndof]."""
        s = b["part"].sam
        B = np.asfortranarray(b["part"].B, F64) if b["part"].B is not None else np.zeros((1, 1), order="F")
        E = np.asfortranarray(b["part"].E, F64) if b["part"].E is not None else np.zeros((1, 1), order="F")
        q = np.ascontiguousarray(q, F64)
        work = np.zeros(s.neq + s.ndof1 + s.ndof2 + 1, F64)
        sv = np.zeros(s.ndof, F64)
        finit = q[:s.ndof2].copy(); vg = q[s.ndof2:].copy() if s.ngen else np.zeros(1)
        self.lib.orc_calc_int_displacements(C.byref(b["sam"]), _dp(B), _dp(E), _dp(finit), _dp(vg), _dp(work), _dp(sv))
        return sv

    def calc_stresses(self, b, sv, nthreads=1):
        """One step of the calcStresses element loop with every measure on."""
        s = b["part"].sam
        npts = max(b["npts"], 1)
        out = dict(resmat=np.zeros((npts, 8), F64), stress=np.zeros((npts, 6), F64),
                   strain=np.zeros((npts, 6), F64), sres=np.zeros((s.nel, 24), F64))
        sv = np.ascontiguousarray(sv, F64)
        out["nfail"] = self.lib.orc_calc_stresses(C.byref(b["sam"]), C.byref(b["elm"]), _dp(sv), _ip(b["ptoff"]),
                                                  _dp(out["resmat"]), _dp(out["stress"]), _dp(out["strain"]),
                                                  _dp(out["sres"]), nthreads)
        return out

    def recover_history(self, b, Q, want_history=True, nthreads=1):
        """The reference's time loop over Q [ndim, nsteps]: von Mises history + envelopes."""
        part = b["part"]
        Q = np.asfortranarray(Q, F64)
        nsteps = Q.shape[1]
        B = np.asfortranarray(part.B, F64) if part.B is not None and part.B.size else np.zeros((1, 1), order="F")
        E = np.asfortranarray(part.E, F64) if part.E is not None and part.E.size else np.zeros((1, 1), order="F")
        npts = b["npts"]
        vm = np.zeros((nsteps, npts), F64) if want_history else None
        mx = np.zeros(max(npts, 1), F64); mn = np.zeros(max(npts, 1), F64)
        rc = self.lib.orc_recover_history(C.byref(b["sam"]), C.byref(b["elm"]), _dp(B), _dp(E), _dp(Q), nsteps,
                                          _ip(b["ptoff"]), _dp(vm), _dp(mx), _dp(mn), nthreads)
        assert rc == 0
        return vm, mx[:npts], mn[:npts]

    # ---- strain rosettes (fedem_gage) -----------------------------------------------------
    def _ros(self, r):
        c = r.to_c()   # fsr_rosette and orc_rosette share their layout
        o = OrcRosette()
        C.memmove(C.byref(o), C.byref(c), C.sizeof(o))
        return o

    def rosette_bcart(self, b, ros):
        """InitStrainRosette: Bcart [3, ndim] from ElDispFromSupElDisp and the rosette geometry."""
        part = b["part"]; s = part.sam
        B = np.asfortranarray(part.B, F64) if part.B is not None and part.B.size else np.zeros((1, 1), order="F")
        E = np.asfortranarray(part.E, F64) if part.E is not None and part.E.size else np.zeros((1, 1), order="F")
        out = np.zeros((s.ndim, 3), F64)
        o = self._ros(ros)
        self.lib.orc_rosette_bcart.argtypes = [C.POINTER(OrcRosette), C.POINTER(OrcSam), _D, _D, _D, _D]
        rc = self.lib.orc_rosette_bcart(C.byref(o), C.byref(b["sam"]), b["elm"].xyz, _dp(B), _dp(E), _dp(out))
        assert rc == 0, rc
        return np.ascontiguousarray(out.T)

    def rosette_history(self, b, ros, Q, Bcart=None):
        """gage.f90:293-355 for one rosette: values [nsteps, 24] (zero-start strains honoured)."""
        Bc = self.rosette_bcart(b, ros) if Bcart is None else Bcart
        Bf = np.asfortranarray(Bc)
        o = self._ros(ros)
        Tg = np.zeros(9, F64)
        self.lib.orc_gage_directions.argtypes = [C.POINTER(OrcRosette), _D]
        self.lib.orc_gage_directions(C.byref(o), _dp(Tg))
        self.lib.orc_calc_rosette_strains.argtypes = [_D, C.c_int, _D, _D, C.c_double, C.c_double, _D, _D, C.c_int, _D]
        Q = np.asfortranarray(Q, F64)
        ns = Q.shape[1]
        out = np.zeros((ns, 24), F64)
        eps0 = np.zeros(3, F64); sig0 = np.zeros(3, F64)
        if ros.zero_init:
            eps0 = -(Bc @ Q[:, 0])          # calcZeroStartRosetteStrains
        for t in range(ns):
            q = np.ascontiguousarray(Q[:, t])
            self.lib.orc_calc_rosette_strains(_dp(Bf), Q.shape[0], _dp(q), _dp(eps0), ros.emod, ros.nu, _dp(sig0),
                                              _dp(Tg), o.ngage, _dp(out[t]))
        return out

    def series_fatigue(self, x, gate, curve, bin_size=0.0, nbins=0):
        """ffp_getdamage + ffp_getnumcycles on one history -> (damage, ncycles, bins, ok)."""
        tp = self.pvx(x, gate)
        cyc = self.rainflow(tp, gate)
        if cyc is None:
            return 0.0, 0, np.full(nbins, -1, I32), False
        d = self.damage(cyc, curve) if len(cyc) else 0.0
        r = np.abs(cyc[:, 0] - cyc[:, 1]) if len(cyc) else np.zeros(0)
        bins = np.zeros(nbins, I32)
        lo = 0.0
        for k in range(nbins):
            hi = lo + bin_size
            bins[k] = -1 if (len(r) == 0 or lo > r.max()) else int(((r >= lo) & (r < hi)).sum())
            lo = hi
        return d, len(cyc), bins, True

    # ---- invariants / fatigue -------------------------------------------------------------
    def von_mises(self, S):
        S = np.ascontiguousarray(S, F64)
        return self.lib.orc_von_mises({1: 1, 3: 2, 6: 3}[len(S)], _dp(S))

    def principal(self, S):
        S = np.ascontiguousarray(S, F64)
        n = {1: 1, 3: 2, 6: 3}[len(S)]
        P = np.zeros(3, F64)
        ok = self.lib.orc_principal_values(n, _dp(S), _dp(P))
        return ok, P[:n]

    def pvx(self, data, gate):
        data = np.ascontiguousarray(data, F64)
        turns = np.zeros(len(data) + 2, F64)
        n = self.lib.orc_pvx(_dp(data), len(data), gate, _dp(turns))
        return turns[:n].copy()

    def rainflow(self, turns, gate):
        turns = np.ascontiguousarray(turns, F64)
        cf = np.zeros(len(turns) + 4, F64); cs = np.zeros(len(turns) + 4, F64)
        n = self.lib.orc_rainflow(_dp(turns), len(turns), gate, _dp(cf), _dp(cs))
        if n < 0:
            return None
        return np.stack([cf[:n], cs[:n]], 1)

    def damage(self, cycles, curve):
        cf = np.ascontiguousarray(cycles[:, 0]); cs = np.ascontiguousarray(cycles[:, 1])
        return self.lib.orc_damage(_dp(cf), _dp(cs), len(cf), *[float(c) for c in curve])


class Reference:
    """The reference's own compiled C++ (oracle/_ref/libfedem_ref.so, built by oracle/Makefile from
    the sources where they lie under /root/reference; prebuilt file travels to the GPU box)."""

    def __init__(self):
        so = os.path.join(ORACLE_DIR, "_ref", "libfedem_ref.so")
        if not os.path.exists(so) and os.path.isdir("/root/reference"):
            subprocess.call(["make", "-s", "-C", ORACLE_DIR, "ref"])
        self.available = os.path.exists(so)
        if not self.available:
            return
        self.lib = C.CDLL(so)
        L = self.lib
        L.ref_von_mises.restype = C.c_double
        L.ref_get_damage.restype = C.c_double
        L.ref_sn_norsok.restype = C.c_double
        L.ref_sn_norsok.argtypes = [C.c_double] * 5
        L.ref_get_damage.argtypes = [_D, C.c_int, C.c_double, _D]
        L.ref_get_num_cycles.argtypes = [C.c_double, C.c_double]
        L.ref_pvx.argtypes = [_D, C.c_int, C.c_double, _D]
        L.ref_rainflow.argtypes = [_D, C.c_int, C.c_double, _D, _D]
        L.ref_cubic_solve.argtypes = [C.c_double] * 4 + [_D]

    def von_mises(self, S):
        S = np.ascontiguousarray(S, F64)
        return self.lib.ref_von_mises({1: 1, 3: 2, 6: 3}[len(S)], _dp(S))

    def principal(self, S):
        S = np.ascontiguousarray(S, F64)
        n = {1: 1, 3: 2, 6: 3}[len(S)]
        P = np.zeros(3, F64)
        ok = self.lib.ref_principal_values(n, _dp(S), _dp(P))
        return ok, P[:n]

    def pvx(self, data, gate):
        data = np.ascontiguousarray(data, F64)
        turns = np.zeros(len(data) + 2, F64)
        n = self.lib.ref_pvx(_dp(data), len(data), gate, _dp(turns))
        return turns[:n].copy()

    def rainflow(self, turns, gate):
        turns = np.ascontiguousarray(turns, F64)
        cf = np.zeros(len(turns) + 4, F64); cs = np.zeros(len(turns) + 4, F64)
        n = self.lib.ref_rainflow(_dp(turns), len(turns), gate, _dp(cf), _dp(cs))
        if n < 0:
            return None
        return np.stack([cf[:n], cs[:n]], 1)

    def get_damage(self, data, gate, curve):
        data = np.ascontiguousarray(data, F64)
        curve = np.ascontiguousarray(curve, F64)
        d = self.lib.ref_get_damage(_dp(data), len(data), gate, _dp(curve))
        return d, self.lib.ref_num_cycles_total()

    def num_cycles(self, low, high):
        return self.lib.ref_get_num_cycles(low, high)
