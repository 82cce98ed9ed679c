"""Host-side mirror of the reference's strain-gage driver (fedem_gage, src/vpmStress/gage.f90) on
top of the C ABI: rosette records as readStrainGageData delivers them
(strainGageModule.f90:107-237), Bcart construction on the GPU (InitStrainRosette), per-step rosette
results (calcRosetteStrains) and rainflow / damage of the max principal stress and every gage leg
(AddFatiguePoints, reportDamage).  Nothing here computes: all arithmetic is in libfedem_b200.so."""
import ctypes as C
from dataclasses import dataclass, field
import numpy as np

from . import _lib
from ._lib import FsrRosette, check

F64 = np.float64
I32 = np.int32

# rosette types of strainGageModule.f90:19-22,208-226: name -> (number of legs, angle between legs)
ROSETTE_TYPES = {"SINGLE_GAGE": (1, 0.0), "DOUBLE_GAGE_90": (2, np.pi / 2.0),
                 "TRIPLE_GAGE_60": (3, np.pi / 3.0), "TRIPLE_GAGE_45": (3, np.pi / 4.0)}
NVAL = 24
VALUE_NAMES = (["epsC_x", "epsC_y", "gammaC_xy", "epsP_max", "epsP_min", "epsP_sam", "gammaMax", "epsVM",
                "alpha1", "alphaGamma", "sigmaC_x", "sigmaC_y", "tauC_xy", "sigmaP_max", "sigmaP_min",
                "sigmaP_sam", "tauMax", "sigmaVM"] + [f"epsGage{i}" for i in (1, 2, 3)] +
               [f"sigGage{i}" for i in (1, 2, 3)])


@dataclass
class Rosette:
    """One &STRAIN_ROSETTE namelist record (solverTests/TimeDomain/SubModelling/globalmodel.fsi:123-153)."""
    id: int
    nodes: list                    # internal node numbers (1-based), 3 or 4
    rpos: np.ndarray               # posInGl [3, 4]: columns = rosette X, Y, Z axis, position
    type: str = "TRIPLE_GAGE_45"
    zpos: float = 0.0
    emod: float = 2.1e11
    nu: float = 0.3
    zero_init: bool = False
    gate: float = 0.0              # <= 0: run default (-gate)
    sncurve: list = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])

    def to_c(self):
        ng, alpha = ROSETTE_TYPES[self.type]
        r = FsrRosette(id=self.id, numnod=len(self.nodes), ngage=ng, zero_init=int(self.zero_init),
                       zpos=self.zpos, emod=self.emod, nu=self.nu, alpha_gages=alpha, gate=self.gate)
        for i, n in enumerate(self.nodes):
            r.nodes[i] = int(n)
        pos = np.asarray(self.rpos, F64)
        for j in range(4):
            for i in range(3):
                r.rpos[3 * j + i] = pos[i, j]
        for k in range(4):
            r.sncurve[k] = float(self.sncurve[k])
        return r


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def write_rosette_file(path, rosettes, link_id, minex=None, user_ids=None, descr=None):
    """The rosette input file of fedem_gage in .fsi format: one &STRAIN_ROSETTE record per rosette as the GUI writes them
    (solverTests/InversePy/shell_strain/fedem_solver.fsi:162-177; rPos = posInGl row by row, 10 significant digits).
    nodes are written as EXTERNAL node numbers (minex[internal - 1])."""
    with open(path, "w") as f:
        for k, r in enumerate(rosettes):
            ext = [int(minex[n - 1]) if minex is not None else int(n) for n in r.nodes]
            pos = np.asarray(r.rpos, F64)
            f.write("&STRAIN_ROSETTE\n")
            f.write(f"  id = {r.id}\n  extId = {user_ids[k] if user_ids is not None else r.id}\n")
            f.write(f"  extDescr = '{descr[k] if descr is not None else 'rosette_%d' % r.id}'\n  linkId = {link_id}\n")
            f.write(f"  type = '{r.type}'\n  zeroInit = {int(r.zero_init)}\n  numnod = {len(ext)}\n")
            f.write("  nodes = " + " ".join(map(str, ext)) + "\n")
            f.write("  rPos = " + "\n         ".join(" ".join(f"{v: .9e}" for v in pos[i]) for i in range(3)) + "\n")
            f.write(f"  zPos = {r.zpos: .9e}\n  Emod = {r.emod: .9e}\n  nu   = {r.nu: .9e}\n")
            if r.gate > 0.0:
                f.write(f"  gateVal = {r.gate: .9e}\n")
            if any(v > 0.0 for v in r.sncurve):
                f.write("  snCurve = " + " ".join(f"{v: .9e}" for v in r.sncurve) + "\n")
            f.write("/\n\n")


def read_rosette_file(path, link_id):
    """(rosettes with EXTERNAL node numbers, user ids, descriptions) through the library's ReadStrainGages."""
    from ._lib import check, load_library
    lib = load_library()
    n = check(lib.fsr_fsi_read_rosettes(path.encode(), int(link_id), None, None, None, 0, 0), "fsr_fsi_read_rosettes")
    arr = (FsrRosette * max(n, 1))()
    uid = np.zeros(max(n, 1), I32)
    buf = C.create_string_buffer(128 * max(n, 1))
    check(lib.fsr_fsi_read_rosettes(path.encode(), int(link_id), arr, uid.ctypes.data_as(C.POINTER(C.c_int)), buf, 128, n),
          "fsr_fsi_read_rosettes")
    names = {round(a, 6): t for t, (g, a) in ROSETTE_TYPES.items()}
    out = []
    for k in range(n):
        c = arr[k]
        rpos = np.array([[c.rpos[3 * j + i] for j in range(4)] for i in range(3)])
        t = "SINGLE_GAGE" if c.ngage == 1 else names[round(c.alpha_gages, 6)]
        out.append(Rosette(id=c.id, nodes=[c.nodes[i] for i in range(c.numnod)], rpos=rpos, type=t, zpos=c.zpos, emod=c.emod,
                           nu=c.nu, zero_init=bool(c.zero_init), gate=c.gate, sncurve=[c.sncurve[i] for i in range(4)]))
    descr = [buf.raw[128 * k:128 * (k + 1)].split(b"\0")[0].decode() for k in range(n)]
    return out, uid[:n].copy(), descr


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class StrainGages:
    """All rosettes of one FE part on one GPU.  `recovery` is the part's StressRecovery with B and E
    loaded (openBandEmatrices); it may be closed after construction, as gage.f90:256 does."""

    def __init__(self, recovery, rosettes):
        self._lib = _lib.load_library()
        self._h = C.c_void_p()
        self.nros = len(rosettes)
        self.ndim = recovery.ndim
        arr = (FsrRosette * max(self.nros, 1))()
        for i, r in enumerate(rosettes):
            arr[i] = r.to_c()
        check(self._lib.fsr_gage_create(C.byref(self._h), recovery._h, arr, self.nros), "fsr_gage_create")

    def bcart(self):
        """Bcart [nros, 3, ndim] (strainRosetteModule.f90:727)."""
        b = np.zeros((self.nros, self.ndim, 3), F64)
        check(self._lib.fsr_gage_get_bcart(self._h, _dp(b)), "fsr_gage_get_bcart")
        return np.ascontiguousarray(b.transpose(0, 2, 1))

    def recover(self, Q):
        """Q [ndim, nsteps] -> values [nsteps, nros, NVAL] (see VALUE_NAMES)."""
        Q = np.asfortranarray(Q, F64)
        ns = Q.shape[1]
        v = np.zeros((ns, self.nros, NVAL), F64)
        check(self._lib.fsr_gage_recover(self._h, _dp(Q), Q.shape[0], ns, _dp(v)), "fsr_gage_recover")
        return v

    def fatigue(self, Q, to_mpa=1.0e-6, gate=25.0, curve=(15.117, 17.146, 4.0, 5.0), bin_size=10.0, nbins=0):
        """Rainflow + damage over the history Q (defaults = the -gate/-loga1/-loga2/-m1 defaults of
        gagemain.C and m2 of FFpSNCurve.H).  Returns dict of [nros, 4] arrays (column 0 = rosette max
        principal stress, 1..3 = gage legs) and bins [nros, 4, nbins]."""
        Q = np.asfortranarray(Q, F64)
        ns = Q.shape[1]
        n = 4 * self.nros
        curve = np.ascontiguousarray(curve, F64)
        dmg = np.zeros(n, F64); ncyc = np.zeros(n, I32); status = np.zeros(n, I32)
        bins = np.zeros((n, nbins), I32) if nbins > 0 else None
        nwarn = check(self._lib.fsr_gage_fatigue(self._h, _dp(Q), Q.shape[0], ns, float(to_mpa), float(gate),
                                                 _dp(curve), float(bin_size), nbins, _dp(dmg), _ip(ncyc),
                                                 _ip(bins), _ip(status)), "fsr_gage_fatigue")
        return dict(damage=dmg.reshape(-1, 4), ncycles=ncyc.reshape(-1, 4), status=status.reshape(-1, 4),
                    bins=bins.reshape(-1, 4, nbins) if bins is not None else None, nwarn=nwarn)

    # ---- device-resident streaming (bench / multi-GPU drivers) ---------------------------------
    def fatigue_begin(self, to_mpa=1.0e-6, gate=25.0, curve=(15.117, 17.146, 4.0, 5.0), bin_size=10.0, nbins=0,
                      stack_cap=0):
        curve = np.ascontiguousarray(curve, F64)
        check(self._lib.fsr_gage_fatigue_begin(self._h, float(to_mpa), float(gate), _dp(curve), float(bin_size),
                                               nbins, stack_cap), "fsr_gage_fatigue_begin")
        self._nbins = nbins

    def fatigue_feed_dev(self, q_ptr, ldq, step0, nsteps, mode, stream=None, want_pending=False):
        n = C.c_int(-1)
        check(self._lib.fsr_gage_fatigue_feed_dev(self._h, C.c_void_p(q_ptr), ldq, step0, nsteps, mode,
                                                  C.byref(n) if want_pending else None,
                                                  C.c_void_p(stream) if stream else None), "fsr_gage_fatigue_feed_dev")
        return n.value

    def fatigue_end(self):
        n = 4 * self.nros
        nb = getattr(self, "_nbins", 0)
        dmg = np.zeros(n, F64); ncyc = np.zeros(n, I32); status = np.zeros(n, I32)
        bins = np.zeros((n, nb), I32) if nb > 0 else None
        nwarn = check(self._lib.fsr_gage_fatigue_end(self._h, _dp(dmg), _ip(ncyc), _ip(bins), _ip(status)),
                      "fsr_gage_fatigue_end")
        return dict(damage=dmg.reshape(-1, 4), ncycles=ncyc.reshape(-1, 4), status=status.reshape(-1, 4),
                    bins=bins.reshape(-1, 4, nb) if bins is not None else None, nwarn=nwarn)

    # ---- strain coat recovery summary (calcStrainCoatData / calcAngleData, strainCoatModule.f90) ----
    COAT_ENV = ("epsMax", "epsMin", "sigMax", "sigMin", "gammaMax", "tauMax", "vmeMax", "vmsMax")
    COAT_SUMMARY = ("stressRange", "strainRange", "popAngle", "angSpread", "biAxMean", "biAxStdDev")

    def coat_summary(self, Q, angle_bins=541, biaxial_gate=10.0, chunk=0):
        """Envelopes, angle bins and biaxiality of every rosette over the history Q (fed in `chunk`-step calls when
        chunk > 0: the state carries over).  Returns dict name -> [nros] arrays (+ 'nBiAxial')."""
        Q = np.asfortranarray(Q, F64)
        ns = Q.shape[1]
        check(self._lib.fsr_coat_begin(self._h, int(angle_bins), float(biaxial_gate)), "fsr_coat_begin")
        step = chunk if chunk > 0 else max(ns, 1)
        for t0 in range(0, ns, step):
            q = np.asfortranarray(Q[:, t0:t0 + step])
            check(self._lib.fsr_coat_feed(self._h, _dp(q), q.shape[0], q.shape[1]), "fsr_coat_feed")
        env, summ, nb = np.zeros((8, self.nros), F64), np.zeros((6, self.nros), F64), np.zeros(self.nros, I32)
        check(self._lib.fsr_coat_end(self._h, _dp(env), _dp(summ), _ip(nb)), "fsr_coat_end")
        out = {k: env[i] for i, k in enumerate(self.COAT_ENV)}
        out.update({k: summ[i] for i, k in enumerate(self.COAT_SUMMARY)})
        out["nBiAxial"] = nb
        return out

    def set_coat_fatigue(self, scf=None):
        """fedem_fpp's fatigue series: signed abs-max principal stress * to_mpa * scf[r] as series 4 r (None: back to sigmaP(1))."""
        a = None if scf is None else np.ascontiguousarray(scf, F64)
        check(self._lib.fsr_gage_set_coat_fatigue(self._h, _dp(a)), "fsr_gage_set_coat_fatigue")

    def recover_dev(self, q_ptr, ldq, nsteps, values_ptr=None, stream=None):
        check(self._lib.fsr_gage_recover_dev(self._h, C.c_void_p(q_ptr), ldq, nsteps,
                                             C.c_void_p(values_ptr) if values_ptr else None,
                                             C.c_void_p(stream) if stream else None), "fsr_gage_recover_dev")

    def close(self):
        if self._h:
            self._lib.fsr_gage_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
