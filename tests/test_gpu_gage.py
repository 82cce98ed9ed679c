"""GPU parity of the strain-gage path (fedem_gage: Bcart construction, per-step rosette results,
rainflow + damage of the max principal stress and the gage legs) through the C ABI against the
oracle.  Bar: <= 1e-10 relative on FP64 strains/stresses, identical cycle counts and histograms."""
import numpy as np
import pytest

from fedem_solvers_b200 import StressRecovery, StrainGages
from fedem_solvers_b200.model import plate_part, reduced_history, rosettes_on_part
from fedem_solvers_b200.gage import NVAL

pytestmark = pytest.mark.gpu
TOL = 1.0e-10
CURVE = (15.117, 17.146, 4.0, 5.0)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("tri,shuffle,types", [(0.0, False, ["TRIPLE_GAGE_45"]),
                                               (0.4, True, ["SINGLE_GAGE", "DOUBLE_GAGE_90", "TRIPLE_GAGE_60", "TRIPLE_GAGE_45"])])
def test_rosette_bcart_and_history(oracle, tri, shuffle, types):
    part = plate_part(7, 6, ngen=5, seed=21, tri_fraction=tri, shuffle_eq=shuffle, n_fixed=2, n_constraints=3,
                      warp=0.03, n_ext=4)
    ros = []
    for k, ty in enumerate(types):
        ros += rosettes_on_part(part, 9, seed=30 + k, rtype=ty, zero_init_fraction=0.3, top_surface=bool(k % 2))
    # one rosette sits on an element that touches an external node (unit rows of H_el)
    b = oracle.bind_part(part)
    rec = StressRecovery(part)
    g = StrainGages(rec, ros)
    rec.close()  # closeBandEmatrices: Bcart no longer needs the part (gage.f90:256)
    Bg = g.bcart()
    Q = reduced_history(part.sam.ndim, 700, seed=8)   # two device tiles
    V = g.recover(Q)
    assert V.shape == (700, len(ros), NVAL)
    for i, r in enumerate(ros):
        Bo = oracle.rosette_bcart(b, r)
        assert rel(Bg[i], Bo) <= TOL, (i, rel(Bg[i], Bo))
        Vo = oracle.rosette_history(b, r, Q, Bo)
        for k in range(NVAL):
            if k in (8, 9):      # angles: atan2, compare absolutely
                assert np.abs(V[:, i, k] - Vo[:, k]).max() <= 1e-9, (i, k)
            else:
                scale = max(np.abs(Vo[:, [0, 1, 2]]).max() if k < 8 or 18 <= k < 21 else np.abs(Vo[:, 10:13]).max(), 1e-300)
                assert np.abs(V[:, i, k] - Vo[:, k]).max() <= TOL * scale, (i, k)
    g.close()


def test_gage_fatigue_matches_reference_chain(oracle):
    """rosette stress histories -> PVX -> rainflow -> damage + histogram on the GPU vs the oracle chain
    fed with the oracle's own stress histories (so a parity slip anywhere changes cycle counts)."""
    part = plate_part(6, 6, ngen=6, seed=4)
    ros = rosettes_on_part(part, 25, seed=5, rtype="TRIPLE_GAGE_45")
    ros += rosettes_on_part(part, 6, seed=6, rtype="DOUBLE_GAGE_90", zero_init_fraction=1.0)
    b = oracle.bind_part(part)
    rec = StressRecovery(part)
    g = StrainGages(rec, ros)
    ns = 2500
    Q = reduced_history(part.sam.ndim, ns, seed=12, amp=2e-3)
    Q[:, :40] = Q[:, [40]]           # the run starts at rest: plateau -> duplicated first turning point
    to_mpa, gate = 1e-6, 3.0
    res = g.fatigue(Q, to_mpa=to_mpa, gate=gate, curve=CURVE, bin_size=5.0, nbins=60)
    ncheck = 0
    for i, r in enumerate(ros):
        Vo = oracle.rosette_history(b, r, Q)
        series = [Vo[:, 13] * to_mpa] + [Vo[:, 21 + k] * to_mpa for k in range(3)]
        ng = r.to_c().ngage
        for k, x in enumerate(series):
            if k > ng:
                assert res["ncycles"][i, k] == 0 and res["damage"][i, k] == 0.0
                continue
            d, n, bins, ok = oracle.series_fatigue(x, gate, CURVE, 5.0, 60)
            if not ok:
                assert res["status"][i, k] == 1
                continue
            # the GPU history differs from the oracle's by ~1e-16 relative: a cycle whose range sits within
            # that distance of the gate or of a bin edge may legitimately flip; none does for this seed
            assert res["ncycles"][i, k] == n, (i, k, res["ncycles"][i, k], n)
            assert np.array_equal(res["bins"][i, k], bins), (i, k)
            assert abs(res["damage"][i, k] - d) <= 1e-9 * max(d, 1e-300), (i, k, res["damage"][i, k], d)
            ncheck += n
    assert ncheck > 1000
    g.close(); rec.close()


def test_in_core_vms_layout(oracle):
    """fedempy's savePartStressState layout: [iel, nenod, nstrp, vm...] per active element with
    stress points (stressRoutines.f90:324-331, stressRecoveryModule.f90:718-747)."""
    import ctypes as C
    part = plate_part(4, 3, ngen=3, seed=9, tri_fraction=0.5)
    part.elm.elmid[1] = 0
    b = oracle.bind_part(part)
    rec = StressRecovery(part)
    lib = rec._lib
    n = lib.fsr_vms_size(rec._h)
    nstrp = part.nstrp()
    assert n == int((3 + nstrp[nstrp > 0]).sum())
    q = reduced_history(part.sam.ndim, 1, seed=2)[:, 0].copy()
    vms = np.zeros(n)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.fsr_get_vms(rec._h, dp(q), dp(vms), n) == 0
    ref = oracle.calc_stresses(b, oracle.expand(b, q))
    k = 0
    for e in range(part.sam.nel):
        if nstrp[e] == 0:
            continue
        assert vms[k] == e + 1 and vms[k + 1] == part.sam.mpmnpc[e + 1] - part.sam.mpmnpc[e] and vms[k + 2] == nstrp[e]
        want = ref["resmat"][b["ptoff"][e]:b["ptoff"][e] + nstrp[e], 0]
        assert np.abs(vms[k + 3:k + 3 + nstrp[e]] - want).max() <= TOL * np.abs(ref["resmat"][:, 0]).max()
        k += 3 + nstrp[e]
    assert k == n
    rec.close()


def test_strain_coat_summary(oracle):
    """calcStrainCoatData / calcAngleData / BiAxMean / BiAxStdDev per rosette on the GPU against the oracle's restatement fed with
    the oracle's own rosette values: envelopes and ranges <= 1e-10, angle-bin derived quantities (most popular angle, angle
    spread) and the gated step count identical; the state carries over between feed calls of ragged size."""
    import ctypes as C
    from oracle_bind import _dp, _ip
    part = plate_part(6, 5, ngen=5, seed=41, tri_fraction=0.3, warp=0.02)
    ros = rosettes_on_part(part, 40, seed=42, rtype="TRIPLE_GAGE_45", zero_init_fraction=0.2)
    b = oracle.bind_part(part)
    rec = StressRecovery(part)
    g = StrainGages(rec, ros)
    ns = 1500
    Q = reduced_history(part.sam.ndim, ns, seed=13, amp=2e-3)
    Vo = [oracle.rosette_history(b, r, Q) for r in ros]
    gate = float(np.median([np.abs(v[:, 15]).max() for v in Vo]) * 0.3)     # about a third of the steps above the gate
    for chunk, nbins in ((0, 541), (333, 541), (97, 12)):      # 12 -> 11 bins: odd count, coarse bins
        res = g.coat_summary(Q, angle_bins=nbins, biaxial_gate=gate, chunk=chunk)
        f = oracle.lib.orc_coat_summary
        f.restype = C.c_int
        n_gated = 0
        for i, r in enumerate(ros):
            env, summ = np.zeros(8), np.zeros(6)
            v = np.ascontiguousarray(Vo[i])
            nb = f(_dp(v), ns, nbins, C.c_double(gate), _dp(env), _dp(summ), None)
            sc_e, sc_s = np.abs(v[:, 3:6]).max(), np.abs(v[:, 13:16]).max()
            for k, name in enumerate(g.COAT_ENV):
                sc = sc_e if name in ("epsMax", "epsMin", "gammaMax", "vmeMax") else sc_s
                assert abs(res[name][i] - env[k]) <= TOL * sc, (i, name, res[name][i], env[k])
            assert abs(res["stressRange"][i] - summ[0]) <= TOL * sc_s and abs(res["strainRange"][i] - summ[1]) <= TOL * sc_e
            assert res["popAngle"][i] == summ[2] and res["angSpread"][i] == summ[3], (i, res["popAngle"][i], summ[2], res["angSpread"][i], summ[3])
            assert res["nBiAxial"][i] == nb
            assert abs(res["biAxMean"][i] - summ[4]) <= 1e-9 and abs(res["biAxStdDev"][i] - summ[5]) <= 1e-9
            n_gated += nb
        assert 0.05 * ns * len(ros) < n_gated < 0.95 * ns * len(ros)
    g.close(); rec.close()
