"""GPU parity of the stress results database written from the device (csrc/io_rdb.cu): the file is read back
through the product's .frs reader and (when built) the reference's own FFrLib, and every value is compared with
the CPU oracle's calcStresses of the same step: <= 1e-10 relative in -double files, float32 rounding of the
oracle's value otherwise."""
import os
import numpy as np
import pytest

from fedem_solvers_b200 import StressRecovery
from fedem_solvers_b200.frs import FrsReader
from fedem_solvers_b200.model import plate_part, tet10_block, hex20_block, linsolid_block, thickshell_panel, reduced_history
from fedem_solvers_b200.rdb import StressRdb, out_mask
from test_rdb_cpu import NAMES, NENOD, MEASURES

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def oracle():
    import oracle_bind
    return oracle_bind.Oracle()


def _close(a, b, scale, double):
    if double:
        return np.abs(a - b).max() <= TOL * scale
    # float file: the stored value is the float32 rounding of a double within TOL of the oracle's
    return np.abs(a - b).max() <= 1.3e-7 * scale


def _check(oracle, part, tmp_path, mask, double, nsteps, total=False, step_tile=0, rdbinc=2):
    b = oracle.bind_part(part)
    sam = part.sam
    Q = reduced_history(sam.ndim, nsteps, seed=4)
    rng = np.random.default_rng(3)
    rec = StressRecovery(part, step_tile=step_tile)
    stepno = np.arange(10, 10 + nsteps)
    time = 0.01 * stepno
    sup_tr = tr0 = None
    if total:
        from scipy.spatial.transform import Rotation
        tr0 = np.hstack([Rotation.from_rotvec([0.1, -0.2, 0.3]).as_matrix(), [[1.0], [2.0], [3.0]]])
        sup_tr = np.stack([np.hstack([Rotation.from_rotvec(rng.normal(0, 0.5, 3)).as_matrix(), rng.normal(0, 1, (3, 1))])
                           for _ in range(nsteps)])
    path = str(tmp_path / "part_stress.frs")
    with StressRdb(rec, path, mask, double=double, rdbinc=rdbinc, base_id=21, user_id=3, descr=part.name,
                   link_file=part.name + ".ftl", elmid=part.elm.elmid, minex=sam.minex, sup_tr_init=tr0) as rdb:
        assert rdb.path.endswith(f"part_stress_{rdbinc}.frs")
        half = nsteps // 2   # two calls: records append
        rdb.write_steps(Q[:, :half], stepno[:half], time[:half], None if sup_tr is None else sup_tr[:half])
        rdb.write_steps(Q[:, half:], stepno[half:], time[half:], None if sup_tr is None else sup_tr[half:])
        step_bytes, out_path = rdb.step_bytes, rdb.path
    raw = open(out_path, "rb").read()
    hsize = raw.find(b"\nDATA:") + 6
    assert len(raw) - hsize == nsteps * step_bytes
    rd = FrsReader(out_path)
    assert rd.nsteps == nsteps and np.array_equal(rd.step_numbers, stepno) and np.allclose(rd.times, time, rtol=0, atol=0)
    ref = None
    try:
        from test_frs_cpu import RefFrs, REF_LIB
        if os.path.exists(REF_LIB):
            ref = RefFrs([out_path])
    except Exception:
        ref = None
    # the oracle's results of every step
    o = []
    for s in range(nsteps):
        sv = oracle.expand(b, Q[:, s])
        r = oracle.calc_stresses(b, sv)
        r["sv"] = sv
        o.append(r)
    ptoff = b["ptoff"]
    sel = [j for j in range(8) if mask & (1 << j)]
    sr_on, st_on, sn_on, def_on = bool(mask & 0x400), bool(mask & 0x100), bool(mask & 0x200), bool(mask & 0x800)
    scale = {k: max(np.abs(np.stack([x[k] for x in o])).max(), 1e-300) for k in ("stress", "strain", "sv")}
    shells = np.nonzero(sam.melcon > 20)[0]
    if len(shells) and sam.melcon[shells[0]] < 30:
        allsr = np.stack([x["sres"][shells] for x in o]).reshape(nsteps, len(shells), 4, 6)
        scale["srf"], scale["srm"] = np.abs(allsr[..., :3]).max(), np.abs(allsr[..., 3:]).max()
    n_checked = 0

    def check(p, want, sc):
        nonlocal n_checked
        h = rd.find(p, "Part", 21)
        assert h is not None, p
        got = rd.read(h)
        assert got.shape == want.shape, p
        assert _close(got, want, sc, double), (p, np.abs(got - want).max(), sc)
        if ref is not None and n_checked % 7 == 0:
            ok, g2 = ref.read(p, "Part", 21, ref.keys(), want.shape[1])
            assert ok == nsteps and np.array_equal(g2, got), p
        n_checked += 1

    elems = list(range(sam.nel))
    if sam.nel > 60:
        elems = sorted(set(rng.choice(sam.nel, 50, replace=False)) | {0, sam.nel - 1} | set(np.nonzero(sam.melcon == 11)[0][:5]))
    for e in elems:
        t = int(sam.melcon[e])
        if part.elm.elmid is not None and part.elm.elmid[e] <= 0:
            assert rd.find(f"Elements|{abs(part.elm.elmid[e])}|{NAMES[t]}|Element nodes|Basic|1|Stress", "Part", 21) is None
            continue
        p = f"Elements|{part.elm.elmid[e]}|{NAMES[t]}|Element nodes|"
        nn = NENOD[t]
        if t == 11:
            if sr_on:
                sf = np.stack([x["sres"][e, :12] for x in o])     # [nsteps, 12]
                for n in range(2):
                    check(p + f"Basic|{n + 1}|Beam sectional force", sf[:, 6 * n:6 * n + 3], np.abs(sf[:, 0::6]).max() + np.abs(sf[:, :3]).max())
                    check(p + f"Basic|{n + 1}|Beam sectional moment", sf[:, 6 * n + 3:6 * n + 6], np.abs(sf[:, 3:6]).max() + np.abs(sf[:, 9:]).max())
            continue
        shell = t < 40
        ncmp = 3 if t < 30 else 6
        sides = ("Top", "Bottom") if shell else ("Basic",)
        if shell and sr_on and t > 30:      # thick shells: SR = 0 for every node (STR31 / STR32 "maybe later")
            for n in range(nn):
                check(p + f"Basic|{n + 1}|Shell stress resultant force", np.zeros((nsteps, 3)), 1.0)
                check(p + f"Basic|{n + 1}|Shell stress resultant moment", np.zeros((nsteps, 3)), 1.0)
        elif shell and sr_on:
            sr = np.stack([x["sres"][e] for x in o])      # [nsteps, 24] = SR(6, node)
            for n in range(nn):
                check(p + f"Basic|{n + 1}|Shell stress resultant force", sr[:, 6 * n:6 * n + 3], scale["srf"])
                check(p + f"Basic|{n + 1}|Shell stress resultant moment", sr[:, 6 * n + 3:6 * n + 6], scale["srm"])
        for si, side in enumerate(sides):
            for n in (0, nn - 1):
                pt = ptoff[e] + si * nn + n
                q = p + f"{side}|{n + 1}|"
                if st_on:
                    check(q + "Stress", np.stack([x["stress"][pt, :ncmp] for x in o]), scale["stress"])
                if sn_on:
                    check(q + "Strain", np.stack([x["strain"][pt, :ncmp] for x in o]), scale["strain"])
                for j in sel:
                    # principal values: trigonometric cubic, 1e-9
                    sc = (scale["stress"] if j < 4 else scale["strain"]) * (1.0 if j in (0, 4) or t < 30 else 10.0)
                    check(q + MEASURES[j], np.stack([x["resmat"][pt, j:j + 1] for x in o]), sc)
    if def_on:
        from oracle_bind import _dp
        import ctypes as C
        for n in sorted(set(rng.choice(sam.nnod, min(sam.nnod, 25), replace=False)) | {0, sam.nnod - 1}):
            j0, nd = sam.madof[n] - 1, sam.madof[n + 1] - sam.madof[n]
            p = f"Nodes|{sam.minex[n]}|Dynamic response|"
            u = np.stack([x["sv"][j0:j0 + nd] for x in o])
            check(p + "Translational deformation", u[:, :3], scale["sv"])
            if nd > 5:
                check(p + "Angular deformation", u[:, 3:6], scale["sv"])
            if total:
                ut = np.zeros((nsteps, 6))
                for s in range(nsteps):
                    x0 = np.ascontiguousarray(part.elm.xyz[n])
                    oracle.lib.orc_total_nodal_displacement(_dp(x0), _dp(np.ascontiguousarray(u[s])), 6 if nd > 5 else 3,
                                                            _dp(np.ascontiguousarray(sup_tr[s].T.ravel())),
                                                            _dp(np.ascontiguousarray(tr0.T.ravel())), _dp(ut[s]))
                check(p + "Total translation", ut[:, :3], np.abs(ut[:, :3]).max())
                if nd > 5:
                    check(p + "Total rotation", ut[:, 3:], np.abs(ut[:, 3:]).max())
    assert n_checked > 10
    if ref is not None:
        ref.close()
    rec.close()
    return n_checked


def test_plate_von_mises_float_file(oracle, tmp_path):
    part = plate_part(9, 8, ngen=6, seed=2, tri_fraction=0.3, warp=0.02)
    _check(oracle, part, tmp_path, out_mask(vmStress=True), False, nsteps=11)


def test_plate_everything_double_with_total_displacements(oracle, tmp_path):
    part = plate_part(7, 6, ngen=5, seed=3, tri_fraction=0.5, warp=0.02)
    part.elm.elmid = part.elm.elmid.copy()
    part.elm.elmid[[3, 17]] *= -1     # outside the -group selection: absent from the file
    mask = out_mask(SR=True, stress=True, strain=True, vmStress=True, maxPStress=True, minPStress=True, maxSStress=True,
                    vmStrain=True, maxPStrain=True, minPStrain=True, maxSStrain=True, deformation=True)
    _check(oracle, part, tmp_path, mask, True, nsteps=9, total=True)


def test_tets_and_beams_multi_tile(oracle, tmp_path):
    part = tet10_block(3, 2, 2, ngen=6, seed=5, n_beams=7)
    mask = out_mask(SR=True, stress=True, vmStress=True, maxPStress=True, deformation=True)
    _check(oracle, part, tmp_path, mask, True, nsteps=150, step_tile=64)   # 3 record tiles, the last one ragged


def test_hex20_strain_measures_float(oracle, tmp_path):
    part = hex20_block(2, 2, 1, ngen=4, seed=6)
    _check(oracle, part, tmp_path, out_mask(strain=True, vmStrain=True, maxSStrain=True), False, nsteps=5)


def test_linear_solids_all_measures_double(oracle, tmp_path):
    part = linsolid_block(2, 2, 1, ngen=4, seed=7)
    mask = out_mask(stress=True, strain=True, vmStress=True, maxPStress=True, minPStress=True, maxSStress=True, vmStrain=True)
    _check(oracle, part, tmp_path, mask, True, nsteps=6)


def test_thick_shells_everything_double(oracle, tmp_path):
    part = thickshell_panel(3, 2, ngen=4, seed=13)
    mask = out_mask(SR=True, stress=True, strain=True, vmStress=True, maxPStress=True, minPStress=True, maxSStress=True,
                    vmStrain=True, maxPStrain=True, minPStrain=True, maxSStrain=True, deformation=True)
    _check(oracle, part, tmp_path, mask, True, nsteps=7)


def test_thick_shells_von_mises_float(oracle, tmp_path):
    part = thickshell_panel(2, 3, ngen=3, seed=14)
    _check(oracle, part, tmp_path, out_mask(vmStress=True, strain=True), False, nsteps=5)


def test_pipeline_with_many_small_tiles(oracle, tmp_path, monkeypatch):
    """record tiles of 8 steps: the double-buffered device / PCIe / writer-thread pipeline wraps around several times,
    two writer helpers split the steps of a tile, the last tile of each call is ragged"""
    monkeypatch.setenv("FSR_RDB_TILE", "8")
    monkeypatch.setenv("FSR_RDB_WRITERS", "3")
    part = plate_part(8, 7, ngen=5, seed=9, tri_fraction=0.4, warp=0.02)
    mask = out_mask(SR=True, stress=True, strain=True, vmStress=True, maxPStress=True, deformation=True)
    _check(oracle, part, tmp_path, mask, False, nsteps=45, total=True)
    part = hex20_block(2, 1, 1, ngen=3, seed=10)
    _check(oracle, part, tmp_path, out_mask(stress=True, vmStress=True, minPStress=True), True, nsteps=19)


def test_pipeline_with_pwritev_file_writes(oracle, tmp_path, monkeypatch):
    """FSR_RDB_MMAP=0: the step records go out with pwritev from the helper threads instead of copies into a shared mapping of
    the file (the default, and what every other case runs: byte ranges that cut through keys and records); same files, checked
    value by value like every other case"""
    monkeypatch.setenv("FSR_RDB_MMAP", "0")
    monkeypatch.setenv("FSR_RDB_TILE", "8")
    monkeypatch.setenv("FSR_RDB_WRITERS", "3")
    part = plate_part(8, 7, ngen=5, seed=9, tri_fraction=0.4, warp=0.02)
    mask = out_mask(SR=True, stress=True, strain=True, vmStress=True, maxPStress=True, deformation=True)
    _check(oracle, part, tmp_path, mask, False, nsteps=45, total=True)
    monkeypatch.delenv("FSR_RDB_TILE")
    part = hex20_block(2, 1, 1, ngen=3, seed=10)
    _check(oracle, part, tmp_path, out_mask(stress=True, vmStress=True, minPStress=True), True, nsteps=19)


def test_solver_recovery_switches_and_frs3_files(oracle, tmp_path):
    """-recovery / -partVMStress / -partDeformation / -frs3file of the dynamics solver (solverInterface.C:447-452) over the
    part registry: state arrays only for -partVMStress >= 2 (getStressSize: -1 otherwise), one frs file per recovered part
    with the deformations and / or the von Mises stresses of every saved step (recoverAndSave), nothing for a step with
    doSave off, nothing at all for an even -recovery."""
    import ctypes as C
    from fedem_solvers_b200 import _lib
    lib = _lib.load_library()
    parts = [plate_part(5, 4, ngen=3, seed=71, tri_fraction=0.3), tet10_block(2, 2, 1, ngen=2, seed=72, curved="surface")]
    recs = [StressRecovery(p) for p in parts]
    binds = [oracle.bind_part(p) for p in parts]
    f1, f2 = str(tmp_path / "th_p1.frs"), str(tmp_path / "th_p2.frs")
    opts = f'-recovery 3 -partVMStress 1 -partDeformation 1 -double -frs3file <"{f1}","{f2}">'
    assert lib.fsr_recovery_options(opts.encode()) == 0
    ids = np.array([11, 12], np.int32)
    for k, (p, r) in enumerate(zip(parts, recs)):
        mx = np.ascontiguousarray(p.sam.minex, np.int32)
        assert lib.fsr_recovery_register_part(int(ids[k]), 100 + k, p.name.encode(), r._h, mx.ctypes.data_as(_lib._I), None) == 0, lib.fsr_last_error()
    assert lib.getPartStressStateSize(11) == -1 and lib.getPartDeformationStateSize(11) == 3 * parts[0].sam.nnod + 4
    buf = C.create_string_buffer(512)
    assert lib.fsr_recovery_file(12, buf, 512) > 0 and buf.value.decode() == f2
    Qs = [reduced_history(p.sam.ndim, 4, seed=73 + k) for k, p in enumerate(parts)]
    saved = []
    for s in range(4):
        cols = [np.ascontiguousarray(Q[:, s]) for Q in Qs]
        qs = (_lib._D * 2)(*[c.ctypes.data_as(_lib._D) for c in cols])
        do_save = s != 2
        assert lib.fsr_recovery_update_parts_save(2, ids.ctypes.data_as(_lib._I), 5 + s, 0.1 * s, 0.1, qs, None, int(do_save)) == 0, lib.fsr_last_error()
        if do_save:
            saved.append(s)
    # the state of the last step is in core whether it was saved or not
    d = np.zeros(3 * parts[0].sam.nnod + 4)
    assert lib.savePartDeformationState(11, d.ctypes.data_as(_lib._D), len(d)) and d[0] == 8.0
    assert lib.fsr_recovery_close() == 0
    assert lib.getPartDeformationStateSize(11) == -999
    for k, (p, b, path) in enumerate(zip(parts, binds, (f1, f2))):
        rd = FrsReader(path)
        assert rd.nsteps == len(saved) and list(rd.step_numbers) == [5 + s for s in saved]
        svs = [oracle.expand(b, Qs[k][:, s]) for s in saved]
        res = [oracle.calc_stresses(b, sv) for sv in svs]
        scale = max(np.abs(np.stack(svs)).max(), 1e-300)
        for n in (0, p.sam.nnod - 1):
            h = rd.find(f"Nodes|{p.sam.minex[n]}|Dynamic response|Translational deformation", "Part", int(ids[k]))
            assert h is not None
            j0 = p.sam.madof[n] - 1
            assert np.abs(rd.read(h) - np.stack([sv[j0:j0 + 3] for sv in svs])).max() <= TOL * scale
        e = int(np.nonzero(p.sam.melcon > 20)[0][0])
        t = int(p.sam.melcon[e])
        side = "Top" if t < 40 else "Basic"
        h = rd.find(f"Elements|{e + 1}|{NAMES[t]}|Element nodes|{side}|1|{MEASURES[0]}", "Part", int(ids[k]))
        assert h is not None
        pt = b["ptoff"][e]
        want = np.stack([r["resmat"][pt, 0:1] for r in res])
        assert np.abs(rd.read(h) - want).max() <= TOL * np.abs(np.stack([r["resmat"][:, 0] for r in res])).max()
    # gages only: no stress recovery, no files, no state
    assert lib.fsr_recovery_options(b"-recovery 2 -frs3file " + str(tmp_path / "none.frs").encode()) == 0
    assert lib.fsr_recovery_register(11, recs[0]._h, None) == 0
    q = np.ascontiguousarray(Qs[0][:, 0])
    assert lib.fsr_recovery_update(11, 1, 0.0, 0.1, q.ctypes.data_as(_lib._D)) == 0
    assert lib.getPartStressStateSize(11) == -1 and not os.path.exists(str(tmp_path / "none.frs"))
    assert lib.fsr_recovery_close() == 0
    # the library's own default again (no switches given): state arrays on
    assert lib.fsr_recovery_register(11, recs[0]._h, None) == 0
    assert lib.getPartStressStateSize(11) > 0
    assert lib.fsr_recovery_close() == 0
    assert lib.fsr_recovery_options(b"-partVMStress 7") < 0
    for r in recs:
        r.close()
