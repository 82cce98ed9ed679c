"""End-to-end drop-in test of bin/fedem_stress on the GPU: the files a reducer + dynamics solver run leaves behind
(.ftl FE part, _SAM.fsm, _B.fmx, _E.fmx, _V.fmx gravitation modes, fedem_solver.fsi, th_p_*.frs time history) are
generated, the executable is run with the reference's command-line options (some of them through -fco/-fop
option files), and its stress results database is read back and compared with the CPU oracle driven by the
same reference-order recipe: readSupElDisplacements -> BuildFinit -> calcIntDisplacements -> calcStresses."""
import os
import subprocess
import numpy as np
import pytest

from fedem_solvers_b200.files import save_part, write_fmx
from fedem_solvers_b200.frs import FrsReader
from fedem_solvers_b200.fsi import SolverPart, write_fsi
from fedem_solvers_b200.ftl import write_ftl
from fedem_solvers_b200 import StressRecovery
from fedem_solvers_b200.model import plate_part, tet10_block, thickshell_panel, reduced_history
from test_frs_cpu import _write_solver_file, _build_finit_numpy
from test_rdb_cpu import NAMES, NENOD, MEASURES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "fedem_solvers_b200", "bin", "fedem_stress")
TOL = 1e-10


@pytest.fixture(scope="module")
def oracle():
    import oracle_bind
    return oracle_bind.Oracle()


def _make_case(tmp_path, part, name, nsteps, gravity=None, seed=1):
    """Writes every input file of a stress run for `part` into tmp_path; returns what the checks need."""
    rng = np.random.default_rng(seed)
    sam = part.sam
    ntriads, ngen, base = sam.ndof2 // 6, sam.ngen, 40
    assert ntriads * 6 == sam.ndof2
    write_ftl(str(tmp_path / f"{name}.ftl"), part, groups={3: [int(i) for i in part.elm.elmid[: sam.nel // 2]]})
    save_part(str(tmp_path / name), part, checksum=1234, part_id=base)
    triads, tr_undef, sup, tri, gen = _write_solver_file(str(tmp_path / "th_p_1.frs"), rng, nsteps, ntriads, ngen, t0=0.0, dt=0.01,
                                                         step0=1, sup_base=base)
    ndofs, first = np.full(ntriads, 6), 1 + 6 * np.arange(ntriads)
    sp = SolverPart(base_id=base, user_id=2, descr="Flexible part", ngen=ngen, sup_pos=sup[0], gravity=np.zeros(3), model_file="",
                    triad_base_id=np.array([t[0] for t in triads]), triad_user_id=np.array([t[1] for t in triads]), ndofs=ndofs,
                    first_dof=first, tr_undef=tr_undef, triad_ur=tri[0], gen_first_dof=0)
    write_fsi(str(tmp_path / "fedem_solver.fsi"), [sp], gravity=gravity if gravity is not None else (0, 0, 0),
              model_file="generated.fmm")
    # the .fsi file carries 10 significant digits: use the undeformed triad positions exactly as the program reads them
    from fedem_solvers_b200.fsi import read_fsi
    tr_undef = read_fsi(str(tmp_path / "fedem_solver.fsi"), base).tr_undef
    Q = _build_finit_numpy(sup, tri, tr_undef, ndofs, first, gen, sam.ndof2 + 1, sam.ndof2 + ngen)
    V = None
    if gravity is not None:
        V = rng.normal(0, 1e-6, (sam.ndof1, 3))
        write_fmx(str(tmp_path / f"{name}_V.fmx"), V, tag="#FEDEM displacement matrix", checksum=1234)
    return dict(Q=Q, sup=sup, V=V, base=base, times=0.01 * np.arange(nsteps), stepno=1 + np.arange(nsteps))


def _oracle_steps(oracle, part, case, steps, gravity=None):
    b = oracle.bind_part(part)
    out = []
    for s in steps:
        sv = oracle.expand(b, case["Q"][:, s])
        if gravity is not None:   # vi += vii . g,  g = matmul(grv, supTr(:,1:3))  (stress.f90:412, displacementModule.f90:992-995)
            g = np.asarray(gravity) @ case["sup"][s][:, :3]
            vi = case["V"] @ g
            sam = part.sam
            sveq = np.zeros(sam.neq)
            sveq[sam.meqn1 - 1] = vi
            add = np.where(sam.meqn > 0, sveq[np.maximum(sam.meqn, 1) - 1], 0.0)
            assert sam.nceq == 0
            sv = sv + add
        r = oracle.calc_stresses(b, sv)
        r["sv"] = sv
        out.append(r)
    return b, out


def _run(tmp_path, args):
    r = subprocess.run([EXE, "-cwd", str(tmp_path)] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Stress calculation successfully completed" in r.stdout
    return r.stdout


def test_plate_all_results_with_option_files(oracle, tmp_path):
    part = plate_part(8, 7, ngen=5, seed=21, tri_fraction=0.3, warp=0.02, n_ext=4)
    case = _make_case(tmp_path, part, "plate", nsteps=40)
    with open(tmp_path / "stress.fco", "w") as f:
        f.write("# calculation options\n-linkfile plate.ftl -samfile plate_SAM.fsm\n-Bmatfile plate_B.fmx -eigfile plate_E.fmx\n"
                "-fsifile fedem_solver.fsi -frsfile th_p_1.frs\n-statm 0.05 -stotm 0.30 -tinc 0.0\n")
    with open(tmp_path / "stress.fop", "w") as f:
        f.write("-rdbfile results/plate.frs -rdbinc 3 -double\n-SR -stress -strain -vmStress -maxPStress -minPStress -maxSStress\n"
                "-vmStrain -maxPStrain -minPStrain -maxSStrain -deformation\n")
    os.makedirs(tmp_path / "results")
    _run(tmp_path, ["-fco", "stress.fco", "-fop", "stress.fop"])
    rd = FrsReader(str(tmp_path / "results" / "plate_3.frs"))
    steps = np.nonzero((case["times"] > 0.05 - 1e-9) & (case["times"] < 0.30 + 1e-9))[0]
    assert np.array_equal(rd.step_numbers, case["stepno"][steps]) and np.allclose(rd.times, case["times"][steps], rtol=0, atol=1e-15)
    b, o = _oracle_steps(oracle, part, case, steps)
    sc = {k: np.abs(np.stack([x[k] for x in o])).max() for k in ("stress", "strain", "resmat", "sv", "sres")}
    n = 0
    for e in (0, 5, part.sam.nel // 2, part.sam.nel - 1):
        t = int(part.sam.melcon[e]); nn = NENOD[t]
        p = f"Elements|{part.elm.elmid[e]}|{NAMES[t]}|Element nodes|"
        for si, side in enumerate(("Top", "Bottom")):
            for k in range(nn):
                pt = b["ptoff"][e] + si * nn + k
                for var, key, w in (("Stress", "stress", 3), ("Strain", "strain", 3)):
                    got = rd.read(rd.find(p + f"{side}|{k + 1}|{var}", "Part", case["base"]))
                    want = np.stack([x[key][pt, :w] for x in o])
                    assert np.abs(got - want).max() <= TOL * sc[key]
                for j in range(8):
                    got = rd.read(rd.find(p + f"{side}|{k + 1}|{MEASURES[j]}", "Part", case["base"]))
                    want = np.stack([x["resmat"][pt, j:j + 1] for x in o])
                    assert np.abs(got - want).max() <= TOL * (sc["stress"] if j < 4 else sc["strain"])
                    n += 1
        for k in range(nn):
            got = rd.read(rd.find(p + f"Basic|{k + 1}|Shell stress resultant moment", "Part", case["base"]))
            assert np.abs(got - np.stack([x["sres"][e, 6 * k + 3:6 * k + 6] for x in o])).max() <= TOL * sc["sres"]
    for nd in (0, part.sam.nnod - 1):
        got = rd.read(rd.find(f"Nodes|{part.sam.minex[nd]}|Dynamic response|Angular deformation", "Part", case["base"]))
        j0 = part.sam.madof[nd] - 1
        assert np.abs(got - np.stack([x["sv"][j0 + 3:j0 + 6] for x in o])).max() <= TOL * sc["sv"]
        assert rd.find(f"Nodes|{part.sam.minex[nd]}|Dynamic response|Total translation", "Part", case["base"]) is not None
    assert n > 100


def test_tets_group_selection_gravity_and_time_increment(oracle, tmp_path):
    part = tet10_block(2, 2, 2, ngen=4, seed=22, n_ext=4, n_beams=0)
    # external nodes of the generator have 3 DOFs; triads need 6: use beams on the external nodes instead? keep it simple:
    if part.sam.ndof2 % 6:
        pytest.skip("generated part has 3-DOF external nodes")
    case = _make_case(tmp_path, part, "block", nsteps=30, gravity=(0.0, 0.0, -9.81), seed=5)
    _run(tmp_path, ["-linkfile", "block.ftl", "-samfile", "block_SAM.fsm", "-fsifile", "fedem_solver.fsi", "-frsfile", "<th_p_1.frs>", "-group", "3", "-vmStress", "-stress", "-statm", "0.0",
                    "-stotm", "1.0", "-tinc", "0.03", "-dispfile", "block_V.fmx", "-rdbinc", "1"])
    rd = FrsReader(str(tmp_path / "block_1.frs"))
    steps = [0, 3, 6, 9, 12, 15, 18, 21, 24, 27, 29]   # ffr_setposition clamps 0.30 to the last key on file, like the reference
    assert np.array_equal(rd.step_numbers, case["stepno"][steps])
    b, o = _oracle_steps(oracle, part, case, steps, gravity=(0.0, 0.0, -9.81))
    sc = np.abs(np.stack([x["stress"] for x in o])).max()
    half = part.sam.nel // 2
    for e in (0, half - 1):
        for k in (0, 9):
            pt = b["ptoff"][e] + k
            got = rd.read(rd.find(f"Elements|{part.elm.elmid[e]}|TET10|Element nodes|Basic|{k + 1}|Stress", "Part", case["base"]))
            assert np.abs(got - np.stack([x["stress"][pt] for x in o])).max() <= 1.3e-7 * sc   # float file
    assert rd.find(f"Elements|{part.elm.elmid[half]}|TET10|Element nodes|Basic|1|Stress", "Part", case["base"]) is None


def test_thick_shell_panel_through_the_executable(oracle, tmp_path):
    """TRI6 / QUAD8 part: the .ftl lists a TRI6 around its perimeter (ffl_gettopol reorders the mid-side nodes last,
    FFlLinkHandler_F.C:657-664); stress tensors (TENSOR3 in Top / Bottom groups) and von Mises in a float file."""
    part = thickshell_panel(3, 3, ngen=4, seed=23, n_ext=4)
    case = _make_case(tmp_path, part, "panel", nsteps=12)
    out = _run(tmp_path, ["-linkfile", "panel.ftl", "-samfile", "panel_SAM.fsm", "-Bmatfile", "panel_B.fmx", "-eigfile", "panel_E.fmx",
                          "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs", "-rdbfile", "panel.frs", "-tinc", "0", "-stress", "-vmStress"])
    assert "Warning" not in out
    rd = FrsReader(str(tmp_path / "panel_1.frs"))
    steps = np.arange(12)
    assert np.array_equal(rd.step_numbers, case["stepno"])
    b, o = _oracle_steps(oracle, part, case, steps)
    sc = np.abs(np.stack([x["stress"] for x in o])).max()
    n = 0
    for e in range(part.sam.nel):
        t = int(part.sam.melcon[e]); nn = NENOD[t]
        p = f"Elements|{part.elm.elmid[e]}|{NAMES[t]}|Element nodes|"
        for si, side in enumerate(("Top", "Bottom")):
            for k in (0, nn // 2, nn - 1):
                pt = b["ptoff"][e] + si * nn + k
                got = rd.read(rd.find(p + f"{side}|{k + 1}|Stress", "Part", case["base"]))
                want = np.stack([x["stress"][pt] for x in o])
                assert got.shape == want.shape and np.abs(got - want).max() <= 1.3e-7 * sc
                got = rd.read(rd.find(p + f"{side}|{k + 1}|Von Mises stress", "Part", case["base"]))
                assert np.abs(got - np.stack([x["resmat"][pt, :1] for x in o])).max() <= 1.3e-7 * sc
                n += 1
    assert n > 50


def test_fedem_gage_executable(oracle, tmp_path):
    """bin/fedem_gage on generated files: rosette input file (.fsi format, external node numbers, one rosette listed
    against its normal -> checkRosette swap), strain gage results database read back and compared with the oracle's
    calcRosetteStrains per step (float file), fatigue report of the -resfile against the oracle's PVX / rainflow / damage."""
    import re
    from fedem_solvers_b200.gage import write_rosette_file
    from fedem_solvers_b200.model import rosettes_on_part
    part = plate_part(7, 6, ngen=5, seed=31, tri_fraction=0.3, warp=0.02, n_ext=4)
    part.sam.minex = (1000 + 3 * np.arange(part.sam.nnod)).astype(np.int32)     # external != internal node numbers
    nsteps = 400
    case = _make_case(tmp_path, part, "plate", nsteps=nsteps)
    ros = rosettes_on_part(part, 5, seed=32, rtype="TRIPLE_GAGE_45") + rosettes_on_part(part, 3, seed=33, rtype="SINGLE_GAGE", zero_init_fraction=1.0)
    for k, r in enumerate(ros):
        r.id = 70 + k
    ros[2].gate = 2.0
    listed = [type(r)(**{**r.__dict__}) for r in ros]
    listed[1].nodes = listed[1].nodes[::-1]       # against the rosette normal: fedem_gage swaps it back
    write_rosette_file(str(tmp_path / "gages.fsi"), listed, link_id=case["base"], minex=part.sam.minex, user_ids=list(range(1, 9)))
    exe = os.path.join(os.path.dirname(EXE), "fedem_gage")
    r = subprocess.run([exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx",
                        "-eigfile", "plate_E.fmx", "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs", "-rosfile", "gages.fsi",
                        "-rdbfile", "gage.frs", "-rdbinc", "2", "-stotm", "100", "-deformation", "-fatigue", "1", "-gate", "1.0", "-binSize", "2.0",
                        "-stressToMPaScale", "1.0e-6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Strain gage recovery successfully completed" in r.stdout and "Nodal ordering for Rosette 71 has been swapped" in r.stdout
    rd = FrsReader(str(tmp_path / "gage_2.frs"))
    assert rd.nsteps == nsteps and np.array_equal(rd.step_numbers, case["stepno"])
    b = oracle.bind_part(part)
    Q = case["Q"]
    res = open(tmp_path / "plate_gage.res").read()
    blocks = res.split("===== Computed damage in strain rosette =====")[1:]
    assert len(blocks) == len(ros)
    ncyc = 0
    for k, ro in enumerate(ros):
        Vo = oracle.rosette_history(b, ro, Q)
        ng = ro.to_c().ngage
        sc_e, sc_s = np.abs(Vo[:, :3]).max(), np.abs(Vo[:, 10:13]).max()
        def var(name, w):
            h = rd.find(name, "Strain rosette", ro.id)
            assert h is not None, name
            got = rd.read(h)
            assert got.shape == (nsteps, w), name
            return got
        eps = var("Strain tensor", 3)
        want = Vo[:, :3].copy(); want[:, 2] *= 0.5
        assert np.abs(eps - want).max() <= 1.3e-7 * sc_e
        assert np.abs(var("Stress tensor", 3) - Vo[:, 10:13]).max() <= 1.3e-7 * sc_s
        assert np.abs(var("Angle of maximum principal strain/stress", 1)[:, 0] - Vo[:, 8]).max() <= 1e-6
        for j in range(ng):
            assert np.abs(var(f"Gage {j + 1}|Gage strain", 1)[:, 0] - Vo[:, 18 + j]).max() <= 1.3e-7 * sc_e
            assert np.abs(var(f"Gage {j + 1}|Gage stress", 1)[:, 0] - Vo[:, 21 + j]).max() <= 1.3e-7 * sc_s
        assert rd.find(f"Gage {ng + 1}|Gage strain", "Strain rosette", ro.id) is None
        # CalcRosetteDisplacements: node deformations (-deformation), rosette position and Euler angles in the global system
        import ctypes as C
        from oracle_bind import _dp
        nn = len(ro.nodes)
        sam = part.sam
        U = np.stack([oracle.expand(b, Q[:, s_]) for s_ in range(nsteps)])
        disp = np.stack([U[:, sam.madof[n - 1] - 1: sam.madof[n - 1] + 2] for n in ro.nodes], 1)      # [nsteps, nn, 3]
        for i, n in enumerate(ro.nodes):
            got = var(f"Node{n}|Deformation", 3)
            assert np.abs(got - disp[:, i]).max() <= 1.3e-7 * np.abs(U).max()
        X0 = part.elm.xyz[np.asarray(ro.nodes) - 1]

        def axes(X):
            V = [np.zeros(3) for _ in range(3)]
            x, y, z = (np.ascontiguousarray(X[:, k_]) for k_ in range(3))
            assert oracle.lib.orc_shell_element_axes(nn, _dp(x), _dp(y), _dp(z), _dp(V[0]), _dp(V[1]), _dp(V[2])) == 0
            return np.stack(V, 1)                                                                   # columns = axes
        T0 = axes(X0)
        want_pos, want_ang = np.zeros((nsteps, 3)), np.zeros((nsteps, 3))
        ref_lib = os.path.join(ROOT, "oracle", "_ref", "libfedem_ref_frs.so")
        euler = C.CDLL(ref_lib).ffa_glbeulerzyx_ if os.path.exists(ref_lib) else None
        for s_ in range(nsteps):
            S = case["sup"][s_]
            posR = np.asarray(ro.rpos)[:, 3] + disp[s_].mean(0)
            want_pos[s_] = S[:, :3] @ posR + S[:, 3]
            Tg = axes(X0 + disp[s_]) @ T0.T @ S[:, :3]
            if euler is not None:       # the reference's own FaMat33::getEulerZYX
                a, ang = np.ascontiguousarray(Tg.T.reshape(-1)), np.zeros(3)
                euler(_dp(a), _dp(ang))
                want_ang[s_] = ang
            else:
                want_ang[s_] = [np.arctan2(Tg[2, 1], Tg[2, 2]), -np.arctan2(Tg[2, 0], np.hypot(Tg[0, 0], Tg[1, 0])), np.arctan2(Tg[1, 0], Tg[0, 0])]
        assert np.abs(var("Position", 3) - want_pos).max() <= 1.3e-7 * np.abs(want_pos).max()
        assert np.abs(var("Euler angles", 3) - want_ang).max() <= 2e-6
        # fatigue report: damage row of this rosette = max principal + legs
        gate = ro.gate if ro.gate > 0 else 1.0
        row = [l for l in blocks[k].splitlines() if re.search(r"E[+-]\d\d", l) and "gate value" not in l][0]
        dmg = [float(x) for x in re.findall(r"-?\d\.\d{5}E[+-]\d\d", row)]
        assert len(dmg) == 1 + ng
        assert f"gate value :{gate:12.5E}" in blocks[k]
        series = [Vo[:, 13] * 1e-6] + [Vo[:, 21 + j] * 1e-6 for j in range(ng)]
        for j, x in enumerate(series):
            d, n, bins, ok = oracle.series_fatigue(x, gate, (15.117, 17.146, 4.0, 5.0), 2.0, 40)
            assert ok and abs(dmg[j] - d) <= 2e-5 * max(d, 1e-300) + 1e-300, (k, j, dmg[j], d)
            ncyc += n
        # cycle histogram lines: counts of the first series (max principal) in its column
        d, n, bins, ok = oracle.series_fatigue(series[0], gate, (15.117, 17.146, 4.0, 5.0), 2.0, 40)
        for bi, cnt in enumerate(bins):
            m = re.search(rf"Stress cycles +{2.0 * bi:.2f} - *{2.0 * bi + 2.0:.2f}( .*)", blocks[k])
            if cnt > 0:
                assert m is not None and int(m.group(1)[:12]) == cnt, (k, bi, cnt)
    assert ncyc > 200


def test_fedem_gage_with_gravitation_modes(oracle, tmp_path):
    """-dispfile: the static gravitation deflection V . g (g = the solver input file's gravitation vector, not turned with
    the part, gage.f90:175-194) adds a constant strain to every rosette that does not start from zero."""
    import copy
    from fedem_solvers_b200.gage import write_rosette_file
    from fedem_solvers_b200.model import rosettes_on_part
    part = plate_part(5, 5, ngen=3, seed=35, warp=0.02, n_ext=4)
    grav = (0.0, 3.0, -9.81)
    nsteps = 20
    case = _make_case(tmp_path, part, "plate", nsteps=nsteps, gravity=grav)
    ros = rosettes_on_part(part, 4, seed=36, rtype="DOUBLE_GAGE_90")
    ros[3].zero_init = True
    write_rosette_file(str(tmp_path / "gages.fsi"), ros, link_id=case["base"], minex=part.sam.minex)
    exe = os.path.join(os.path.dirname(EXE), "fedem_gage")
    r = subprocess.run([exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx",
                        "-eigfile", "plate_E.fmx", "-dispfile", "plate_V.fmx", "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs",
                        "-rosfile", "gages.fsi", "-rdbfile", "gage.frs", "-stotm", "100"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rd = FrsReader(str(tmp_path / "gage_1.frs"))
    assert rd.nsteps == nsteps
    # the oracle with the gravitation modes as three more component modes of constant amplitude g
    gpart = copy.copy(part)
    gpart.sam = copy.copy(part.sam)
    gpart.sam.ngen = part.sam.ngen + 3
    gpart.E = np.asfortranarray(np.hstack([part.E, case["V"]]))
    b = oracle.bind_part(gpart)
    Q = np.vstack([case["Q"], np.tile(np.asarray(grav)[:, None], (1, nsteps))])
    for ro in ros:
        Vo = oracle.rosette_history(b, ro, Q)
        got = rd.read(rd.find("Stress tensor", "Strain rosette", ro.id))
        assert np.abs(got - Vo[:, 10:13]).max() <= 1.3e-7 * np.abs(Vo[:, 10:13]).max()
        got = rd.read(rd.find("Gage 2|Gage strain", "Strain rosette", ro.id))[:, 0]
        assert np.abs(got - Vo[:, 19]).max() <= 1.3e-7 * np.abs(Vo[:, :3]).max()
    # and the offset is really there: without the modes the first three rosettes would differ
    b0 = oracle.bind_part(part)
    V0 = oracle.rosette_history(b0, ros[0], case["Q"])
    V1 = oracle.rosette_history(b, ros[0], Q)
    assert np.abs(V1[:, 10:13] - V0[:, 10:13]).max() > 100 * 1.3e-7 * np.abs(V0[:, 10:13]).max()      # far above the file's float rounding


def _write_modal_file(path, rng, triads, part_base, ngen, steps, times, nmodes, ncomp):
    """A solver modal results file in the grammar of the reference's own (FFrTests/.../eigval_0001/ev_p_3.frs): Eigenvalues under
    the Mechanism, "Eigenvectors|Mode  n" item groups (translational + angular components, global directions) under every
    Triad and the component-mode part under the Part."""
    from fedem_solvers_b200.frs import FrsWriter
    hdr = " Module                  = fedem_solver;\nVARIABLES:\n<1;\"Time step number\";NONE;INT;32;NUMBER>\n<2;\"Physical time\";TIME;FLOAT;64;SCALAR>\n"
    hdr += "<3;\"Eigenvalue\";ANGLE/TIME;FLOAT;64;SCALAR>\n<4;\"Eigenfrequency\";NONE/TIME;FLOAT;32;SCALAR>\n<5;\"Damping ratio\";NONE;FLOAT;32;SCALAR>\n"
    nd = 3 * ncomp
    hdr += f"<6;\"Translational deformation\";LENGTH;FLOAT;64;VECTOR;({nd})>\n<7;\"Angular deformation\";ANGLE;FLOAT;64;VECTOR;({nd})>\n"
    hdr += f"<8;\"Generalized deformation\";NONE;FLOAT;64;VECTOR;({ngen * ncomp})>\n"
    hdr += "[1;\"Eigenvalues\";\n" + "".join(f"  [;\"Mode{m + 1:3d}\";<3><4><5>]\n" for m in range(nmodes)) + "]\n"
    hdr += "[2;\"Eigenvectors\";\n" + "".join(f"  [;\"Mode{m + 1:3d}\";<6><7>]\n" for m in range(nmodes)) + "]\n"
    hdr += "[3;\"Eigenvectors\";\n" + "".join(f"  [;\"Mode{m + 1:3d}\";<8>]\n" for m in range(nmodes)) + "]\n"
    hdr += "DATABLOCKS:\n<1><2>\n{\"Mechanism\";2;1;\"\";[1]}\n"
    for bid, uid, _ in triads:
        hdr += f"{{\"Triad\";{bid};{uid};\"\";[2]}}\n"
    hdr += f"{{\"Part\";{part_base};2;\"\";[3]}}\n"
    nbytes = nmodes * (8 + 4 + 4) + len(triads) * nmodes * 2 * nd * 8 + nmodes * ngen * ncomp * 8
    data = {}
    with FrsWriter(path, hdr, nbytes) as w:
        for st, tm in zip(steps, times):
            rec = b""
            for m in range(nmodes):
                rec += np.array([10.0 + m]).tobytes() + np.array([1.6 + m, 0.02], np.float32).tobytes()
            tri = rng.standard_normal((len(triads), nmodes, 6 * ncomp))      # eigVec(6*ncomp): component l at [6*l, 6*l+6)
            for t in range(len(triads)):
                for m in range(nmodes):
                    v = tri[t, m].reshape(ncomp, 6)
                    # the file stores the translational variable (3 x ncomp) and then the angular one: readSupElModes reads n*iComp
                    # values from the group and takes component l at offset n*(l-1), so write them in exactly that order
                    rec += np.ascontiguousarray(tri[t, m]).tobytes()
            gen = rng.standard_normal((nmodes, ngen * ncomp))
            rec += gen.tobytes()
            w.write_step(int(st), float(tm), np.frombuffer(rec, np.uint8))
            data[int(st)] = (tri, gen)
    return data


@pytest.mark.parametrize("damped", [False, True])
def test_fedem_modes_executable(oracle, tmp_path, damped):
    """bin/fedem_modes: the dynamic response and two eigenmodes at two of the solver's time steps, expanded on the GPU and written as
    vector data; read back and compared with the oracle's calcIntDisplacements of BuildFinit / readSupElModes columns."""
    part = plate_part(5, 4, ngen=3, seed=51, tri_fraction=0.3, warp=0.02, n_ext=4)
    nsteps = 12
    case = _make_case(tmp_path, part, "plate", nsteps=nsteps)
    sam = part.sam
    ntriads, ngen, ncomp, nmodes = sam.ndof2 // 6, sam.ngen, 2 if damped else 1, 3
    rng = np.random.default_rng(52)
    triads = [(11 + i, 1 + i, "") for i in range(ntriads)]          # the ids _write_solver_file gives the triads
    steps = [3, 8]
    modal = _write_modal_file(str(tmp_path / "ev_p_1.frs"), rng, triads, case["base"], ngen, [case["stepno"][k] for k in steps],
                              [case["times"][k] for k in steps], nmodes, ncomp)
    exe = os.path.join(os.path.dirname(EXE), "fedem_modes")
    sel = f"<<{case['times'][3]:.4f},1,3>,<{case['times'][8]:.4f},1,3>>"
    args = [exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx", "-eigfile",
            "plate_E.fmx", "-fsifile", "fedem_solver.fsi", "-frsfile", "<th_p_1.frs,ev_p_1.frs>", "-rdbfile", "modes.frs", "-double",
            "-recover_modes", sel] + (["-damped"] if damped else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Modal recovery successfully completed" in r.stdout
    out = str(tmp_path / "modes_1.frs")
    assert open(out, "rb").read(17) == b"#FEDEM modal data"
    rd = FrsReader(out)
    assert rd.nsteps == 2 and np.array_equal(rd.step_numbers, [case["stepno"][k] for k in steps])
    b = oracle.bind_part(part)
    nnod = sam.nnod
    six = np.array([sam.madof[i + 1] - sam.madof[i] >= 6 for i in range(nnod)])

    def vectors(sv):
        tra = np.concatenate([sv[sam.madof[i] - 1: sam.madof[i] + 2] for i in range(nnod)])
        rot = np.concatenate([sv[sam.madof[i] + 2: sam.madof[i] + 5] for i in range(nnod) if six[i]])
        return tra, rot
    for row, k in enumerate(steps):
        tra, rot = vectors(oracle.expand(b, case["Q"][:, k]))
        got = rd.read(rd.find("Vectors|Dynamic response|Translational deformation", "Part", case["base"]))[row]
        assert got.shape == tra.shape and np.abs(got - tra).max() <= TOL * np.abs(tra).max()
        got = rd.read(rd.find("Vectors|Dynamic response|Angular deformation", "Part", case["base"]))[row]
        assert np.abs(got - rot).max() <= TOL * np.abs(rot).max()
        tri, gen = modal[int(case["stepno"][k])]
        tinv = case["sup"][k][:, :3].T
        for m in (1, 3):
            for l in range(ncomp):
                q = np.zeros(sam.ndim)
                for t in range(ntriads):
                    e = tri[t, m - 1][6 * l: 6 * l + 6]
                    q[6 * t: 6 * t + 3] = tinv @ e[:3]
                    q[6 * t + 3: 6 * t + 6] = tinv @ e[3:]
                q[sam.ndof2:] = gen[m - 1][ngen * l: ngen * (l + 1)]
                tra, rot = vectors(oracle.expand(b, q))
                sub = ("Re|" if l == 0 else "Im|") if damped else ""
                got = rd.read(rd.find(f"Vectors|Mode{m:3d}|{sub}Translational deformation", "Part", case["base"]))[row]
                assert np.abs(got - tra).max() <= TOL * np.abs(tra).max(), (k, m, l)
                got = rd.read(rd.find(f"Vectors|Mode{m:3d}|{sub}Angular deformation", "Part", case["base"]))[row]
                assert np.abs(got - rot).max() <= TOL * np.abs(rot).max(), (k, m, l)
    assert rd.find("Vectors|Mode  2|Translational deformation", "Part", case["base"]) is None
    # the reference's own FFrExtractor finds the same arrays in the file (header grammar of writeModesHeader / writeNodesHeader)
    from test_frs_cpu import RefFrs, REF_LIB
    if os.path.exists(REF_LIB):
        # FFrResultContainer opens the data of a "fedem_modes" file only on demand (FFrResultContainer.C:153-161): give the copy
        # another module name of the same length so that the extractor indexes its time steps right away
        raw = open(out, "rb").read()
        assert raw.count(b"= fedem_modes;") == 1
        out2 = str(tmp_path / "modes_copy.frs")
        open(out2, "wb").write(raw.replace(b"= fedem_modes;", b"= fedem_m0des;"))
        ref = RefFrs([out2])
        keys = ref.keys()
        assert len(keys) == 2
        for name in ("Vectors|Dynamic response|Angular deformation", f"Vectors|Mode  3|{'Im|' if damped else ''}Translational deformation"):
            mine = rd.read(rd.find(name, "Part", case["base"]))
            ok, theirs = ref.read(name, "Part", case["base"], keys, mine.shape[1])
            assert ok == 2 and np.array_equal(theirs, mine), name
        ref.close()


def test_fedem_modes_one_file_per_mode_nodes_and_energy_density(oracle, tmp_path):
    """modes.f90:258-301,428-452: different mode lists per time, -write_nodes and -energy_density give one results file for the dynamic
    response and one per mode (file increments in creation order); the mode files hold records only for the times at which the mode was
    asked for, the nodal form next to the vector form, and the scaled strain energy density (calcStrainEnergyDensity) per result point."""
    part = plate_part(5, 4, ngen=3, seed=53, tri_fraction=0.3, warp=0.02, n_ext=4)
    case = _make_case(tmp_path, part, "plate", nsteps=12)
    sam = part.sam
    ntriads, ngen, nmodes = sam.ndof2 // 6, sam.ngen, 3
    rng = np.random.default_rng(54)
    triads = [(11 + i, 1 + i, "") for i in range(ntriads)]
    steps = [3, 8]
    modal = _write_modal_file(str(tmp_path / "ev_p_1.frs"), rng, triads, case["base"], ngen, [case["stepno"][k] for k in steps],
                              [case["times"][k] for k in steps], nmodes, 1)
    exe = os.path.join(os.path.dirname(EXE), "fedem_modes")
    sel = f"<<{case['times'][3]:.4f},2,1>,<{case['times'][8]:.4f},2,3>>"       # mode order of first appearance: 2, 1, 3
    args = [exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx", "-eigfile",
            "plate_E.fmx", "-fsifile", "fedem_solver.fsi", "-frsfile", "<th_p_1.frs,ev_p_1.frs>", "-rdbfile", "modes.frs", "-double",
            "-rdbinc", "4", "-write_nodes", "-energy_density", "-recover_modes", sel]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    b = oracle.bind_part(part)
    base = case["base"]

    def mode_q(k, m):
        tri, gen = modal[int(case["stepno"][k])]
        tinv = case["sup"][k][:, :3].T
        q = np.zeros(sam.ndim)
        for t in range(ntriads):
            e = tri[t, m - 1][:6]
            q[6 * t: 6 * t + 3] = tinv @ e[:3]
            q[6 * t + 3: 6 * t + 6] = tinv @ e[3:]
        q[sam.ndof2:] = gen[m - 1][:ngen]
        return q
    # file 4 = dynamic response (both times), 5 = mode 2 (both), 6 = mode 1 (first time only), 7 = mode 3 (second time only)
    expect = {4: (None, [3, 8]), 5: (2, [3, 8]), 6: (1, [3]), 7: (3, [8])}
    for inc, (m, ks) in expect.items():
        out = str(tmp_path / f"modes_{inc}.frs")
        assert open(out, "rb").read(17) == b"#FEDEM modal data"
        rd = FrsReader(out)
        assert rd.nsteps == len(ks) and np.array_equal(rd.step_numbers, [case["stepno"][k] for k in ks]), inc
        name = "Dynamic response" if m is None else f"Mode{m:3d}"
        for row, k in enumerate(ks):
            q = case["Q"][:, k] if m is None else mode_q(k, m)
            sv = oracle.expand(b, q)
            tra = np.concatenate([sv[sam.madof[i] - 1: sam.madof[i] + 2] for i in range(sam.nnod)])
            got = rd.read(rd.find(f"Vectors|{name}|Translational deformation", "Part", base))[row]
            assert np.abs(got - tra).max() <= TOL * np.abs(tra).max(), (inc, k)
            for node in (0, sam.nnod // 2, sam.nnod - 1):       # the nodal form
                j = sam.madof[node] - 1
                got = rd.read(rd.find(f"Nodes|{int(sam.minex[node])}|{name}|Translational deformation", "Part", base))[row]
                assert np.abs(got - sv[j:j + 3]).max() <= TOL * np.abs(tra).max(), (inc, k, node)
                got = rd.read(rd.find(f"Nodes|{int(sam.minex[node])}|{name}|Angular deformation", "Part", base))[row]
                assert np.abs(got - sv[j + 3:j + 6]).max() <= TOL * np.abs(sv).max(), (inc, k, node)
            if m is None:
                assert rd.find(f"Elements|{part.elm.elmid[0]}|QUAD4|Element nodes|Top|1|Scaled strain energy density", "Part", base) is None
                continue
            o = oracle.calc_stresses(b, sv)
            dens = (o["stress"][:, :3] * o["strain"][:, :3]).sum(axis=1) + o["stress"][:, 2] * o["strain"][:, 2]
            for e in (0, sam.nel // 2, sam.nel - 1):
                t = int(sam.melcon[e])
                tn, nn = ("QUAD4", 4) if t == 24 else ("TRI3", 3)
                for side, off in (("Top", 0), ("Bottom", nn)):
                    for i in range(nn):
                        h = rd.find(f"Elements|{part.elm.elmid[e]}|{tn}|Element nodes|{side}|{i + 1}|Scaled strain energy density", "Part", base)
                        assert h is not None, (inc, e, side, i)
                        got = rd.read(h)[row]
                        want = dens[b["ptoff"][e] + off + i]
                        assert abs(got - want) <= 1e-9 * np.abs(dens).max(), (inc, e, side, i, got, want)
    assert not os.path.exists(tmp_path / "modes_8.frs")


def _globalized_axes(X):
    """getShellElementAxes with doGlobalize (strainAndStressUtils.f90:297-434)"""
    n = np.cross(X[2] - X[0], X[3] - X[1]) if len(X) == 4 else np.cross(X[1] - X[0], X[2] - X[0])
    n = n / np.linalg.norm(n)
    if abs(n[1]) > 0.01 or abs(n[2]) > 0.01:
        v1 = np.array([n[1] ** 2 + n[2] ** 2, -n[0] * n[1], -n[0] * n[2]])
    else:
        v1 = np.cross(np.array([-n[1] * n[0], n[0] ** 2 + n[2] ** 2, -n[1] * n[2]]), n)
    v1 = v1 / np.linalg.norm(v1)
    return v1, np.cross(n, v1), n


@pytest.mark.parametrize("surface", [0, 3])
def test_fedem_fpp_executable(oracle, tmp_path, surface):
    """bin/fedem_fpp (fpp.f90): strain coat elements of the FE part -> rosettes in the element systems -> running envelopes, angle
    bins, biaxiality and rainflow damage on the GPU -> ONE summary record on the strain coat results database.  Every value read back
    and compared with the oracle's calcRosetteStrains / calcStrainCoatData / calcAngleData restatement and its PVX + rainflow + Miner
    sum on the S-N curve of the library file."""
    import ctypes as C
    from oracle_bind import _dp
    from test_fpp_cpu import SN_TEXT, coats_for
    from fedem_solvers_b200.gage import Rosette
    part = plate_part(6, 5, ngen=3, seed=81, tri_fraction=0.3, warp=0.02, n_ext=4)
    ns = 400
    case = _make_case(tmp_path, part, "plate", nsteps=ns)
    coats = coats_for(part)
    write_ftl(str(tmp_path / "plate.ftl"), part, strain_coats=coats)
    open(tmp_path / "curves.fsn", "w").write(SN_TEXT)
    sam, base = part.sam, case["base"]
    b = oracle.bind_part(part)
    # oracle side first: the gates are set from the response so that both the biaxiality gate and the PVX gate bite
    to_mpa = 1.0e-6
    curves = {(0, 0): (15.117, 17.146, 4.0, 5.0), (0, 1): (12.592, 16.320, 3.0, 5.0), (1, 0): (12.1818 - 0.2095 * 2.0,) * 2 + (3.0, 3.0),
              (1, 1): (12.0128 - 0.2509 * 2.0,) * 2 + (3.0, 3.0)}
    want = {}
    for sc in coats:
        e = sc["elm"]
        nodes = [int(k) for k in sam.mmnpc[sam.mpmnpc[e] - 1: sam.mpmnpc[e + 1] - 1]]
        X = part.elm.xyz[np.array(nodes) - 1]
        v1, v2, v3 = _globalized_axes(X)
        for name, h in sc["sets"]:
            if surface and {"Bottom": 1, "Mid": 2, "Top": 3}[name] != surface:
                continue
            z = h if h is not None else {"Bottom": -0.5, "Top": 0.5}[name] * float(part.elm.thk[e])
            r = Rosette(id=1, nodes=nodes, rpos=np.stack([v1, v2, v3, X.mean(0)], 1), type="SINGLE_GAGE", zpos=z, emod=float(part.elm.emod[e]),
                        nu=float(part.elm.rny[e]))
            want[(sc["id"], name)] = oracle.rosette_history(b, r, case["Q"])
    smax = max(np.abs(v[:, 15]).max() for v in want.values())
    bgate = 0.3 * float(np.median([np.abs(v[:, 15]).max() for v in want.values()]))
    pvx = float(0.05 * smax * to_mpa)
    exe = os.path.join(os.path.dirname(EXE), "fedem_fpp")
    args = [exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx", "-eigfile", "plate_E.fmx",
            "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs", "-rdbfile", "coat.frs", "-rdbinc", "2", "-double", "-debug", "1",
            "-statm", "0", "-stotm", "100", "-HistDataType", "1", "-SNfile", "curves.fsn", "-PVXGate", repr(pvx), "-biAxialGate", repr(bgate),
            "-stressToMPaScale", repr(to_mpa), "-angleBins", "181", "-surface", str(surface)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Invalid option value" not in r.stderr, r.stdout + r.stderr
    assert "Strain coat calculation successfully completed" in r.stdout and "STRAIN COAT RECOVERY SUMMARY" in r.stdout
    out = str(tmp_path / "coat_2.frs")
    assert open(out, "rb").read(23) == b"#FEDEM strain coat data"
    rd = FrsReader(out)
    assert rd.nsteps == 1 and rd.step_numbers[0] == 1 and rd.times[0] == 100.0
    f = oracle.lib.orc_coat_summary
    f.restype = C.c_int
    pvx_used = float(np.float32(pvx))          # -PVXGate is a float option (fppmain.C:47, ffa_cmdlinearg_getfloat)
    n_damaged = n_biax = 0
    for (cid, name), v in want.items():
        sc = coats[cid - 1000]
        nn = len(sam.mmnpc[sam.mpmnpc[sc["elm"]] - 1: sam.mpmnpc[sc["elm"] + 1] - 1])
        tn = "STRCT3" if nn == 3 else "STRCQ4"
        env, summ = np.zeros(8), np.zeros(6)
        nb = f(_dp(np.ascontiguousarray(v)), ns, 181, C.c_double(bgate), _dp(env), _dp(summ), None)

        def val(var):
            h = rd.find(f"Elements|{cid}|{tn}|Element|{name}|{var}", "Part", base)
            return None if h is None else float(rd.read(h)[0, 0])
        sc_s, sc_e = np.abs(v[:, 13:16]).max(), np.abs(v[:, 3:6]).max()
        # env: epsMax, epsMin, sigMax, sigMin, gammaMax, tauMax, vmeMax, vmsMax; summ: stress range, strain range, popAngle, angSpread, mean, std
        assert abs(val("Max principal stress") - (env[2] if abs(env[2]) > abs(env[3]) else env[3])) <= TOL * sc_s
        assert abs(val("Max shear stress") - env[5]) <= TOL * sc_s and abs(val("Max von Mises stress") - env[7]) <= TOL * sc_s
        assert abs(val("Max stress range") - summ[0]) <= TOL * sc_s and abs(val("Max strain range") - summ[1]) <= TOL * sc_e
        assert abs(val("Max principal strain") - (env[0] if abs(env[0]) > abs(env[1]) else env[1])) <= TOL * sc_e
        assert abs(val("Max shear strain") - env[4]) <= TOL * sc_e and abs(val("Max von Mises strain") - env[6]) <= TOL * sc_e
        assert val("Most popular angle") == summ[2] and val("Angle spread") == summ[3]
        if nb > 0:
            assert abs(val("Mean bi-axiality") - summ[4]) <= 1e-9 and abs(val("Biaxiality standard deviation") - summ[5]) <= 1e-9
            n_biax += 1
        else:
            assert val("Mean bi-axiality") is None
        if sc["fatigue"] is None:
            assert val("Damage") is None and val("Life (repeats)") is None
        else:
            a, bb, scf = sc["fatigue"]
            dmg, ncyc, _, ok = oracle.series_fatigue(v[:, 15] * to_mpa * scf, pvx_used, curves[(a, bb)])
            assert ok
            got = val("Damage")
            if dmg > 0:
                assert abs(got - dmg) <= 1e-9 * dmg and abs(val("Life (repeats)") - 1.0 / dmg) <= 1e-9 / dmg
                assert abs(val("Life (equnits)") - 100.0 / dmg) <= 1e-9 * 100.0 / dmg
                n_damaged += 1
            else:
                assert got == 1.0e20 and val("Life (repeats)") == 1.0e20
    assert n_damaged >= 3 and n_biax >= 3
    # the reference's own FFrExtractor reads the file (header grammar of saveStrainCoatModule)
    from test_frs_cpu import RefFrs, REF_LIB
    if os.path.exists(REF_LIB):
        ref = RefFrs([out])
        keys = ref.keys()
        assert len(keys) == 1
        (cid, name), v = next(iter(want.items()))
        nn = len(sam.mmnpc[sam.mpmnpc[coats[cid - 1000]["elm"]] - 1: sam.mpmnpc[coats[cid - 1000]["elm"] + 1] - 1])
        var = f"Elements|{cid}|{'STRCT3' if nn == 3 else 'STRCQ4'}|Element|{name}|Max von Mises stress"
        mine = rd.read(rd.find(var, "Part", base))
        ok, theirs = ref.read(var, "Part", base, keys, 1)
        assert ok == 1 and np.array_equal(theirs, mine)
        ref.close()


def test_fedem_fpp_history_files(oracle, tmp_path):
    """-writeHistory: per-step records of sigmaP(1:3), tauMax, sigmaVM, epsP(1:3), gammaMax, epsVM of every coat result point
    (writeHistoryHeader / writeRosetteDB)"""
    from test_fpp_cpu import coats_for
    from fedem_solvers_b200.gage import Rosette
    part = plate_part(5, 4, ngen=2, seed=83, tri_fraction=0.3, warp=0.02, n_ext=4)
    ns = 30
    case = _make_case(tmp_path, part, "plate", nsteps=ns)
    coats = coats_for(part, every=3)
    write_ftl(str(tmp_path / "plate.ftl"), part, strain_coats=coats)
    sam, base = part.sam, case["base"]
    exe = os.path.join(os.path.dirname(EXE), "fedem_fpp")
    args = [exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx", "-eigfile", "plate_E.fmx",
            "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs", "-rdbfile", "coat.frs", "-double", "-statm", "0", "-stotm", "100", "-writeHistory"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rd = FrsReader(str(tmp_path / "coat_1.frs"))
    assert rd.nsteps == ns
    b = oracle.bind_part(part)
    sc = coats[1]
    e = sc["elm"]
    nodes = [int(k) for k in sam.mmnpc[sam.mpmnpc[e] - 1: sam.mpmnpc[e + 1] - 1]]
    X = part.elm.xyz[np.array(nodes) - 1]
    v1, v2, v3 = _globalized_axes(X)
    rr = Rosette(id=1, nodes=nodes, rpos=np.stack([v1, v2, v3, X.mean(0)], 1), type="SINGLE_GAGE", zpos=0.5 * float(part.elm.thk[e]),
                 emod=float(part.elm.emod[e]), nu=float(part.elm.rny[e]))
    v = oracle.rosette_history(b, rr, case["Q"])
    tn = "STRCT3" if len(nodes) == 3 else "STRCQ4"
    for var, col in (("Max principal stress", 13), ("Min principal stress", 14), ("Signed abs max stress", 15), ("Max shear stress", 16),
                     ("Von Mises stress", 17), ("Max principal strain", 3), ("Signed abs max strain", 5), ("Max shear strain", 6), ("Von Mises strain", 7)):
        got = rd.read(rd.find(f"Elements|{sc['id']}|{tn}|Element|Top|{var}", "Part", base))[:, 0]
        assert np.abs(got - v[:, col]).max() <= TOL * np.abs(v[:, col]).max(), var


def test_direct_solution_without_solver_input_file(oracle, tmp_path):
    """fedem_stress without -fsifile (stress.f90:131-135,397): the results files hold the nodal displacements of the part
    themselves ("Vectors|Dynamic response|Displacement", readIntDisplacements) -- no B / E matrices, no expansion, the element
    kernels run on what was read; stresses, von Mises and the deformations come out as if the displacements had been expanded"""
    from fedem_solvers_b200.frs import FrsWriter
    part = plate_part(6, 5, ngen=3, seed=31, tri_fraction=0.3, warp=0.02)
    sam = part.sam
    base, ns = 77, 21
    write_ftl(str(tmp_path / "plate.ftl"), part)
    save_part(str(tmp_path / "plate"), part, checksum=99, part_id=base)
    b = oracle.bind_part(part)
    Q = reduced_history(sam.ndim, ns, seed=3)
    SV = np.stack([oracle.expand(b, Q[:, s]) for s in range(ns)])          # any nodal displacement field will do
    hdr = (" Module                  = fedem_solver;\nVARIABLES:\n<1;\"Time step number\";NONE;INT;32;NUMBER>\n<2;\"Physical time\";TIME;FLOAT;64;SCALAR>\n"
           f"<3;\"Displacement\";LENGTH;FLOAT;64;VECTOR;({sam.ndof})>\n[1;\"Vectors\";\n  [;\"Dynamic response\";<3>]\n]\n"
           f"DATABLOCKS:\n<1><2>\n{{\"Part\";{base};1;\"plate\";[1]}}\n")
    with FrsWriter(str(tmp_path / "lin_p_1.frs"), hdr, 8 * sam.ndof) as w:
        for s in range(ns):
            w.write_step(s + 1, 0.05 * s, SV[s])
    out = _run(tmp_path, ["-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-linkId", str(base), "-frsfile", "lin_p_1.frs", "-vmStress", "-stress",
                          "-deformation", "-double", "-statm", "0", "-stotm", "10", "-tinc", "0", "-rdbfile", "res.frs", "-rdbinc", "0"])
    assert "nodal displacements are read from the results database" in out
    rd = FrsReader(str(tmp_path / "res.frs"))
    assert rd.nsteps == ns
    o = [oracle.calc_stresses(b, SV[s]) for s in range(ns)]
    sc = np.abs(np.stack([x["stress"] for x in o])).max()
    for e in (0, sam.nel // 2, sam.nel - 1):
        t = int(sam.melcon[e])
        name, nn = ("QUAD4", 4) if t == 24 else ("TRI3", 3)
        pt = b["ptoff"][e] + nn          # first bottom point
        got = rd.read(rd.find(f"Elements|{part.elm.elmid[e]}|{name}|Element nodes|Bottom|1|Stress", "Part", base))
        assert np.abs(got - np.stack([x["stress"][pt, :3] for x in o])).max() <= TOL * sc
        got = rd.read(rd.find(f"Elements|{part.elm.elmid[e]}|{name}|Element nodes|Bottom|1|Von Mises stress", "Part", base))
        assert np.abs(got[:, 0] - np.array([x["resmat"][pt, 0] for x in o])).max() <= TOL * sc
    n = 7
    j0 = sam.madof[n] - 1
    got = rd.read(rd.find(f"Nodes|{sam.minex[n]}|Dynamic response|Translational deformation", "Part", base))
    assert np.abs(got - SV[:, j0:j0 + 3]).max() <= TOL * np.abs(SV).max()
    rd.close()
    # the library call behind it, with the history and the envelope
    from fedem_solvers_b200 import _lib
    part.B = part.E = None
    rec = StressRecovery(part)
    vm = np.zeros((ns, rec.npts))
    _lib.check(rec._lib.fsr_recover_displacements(rec._h, np.ascontiguousarray(SV).ctypes.data_as(_lib._D), ns, vm.ctypes.data_as(_lib._D)),
               "fsr_recover_displacements")
    want = np.stack([x["resmat"][:, 0] for x in o])
    assert np.abs(vm - want).max() <= TOL * np.abs(want).max()
    mx, mn = rec.envelope()
    assert np.abs(mx - want.max(0)).max() <= TOL * np.abs(want).max()
    rec.close()


def _ortho_norm3(M):
    """orthoNorm3 = mat_to_quat + quat_to_mat (rotationModule.f90:441-497,549-558), restated for the check"""
    q = np.zeros(5)
    tr = M[0, 0] + M[1, 1] + M[2, 2]
    imax = int(np.argmax([M[0, 0], M[1, 1], M[2, 2]])) + 1
    m = lambda i, j: M[i - 1, j - 1]
    if tr > m(imax, imax):
        q[1] = np.sqrt(1.0 + tr) * 0.5
        q[2], q[3], q[4] = (m(3, 2) - m(2, 3)) / (4 * q[1]), (m(1, 3) - m(3, 1)) / (4 * q[1]), (m(2, 1) - m(1, 2)) / (4 * q[1])
    else:
        i, j, k = imax, imax % 3 + 1, (imax + 1) % 3 + 1
        q[i + 1] = np.sqrt(m(i, i) * 0.5 + (1.0 - tr) * 0.25)
        q[1] = (m(k, j) - m(j, k)) / (4 * q[i + 1])
        q[j + 1] = (m(j, i) + m(i, j)) / (4 * q[i + 1])
        q[k + 1] = (m(k, i) + m(i, k)) / (4 * q[i + 1])
    q /= np.sqrt((q[1:] ** 2).sum())
    R = np.empty((3, 3))
    R[0, 0] = 2 * (q[2] * q[2] + q[1] * q[1]) - 1; R[1, 1] = 2 * (q[3] * q[3] + q[1] * q[1]) - 1; R[2, 2] = 2 * (q[4] * q[4] + q[1] * q[1]) - 1
    R[0, 1] = 2 * (q[2] * q[3] - q[4] * q[1]); R[0, 2] = 2 * (q[2] * q[4] + q[3] * q[1]); R[1, 2] = 2 * (q[3] * q[4] - q[2] * q[1])
    R[1, 0] = 2 * (q[3] * q[2] + q[4] * q[1]); R[2, 0] = 2 * (q[4] * q[2] - q[3] * q[1]); R[2, 1] = 2 * (q[4] * q[3] + q[2] * q[1])
    return R


def test_fedem_gage_old_rosette_definition_file(oracle, tmp_path):
    """-rosfile in the old free-format layout (ReadStrainGageOldData, strainGageModule.f90:246-476): id type link nnod nodes zPos
    X Z Emod nu per rosette, '#' comments, END; rosettes of other links are skipped, the position matrix is recomputed from the
    element (centroid, element Z, X projected into the plane, orthoNorm3), dummy base ids idIn + 1000 link + 1e7 rdbinc, and
    the results database carries no rosette displacement state."""
    from fedem_solvers_b200.model import rosettes_on_part
    part = plate_part(6, 5, ngen=4, seed=41, tri_fraction=0.3, warp=0.02, n_ext=4)
    part.sam.minex = (500 + 2 * np.arange(part.sam.nnod)).astype(np.int32)
    nsteps = 60
    case = _make_case(tmp_path, part, "plate", nsteps=nsteps)
    ros = rosettes_on_part(part, 2, seed=42, rtype="TRIPLE_GAGE_45") + rosettes_on_part(part, 1, seed=43, rtype="TRIPLE_GAGE_60") \
        + rosettes_on_part(part, 1, seed=44, rtype="DOUBLE_GAGE_90") + rosettes_on_part(part, 1, seed=45, rtype="SINGLE_GAGE")
    tcode = {"SINGLE_GAGE": 1, "DOUBLE_GAGE_90": 2, "TRIPLE_GAGE_60": 3, "TRIPLE_GAGE_45": 4}
    lines = ["# id type link nnod nodes... zPos Xx Xy Xz Zx Zy Zz Emod nu"]
    given_x = []
    for k, r in enumerate(ros):
        x = 2.5 * r.rpos[:, 0] + (0.05 * r.rpos[:, 2] if k == 1 else 0.0)      # not unit; one with an out-of-plane part
        z = 0.7 * r.rpos[:, 2]
        given_x.append(x / np.linalg.norm(x))
        ext = [int(part.sam.minex[n - 1]) for n in r.nodes]
        lines.append(f"{900 + k} {tcode[r.type]} 2 {len(ext)} " + " ".join(map(str, ext)) + f" {r.zpos:.12e}  # comment")
        lines.append("   " + " ".join(f"{v:.15e}" for v in x) + "   " + " ".join(f"{v:.15e}" for v in z) + f" {r.emod:.6e} {r.nu:.4f}")
        if k == 0:   # a rosette of another part in between
            lines.append(f"77 1 5 3 {ext[0]} {ext[1]} {ext[2]} 0.0 1 0 0 0 0 1 2.1D11 0.3")
    lines.append("END")
    lines.append("this is never read")
    (tmp_path / "rosettes.dat").write_text("\n".join(lines) + "\n")
    exe = os.path.join(os.path.dirname(EXE), "fedem_gage")
    r = subprocess.run([exe, "-cwd", str(tmp_path), "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-Bmatfile", "plate_B.fmx",
                        "-eigfile", "plate_E.fmx", "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs", "-rosfile", "rosettes.dat",
                        "-rdbfile", "gage.frs", "-rdbinc", "1", "-stotm", "100", "-deformation"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Number of strain rosettes on this link =     5" in r.stdout
    rd = FrsReader(str(tmp_path / "gage_1.frs"))
    assert rd.nsteps == nsteps
    b = oracle.bind_part(part)
    for k, ro in enumerate(ros):
        base_id = (k + 1) + 2 * 1000 + 1 * 10000000
        # the position matrix the program must have built
        Z = ro.rpos[:, 2]
        Y = np.cross(Z, given_x[k]); X = np.cross(Y, Z)
        P = _ortho_norm3(np.stack([X, Y, Z], 1))
        want = type(ro)(**{**ro.__dict__})
        want.rpos = np.concatenate([P, ro.rpos[:, 3:4]], 1)
        Vo = oracle.rosette_history(b, want, case["Q"])
        ng = ro.to_c().ngage
        sc_e, sc_s = np.abs(Vo[:, :3]).max(), np.abs(Vo[:, 10:13]).max()
        h = rd.find("Strain tensor", "Strain rosette", base_id)
        assert h is not None, (k, base_id)
        eps = rd.read(h)
        wante = Vo[:, :3].copy(); wante[:, 2] *= 0.5
        assert np.abs(eps - wante).max() <= 1.3e-7 * sc_e, k
        assert np.abs(rd.read(rd.find("Stress tensor", "Strain rosette", base_id)) - Vo[:, 10:13]).max() <= 1.3e-7 * sc_s
        for j in range(ng):
            assert np.abs(rd.read(rd.find(f"Gage {j + 1}|Gage strain", "Strain rosette", base_id))[:, 0] - Vo[:, 18 + j]).max() <= 1.3e-7 * sc_e
        assert rd.find("Position", "Strain rosette", base_id) is None and rd.find("Euler angles", "Strain rosette", base_id) is None
    assert rd.find("Strain tensor", "Strain rosette", 6 + 2000 + 10000000) is None
