"""Host logic of the fedem_stress driver (csrc/stress_driver.cu, cmdline.hpp, io_fsi.cu), no GPU:
 * the option parser against the reference's OWN FFaCmdLineArg (compiled into oracle/_ref/libfedem_ref_ffl.so)
   on hand-picked and fuzzed argument lists and option files;
 * the time-step selection against the Fortran ffr_getNextStep loop replayed over the reference's own
   FFrExtractor::positionRDB / incrementRDB (oracle/_ref/libfedem_ref_frs.so);
 * the solver-input (.fsi) reader on the reference's sample files and on files written by this repo;
 * the executable: option table, -help, and loud failures without inputs / without a GPU."""
import ctypes as C
import glob
import os
import subprocess
import numpy as np
import pytest

from fedem_solvers_b200 import _lib
from fedem_solvers_b200.frs import FrsWriter, solver_header
from fedem_solvers_b200.fsi import read_fsi, write_fsi, SolverPart

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FFL = os.path.join(ROOT, "oracle", "_ref", "libfedem_ref_ffl.so")
REF_FRS = os.path.join(ROOT, "oracle", "_ref", "libfedem_ref_frs.so")
EXE = os.path.join(ROOT, "fedem_solvers_b200", "bin", "fedem_stress")

OPTIONS = [("stress", "bool", False), ("stressForm", "int", 0), ("strain", "bool", False), ("SR", "bool", False),
           ("VTFavgelm", "bool", True), ("double", "bool", False), ("statm", "double", 0.0), ("stotm", "double", 1.0),
           ("tinc", "double", 0.1), ("linkfile", "string", ""), ("group", "string", ""), ("rdbinc", "int", 1),
           ("linkId", "int", 0), ("frsfile", "string", ""), ("fsifile", "string", "fedem_solver.fsi"), ("debug", "int", 0)]


def _argv(args):
    arr = (C.c_char_p * (len(args) + 1))(b"prog", *[a.encode() for a in args])
    return len(args) + 1, arr


class Mine:
    def __init__(self, args, files=()):
        self.L = _lib.load_library()
        self.L.fsr_cmdline_reset()
        n, arr = _argv(args)
        self.L.fsr_cmdline_init(n, arr)
        for name, typ, d in OPTIONS:
            getattr(self.L, f"fsr_cmdline_add_{typ}")(name.encode(), d.encode() if typ == "string" else d)
        self.files = files

    def values(self):
        out = {}
        first = True
        for name, typ, _ in OPTIONS:
            if typ == "string":
                b = C.create_string_buffer(512)
                self.L.fsr_cmdline_get_string(name.encode(), b, 512)
                v = b.value.decode()
            else:
                v = getattr(self.L, f"fsr_cmdline_get_{typ}")(name.encode())
            if first:   # option files are appended after the command line has been evaluated once
                for f in self.files:
                    self.L.fsr_cmdline_read_file(os.fsencode(f))
                first = False
            out[name] = (v, self.L.fsr_cmdline_is_set(name.encode()))
        return out


class Ref:
    def __init__(self, args, files=()):
        self.L = C.CDLL(REF_FFL)
        self.L.ref_cmdline_get_double.restype = C.c_double
        self.L.ref_cmdline_add_double.argtypes = [C.c_char_p, C.c_double]
        n, arr = _argv(args)
        self.L.ref_cmdline_init(n, arr)
        for name, typ, d in OPTIONS:
            if name in ("linkfile", "group"):
                continue   # the shim defines those itself
            getattr(self.L, f"ref_cmdline_add_{typ}")(name.encode(), d.encode() if typ == "string" else d)
        self.files = files

    def values(self):
        out = {}
        first = True
        for name, typ, _ in OPTIONS:
            if typ == "string":
                b = C.create_string_buffer(512)
                self.L.ref_cmdline_get_string(name.encode(), b, 512)
                v = b.value.decode()
            else:
                v = getattr(self.L, f"ref_cmdline_get_{typ}")(name.encode())
            if first:
                for f in self.files:
                    self.L.ref_cmdline_read_file(os.fsencode(f))
                first = False
            out[name] = (v, self.L.ref_cmdline_is_set(name.encode()))
        return out


CASES = [
    ["-stress", "-stressForm", "2", "-linkfile", "part one.ftl", "-statm", "-0.5"],
    ["-stressForm=1", "-strain-", "-VTFavgelm-", "-group", "<1,", "2,", "PMAT", "3>", "-double"],
    ["-STRESSFORM", "1", "-Stress", "-tinc0.25", "-rdbinc", "7", "-rdbinc", "9"],
    ["-linkfile", '"quoted name.ftl"', "-stotm", "1e-3", "-unknownOption", "3", "-SR+", "-debug", "abc"],
    ["-frsfile", "<a.frs,", "b.frs>", "-statm", "-1", "-2", "-linkId", "-4"],
    ["stray", "-stress=+", "-strain=-", "-double", "x", "-fsifile", "model.fsi"],
    [],
]


@pytest.mark.skipif(not os.path.exists(REF_FFL), reason="oracle/_ref/libfedem_ref_ffl.so not built")
@pytest.mark.parametrize("args", CASES, ids=[str(i) for i in range(len(CASES))])
def test_option_parser_matches_reference_parser(args, capfd):
    assert Mine(args).values() == Ref(args).values()


@pytest.mark.skipif(not os.path.exists(REF_FFL), reason="oracle/_ref/libfedem_ref_ffl.so not built")
def test_option_parser_fuzz_and_option_files(tmp_path, capfd):
    rng = np.random.default_rng(1)
    names = [o[0] for o in OPTIONS]
    vals = ["1", "-2", "0.5", "-1e-2", "+", "-", "abc", '"a b"', "<1,2>", "x.ftl", "", "3x", "=4"]
    for trial in range(300):
        args = []
        for _ in range(rng.integers(0, 7)):
            n = names[rng.integers(len(names))]
            form = rng.integers(0, 5)
            if form == 0:
                n = n.upper() if rng.random() < 0.5 else n.lower()
            v = vals[rng.integers(len(vals))]
            if form <= 1:
                args += ["-" + n] + ([v] if v and rng.random() < 0.8 else [])
            elif form == 2:
                args.append(f"-{n}={v}")
            elif form == 3:
                args.append(f"-{n}{v}")
            else:
                args += ["-" + n, v, vals[rng.integers(len(vals))]]
        args = [a for a in args if a]
        files = []
        if trial % 3 == 0:
            f = str(tmp_path / f"opt{trial}.fco")
            with open(f, "w") as fh:
                fh.write("# calculation options\n-tinc 0.01 -statm 2   # trailing comment\n-linkfile \"my part.ftl\"\n"
                         f"-stressForm {trial % 3}\n-group <1, 2>\n-double")
            files = [f]
        a, b = Mine(args, files).values(), Ref(args, files).values()
        assert a == b, (args, files)


def _ref_steps(path, start, stop, tinc):
    """ffr_getNextStep (FFrExtractorInterface.f90:134-170) replayed over the reference's extractor."""
    from test_frs_cpu import RefFrs
    ref = RefFrs([path])
    ref.lib.ref_frs_setposition.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    ref.lib.ref_frs_increment.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    tol, huge = 1e-12, np.finfo(float).max
    curr, last, out = start - 1.0, -huge, []
    bt = C.c_double()
    while True:
        if curr > stop - tol:
            break
        if curr < start - tol:
            rc = ref.lib.ref_frs_setposition(ref.h, start, C.byref(bt)); curr = bt.value
            start, last = curr, -huge
        elif tinc < tol:
            rc = ref.lib.ref_frs_increment(ref.h, C.byref(bt)); curr = bt.value
        else:
            rc = ref.lib.ref_frs_setposition(ref.h, curr + tinc, C.byref(bt)); curr = bt.value
        ok = curr < stop + tol and rc >= 0 and curr > last + tol
        last = curr
        if not ok:
            break
        out.append(curr)
    ref.close()
    return out


@pytest.mark.skipif(not os.path.exists(REF_FRS), reason="oracle/_ref/libfedem_ref_frs.so not built")
def test_time_step_selection_matches_reference_loop(tmp_path):
    lib = _lib.load_library()
    rng = np.random.default_rng(2)
    times = np.round(np.cumsum(rng.choice([0.001, 0.0025, 0.01], 200)), 6) + 0.1
    text, nb = solver_header([(5, 1, "t")], [])
    path = str(tmp_path / "th.frs")
    with FrsWriter(path, text, nb) as w:
        for s, t in enumerate(times):
            w.write_step(s, t, np.zeros(12))
    dp = times.ctypes.data_as(C.POINTER(C.c_double))
    for start, stop, tinc in [(0.0, 1.0, 0.0), (0.0, 1.0, 0.1), (0.2, 0.5, 0.0), (0.25, 0.3, 0.004), (0.0, 100.0, 0.05),
                              (5.0, 6.0, 0.1), (float(times[17]), float(times[40]), 0.0), (0.1, 0.1, 0.0),
                              (float(times[3]), 0.9, 0.0071), (0.0, 0.05, 0.0), (0.3, 0.2, 0.0), (0.0, float(times[-1]), 1e-3)]:
        idx = np.zeros(len(times), np.int32)
        n = lib.fsr_select_steps(dp, len(times), start, stop, tinc, idx.ctypes.data_as(C.POINTER(C.c_int)), len(idx))
        want = _ref_steps(path, start, stop, tinc)
        assert n == len(want), (start, stop, tinc, n, len(want))
        assert np.array_equal(times[idx[:n]], np.array(want)), (start, stop, tinc)
    assert lib.fsr_select_steps(None, 0, 0.0, 1.0, 0.0, None, 0) == 0


def test_fsi_reader_on_written_file(tmp_path):
    rng = np.random.default_rng(3)
    def part(base, triads, ngen):
        n = len(triads)
        return SolverPart(base_id=base, user_id=base - 10, descr=f"Part {base}", ngen=ngen, sup_pos=rng.normal(size=(3, 4)),
                          gravity=np.zeros(3), model_file="", triad_base_id=np.array(triads), triad_user_id=np.arange(1, n + 1),
                          ndofs=np.array([3 if t == 105 else 6 for t in triads]), first_dof=np.zeros(n, int), tr_undef=rng.normal(size=(n, 3, 4)),
                          triad_ur=rng.normal(size=(n, 3, 4)), gen_first_dof=0)
    a, b = part(21, [101, 105, 103], 4), part(22, [105, 107], 0)
    b.triad_ur[0] = a.triad_ur[1]   # triad 105 is shared by the two parts: one &TRIAD record
    f = str(tmp_path / "fedem_solver.fsi")
    write_fsi(f, [a, b], gravity=(0.0, -9.81, 0.0), model_file=r"C:\models\crane.fmm")
    for want in (a, b):
        got = read_fsi(f, want.base_id)
        assert (got.user_id, got.descr, got.ngen, got.model_file) == (want.user_id, want.descr, want.ngen, r"C:\models\crane.fmm")
        assert np.array_equal(got.triad_base_id, want.triad_base_id) and np.array_equal(got.ndofs, want.ndofs)
        assert np.array_equal(got.first_dof, 1 + np.concatenate([[0], np.cumsum(want.ndofs)[:-1]]))
        assert got.gen_first_dof == 1 + want.ndofs.sum() and got.ndim == want.ndofs.sum() + want.ngen
        np.testing.assert_allclose(got.sup_pos, want.sup_pos, rtol=1e-9)
        np.testing.assert_allclose(got.tr_undef, want.tr_undef, rtol=1e-9)
        np.testing.assert_allclose(got.triad_ur, want.triad_ur, rtol=1e-9)
        assert np.array_equal(got.gravity, [0.0, -9.81, 0.0])
    with pytest.raises(_lib.FsrError, match="baseID 99 was not found"):
        read_fsi(f, 99)
    with pytest.raises(_lib.FsrError, match="Unable to open"):
        read_fsi(str(tmp_path / "none.fsi"), 1)


REF_FSI = "/root/reference/solverTests/InversePy/shell_strain/fedem_solver.fsi"


@pytest.mark.skipif(not os.path.exists(REF_FSI), reason="reference checkout not present")
def test_fsi_reader_on_reference_samples():
    p = read_fsi(REF_FSI, 16)
    assert (p.user_id, p.descr, p.ngen, p.ndim) == (1, "#recover-gages shell", 12, 36)
    assert list(p.triad_base_id) == [19, 21, 20, 22] and list(p.triad_user_id) == [1, 3, 2, 4] and list(p.first_dof) == [1, 7, 13, 19]
    assert np.array_equal(p.tr_undef[:, :, 3], [[0, 0, 0], [4, 0, 0], [0, 0.2, 0], [4, 0.2, 0]])
    assert p.model_file.endswith("s41_gage_twin.fmm") and np.array_equal(p.sup_pos, np.eye(3, 4))
    n = 0
    for f in glob.glob("/root/reference/solverTests/**/*.fsi", recursive=True):
        txt = open(f, errors="replace").read()
        import re
        for m in re.finditer(r"&SUP_EL\s+id\s*=\s*(\d+)", txt):
            q = read_fsi(f, int(m.group(1)))
            assert q.ndim >= 0 and len(q.triad_base_id) == len(q.ndofs)
            n += 1
    assert n >= 5


def test_executable_options_and_loud_failures(tmp_path):
    assert os.path.exists(EXE), "build.sh did not produce bin/fedem_stress"
    r = subprocess.run([EXE, "-help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    for opt in ("-linkfile", "-Bmatfile", "-eigfile", "-samfile", "-fsifile", "-frsfile", "-rdbfile", "-rdbinc", "-vmStress",
                "-maxPStrain", "-SR", "-deformation", "-double", "-group", "-statm", "-stotm", "-tinc", "-fco", "-fop"):
        assert opt + " " in r.stdout, opt
    assert "-stressForm" not in r.stdout   # private option, only with -helpAll
    r = subprocess.run([EXE, "-helpAll"], capture_output=True, text=True, timeout=60)
    assert "-stressForm" in r.stdout and "-ffqStressForm" in r.stdout
    r = subprocess.run([EXE, "-cwd", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "FE data file must be specified through -linkfile" in r.stdout
    assert os.path.exists(tmp_path / "fedem_stress.res")
    r = subprocess.run([EXE, "-cwd", str(tmp_path), "-linkfile", "nothere.ftl"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Can not open FE data file nothere.ftl" in r.stdout
    assert os.path.exists(tmp_path / "nothere_stress.res")


def test_rosette_file_reader(tmp_path):
    """ReadStrainGages: written &STRAIN_ROSETTE records come back value for value (10 digits); the reference's own sample
    (solverTests/InversePy/shell_strain/fedem_solver.fsi) parses; a record of another part or an unknown type is an error."""
    from fedem_solvers_b200.gage import Rosette, write_rosette_file, read_rosette_file
    rng = np.random.default_rng(5)
    ros = []
    for k, t in enumerate(["SINGLE_GAGE", "DOUBLE_GAGE_90", "TRIPLE_GAGE_60", "TRIPLE_GAGE_45"]):
        Qm, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        ros.append(Rosette(id=40 + k, nodes=[3 + k, 9, 12] + ([20] if k % 2 else []), rpos=np.hstack([Qm, rng.standard_normal((3, 1))]),
                           type=t, zpos=0.01 * k, emod=2.0e11 + k, nu=0.29, zero_init=bool(k & 1), gate=10.0 * k,
                           sncurve=[12.0, 15.0, 3.0, 5.0] if k == 2 else [0.0] * 4))
    path = str(tmp_path / "gages.fsi")
    write_rosette_file(path, ros, link_id=16, user_ids=[7, 8, 9, 10], descr=["a b", "", "third", "x"])
    back, uid, descr = read_rosette_file(path, 16)
    assert len(back) == 4 and list(uid) == [7, 8, 9, 10] and descr == ["a b", "", "third", "x"]
    for a, b in zip(ros, back):
        assert (a.id, a.nodes, a.type, a.zero_init) == (b.id, b.nodes, b.type, b.zero_init)
        assert np.allclose(a.rpos, b.rpos, rtol=0, atol=5e-10 * np.abs(a.rpos).max())
        assert abs(a.emod - b.emod) <= 1e-9 * a.emod and abs(a.nu - b.nu) < 1e-9 and abs(a.gate - b.gate) < 1e-8
        assert np.allclose(a.sncurve, b.sncurve)
    lib = _lib.load_library()
    assert lib.fsr_fsi_read_rosettes(path.encode(), 17, None, None, None, 0, 0) == 4      # counting does not look at the part
    arr = (_lib.FsrRosette * 4)()
    assert lib.fsr_fsi_read_rosettes(path.encode(), 17, arr, None, None, 0, 4) < 0
    assert b"does not match" in lib.fsr_last_error()
    open(path, "a").write("&STRAIN_ROSETTE\n id = 50\n linkId = 16\n type = 'QUAD_GAGE'\n numnod = 3\n nodes = 1 2 3\n rPos = 12*0.0\n/\n")
    arr = (_lib.FsrRosette * 5)()
    assert lib.fsr_fsi_read_rosettes(path.encode(), 16, arr, None, None, 0, 5) < 0
    assert b"invalid rosette-type" in lib.fsr_last_error()
    sample = "/root/reference/solverTests/InversePy/shell_strain/fedem_solver.fsi"
    if os.path.exists(sample):
        back, uid, descr = read_rosette_file(sample, 16)
        assert len(back) >= 1 and back[0].id == 44 and back[0].type == "SINGLE_GAGE" and back[0].nodes == [4, 5, 10, 9]
        assert descr[0] == "straingage_1" and uid[0] == 1
        assert np.allclose(back[0].rpos, [[1, 0, 0, 3.5], [0, 1, 0, 0.1], [0, 0, 1, 0.1]]) and back[0].zpos == 0.1


def test_gage_executable_options_and_loud_failures(tmp_path):
    exe = os.path.join(os.path.dirname(EXE), "fedem_gage")
    assert os.path.exists(exe), "build.sh did not produce bin/fedem_gage"
    r = subprocess.run([exe, "-help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    for opt in ("-rosfile", "-fatigue", "-gate", "-binSize", "-loga1", "-loga2", "-m1", "-stressToMPaScale", "-nullify_start_rosettestrains",
                "-linkfile", "-frsfile", "-rdbfile", "-tinc"):
        assert opt + " " in r.stdout, opt
    assert "-vmStress" not in r.stdout and "-group" not in r.stdout      # fedem_stress options are not fedem_gage options
    r = subprocess.run([exe, "-cwd", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "FE data file must be specified through -linkfile" in r.stdout and "Strain gage recovery failed" in r.stdout
    assert os.path.exists(tmp_path / "fedem_gage.res")


def test_modes_executable_options_and_loud_failures(tmp_path):
    exe = os.path.join(os.path.dirname(EXE), "fedem_modes")
    assert os.path.exists(exe), "build.sh did not produce bin/fedem_modes"
    r = subprocess.run([exe, "-help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    for opt in ("-recover_modes", "-damped", "-energy_density", "-write_vector", "-write_nodes", "-linkfile", "-frsfile", "-rdbfile", "-double"):
        assert opt + " " in r.stdout, opt
    assert "-vmStress" not in r.stdout and "-rosfile" not in r.stdout
    r = subprocess.run([exe, "-cwd", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "FE data file must be specified through -linkfile" in r.stdout and "Modal recovery failed" in r.stdout
    assert os.path.exists(tmp_path / "fedem_modes.res")
