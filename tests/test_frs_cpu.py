"""CPU tests of the .frs layer (csrc/io_frs.cu): the reader is checked value by value against the
reference's OWN FFrLib reader (compiled unmodified into oracle/_ref/libfedem_ref_frs.so by oracle/Makefile)
on the reference's .frs fixtures (fedem-foundation/src/FFrLib/FFrTests, read in place when /root/reference
is present) and on files produced by this repo's writer; readSupElDisplacements + BuildFinit over a
window of steps is checked against a numpy restatement of supElTypeModule.f90:1067-1114."""
import ctypes as C
import glob
import os
import re
import numpy as np
import pytest

from fedem_solvers_b200.frs import FrsReader, FrsWriter, solver_header
from fedem_solvers_b200 import FsrError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libfedem_ref_frs.so")
FIXTURES = "/root/reference/fedem-foundation/src/FFrLib/FFrTests"
_D = C.POINTER(C.c_double)


class RefFrs:
    """ctypes face of oracle/ref_shim_frs.cpp (the reference's FFrExtractor)."""

    def __init__(self, paths):
        if not os.path.exists(REF_LIB):
            pytest.skip("oracle/_ref/libfedem_ref_frs.so not built (reference sources absent)")
        self.lib = C.CDLL(REF_LIB)
        self.lib.ref_frs_open.restype = C.c_void_p
        self.lib.ref_frs_open.argtypes = [C.POINTER(C.c_char_p), C.c_int]
        self.lib.ref_frs_close.argtypes = [C.c_void_p]
        self.lib.ref_frs_keys.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, _D, C.c_int]
        self.lib.ref_frs_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, _D, C.c_int, C.c_int, _D]
        self.arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        self.n = len(paths)
        self.h = self.lib.ref_frs_open(self.arr, self.n)
        assert self.h, "the reference reader rejected the files"

    def keys(self):
        n = self.lib.ref_frs_keys(self.h, self.arr, self.n, None, 0)
        k = np.zeros(max(n, 1))
        self.lib.ref_frs_keys(self.h, self.arr, self.n, k.ctypes.data_as(_D), n)
        return k[:n]

    def read(self, path, og, base, keys, nw):
        out = np.full((len(keys), nw), np.nan)
        k = np.ascontiguousarray(keys, np.float64)
        ok = self.lib.ref_frs_read(self.h, path.encode(), og.encode(), base, k.ctypes.data_as(_D), len(k), nw,
                                   out.ctypes.data_as(_D))
        return ok, out

    def close(self):
        self.lib.ref_frs_close(self.h)


def _header_text(path):
    b = open(path, "rb").read()
    i = b.find(b"\nDATA:")
    return b[:i + 1].decode("latin1"), len(b), i + 6


def _fixture_files():
    return sorted(glob.glob(os.path.join(FIXTURES, "**", "*.frs"), recursive=True))


@pytest.mark.skipif(not os.path.isdir(FIXTURES), reason="reference fixtures not present")
def test_reader_matches_reference_reader_on_reference_fixtures():
    files = _fixture_files()
    assert len(files) >= 10
    checked = 0
    for f in files:
        hdr, _, _ = _header_text(f)
        ours = FrsReader(f)
        ref = RefFrs([f])
        keys = ref.keys()
        assert np.array_equal(ours.times, keys), f
        # plain object groups {"Type";base;user;"descr";<v>...}: every directly held variable
        var_names = {int(m.group(1)): (m.group(2), m.group(3)) for m in
                     re.finditer(r'<\s*(\d+);"([^"]*)";[^;]*;[A-Z]+;\d+;[A-Z0-9]+(?:;\(([\d,]+)\))?', hdr)}
        objs = re.findall(r'\{"([^"]+)";(\d+);(\d+);"[^"]*";((?:<\s*\d+\s*>)+)\}', hdr)
        for og, base, _, refs in objs[:60]:
            for vid in re.findall(r"<\s*(\d+)\s*>", refs):
                name, dims = var_names[int(vid)]
                nw = int(np.prod([int(x) for x in dims.split(",")])) if dims else 1
                h = ours.find(name, og, int(base))
                assert h is not None, (f, og, base, name)
                assert ours.var_size(h) == nw
                a = ours.read(h)
                ok, b = ref.read(name, og, int(base), keys, nw)
                assert ok == len(keys)
                assert np.array_equal(a, b), (f, og, base, name)
                checked += 1
        ref.close()
        ours.close()
    assert checked > 100


@pytest.mark.skipif(not os.path.isdir(FIXTURES), reason="reference fixtures not present")
def test_nested_item_groups_of_a_fedem_stress_file():
    """Boom_1.frs is fedem_stress output (Part -> Nodes|n|Dynamic response|..., Elements|e|TRI3|Element
    nodes|Top|k|Von Mises stress): nested, referenced and inlined item groups, FLOAT 32 data."""
    f = os.path.join(FIXTURES, "response_0001/timehist_rcy_0001/2_Boom_0001/Boom_1.frs")
    hdr, fsize, hsize = _header_text(f)
    ours, ref = FrsReader(f), RefFrs([f])
    keys = ref.keys()
    assert ours.nsteps == len(keys) > 0
    base = int(re.search(r'\{"Part";(\d+);', hdr).group(1))
    nodes = [int(x) for x in re.findall(r"\[;\s*(\d+);\[\s*1\]\]", hdr)]
    elems = [int(x) for x in re.findall(r"\[;\s*(\d+);\[\s*2\]\]", hdr)]
    assert len(nodes) > 1000 and len(elems) > 1000
    rng = np.random.default_rng(0)
    n = 0
    for nd in list(rng.choice(nodes, 12, replace=False)) + [nodes[0], nodes[-1]]:
        for var in ("Translational deformation", "Angular deformation"):
            p = f"Nodes|{nd}|Dynamic response|{var}"
            h = ours.find(p, "Part", base)
            assert h is not None, p
            ok, b = ref.read(p, "Part", base, keys, 3)
            assert ok == len(keys) and np.array_equal(ours.read(h), b), p
            n += 1
    for el in list(rng.choice(elems, 12, replace=False)) + [elems[0], elems[-1]]:
        for side in ("Top", "Bottom"):
            for k in (1, 2, 3):
                for var, nw in (("Stress", 3), ("Von Mises stress", 1)):
                    p = f"Elements|{el}|TRI3|Element nodes|{side}|{k}|{var}"
                    h = ours.find(p, "Part", base)
                    assert h is not None, p
                    ok, b = ref.read(p, "Part", base, keys, nw)
                    assert ok == len(keys) and np.array_equal(ours.read(h), b), p
                    n += 1
    assert n == 28 + 14 * 12
    # the record layout accounts for the whole file
    rec = (fsize - hsize) / ours.nsteps
    assert rec == int(rec)
    # unknown paths: like a null pointer from ffr_findptr
    assert ours.find("Nodes|1|Dynamic response|No such variable", "Part", base) is None
    assert ours.find("Position matrix", "Part", 99999) is None
    ref.close()


def _rot(rng, small=0.05):
    w = rng.normal(size=3) * small
    th = np.linalg.norm(w)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _build_finit_numpy(sup_tr, triad_ur, tr_undef, ndofs, first_dof, gen_ur, gen_first, ndim):
    """supElTypeModule.f90:1067-1114 restated with numpy (3x4 matrices as [3,4] arrays)."""
    ns = sup_tr.shape[0]
    Q = np.zeros((ndim, ns))
    for s in range(ns):
        R, x = sup_tr[s][:, :3], sup_tr[s][:, 3]
        for t in range(len(ndofs)):
            u = triad_ur[s, t]
            ul = np.concatenate([R.T @ u[:, :3], (R.T @ (u[:, 3] - x))[:, None]], axis=1)
            j = first_dof[t] - 1
            Q[j:j + 3, s] = ul[:, 3] - tr_undef[t][:, 3]
            if ndofs[t] >= 6:
                dR = ul[:, :3] @ tr_undef[t][:, :3]      # matmul(urLocal(:,1:3),TrUndeformed(:,1:3,i)), :1096
                Q[j + 3:j + 6, s] = (dR[2, 1], dR[0, 2], dR[1, 0])
        if gen_ur is not None:
            Q[gen_first - 1:gen_first - 1 + gen_ur.shape[1], s] = gen_ur[s]
    return Q


def _write_solver_file(path, rng, nsteps, ntriads, ngen, t0=0.0, dt=0.01, step0=0, sup_base=7, with_other_part=True):
    triads = [(11 + i, 1 + i, f"triad {i}") for i in range(ntriads)]
    parts = [(sup_base, 2, "Flexible; part", ngen)] + ([(sup_base + 1, 3, "other", 0)] if with_other_part else [])
    text, nbytes = solver_header(triads, parts)
    tr_undef = np.zeros((ntriads, 3, 4))
    for t in range(ntriads):
        tr_undef[t][:, :3] = _rot(rng, 1.0)
        tr_undef[t][:, 3] = rng.normal(size=3)
    sup = np.zeros((nsteps, 3, 4))
    tri = np.zeros((nsteps, ntriads, 3, 4))
    gen = rng.normal(size=(nsteps, ngen)) * 1e-3
    with FrsWriter(path, text, nbytes) as w:
        for s in range(nsteps):
            sup[s][:, :3] = _rot(rng, 0.8)
            sup[s][:, 3] = rng.normal(size=3) * 3
            rec = []
            for t in range(ntriads):
                # triad position = part position . undeformed . small deformation
                d = np.zeros((3, 4)); d[:, :3] = _rot(rng, 0.01); d[:, 3] = rng.normal(size=3) * 1e-3
                loc = np.zeros((3, 4)); loc[:, :3] = d[:, :3] @ tr_undef[t][:, :3].T; loc[:, 3] = tr_undef[t][:, 3] + d[:, 3]
                tri[s, t][:, :3] = sup[s][:, :3] @ loc[:, :3]
                tri[s, t][:, 3] = sup[s][:, :3] @ loc[:, 3] + sup[s][:, 3]
                rec.append(tri[s, t].T.ravel())      # column-major 3x4
            rec.append(sup[s].T.ravel())
            rec.append(gen[s])
            if with_other_part:
                rec.append(np.arange(12.0))
            w.write_step(step0 + s, t0 + dt * s, np.concatenate(rec))
    return triads, tr_undef, sup, tri, gen


def test_writer_reader_roundtrip_and_reduced_history(tmp_path):
    rng = np.random.default_rng(11)
    ntriads, ngen, ns = 5, 7, 23
    f1 = str(tmp_path / "th_p_1.frs")
    triads, tr_undef, sup, tri, gen = _write_solver_file(f1, rng, ns, ntriads, ngen)
    rd = FrsReader(f1)
    assert rd.nsteps == ns
    assert np.array_equal(rd.step_numbers, np.arange(ns))
    assert np.allclose(rd.times, 0.01 * np.arange(ns), rtol=0, atol=1e-15)
    # raw values, bit-exact, against what was written and against the reference's reader
    ref = RefFrs([f1])
    keys = ref.keys()
    assert np.array_equal(keys, rd.times)
    for t, (b, _, _) in enumerate(triads):
        h = rd.find("Position matrix", "Triad", b)
        a = rd.read(h)
        assert np.array_equal(a, np.swapaxes(tri[:, t], -1, -2).reshape(ns, 12))
        ok, r = ref.read("Position matrix", "Triad", b, keys, 12)
        assert ok == ns and np.array_equal(a, r)
    hg = rd.find("Generalized displacement", "Part", 7)
    assert rd.var_size(hg) == ngen and np.array_equal(rd.read(hg), gen)
    ok, r = ref.read("Generalized displacement", "Part", 7, keys, ngen)
    assert ok == ns and np.array_equal(r, gen)
    ref.close()
    # readSupElDisplacements + BuildFinit over windows
    ndofs = np.array([6, 6, 3, 6, 6]); first = np.array([1, 7, 13, 16, 22]); ndim = 27 + ngen
    # a 3-DOF triad is stored under "Position" by the solver; this file has none -> expect the error
    with pytest.raises(FsrError, match="Triad"):
        rd.reduced_history(7, [t[0] for t in triads], ndofs, first, tr_undef, ngen, 28)
    ndofs = np.array([6, 6, 6, 6, 6]); first = np.array([1, 7, 13, 19, 25]); ndim = 30 + ngen
    Q = rd.reduced_history(7, [t[0] for t in triads], ndofs, first, tr_undef, ngen, 31)
    Qn = _build_finit_numpy(sup, tri, tr_undef, ndofs, first, gen, 31, ndim)
    assert Q.shape == (ndim, ns)
    assert np.abs(Q - Qn).max() <= 1e-14 * max(1.0, np.abs(Qn).max())
    Qw = rd.reduced_history(7, [t[0] for t in triads], ndofs, first, tr_undef, ngen, 31, step0=5, nsteps=9)
    assert np.array_equal(Qw, Q[:, 5:14])
    # the deformational displacements are small although the part moves by O(1)
    assert np.abs(Q[:30]).max() < 0.1
    with pytest.raises(FsrError, match="Part"):
        rd.reduced_history(99, [t[0] for t in triads], ndofs, first, tr_undef, ngen, 31)
    rd.close()


def test_multiple_files_are_merged_on_the_time_key(tmp_path):
    """The solver continues in a new file after a restart / size limit: same variables, later times."""
    rng = np.random.default_rng(5)
    f1, f2 = str(tmp_path / "th_p_1.frs"), str(tmp_path / "th_p_2.frs")
    a = _write_solver_file(f1, rng, 10, 2, 3, t0=0.0)
    b = _write_solver_file(f2, rng, 6, 2, 3, t0=0.1, step0=10)
    rd = FrsReader([f2, f1])
    assert rd.nsteps == 16 and np.all(np.diff(rd.times) > 0)
    assert np.array_equal(rd.step_numbers, np.arange(16))
    h = rd.find("Generalized displacement", "Part", 7)
    assert np.array_equal(rd.read(h), np.concatenate([a[4], b[4]]))
    ref = RefFrs([f1, f2])
    ok, r = ref.read("Generalized displacement", "Part", 7, rd.times, 3)
    assert ok == 16 and np.array_equal(r, rd.read(h))
    ref.close()
    rd.close()


def test_malformed_files_are_rejected(tmp_path):
    p = tmp_path / "bad.frs"
    p.write_bytes(b"#FEDEM response data          " + b"\x34\x12" + b"\0" * 8 + b";1.0;\nVARIABLES:\n<1;\"x\";NONE;INT;32;NUMBER>\n")
    with pytest.raises(FsrError, match="DATA"):
        FrsReader(str(p))
    p.write_bytes(b"#FEDEM disk matrix            " + b"\x34\x12" + b"\0" * 8 + b";1.0;\n")
    with pytest.raises(FsrError, match="not a results database"):
        FrsReader(str(p))
    with pytest.raises(FsrError, match="cannot open"):
        FrsReader(str(tmp_path / "missing.frs"))


def test_big_endian_file_is_swapped(tmp_path):
    """Files written on the other endianness (0x1234 mark reversed) are byte-swapped on read."""
    text, nbytes = solver_header([(11, 1, "t")], [(7, 2, "p", 2)])
    rng = np.random.default_rng(2)
    vals = rng.normal(size=(4, nbytes // 8))
    p = tmp_path / "be.frs"
    with open(p, "wb") as f:
        f.write(b"#FEDEM response data".ljust(30) + b"\x12\x34" + b"\0" * 8 + b";1.0;\n" + text.encode() + b"DATA:")
        for s in range(4):
            f.write(np.array([s], ">i4").tobytes() + np.array([0.5 * s], ">f8").tobytes() + vals[s].astype(">f8").tobytes())
    rd = FrsReader(str(p))
    assert np.array_equal(rd.times, 0.5 * np.arange(4)) and np.array_equal(rd.step_numbers, np.arange(4))
    assert np.array_equal(rd.read(rd.find("Position matrix", "Triad", 11)), vals[:, :12])
    assert np.array_equal(rd.read(rd.find("Generalized displacement", "Part", 7)), vals[:, 24:26])
    rd.close()


def test_eigenvector_groups_of_a_real_solver_file_and_mode_finit():
    """fedem_modes' inputs: readModesPointers asks for the ITEM GROUP "Eigenvectors|Mode  n" of every triad and reads it as one
    array (translational + angular components); on the reference's own modal results file the product reader must return what
    the reference's FFrExtractor returns.  fsr_build_mode_finit (readSupElModes) then turns the components into the part system."""
    import ctypes as C
    path = os.path.join(FIXTURES, "response_0001", "eigval_0001", "ev_p_3.frs")
    if not os.path.exists(path):
        pytest.skip("reference fixtures not present")
    rd = FrsReader(path)
    assert rd.nsteps == 2
    ref = RefFrs([path]) if os.path.exists(REF_LIB) else None
    keys = ref.keys() if ref else None
    triads = [11, 12, 20, 48]
    eig = {}
    for mode in (1, 6):
        for t in triads:
            h = rd.find(f"Eigenvectors|Mode{mode:3d}", "Triad", t)
            assert h is not None
            got = rd.read(h)
            assert got.shape == (2, 6)
            tr = rd.read(rd.find(f"Eigenvectors|Mode{mode:3d}|Translational deformation", "Triad", t))
            ro = rd.read(rd.find(f"Eigenvectors|Mode{mode:3d}|Angular deformation", "Triad", t))
            assert np.array_equal(got, np.hstack([tr, ro]))
            if ref:
                ok, want = ref.read(f"Eigenvectors|Mode{mode:3d}", "Triad", t, keys, 6)
                assert ok == 2 and np.array_equal(want, got)
            eig[(mode, t)] = got
    assert rd.find("Eigenvectors|Mode  7", "Triad", 11) is None
    freq = rd.read(rd.find("Eigenvalues|Mode  1|Eigenfrequency", "Mechanism", 2))
    assert freq.shape == (2, 1) and freq[0, 0] > 0
    if ref:
        ref.close()
    # readSupElModes on the first time step of mode 1: four 6-DOF triads + 3 generalized DOFs, two components (damped form)
    from fedem_solvers_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(7)
    from scipy.spatial.transform import Rotation
    sup = np.hstack([Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix(), [[1.0], [2.0], [3.0]]])
    ndofs = np.array([6, 6, 3, 6], np.int32)
    first = np.array([1, 7, 13, 16], np.int32)
    ncomp, ngen, gfirst = 2, 3, 22
    vecs = [np.concatenate([eig[(1, t)][0][:n], eig[(6, t)][1][:n]]) for t, n in zip(triads, ndofs)]      # component 1, component 2
    tri = np.ascontiguousarray(np.concatenate(vecs))
    gen = rng.standard_normal(ngen * ncomp)
    Q = np.zeros((24, ncomp), order="F")
    _dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib.fsr_build_mode_finit(4, _dp(np.ascontiguousarray(sup.T.reshape(-1))), _ip(ndofs), _ip(first), _dp(tri), ngen, gfirst, _dp(gen), ncomp, _dp(Q), 24)
    assert rc == 0
    tinv = sup[:, :3].T
    for l in range(ncomp):
        for v, n, k in zip(vecs, ndofs, first):
            e = v[n * l: n * (l + 1)]
            assert np.allclose(Q[k - 1:k + 2, l], tinv @ e[:3], rtol=0, atol=1e-16)
            if n == 6:
                assert np.allclose(Q[k + 2:k + 5, l], tinv @ e[3:6], rtol=0, atol=1e-16)
        assert np.array_equal(Q[gfirst - 1:gfirst - 1 + ngen, l], gen[ngen * l: ngen * (l + 1)])
    assert lib.fsr_build_mode_finit(4, _dp(np.ascontiguousarray(sup.T.reshape(-1))), _ip(ndofs), _ip(first), _dp(tri), ngen, 23, _dp(gen), ncomp, _dp(Q), 24) < 0


def test_no_system_level_response_is_a_warning_not_an_error(tmp_path):
    """readResponsePointers (displacementModule.f90:292-303): when NONE of the position / generalized-displacement variables
    is on the results files the reference warns and recovers from local deformations relative to the modelling configuration
    (finit = 0); only a partial set is an error"""
    rng = np.random.default_rng(3)
    f1 = str(tmp_path / "th_p_1.frs")
    triads, tr_undef, sup, tri, gen = _write_solver_file(f1, rng, 6, 3, 4)
    rd = FrsReader(f1)
    ndofs, first = np.array([6, 6, 6]), np.array([1, 7, 13])
    Q = rd.reduced_history(500, [901, 902, 903], ndofs, first, tr_undef, 4, 19)       # nothing of this part is on the file
    assert Q.shape == (22, 6) and not Q.any()
    with pytest.raises(FsrError, match="Triad"):                                       # the part is there, one triad is not
        rd.reduced_history(7, [triads[0][0], 902, triads[2][0]], ndofs, first, tr_undef, 4, 19)
    rd.close()
