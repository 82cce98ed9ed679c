"""The .ftl reader (csrc/io_ftl.cu) against the reference's OWN FE-model reader and Fortran accessor layer
(FFlLib + FFlLinkHandler_F.C compiled unmodified into oracle/_ref/libfedem_ref_ffl.so): every number
fedem_stress would get from ffl_getsize / ffl_getnodes / ffl_gettopol / ffl_getelmid / ffl_getcoor /
ffl_getmat / ffl_getthick / ffl_getbeamsection / ffl_getpinflags must be identical (bit-exact: both sides
run strtod on the same text and the same handful of additions)."""
import ctypes as C
import glob
import os
import numpy as np
import pytest

from fedem_solvers_b200 import _lib
from fedem_solvers_b200.ftl import FtlPart, write_ftl
from fedem_solvers_b200.model import plate_part, tet10_block, hex20_block

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfedem_ref_ffl.so")
I32, F64 = np.int32, np.float64
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libfedem_ref_ffl.so not built")


class RefFfl:
    """ctypes face of the reference's ffl_* Fortran entry points (all arguments by reference)."""

    def __init__(self, path, groups=""):
        self.L = C.CDLL(REF_SO)
        rc = self.L.ref_ffl_load(os.fsencode(path), groups.encode())
        if rc < 0:
            raise RuntimeError(f"reference failed to load {path}: {rc}")

    def close(self):
        self.L.ref_ffl_release()

    def sizes(self):
        v = [C.c_int() for _ in range(13)]
        self.L.ffl_getsize_(*[C.byref(x) for x in v])
        keys = ("nnod", "nel", "ndof", "nmnpc", "nmat", "nxnod", "npbeam", "nrgd", "nrbar", "nwavgm", "nprop", "ncons", "nael")
        return {k: x.value for k, x in zip(keys, v)}

    def nodes(self, s):
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        nnod, ndof, ierr = C.c_int(s["nnod"]), C.c_int(0), C.c_int(0)
        madof, minex, mnode = np.zeros(s["nnod"] + 1, I32), np.zeros(s["nnod"], I32), np.zeros(s["nnod"], I32)
        msc = np.zeros(s["ndof"], I32)
        X, Y, Z = (np.zeros(s["nnod"], F64) for _ in range(3))
        self.L.ffl_getnodes_(C.byref(nnod), C.byref(ndof), ip(madof), ip(minex), ip(mnode), ip(msc), dp(X), dp(Y), dp(Z),
                             C.byref(ierr))
        assert ierr.value == 0 and ndof.value == s["ndof"]
        return madof, minex, mnode, msc, np.stack([X, Y, Z], 1)

    def topology(self, s):
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        nel, nmnpc, ierr = C.c_int(0), C.c_int(0), C.c_int(0)
        mekn, mmnpc, mpmnpc = np.zeros(s["nel"], I32), np.zeros(max(s["nmnpc"], 1), I32), np.zeros(s["nel"] + 1, I32)
        z = lambda n: np.zeros(max(n, 1), I32)
        self.L.ffl_gettopol_(C.byref(nel), C.byref(nmnpc), ip(mekn), ip(mmnpc), ip(mpmnpc), ip(z(s["npbeam"])),
                             ip(z(s["nrgd"])), ip(z(s["nrbar"])), ip(z(s["nwavgm"])), C.byref(ierr))
        assert ierr.value == 0 and nel.value == s["nel"]
        return mekn, mpmnpc, mmnpc[:nmnpc.value]

    def element(self, iel, nenod, shell, beam):
        """(elmid, E, nu, rho, ierr_mat, thk, X, Y, Z, bsec, pins)"""
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        i = C.c_int(iel)
        self.L.ffl_getelmid_.restype = C.c_int
        eid = self.L.ffl_getelmid_(C.byref(i))
        E, nu, rho, ierr = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        self.L.ffl_getmat_(C.byref(E), C.byref(nu), C.byref(rho), C.byref(i), C.byref(ierr))
        out = dict(elmid=eid, E=E.value, nu=nu.value, rho=rho.value, ierr_mat=ierr.value)
        if shell:
            th = np.zeros(nenod, F64)
            self.L.ffl_getthick_(dp(th), C.byref(i), C.byref(ierr))
            out["thk"], out["ierr_thk"] = th, ierr.value
        if beam:
            X, Y, Z, sec = np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(14)
            self.L.ffl_getcoor_(dp(X), dp(Y), dp(Z), C.byref(i), C.byref(ierr))
            out["ierr_coor"] = ierr.value
            self.L.ffl_getbeamsection_(dp(sec), C.byref(i), C.byref(ierr))
            pa, pb = C.c_int(), C.c_int()
            self.L.ffl_getpinflags_(C.byref(pa), C.byref(pb), C.byref(i), C.byref(ierr))
            out.update(X=X, Y=Y, Z=Z, sec=sec, ierr_sec=ierr.value, pins=(pa.value, pb.value))
        return out


def compare_with_reference(path, groups=""):
    ref = RefFfl(path, groups)
    try:
        mine = FtlPart(path, groups)
        s_ref, s = ref.sizes(), mine.sizes()
        assert s == s_ref
        for a, b in zip(mine.nodes(), ref.nodes(s_ref)):
            assert a.shape == b.shape and (a == b).all()
        melcon, mpmnpc, mmnpc = mine.topology()
        mekn_r, mpmnpc_r, mmnpc_r = ref.topology(s_ref)
        assert (melcon == mekn_r).all() and (mpmnpc == mpmnpc_r).all() and (mmnpc == mmnpc_r).all()
        elm, rho, status = mine.element_data()
        for e in range(s["nel"]):
            t = int(melcon[e])
            structural = t in (11, 21, 22, 31, 32, 41, 42, 43, 44, 45, 46)
            r = ref.element(e + 1, int(mpmnpc[e + 1] - mpmnpc[e]), t in (21, 22, 31, 32), t == 11)
            assert r["elmid"] == elm.elmid[e]
            if not structural:
                continue
            if r["ierr_mat"] == 0:
                assert (r["E"], r["nu"], r["rho"]) == (elm.emod[e], elm.rny[e], rho[e])
            else:
                assert status[e] != 0
            if "thk" in r:
                assert r["ierr_thk"] == 0 and (r["thk"] == elm.thk[e]).all()
            if t == 11:
                b = elm.beam[e]
                assert (b[0:5] == r["X"]).all() and (b[5:10] == r["Y"]).all() and (b[10:15] == r["Z"]).all()
                if r["ierr_sec"] == 0:
                    assert (b[15:29] == r["sec"]).all()
                assert (int(b[29]), int(b[30])) == r["pins"]
        mine.close()
        return s
    finally:
        ref.close()


REF_FTL = sorted(glob.glob("/root/reference/solverTests/**/*.ftl", recursive=True))


@pytest.mark.skipif(not REF_FTL, reason="reference checkout not present")
@pytest.mark.parametrize("path", REF_FTL, ids=[os.path.basename(p) for p in REF_FTL])
def test_reference_sample_parts(path):
    s = compare_with_reference(path)
    assert s["nel"] > 0 and s["nnod"] > 0


def _roundtrip(part, tmp_path, name, groups=None, select=""):
    p = str(tmp_path / name)
    write_ftl(p, part, groups=groups)
    s = compare_with_reference(p, select)
    assert (s["nnod"], s["nel"], s["ndof"]) == (part.sam.nnod, part.sam.nel, part.sam.ndof)
    mine = FtlPart(p, select)
    elm, _, status = mine.element_data()
    melcon, mpmnpc, mmnpc = mine.topology(use_andes=True)
    madof, minex, _, msc, xyz = mine.nodes()
    assert (status == 0).all()
    assert (melcon == part.sam.melcon).all() and (mpmnpc == part.sam.mpmnpc).all() and (mmnpc == part.sam.mmnpc).all()
    assert (madof == part.sam.madof).all() and (xyz == part.elm.xyz).all()
    free = part.sam.meqn >= 0   # constraint-equation DOFs are SAM's business, not the file's
    assert (msc[free] == part.sam.msc[free]).all()
    solid_or_shell = part.sam.melcon != 11
    assert (elm.emod[solid_or_shell] == part.elm.emod[solid_or_shell]).all()
    assert (elm.rny[solid_or_shell] == part.elm.rny[solid_or_shell]).all() and (elm.thk == part.elm.thk).all()
    if part.elm.beam is not None:
        np.testing.assert_allclose(elm.beam, part.elm.beam, rtol=1e-13, atol=1e-15)
    return mine, elm


def test_generated_plate_with_triangles_and_groups(tmp_path):
    part = plate_part(9, 7, ngen=4, seed=11, tri_fraction=0.4, warp=0.03)
    groups = {5: list(range(1, 20)), 9: [30, 31, 32]}
    mine, elm = _roundtrip(part, tmp_path, "plate.ftl", groups, "<5, 9, 77>")
    assert mine.ignored_groups == 1   # group 77 does not exist: ignored with a message, like the reference
    active = set(range(1, 20)) | {30, 31, 32}
    assert {int(i) for i in elm.elmid if i > 0} == active and (np.abs(elm.elmid) == part.elm.elmid).all()


def test_generated_tet10_block_with_beams(tmp_path):
    part = tet10_block(3, 2, 2, ngen=4, seed=12, n_beams=9)
    _roundtrip(part, tmp_path, "tets.ftl")


def test_generated_thick_shell_panel(tmp_path):
    """TRI6 / QUAD8: the reference's ffl_gettopol moves the three mid-side nodes of a TRI6 last (FFlLinkHandler_F.C:657-664)"""
    from fedem_solvers_b200.model import thickshell_panel
    part = thickshell_panel(4, 3, ngen=3, seed=14, with_recovery=False)
    _roundtrip(part, tmp_path, "panel.ftl")


def test_generated_hex20_block_implicit_group(tmp_path):
    part = hex20_block(2, 2, 1, ngen=3, seed=13)
    mine, elm = _roundtrip(part, tmp_path, "hex.ftl", None, "<PMAT 1>")
    assert (elm.elmid > 0).all()


TRICKY = """FTLVERSION{7 ASCII}
# File checksum: 12345
# a hand-written part: obsolete keywords, comments inside records, Nastran-style exponents, loose nodes,
# constraint elements, a concentrated mass, strain coat elements and pinned/eccentric beams
NODE{1 1 0 0 0}
NODE{2 0 1.0 0 0} NODE{3 0 2.0 0 0}
node{4 0 2.0 1.0 0}   # lower-case label
NODE{5 0 1.0 1.0 0}
NODE{6 -7 0 1.0 0}            # x, y, z translations fixed
NODE{7 0 0.5 0.5 1.5-1}       # = 0.15, only referenced by the RGD: gets its DOFs from being the master
NODE{8 0 9 9 9}               # loose
NODE{9 1 3.0 0.5 0}           # external, only in the WAVGM
NODE{10 0 3.0 0.0 0}
NODE{11 0 5 5 5}              # WAVGM master without DOFs: dropped from the element
NODE{20 0 2.0 0.0 1.0}
FFQ4{3 1 2 5 6 {PTHICK 2} {PMAT 1}}
QUAD4{1 2 3 4 5 {PTHICK 1} {PMAT 1}
      {VDETAIL 3}}
FFT3{7 3 10 4 {PMAT 2} {PTHICK 1}}
BEAM2{12 3 20 {PMAT 2} {PBEAMSECTION 4} {PBEAMECCENT 1} {PBEAMPIN 1} {PORIENT 2}}
BEAM2{11 10 20 {PMAT 2} {PBEAMSECTION 4}}      # no orientation: globalized Z axis
BEAM2{13 20 4 {PMAT 2} {PBEAMSECTION 5} {PBEAMORIENT 2} {PEFFLENGTH 1}}
RGD{20 7 1 2 5}
WAVGM{21 9 3 4 11 10}
CMASS{30 20}
CMASS{31 8}
STRCQ4{40 2 3 4 5 {PSTRC 1} {FE 1}}
PMAT{1 2.1e+11 8.0e10 0.3 7850 {NAME "steel"}}
PMAT{2 7.0+10 2.6+10 0.33 2700}
PTHICK{1 0.01}
PTHICK{2 2.0-2}
PBEAMSECTION{4 1.0e-4 2.0e-9 1.0e-9 2.5e-9 0.85 0.8 0.004 -0.003}
PBEAMSECTION{5 1.0e-4 0 0 2.5e-9 0 0 0 0 25.0}
PBEAMECCENT{1 0.01 0.02 0.03 -0.01 0 0.005}
PBEAMPIN{1 456 23}
PORIENT{2 0 0.3 1}
PEFFLENGTH{1 0.9}
PSTRC{1 "shell" 0.5}
GROUP{1 1 3 7 {NAME "shells"}}
GROUP{2 11 12 13}
"""


def test_handwritten_part_with_constraints_and_quirks(tmp_path):
    p = str(tmp_path / "tricky.ftl")
    with open(p, "w") as f:
        f.write(TRICKY)
    s = compare_with_reference(p)
    assert s["nxnod"] == 2 and s["npbeam"] == 1 and s["nrgd"] == 1 and s["nwavgm"] == 1
    for sel in ("2", "<1>", "<PTHICK 1, 2>", "<PMAT 2>"):
        compare_with_reference(p, sel)
    mine = FtlPart(p)
    assert mine.ext2int(11) == -1 and mine.ext2int(1) == 1 and mine.ext2int(12, node=False) > 0
    lib = _lib.load_library()
    assert lib.fsr_ftl_version(mine.h) == 7


def test_errors_are_reported(tmp_path):
    p = str(tmp_path / "bad.ftl")
    with open(p, "w") as f:
        f.write("FTLVERSION{7 ASCII}\nNODE{1 0 0 0 0}\nNODE{2 0 1 0 0}\nNODE{3 0 1 1 0}\nTRI3{1 1 2 4 {PMAT 1}}\nPMAT{1 1 1 0.3 1}\n")
    with pytest.raises(_lib.FsrError, match="Resolving TRI3 element 1 failed"):
        FtlPart(p)
    with open(p, "w") as f:
        f.write("FTLVERSION{7 ASCII}\nNODE{1 0 0 0 0}\nNODE{2 0 1 0 0}\nNODE{3 0 1 1 0}\nTRI3{1 1 2 3 {PMAT 9}}\n")
    with pytest.raises(_lib.FsrError, match="PMAT 9"):
        FtlPart(p)
    with open(p, "w") as f:
        f.write("FTLVERSION{7 ASCII}\nNODE{1 0 0 0 0}\nNODE{2 0 1 0 0}\nNODE{3 0 1 1 0}\nTRI3{1 1 2 3}\nGROUP{4 1 2}\n")
    with pytest.raises(_lib.FsrError, match="Resolving element group 4 failed"):
        FtlPart(p)
    with pytest.raises(_lib.FsrError, match="Can not open"):
        FtlPart(str(tmp_path / "missing.ftl"))
    with open(p, "w") as f:
        f.write("FTLVERSION{7 ASCII}\nNODE{1 0 0 0 0")
    with pytest.raises(_lib.FsrError, match="corrupt"):
        FtlPart(p)
