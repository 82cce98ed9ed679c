"""fedem_fpp inputs (host side, no GPU): the strain coat elements of the FE part and the S-N curve library file, both checked value by
value against the reference's OWN code compiled unmodified into oracle/_ref/libfedem_ref_ffl.so (ffl_getnostrc / ffl_getstraincoat,
FFlLinkHandler_F.C:1587-1762; FFpSNCurveLib::readSNCurves + FFpSNCurve::getValue)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fedem_solvers_b200 import _lib  # noqa: E402
from fedem_solvers_b200.ftl import FtlPart, write_ftl  # noqa: E402
from fedem_solvers_b200.model import plate_part  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfedem_ref_ffl.so")
EXE = os.path.join(ROOT, "fedem_solvers_b200", "bin", "fedem_fpp")
I32, F64 = np.int32, np.float64
needs_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libfedem_ref_ffl.so not built")


def coats_for(part, every=2, fatigue_every=3):
    """strain coats on every `every`-th shell: Bottom + Top by thickness reference, every third one also Mid by height"""
    sam = part.sam
    out = []
    for k, e in enumerate(range(0, sam.nel, every)):
        sets = [("Bottom", None), ("Top", None)]
        if k % 3 == 2:
            sets.insert(1, ("Mid", 0.001 * (k + 1)))
        fat = (k % 2, (k // 2) % 2, 1.0 + 0.25 * (k % 4)) if k % fatigue_every else None
        out.append(dict(id=1000 + k, elm=e, sets=sets, fatigue=fat))
    return out


_REF_SCRIPT = r"""
import ctypes as C, json, os, sys
import numpy as np
I32, F64 = np.int32, np.float64
L = C.CDLL(sys.argv[1])
assert L.ref_ffl_load(os.fsencode(sys.argv[2]), sys.argv[3].encode()) >= 0
L.ffl_getnostrc_.restype = C.c_int
n = L.ffl_getnostrc_()
out = []
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
while True:
    cid, nnod, npts, eid, ierr = (C.c_int() for _ in range(5))
    nodes, mat, rset, snc = np.zeros(8, I32), np.zeros(3, I32), np.zeros(3, I32), np.zeros(6, I32)
    E, nu, Z, scf = (np.zeros(3, F64) for _ in range(4))
    L.ffl_getstraincoat_(C.byref(cid), C.byref(nnod), C.byref(npts), ip(nodes), ip(mat), dp(E), dp(nu), dp(Z), ip(rset), dp(scf),
                         ip(snc), C.byref(eid), C.byref(ierr))
    assert ierr.value >= 0
    if ierr.value > 0:
        break
    k = npts.value
    out.append(dict(id=cid.value, nodes=[int(x) for x in nodes[:nnod.value]], npts=k, elm_id=eid.value, mat_id=[int(x) for x in mat[:k]],
                    res_set=[int(x) for x in rset[:k]], sn_curve=[[int(snc[2 * j]), int(snc[2 * j + 1])] for j in range(k)],
                    emod=[float(x).hex() for x in E[:k]], nu=[float(x).hex() for x in nu[:k]], zpos=[float(x).hex() for x in Z[:k]],
                    scf=[float(x).hex() for x in scf[:k]]))
print(json.dumps([n, out]))
"""


def ref_strain_coats(path, groups=""):
    """the reference's ffl_getnostrc / ffl_getstraincoat in a process of its own: ffl_getstraincoat keeps a static element iterator
    (FFlLinkHandler_F.C:1656) that dangles once a second FE part has been loaded into the same process"""
    import json
    r = subprocess.run([sys.executable, "-c", _REF_SCRIPT, REF_SO, path, groups], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    n, out = json.loads(r.stdout.strip().splitlines()[-1])
    for d in out:
        d["sn_curve"] = [tuple(x) for x in d["sn_curve"]]
        for key in ("emod", "nu", "zpos", "scf"):
            d[key] = [float.fromhex(x) for x in d[key]]
    return n, out


@needs_ref
@pytest.mark.parametrize("groups", ["", "<3>"])
def test_strain_coats_match_the_reference_reader(tmp_path, groups):
    part = plate_part(6, 5, ngen=2, seed=71, tri_fraction=0.35, warp=0.02)
    coats = coats_for(part)
    p = str(tmp_path / "coated.ftl")
    ids = np.abs(part.elm.elmid)
    # group 3: some shells AND some of the strain coat elements (the calculation flag decides which coats are processed)
    write_ftl(p, part, groups={3: [int(ids[0]), int(ids[2]), 1000, 1002, 1003]}, strain_coats=coats)
    n_ref, ref = ref_strain_coats(p, groups)
    mine = FtlPart(p, groups)
    got = mine.strain_coats()
    assert len(got) == n_ref == len(ref) and n_ref == (3 if groups else len(coats))
    for a, b in zip(got, ref):
        assert a["id"] == b["id"] and a["nodes"] == b["nodes"] and a["npts"] == b["npts"] and a["elm_id"] == b["elm_id"], (a, b)
        assert a["mat_id"] == b["mat_id"] and a["res_set"] == b["res_set"] and a["sn_curve"] == b["sn_curve"], (a, b)
        for key in ("emod", "nu", "zpos", "scf"):
            assert a[key] == b[key], (key, a, b)
    # and they are what was written: Bottom / Top at -+ t/2, Mid at the given height, the coat sits on its shell's nodes
    sam = part.sam
    for a in got:
        sc = coats[a["id"] - 1000]
        e = sc["elm"]
        assert a["elm_id"] == int(ids[e])
        assert a["nodes"] == [int(k) for k in sam.mmnpc[sam.mpmnpc[e] - 1: sam.mpmnpc[e + 1] - 1]]
        for (name, h), rs, z in zip(sc["sets"], a["res_set"], a["zpos"]):
            assert rs == {"Bottom": 1, "Mid": 2, "Top": 3}[name]
            want = h if h is not None else {"Bottom": -0.5, "Top": 0.5}[name] * float(part.elm.thk[e])
            assert z == want
        if sc["fatigue"] is None:
            assert all(s == (-1, -1) for s in a["sn_curve"])
        else:
            assert all(s == tuple(sc["fatigue"][:2]) for s in a["sn_curve"]) and a["scf"] == [sc["fatigue"][2]] * a["npts"]
    # the finite element side of the part is unchanged by the coats
    assert mine.sizes()["nel"] == sam.nel


SN_TEXT = """# S-N curves for the fpp tests
<"NorSok air", 0,
  <B1, <15.117, 4.0, 17.146, 5.0>, 0.0>,
  <"C 1", <12.592, 3.0, 16.320, 5.0>, 0.15>,
  <bad_parallel, <12.0, 3.0, 13.0, 3.0>, 0.0>,
  <one_segment, <12.164, 3.0>, 0.25>,
  <three, <12.0, 3.0, 16.0, 5.0, 17.5, 6.0>, 0.0>,
  <neg, <12.0, -3.0>, 0.0>,
  <wrong_count, <12.0, 3.0>>
>
  # indented comment
junk line that the reader skips
<British, 1,
  <"Class D", <12.1818, 3.0, 0.2095, 2.0>>,
  <E, <12.0128, 3.0, 0.2509, 2.0>>,
  <odd, <12.0, 3.0, 1.0>>
>
<Unknown standard, 7, <X, <1.0, 2.0>>>
<"Empty", 0>
"""


@needs_ref
def test_sn_curve_library_matches_the_reference_reader(tmp_path):
    p = str(tmp_path / "curves.fsn")
    open(p, "w").write(SN_TEXT)
    R = C.CDLL(REF_SO)
    R.ref_sn_value.restype = C.c_double
    R.ref_sn_value.argtypes = [C.c_int, C.c_int, C.c_double]
    assert R.ref_sn_read(os.fsencode(p)) == 1
    L = _lib.load_library()
    h = C.c_void_p()
    assert L.fsr_sn_read(C.byref(h), os.fsencode(p)) == 0
    nstd = R.ref_sn_num_standards()
    assert L.fsr_sn_num_standards(h) == nstd == 2
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    total = 0
    for i in range(nstd):
        nc = R.ref_sn_num_curves(i)
        assert L.fsr_sn_num_curves(h, i) == nc
        for j in range(nc):
            la_r, m_r, la, m, n0 = (np.zeros(8) for _ in range(5))
            sid_r, sid = C.c_int(), C.c_int()
            ns_r = R.ref_sn_get(i, j, C.byref(sid_r), dp(la_r), dp(m_r), 8)
            ns = L.fsr_sn_get(h, i, j, C.byref(sid), dp(la), dp(m), dp(n0), 8)
            assert ns == ns_r and sid.value == sid_r.value and np.array_equal(la, la_r) and np.array_equal(m, m_r), (i, j)
            for s in (0.5, 1.0, 7.3, 25.0, 52.6, 83.0, 140.0, 400.0, 2.0e3):
                assert L.fsr_sn_value(h, i, j, s) == R.ref_sn_value(i, j, s), (i, j, s)
            total += 1
    assert total == 6          # NorSok: B1, C 1, one_segment, three; British: Class D, E
    assert L.fsr_sn_get(h, 0, 9, None, None, None, None, 0) < 0 and L.fsr_sn_value(h, 5, 0, 10.0) < 0
    L.fsr_sn_free(h)
    assert L.fsr_sn_read(C.byref(h), os.fsencode(str(tmp_path / "missing.fsn"))) < 0


def test_fpp_executable_options_and_loud_failures(tmp_path):
    assert os.path.exists(EXE), "build.sh did not produce bin/fedem_fpp"
    r = subprocess.run([EXE, "-help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    for opt in ("-surface", "-angleBins", "-biAxialGate", "-PVXGate", "-HistDataType", "-SNfile", "-stressToMPaScale", "-blockSize", "-rdbfile",
                "-group", "-double"):
        assert opt + " " in r.stdout, opt
    assert "-vmStress" not in r.stdout and "-recover_modes" not in r.stdout and "-writeHistory" not in r.stdout   # private option
    r = subprocess.run([EXE, "-helpAll"], capture_output=True, text=True, timeout=60)
    assert "-writeHistory" in r.stdout and "-oldRange" in r.stdout
    r = subprocess.run([EXE, "-cwd", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "FE data file must be specified through -linkfile" in r.stdout and "Strain coat calculation failed" in r.stdout
    assert os.path.exists(tmp_path / "fedem_fpp.res")
