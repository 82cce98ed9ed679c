"""Host side of the stress results database writer (csrc/io_rdb.cu):
 * the generated header is compared, byte for byte from VARIABLES: on, with a REAL fedem_stress output file
   of the reference (FFrTests/.../2_Boom_0001/Boom_1.frs, read in place when /root/reference is present);
 * files with that header and records laid out by the rules of calcStresses / writeStrMeasureDB are read
   back through the reference's OWN FFrLib reader (oracle/_ref/libfedem_ref_frs.so) and through the product's
   reader: every variable of every element/node must land on the slot the layout rules give it;
 * the total nodal displacement (calcTotalNodalDisplacement) of the record kernel against the oracle's
   statement-level restatement and an independent scipy composition of the rotations."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import oracle_bind
from fedem_solvers_b200 import _lib
from fedem_solvers_b200.frs import FrsReader, FrsWriter
from fedem_solvers_b200.rdb import build_header, out_mask, OUT
from test_frs_cpu import RefFrs, FIXTURES, _header_text

I32, F64 = np.int32, np.float64
NAMES = {11: "BEAM2", 21: "TRI3", 23: "TRI3", 22: "QUAD4", 24: "QUAD4", 31: "TRI6", 32: "QUAD8", 41: "TET10", 42: "WEDG15", 43: "HEX20", 44: "HEX8", 45: "TET4", 46: "WEDG6"}
NENOD = {11: 2, 21: 3, 23: 3, 22: 4, 24: 4, 31: 6, 32: 8, 41: 10, 42: 15, 43: 20, 44: 8, 45: 4, 46: 6}
MEASURES = ["Von Mises stress", "Max principal stress", "Min principal stress", "Max shear stress",
            "Von Mises strain", "Max principal strain", "Min principal strain", "Max shear strain"]


@pytest.mark.skipif(not os.path.isdir(FIXTURES), reason="reference fixtures not present")
def test_header_is_identical_to_real_fedem_stress_output():
    f = os.path.join(FIXTURES, "response_0001/timehist_rcy_0001/2_Boom_0001/Boom_1.frs")
    hdr, fsize, hsize = _header_text(f)
    body = hdr[hdr.index("VARIABLES:"):]
    base, user, descr = re.search(r'\{"Part";(\d+);(\d+);"([^"]*)";', hdr).groups()
    nodes = [int(x) for x in re.findall(r"\[;\s*(\d+);\[\s*1\]\]", hdr)]
    elems = [int(x) for x in re.findall(r"\[;\s*(\d+);\[\s*2\]\]", hdr)]
    # that run: -deformation (old iDef = 1: deformational displacements only) -stress -vmStress, all nodes 6 DOFs
    madof = 1 + 6 * np.arange(len(nodes) + 1)
    text, step_bytes = build_header(madof, np.full(len(elems), 21), out_mask(deformation=True, stress=True, vmStress=True),
                                    base_id=int(base), user_id=int(user), descr=descr, elmid=elems, minex=nodes)
    mine = text[text.index("VARIABLES:"):]
    assert mine == body
    nsteps = FrsReader(f).nsteps
    assert (fsize - hsize) == nsteps * step_bytes   # the record size accounts for the whole reference file


def _expected_slots(madof, melcon, elmid, minex, mask, total):
    """{variable path: (first slot, width)} by the rules of writeDisplacementDB / calcStresses / writeStrMeasureDB."""
    slots, k = {}, 0
    if mask & OUT["deformation"]:
        for n in range(len(madof) - 1):
            nd = madof[n + 1] - madof[n]
            if nd < 3:
                continue
            six = nd > 5
            p = f"Nodes|{minex[n]}|Dynamic response|"
            slots[p + "Translational deformation"] = (k, 3); k += 3
            if six:
                slots[p + "Angular deformation"] = (k, 3); k += 3
            if total:
                slots[p + "Total translation"] = (k, 3); k += 3
                if six:
                    slots[p + "Total rotation"] = (k, 3); k += 3
    sr, st, sn = bool(mask & OUT["SR"]), bool(mask & OUT["stress"]), bool(mask & OUT["strain"])
    sel = [j for j in range(8) if mask & (1 << j)]
    if not (sr or st or sn or sel):
        return slots, k
    for e, t in enumerate(melcon):
        if elmid[e] <= 0 or t not in NAMES:
            continue
        p = f"Elements|{elmid[e]}|{NAMES[t]}|Element nodes|"
        nn = NENOD[t]
        if t == 11:
            if not sr:
                continue
            for n in range(nn):
                slots[p + f"Basic|{n + 1}|Beam sectional force"] = (k, 3); k += 3
                slots[p + f"Basic|{n + 1}|Beam sectional moment"] = (k, 3); k += 3
            continue
        shell = t < 40
        ncmp = 3 if t < 30 else 6
        if shell and sr:
            for n in range(nn):
                slots[p + f"Basic|{n + 1}|Shell stress resultant force"] = (k, 3); k += 3
                slots[p + f"Basic|{n + 1}|Shell stress resultant moment"] = (k, 3); k += 3
        for side in (("Top", "Bottom") if shell else ("Basic",)):
            for n in range(nn):
                q = p + f"{side}|{n + 1}|"
                if st:
                    slots[q + "Stress"] = (k, ncmp); k += ncmp
                if sn:
                    slots[q + "Strain"] = (k, ncmp); k += ncmp
                for j in sel:
                    slots[q + MEASURES[j]] = (k, 1); k += 1
    return slots, k


CASES = [
    dict(mask=out_mask(vmStress=True), double=False, total=False),
    dict(mask=out_mask(SR=True, stress=True, strain=True, vmStress=True, maxPStress=True, minPStress=True, maxSStress=True,
                       vmStrain=True, maxPStrain=True, minPStrain=True, maxSStrain=True, deformation=True), double=True, total=True),
    dict(mask=out_mask(SR=True), double=False, total=False),
    dict(mask=out_mask(deformation=True), double=False, total=False),
    dict(mask=out_mask(stress=True, strain=True), double=False, total=False),
    dict(mask=out_mask(strain=True, maxPStrain=True, deformation=True), double=False, total=True),
]


@pytest.mark.parametrize("case", CASES, ids=[f"mask{c['mask']:03x}{'d' if c['double'] else 'f'}" for c in CASES])
def test_records_are_where_the_reference_reader_looks_for_them(tmp_path, case):
    rng = np.random.default_rng(case["mask"])
    melcon = np.array([24, 24, 23, 11, 41, 22, 43, 21, 11, 51, 24, 44, 45, 46, 42, 32, 31], I32)   # 51: a mass element, never written
    elmid = np.array([10, 11, 12, 13, 14, -15, 16, 17, 18, 19, 120, 121, 122, 123, 124, 125, 126], I32)   # -15: outside the -group selection
    ndofs = np.array([6, 6, 3, 6, 3, 3, 6, 0, 6], I32)                         # node 8 has no DOFs left
    madof = np.concatenate([[1], 1 + np.cumsum(ndofs)]).astype(I32)
    minex = np.array([1, 2, 5, 7, 8, 9, 20, 21, 300], I32)
    tr0 = np.hstack([np.eye(3), np.zeros((3, 1))]) if case["total"] else None
    text, step_bytes = build_header(madof, melcon, case["mask"], double=case["double"], base_id=33, user_id=4, descr="mixed part",
                                    model_file="m.fmm", link_file="mixed.ftl", elmid=elmid, minex=minex, sup_tr_init=tr0)
    slots, nslot = _expected_slots(madof, melcon, elmid, minex, case["mask"], case["total"])
    vb = 8 if case["double"] else 4
    assert step_bytes == 12 + nslot * vb
    path = str(tmp_path / "stress_1.frs")
    nsteps = 3
    data = rng.integers(-1000, 1000, (nsteps, nslot)).astype(F64 if case["double"] else np.float32)
    with FrsWriter(path, text, nslot * vb) as w:
        for s in range(nsteps):
            w.write_step(s + 1, 0.25 * s, data[s])
    ours, ref = FrsReader(path), RefFrs([path])
    keys = ref.keys()
    assert ours.nsteps == nsteps == len(keys)
    for p, (k, n) in slots.items():
        h = ours.find(p, "Part", 33)
        assert h is not None, p
        assert np.array_equal(ours.read(h), data[:, k:k + n]), p
        ok, b = ref.read(p, "Part", 33, keys, n)
        assert ok == nsteps and np.array_equal(b, data[:, k:k + n]), p
    assert ours.find("Elements|15|QUAD4|Element nodes|Top|1|Von Mises stress", "Part", 33) is None
    assert ours.find("Elements|19|CMASS|Element nodes|Basic|1|Stress", "Part", 33) is None
    ref.close()


def test_no_output_requested_is_an_error():
    with pytest.raises(_lib.FsrError, match="no result output requested"):
        build_header([1, 7], [24], 0)


def test_total_nodal_displacement_matches_oracle_and_scipy():
    from scipy.spatial.transform import Rotation
    lib = _lib.load_library()
    orc = C.CDLL(oracle_bind.ensure_built())
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(400):
        scale = [1e-9, 1e-5, 4e-4, 1e-3, 0.3, 2.0, 3.1][trial % 7]
        x0, u = rng.normal(0, 1, 3), np.concatenate([rng.normal(0, 1e-3, 3), rng.normal(0, scale, 3)])
        R, R0 = Rotation.from_rotvec(rng.normal(0, 1.0, 3)), Rotation.from_rotvec(rng.normal(0, 0.3, 3))
        T = np.hstack([R.as_matrix(), rng.normal(0, 1, (3, 1))])
        T0 = np.hstack([R0.as_matrix(), rng.normal(0, 1, (3, 1))])
        Tc, T0c = np.ascontiguousarray(T.T.ravel()), np.ascontiguousarray(T0.T.ravel())
        for nd in (3, 6):
            a, b = np.zeros(6), np.zeros(6)
            lib.fsr_total_nodal_displacement(dp(x0), dp(u), nd, dp(Tc), dp(T0c), dp(a))
            orc.orc_total_nodal_displacement(dp(x0), dp(u), nd, dp(Tc), dp(T0c), dp(b))
            assert np.abs(a - b).max() <= 1e-14 * max(1.0, np.abs(b).max())
            exact = T[:, :3] @ (x0 + u[:3]) + T[:, 3] - (T0[:, :3] @ x0 + T0[:, 3])
            assert np.abs(a[:3] - exact).max() <= 1e-13
            if nd == 6:
                want = Rotation.from_rotvec(u[3:]) * R * R0.inv()
                # the reference does not reduce the angle to [0, pi] (acos of a negative quaternion scalar), so
                # compare the rotations, not the vectors; it also interpolates sin(x)/x linearly below 5e-4 rad,
                # which costs ~1e-8 relative on the incoming rotation vector
                err = (Rotation.from_rotvec(a[3:]) * want.inv()).magnitude()
                worst = max(worst, err)
                assert err <= 2e-8 * max(scale, 1e-3) + 1e-13
    assert worst > 0.0
