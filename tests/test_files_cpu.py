"""CPU tests of the file layer on the drop-in surface (.fmx / .fsm) and of BuildFinit, against an
independent byte-level restatement of the reference's formats (FFaTag.C:192-297, binaryDB.c:643-733,
saveSAM samReducerModule.f90:586-672) written with numpy/struct in this file."""
import os
import struct
import numpy as np
import pytest

from fedem_solvers_b200 import files
from fedem_solvers_b200.model import plate_part, reduced_history

F64 = np.float64


def _ref_header(tag, checksum, little=True):
    """46 bytes: tag padded to 30 chars, 16-bit 0x1234, 4 zero bytes + 32-bit checksum, ';1.0;\\n'"""
    e = "<" if little else ">"
    return tag.ljust(30).encode() + struct.pack(e + "H", 0x1234) + struct.pack(e + "II", 0, checksum) + b";1.0;\n"


def test_fmx_bytes_and_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.normal(size=(17, 5)))
    p = str(tmp_path / "x_B.fmx")
    files.write_fmx(p, A, checksum=123456789)
    raw = open(p, "rb").read()
    assert len(raw) == 46 + 8 * A.size                      # dmOpen reports 46 + nbd*dmSize (diskMatrixModule.f90:319)
    assert raw[:46] == _ref_header("#FEDEM disk matrix", 123456789)
    assert np.array_equal(np.frombuffer(raw[46:], "<f8"), A.ravel(order="F"))
    B, tag, cs, sp = files.read_fmx(p, 17, 5)
    assert np.array_equal(A, B) and tag == "#FEDEM disk matrix" and cs == 123456789 and not sp
    # single-precision storage: tag suffix ' SP', float payload (diskMatrixModule.f90:271-290)
    files.write_fmx(p, A, checksum=7, single_precision=True)
    raw = open(p, "rb").read()
    assert raw[:30] == b"#FEDEM disk matrix SP".ljust(30) and len(raw) == 46 + 4 * A.size
    B, tag, cs, sp = files.read_fmx(p, 17, 5)
    assert sp and tag == "#FEDEM disk matrix" and np.array_equal(B, A.astype(np.float32).astype(F64))
    # generalized-modes tag is checked like dmOpen's wantTag
    files.write_fmx(p, A, tag=files.GM_TAG)
    with pytest.raises(Exception):
        files.read_fmx(p, 17, 5, want_tag=files.DM_TAG)
    assert files.read_fmx(p, 17, 5, want_tag=files.GM_TAG)[1] == files.GM_TAG


def test_fmx_big_endian_file_is_swapped(tmp_path):
    A = np.arange(12, dtype=F64).reshape(4, 3, order="F") * 0.37
    p = str(tmp_path / "be.fmx")
    with open(p, "wb") as f:
        f.write(_ref_header("#FEDEM disk matrix", 42, little=False))
        f.write(A.ravel(order="F").astype(">f8").tobytes())
    B, tag, cs, sp = files.read_fmx(p, 4, 3)
    assert np.array_equal(A, B) and cs == 42


def test_fmx_errors(tmp_path):
    p = str(tmp_path / "bad.fmx")
    open(p, "wb").write(b"hello world, not a fedem file" * 3)
    with pytest.raises(Exception, match="tagged"):
        files.read_fmx(p, 1, 1)
    files.write_fmx(p, np.zeros((2, 2)))
    with pytest.raises(Exception, match="end of file"):
        files.read_fmx(p, 3, 3)


def test_fsm_layout_and_roundtrip(tmp_path):
    part = plate_part(4, 3, ngen=3, seed=2, tri_fraction=0.4, shuffle_eq=True, n_fixed=2, n_constraints=2)
    s = part.sam
    p = str(tmp_path / "p_SAM.fsm")
    files.write_fsm(p, s, checksum=99, part_id=17)
    raw = open(p, "rb").read()
    assert raw[:46] == _ref_header("#SAM data", 99)
    body = raw[46:]
    npar = struct.unpack("<i", body[:4])[0]
    assert npar == 50
    mpar = np.frombuffer(body[4:4 + 4 * npar], "<i4")
    assert (mpar[0], mpar[1], mpar[2], mpar[3], mpar[4], mpar[6], mpar[10]) == (s.nnod, s.nel, s.ndof, s.ndof1, s.ndof2, s.nceq, s.neq)
    assert mpar[17] == 17 and mpar[21] == 3 and mpar[23] == s.ndof2 + 3
    # array order of saveSAM: madof, minex, mnnn, msc, mpmnpc, mmnpc, melcon, mpmceq, mmceq, ttcc, meqn, meqn1, meqn2
    off = 4 + 4 * npar
    def take(n, dt="<i4"):
        nonlocal off
        a = np.frombuffer(body[off:off + n * int(dt[-1])], dt); off += n * int(dt[-1]); return a
    assert np.array_equal(take(s.nnod + 1), s.madof)
    take(s.nnod); take(s.nnod)
    assert np.array_equal(take(s.ndof), s.msc)
    assert np.array_equal(take(s.nel + 1), s.mpmnpc)
    assert np.array_equal(take(len(s.mmnpc)), s.mmnpc)
    assert np.array_equal(take(s.nel), s.melcon)
    assert np.array_equal(take(s.nceq + 1), s.mpmceq)
    assert np.array_equal(take(len(s.mmceq)), s.mmceq)
    assert np.array_equal(take(len(s.ttcc), "<f8"), s.ttcc)
    assert np.array_equal(take(s.ndof), s.meqn)
    assert np.array_equal(take(s.ndof1), s.meqn1)
    assert np.array_equal(take(s.ndof2), s.meqn2)
    assert off == len(body)
    s2, mp2, cs = files.read_fsm(p)
    assert cs == 99 and s2.ngen == 3
    for k in ("nnod", "nel", "ndof", "ndof1", "ndof2", "neq", "nceq"):
        assert getattr(s2, k) == getattr(s, k)
    for k in ("madof", "msc", "mpmnpc", "mmnpc", "melcon", "meqn", "meqn1", "meqn2", "mpmceq", "mmceq", "ttcc"):
        assert np.array_equal(getattr(s2, k), getattr(s, k)), k


def test_part_files_roundtrip_feeds_the_oracle(tmp_path, oracle):
    """reducer-style files -> load_part -> same recovery as the in-memory part"""
    part = plate_part(5, 4, ngen=4, seed=3, n_constraints=2)
    prefix = str(tmp_path / "plate")
    files.save_part(prefix, part, checksum=5)
    assert sorted(os.listdir(tmp_path)) == ["plate_B.fmx", "plate_E.fmx", "plate_SAM.fsm"]
    p2 = files.load_part(prefix, part.elm)
    Q = reduced_history(part.sam.ndim, 3, seed=1)
    a = oracle.recover_history(oracle.bind_part(part), Q)[0]
    b = oracle.recover_history(oracle.bind_part(p2), Q)[0]
    assert np.array_equal(a, b)


def _rot(axis, ang):
    axis = np.asarray(axis, F64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def test_build_finit():
    """BuildFinit (supElTypeModule.f90:1067-1114) restated with numpy: urLocal = supTr^-1 o ur,
    translations minus the undeformed position, rotations = (dR32, dR13, dR21) of
    urLocal(:,1:3) . TrUndeformed(:,1:3); a rigid motion of the whole superelement gives zero."""
    rng = np.random.default_rng(4)
    nt, ns, ngen = 5, 7, 3
    tr_undef = np.zeros((nt, 3, 4))
    for i in range(nt):
        tr_undef[i, :, :3] = _rot(rng.normal(size=3), rng.uniform(0, 2))
        tr_undef[i, :, 3] = rng.normal(size=3)
    ndofs = np.array([6, 6, 3, 6, 0])
    first = np.array([1, 7, 13, 16, 0])
    gen_first = 22
    sup = np.zeros((ns, 3, 4)); ur = np.zeros((ns, nt, 3, 4)); gen = rng.normal(size=(ns, ngen))
    want = np.zeros((gen_first - 1 + ngen, ns))
    for s in range(ns):
        R, x = _rot(rng.normal(size=3), rng.uniform(0, 3)), rng.normal(size=3) * 10
        sup[s, :, :3], sup[s, :, 3] = R, x
        for i in range(nt):
            dRl = _rot(rng.normal(size=3), 1e-3 * s)           # small deformational rotation
            dul = rng.normal(size=3) * 1e-3 * s                # deformational translation
            Tl = np.zeros((3, 4)); Tl[:, :3] = dRl @ tr_undef[i, :, :3]; Tl[:, 3] = tr_undef[i, :, 3] + dul
            ur[s, i, :, :3] = R @ Tl[:, :3]; ur[s, i, :, 3] = R @ Tl[:, 3] + x
            if ndofs[i] >= 3:
                want[first[i] - 1:first[i] + 2, s] = dul
            if ndofs[i] >= 6:
                dR = Tl[:, :3] @ tr_undef[i, :, :3]            # as the reference forms it (no transpose)
                want[first[i] + 2:first[i] + 5, s] = [dR[2, 1], dR[0, 2], dR[1, 0]]
        want[gen_first - 1:, s] = gen[s]
    Q = files.build_finit(sup, ur, tr_undef, ndofs, first, gen, gen_first)
    assert Q.shape == want.shape
    assert np.abs(Q - want).max() <= 1e-13
    # physics: with the triads aligned with the part axes in the undeformed state (the normal case: the
    # FE node DOFs are in the part system), a rigid motion of the whole superelement gives zero and a
    # small relative rotation theta about z gives (0, 0, theta) to first order
    tr_id = tr_undef.copy(); tr_id[:, :, :3] = np.eye(3)
    R, x = _rot([1, 2, 3], 0.7), np.array([5.0, -3.0, 2.0])
    sup1 = np.zeros((2, 3, 4)); sup1[:, :, :3] = R; sup1[:, :, 3] = x
    ur1 = np.zeros((2, nt, 3, 4))
    th = 1e-6
    for i in range(nt):
        for s_, dl in enumerate((np.eye(3), _rot([0, 0, 1], th))):
            ur1[s_, i, :, :3] = R @ dl
            ur1[s_, i, :, 3] = R @ tr_id[i, :, 3] + x
    Q1 = files.build_finit(sup1, ur1, tr_id, ndofs, first, None, 0, ndim=21)
    assert np.abs(Q1[:, 0]).max() <= 1e-14
    assert np.allclose(Q1[3:6, 1], [0, 0, th], atol=1e-12) and np.abs(Q1[0:3, 1]).max() <= 1e-14
