"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerance: BASELINE.json north_star -- <= 1e-10 relative on FP64 stresses, measured PER VALUE:
|gpu - oracle| <= 1e-10 * max(|oracle value|, 1e-3 * max|oracle field|).  The floor is there because a stress component
is a sum of ~24-60 products that partly cancel (rigid-body content of the element displacements): its absolute rounding
error scales with the size of the terms, not of the result, so values far below the field maximum cannot be held to 1e-10
of themselves by ANY summation order -- the oracle's included."""
import numpy as np
import pytest

from fedem_solvers_b200 import StressRecovery
from fedem_solvers_b200.model import thickshell_panel, linsolid_block, wedg15_block, plate_part, tet10_block, hex20_block, reduced_history

pytestmark = pytest.mark.gpu
TOL = 1.0e-10


FLOOR = 1.0e-3


def rel_err(a, b):
    """largest per-value relative difference, each value measured against max(|b|, FLOOR * max|b|)"""
    a, b = np.asarray(a), np.asarray(b)
    if b.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), max(FLOOR * np.abs(b).max(), 1e-300))
    return float((np.abs(a - b) / scale).max())


def _check_part(oracle, part, nsteps, seed, step_tile=0):
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, nsteps, seed=seed)
    vm_o, mx_o, mn_o = oracle.recover_history(b, Q)
    rec = StressRecovery(part, step_tile=step_tile)
    assert rec.npts == b["npts"]
    assert np.array_equal(rec.result_point_offsets(), b["ptoff"])
    vm_g = rec.recover(Q)
    mx_g, mn_g = rec.envelope()
    assert rel_err(vm_g, vm_o) <= TOL, rel_err(vm_g, vm_o)
    assert rel_err(mx_g, mx_o) <= TOL
    assert rel_err(mn_g, mn_o) <= TOL
    # expansion alone (calcIntDisplacements)
    U = rec.calc_int_displacements(Q[:, :3])
    for s in range(3):
        sv = oracle.expand(b, Q[:, s])
        assert rel_err(U[s], sv) <= TOL
    # full result set of one step
    full = rec.calc_stresses(Q[:, 1])
    ref = oracle.calc_stresses(b, oracle.expand(b, Q[:, 1]))
    assert rel_err(full["sv"], oracle.expand(b, Q[:, 1])) <= TOL
    for key in ("stress", "strain"):
        assert rel_err(full[key], ref[key]) <= TOL, (key, rel_err(full[key], ref[key]))
    for k, name in enumerate(["vmStress", "maxPStress", "minPStress", "maxSStress", "vmStrain",
                              "maxPStrain", "minPStrain", "maxSStrain"]):
        e = rel_err(full["resmat"][:, k], ref["resmat"][:, k])
        # principal values and the max shear derived from them: trigonometric cubic, 1e-9
        assert e <= (1e-9 if ("P" in name or "maxS" in name) else TOL), (name, e)
    rec.close()
    return vm_g


def test_quad_plate_small(oracle):
    part = plate_part(9, 7, ngen=5, seed=11, shuffle_eq=True, n_fixed=4, n_constraints=3, warp=0.05)
    _check_part(oracle, part, nsteps=37, seed=3)


def test_quad_plate_multi_tile(oracle):
    """step window larger than the device batch: tiles of 64 steps, ragged last tile"""
    part = plate_part(8, 8, ngen=4, seed=12)
    _check_part(oracle, part, nsteps=150, seed=4, step_tile=64)


def test_tet10_block_small(oracle):
    part = tet10_block(3, 2, 2, ngen=6, seed=5, shuffle_eq=True)
    _check_part(oracle, part, nsteps=21, seed=6)


@pytest.mark.parametrize("curved", ["none", "surface"])
def test_tet10_straight_sided_fast_path(oracle, curved):
    """straight-sided TET10 (constant Jacobian) take k2_tet10_affine_vm_kernel: corner gradients from three-point
    differences, mid-edge points by averaging; "surface" mixes them with curved elements (general DMMA kernel) in one part.
    Ragged step count, several tiles, beams in between."""
    part = tet10_block(4, 3, 3, ngen=6, seed=21, shuffle_eq=True, n_beams=5, curved=curved)
    vm = _check_part(oracle, part, nsteps=150, seed=8, step_tile=64)
    # the same part with the fast path switched off gives the same numbers to rounding
    import os
    os.environ["FSR_TET10_AFFINE"] = "0"
    try:
        rec = StressRecovery(part, step_tile=64)
        vm2 = rec.recover(reduced_history(part.sam.ndim, 150, seed=8))
        rec.close()
    finally:
        del os.environ["FSR_TET10_AFFINE"]
    assert rel_err(vm, vm2) <= 1e-12
    assert not np.array_equal(vm, vm2)        # ... but by a different kernel


def test_tet10_rotated_skewed_straight_sided(oracle):
    """fast path on a sheared and rotated block (J is not diagonal), nu close to incompressible"""
    part = tet10_block(3, 3, 2, ngen=5, seed=22, curved="none", rny=0.49)
    A = np.array([[0.9, 0.3, -0.2], [-0.1, 1.2, 0.4], [0.25, -0.15, 0.8]])
    part.elm.xyz = part.elm.xyz @ A.T + np.array([3.0, -2.0, 1.0])
    _check_part(oracle, part, nsteps=33, seed=9)


@pytest.mark.parametrize("ngen", [84, 110, 300])
def test_large_reduced_dimension(oracle, ngen):
    """ndof2 + ngen > 108 does not fit the resident-K expansion kernel: the K-slab kernel (52-column slabs, accumulators
    kept across the slabs of a step chunk) takes over -- 3 slabs with a narrow last one / 4 slabs / 7 slabs"""
    part = plate_part(7, 6, ngen=ngen, seed=31, n_ext=8, tri_fraction=0.3, n_constraints=2)
    assert part.sam.ndim > 108
    _check_part(oracle, part, nsteps=140, seed=7, step_tile=64)


def test_flat_and_warped_quads_in_one_part(oracle):
    """flat quads take the membrane / bending split (12 DMMA per 8 steps), warped ones the dense operator; one part holds
    both, turned in space so that no global direction is special; the two kernels agree to rounding on the flat ones"""
    import os
    from scipy.spatial.transform import Rotation
    part = plate_part(11, 9, ngen=5, seed=33, jitter=0.2, warp=0.03, tri_fraction=0.15, shuffle_eq=True, n_constraints=2)
    part.elm.xyz[part.elm.xyz[:, 0] < 0.55, 2] = 0.0        # the left half is flat
    part.elm.xyz = part.elm.xyz @ Rotation.from_rotvec([0.3, 0.5, -0.4]).as_matrix().T + np.array([1.0, 2.0, 3.0])
    vm = _check_part(oracle, part, nsteps=100, seed=12, step_tile=64)
    os.environ["FSR_QUAD_FLAT"] = "0"
    try:
        rec = StressRecovery(part, step_tile=64)
        vm2 = rec.recover(reduced_history(part.sam.ndim, 100, seed=12))
        rec.close()
    finally:
        del os.environ["FSR_QUAD_FLAT"]
    assert rel_err(vm, vm2) <= 1e-12 and not np.array_equal(vm, vm2)


def _folded_plate(nx=12, ny=7, angle=0.6, **kw):
    """a plate folded along the node line x = lx / 2 and turned in space: two flat panels that share the fold nodes"""
    from scipy.spatial.transform import Rotation
    part = plate_part(nx, ny, **kw)
    x = part.elm.xyz
    xf = x[:, 0].max() / 2
    right = x[:, 0] > xf + 1e-9
    d = x[right, 0] - xf
    x[right, 0] = xf + d * np.cos(angle)
    x[right, 2] = d * np.sin(angle)
    part.elm.xyz = x @ Rotation.from_rotvec([-0.4, 0.2, 0.7]).as_matrix().T + np.array([-2.0, 0.5, 1.0])
    return part


def test_quads_of_flat_regions_use_inplane_rows(oracle):
    """nodes whose flat quads all lie in one plane get (u, v, theta1, theta2) rows in the axes of the plane (K1 expands 4
    rows instead of 6, the element operator is two 12 x 8 blocks); the fold nodes, the triangles and their neighbours keep
    the global rows.  Same numbers as the oracle and, to rounding, as the run with the in-plane form switched off."""
    import os
    part = _folded_plate(ngen=6, seed=61, jitter=0.0, tri_fraction=0.1, shuffle_eq=True, n_constraints=2)
    nq = int((part.sam.melcon == 24).sum())
    os.environ["FSR_QUAD_PLANAR"] = "2"    # wherever it applies: the cost rule may leave so small a part on the global rows
    try:
        rec = StressRecovery(part, step_tile=64)
        info = rec.vm_path_info()
        rec.close()
        assert info["quads_inplane"] > 0 and info["quads_flat"] > 0 and info["quads_dense"] == 0, info
        assert info["quads_inplane"] + info["quads_flat"] == nq
        assert info["inplane_rows"] % 4 == 0 and 0 < info["inplane_rows"] < 4 * part.sam.nnod
        vm = _check_part(oracle, part, nsteps=100, seed=12, step_tile=64)
    finally:
        del os.environ["FSR_QUAD_PLANAR"]
    os.environ["FSR_QUAD_PLANAR"] = "0"
    try:
        rec = StressRecovery(part, step_tile=64)
        assert rec.vm_path_info()["quads_inplane"] == 0 and rec.vm_path_info()["k1_rows"] >= part.sam.ndof
        vm2 = rec.recover(reduced_history(part.sam.ndim, 100, seed=12))
        rec.close()
    finally:
        del os.environ["FSR_QUAD_PLANAR"]
    assert rel_err(vm, vm2) <= 1e-12 and not np.array_equal(vm, vm2)


def test_inplane_rows_skip_the_global_expansion_of_a_flat_plate(oracle):
    """a flat plate of quads only: no tile of the global operator is expanded on the von Mises path, K1 does 2/3 of the rows;
    the displacement outputs (expand, the solver step) still come from the global operator"""
    part = plate_part(40, 30, ngen=7, seed=62, jitter=0.1, shuffle_eq=True)
    rec = StressRecovery(part, step_tile=64)
    info = rec.vm_path_info()
    assert info["quads_inplane"] == part.sam.nel and info["global_row_tiles"] == 0, info
    assert info["inplane_rows"] == 4 * part.sam.nnod and info["k1_rows"] < info["ndof"]
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 70, seed=3)
    vm = rec.recover(Q)
    sv = rec.calc_int_displacements(Q)
    vm_o, _, _ = oracle.recover_history(b, Q)
    assert rel_err(vm, vm_o) <= TOL
    for t in (0, 33, 69):
        assert rel_err(sv[t], oracle.expand(b, Q[:, t])) <= TOL
    # the displacements given instead of the reduced history (direct solution): the in-plane rows come from U
    rec.reset_envelope()
    vm_d = rec.recover_displacements(sv)
    assert rel_err(vm_d, vm_o) <= TOL
    rec.close()


def test_inplane_rows_only_where_they_pay(oracle):
    """flat quadrilaterals scattered among triangles: every row tile of R is still read by the triangles, the in-plane rows
    would come on top (4 + 6 rows per node in K1) and cost more than the smaller quad operator saves, so the part stays on
    the global rows; FSR_QUAD_PLANAR=2 takes the in-plane form anyway, same numbers to rounding"""
    import os
    part = plate_part(20, 16, ngen=40, seed=71, jitter=0.0, tri_fraction=0.5, shuffle_eq=True)
    rec = StressRecovery(part, step_tile=64)
    info = rec.vm_path_info()
    rec.close()
    assert info["quads_inplane"] == 0 and info["quads_flat"] == int((part.sam.melcon == 24).sum()), info
    vm = _check_part(oracle, part, nsteps=70, seed=13, step_tile=64)
    os.environ["FSR_QUAD_PLANAR"] = "2"
    try:
        rec = StressRecovery(part, step_tile=64)
        assert rec.vm_path_info()["quads_inplane"] > 0
        vm2 = rec.recover(reduced_history(part.sam.ndim, 70, seed=13))
        rec.close()
    finally:
        del os.environ["FSR_QUAD_PLANAR"]
    assert rel_err(vm, vm2) <= 1e-12


def test_tri_quad_mixed_plate(oracle):
    """ANDES triangles (type 23: two LU inversions per element, REAL*4 Gauss rule) mixed with quads"""
    part = plate_part(9, 8, ngen=6, seed=13, tri_fraction=0.5, shuffle_eq=True, n_fixed=3, n_constraints=2, warp=0.04)
    assert (part.sam.melcon == 23).sum() > 20 and (part.sam.melcon == 24).sum() > 20
    _check_part(oracle, part, nsteps=70, seed=5)


def test_all_triangles_ragged_tile(oracle):
    part = plate_part(6, 5, ngen=3, seed=14, tri_fraction=1.0)
    _check_part(oracle, part, nsteps=129, seed=6, step_tile=64)


def test_tet10_with_beams(oracle):
    """config-3-shaped part: TET10 + BEAM2 stiffeners (eccentricities, shear-centre offsets, rotated
    principal axes, effective length).  Beams give section forces only (nstrp = 0)."""
    part = tet10_block(3, 2, 2, ngen=6, seed=5, shuffle_eq=True, n_beams=9)
    _check_part(oracle, part, nsteps=21, seed=6)
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 2, seed=3)
    rec = StressRecovery(part)
    full = rec.calc_stresses(Q[:, 1])
    ref = oracle.calc_stresses(b, oracle.expand(b, Q[:, 1]))
    beams = np.nonzero(part.sam.melcon == 11)[0]
    assert len(beams) == 9
    sf_g, sf_o = full["sres"][beams, :12], ref["sres"][beams, :12]
    assert np.abs(sf_o).max() > 0
    for k in range(12):   # compare per section-force component (forces and moments differ in scale)
        assert np.abs(sf_g[:, k] - sf_o[:, k]).max() <= TOL * np.abs(sf_o[:, k % 6::6]).max(), k
    rec.close()


def test_tet10_gauss_extrapolation(oracle):
    """-stressForm 1: 4 Gauss points with the reference's REAL*4 abscissae, extrapolated to the nodes"""
    part = tet10_block(2, 2, 1, ngen=4, seed=8)
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 9, seed=2)
    rec = StressRecovery(part, stress_form=1)
    vm_g = rec.recover(Q)
    import ctypes as C
    from oracle_bind import _dp
    X = part.elm.xyz
    for e in (0, 5, part.sam.nel - 1):
        nodes = part.sam.mmnpc[part.sam.mpmnpc[e] - 1: part.sam.mpmnpc[e + 1] - 1] - 1
        for s in (0, 8):
            sv = oracle.expand(b, Q[:, s])
            v = np.ascontiguousarray(np.stack([sv[3 * n: 3 * n + 3] for n in nodes]).ravel())
            xg, yg, zg = (np.ascontiguousarray(X[nodes, k]) for k in range(3))
            sig = np.zeros(60); eps = np.zeros(60)
            assert oracle.lib.orc_str41(_dp(xg), _dp(yg), _dp(zg), part.elm.emod[e], part.elm.rny[e], 1, _dp(v), _dp(sig), _dp(eps)) == 0
            vm_o = np.array([oracle.von_mises(sig[6 * p: 6 * p + 6]) for p in range(10)])
            p0 = rec.result_point_offsets()[e]
            assert rel_err(vm_g[s, p0:p0 + 10], vm_o) <= TOL
    rec.close()


def test_hex20_block(oracle):
    """20-node hexahedra (type 43): 120 x 60 operator through the shared-memory solid kernel"""
    part = hex20_block(3, 2, 2, ngen=5, seed=6, shuffle_eq=True)
    _check_part(oracle, part, nsteps=27, seed=7)


def _envelope_only_matches(oracle, part, nsteps, seed, step_tile):
    """envelope-only run (the kernels keep the envelope on the von Mises radicand there) against the oracle's history"""
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, nsteps, seed=seed)
    vm_o, mx_o, mn_o = oracle.recover_history(b, Q)
    rec = StressRecovery(part, step_tile=step_tile)
    rec.recover(Q, want_history=False)
    mx, mn = rec.envelope()
    rec.close()
    assert rel_err(mx, mx_o) <= TOL and rel_err(mn, mn_o) <= TOL


def test_hex20_natural_coordinate_kernel(oracle):
    """tiles of 32 steps and more take k2_hex20_steplane_vm_kernel (the 288 non-zero natural derivatives at the nodes with
    compile-time coefficients, J^-1 per result point, lane = time step); curved (jittered) elements, several tiles, the
    ragged last tile (22 steps) on the DMMA gradient kernel; the same run with the kernel switched off agrees to rounding."""
    import os
    part = hex20_block(3, 2, 2, ngen=5, seed=6, shuffle_eq=True)
    vm = _check_part(oracle, part, nsteps=150, seed=7, step_tile=64)
    os.environ["FSR_HEX20_STEPLANE"] = "0"
    try:
        rec = StressRecovery(part, step_tile=64)
        vm2 = rec.recover(reduced_history(part.sam.ndim, 150, seed=7))
        rec.close()
    finally:
        del os.environ["FSR_HEX20_STEPLANE"]
    assert rel_err(vm, vm2) <= 1e-11
    assert not np.array_equal(vm[:64], vm2[:64])   # ... by a different kernel
    assert np.array_equal(vm[128:], vm2[128:])     # the short last tile stays on the gradient kernel
    _envelope_only_matches(oracle, part, nsteps=97, seed=8, step_tile=0)
    # a sheared, rotated block: the inverses are full matrices
    part = hex20_block(2, 2, 2, ngen=4, seed=16, rny=0.45)
    A = np.array([[0.9, 0.3, -0.2], [-0.1, 1.2, 0.4], [0.25, -0.15, 0.8]])
    part.elm.xyz = part.elm.xyz @ A.T + np.array([3.0, -2.0, 1.0])
    _check_part(oracle, part, nsteps=45, seed=9)


@pytest.mark.parametrize("curved", ["none", "all"])
def test_tet10_step_lane_kernel(oracle, curved):
    """tiles of 32 steps and more take k2_tet10_steplane_vm_kernel (lane = time step); cross-check against the
    (corner, step) kernels and the envelope-only mode"""
    import os
    part = tet10_block(3, 3, 2, ngen=5, seed=31, shuffle_eq=True, curved=curved)
    vm = _check_part(oracle, part, nsteps=100, seed=4, step_tile=0)
    os.environ["FSR_TET10_STEPLANE"] = "0"
    try:
        rec = StressRecovery(part)
        vm2 = rec.recover(reduced_history(part.sam.ndim, 100, seed=4))
        rec.close()
    finally:
        del os.environ["FSR_TET10_STEPLANE"]
    assert rel_err(vm, vm2) <= 1e-11
    _envelope_only_matches(oracle, part, nsteps=77, seed=5, step_tile=0)


def test_hex20_gauss_extrapolation(oracle):
    import ctypes as C
    from oracle_bind import _dp, _D
    part = hex20_block(2, 1, 2, ngen=3, seed=9)
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 5, seed=2)
    rec = StressRecovery(part, stress_form=2)
    vm_g = rec.recover(Q)
    oracle.lib.orc_str43.argtypes = [_D, _D, _D, C.c_double, C.c_double, C.c_int, _D, _D, _D]
    X = part.elm.xyz
    for e in range(part.sam.nel):
        nodes = part.sam.mmnpc[part.sam.mpmnpc[e] - 1: part.sam.mpmnpc[e + 1] - 1] - 1
        sv = oracle.expand(b, Q[:, 3])
        v = np.ascontiguousarray(np.stack([sv[3 * n: 3 * n + 3] for n in nodes]).ravel())
        xg, yg, zg = (np.ascontiguousarray(X[nodes, k]) for k in range(3))
        sig = np.zeros(120); eps = np.zeros(120)
        assert oracle.lib.orc_str43(_dp(xg), _dp(yg), _dp(zg), part.elm.emod[e], part.elm.rny[e], 2, _dp(v), _dp(sig), _dp(eps)) == 0
        vm_o = np.array([oracle.von_mises(sig[6 * p: 6 * p + 6]) for p in range(20)])
        p0 = rec.result_point_offsets()[e]
        assert rel_err(vm_g[3, p0:p0 + 20], vm_o) <= TOL
    rec.close()


def test_config1_plate(oracle):
    """BASELINE config 1: 70x70 ANDES quads, 5,041 nodes, 4 external nodes, 10 modes; the oracle
    covers a 40-step sample of the 1,000-step history (it rebuilds every element every step)."""
    part = plate_part(70, 70, ngen=10, seed=1)
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 1000, seed=1)
    rec = StressRecovery(part)
    vm_g = rec.recover(Q)
    mx_g, mn_g = rec.envelope()
    sample = np.arange(0, 1000, 25)
    vm_o, _, _ = oracle.recover_history(b, Q[:, sample], nthreads=8)
    assert rel_err(vm_g[sample], vm_o) <= TOL
    # envelope is the running max/min of the history it produced (max starts at 0, min at huge)
    assert np.array_equal(mx_g, np.maximum(vm_g.max(0), 0.0))
    assert np.array_equal(mn_g, vm_g.min(0))
    rec.close()


def test_envelope_accumulates_and_resets(oracle):
    part = plate_part(5, 5, ngen=3, seed=2)
    Q = reduced_history(part.sam.ndim, 64, seed=9)
    rec = StressRecovery(part)
    vm1 = rec.recover(Q[:, :32])
    vm2 = rec.recover(Q[:, 32:])
    mx, mn = rec.envelope()
    allvm = np.vstack([vm1, vm2])
    assert np.array_equal(mx, allvm.max(0)) and np.array_equal(mn, allvm.min(0))
    rec.reset_envelope()
    rec.recover(Q[:, :1], want_history=False)
    mx, mn = rec.envelope()
    assert np.array_equal(mx, vm1[0]) and np.array_equal(mn, vm1[0])
    rec.close()


def test_degenerate_element_gets_huge(oracle):
    """failed element -> hugeVal results, run continues (stressRoutines.f90:237-241,264-268)"""
    part = plate_part(4, 4, ngen=2, seed=8)
    n = part.sam.mmnpc[part.sam.mpmnpc[5] - 1: part.sam.mpmnpc[6] - 1]
    part.elm.xyz[n[2] - 1] = part.elm.xyz[n[0] - 1]   # collapse a diagonal: zero normal
    part.elm.xyz[n[3] - 1] = part.elm.xyz[n[1] - 1]
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 4, seed=1)
    vm_o, _, _ = oracle.recover_history(b, Q)
    rec = StressRecovery(part)
    assert rec.n_failed >= 1
    vm_g = rec.recover(Q)
    bad = vm_o >= 1e300
    assert bad.any() and np.array_equal(bad, vm_g >= 1e300)
    assert rel_err(vm_g[~bad], vm_o[~bad]) <= TOL
    rec.close()


def test_fedempy_part_state_entry_points(oracle):
    """getPartStressStateSize / savePartStressState / savePartDeformationState (solverInterface.C:940-1001) over a
    registered part: header [step, time, dt, baseId], then vms = [iel, nenod, nstrp, vm...] per element with stress
    points / 3 deformations per node."""
    import ctypes as C
    from fedem_solvers_b200 import _lib
    lib = _lib.load_library()
    part = plate_part(6, 5, ngen=4, seed=9, tri_fraction=0.4)
    b = oracle.bind_part(part)
    rec = StressRecovery(part)
    Q = reduced_history(part.sam.ndim, 3, seed=7)
    minex = np.ascontiguousarray(part.sam.minex, np.int32).copy()
    minex[2] = -3   # an internal beam-pin node: no output
    assert lib.fsr_recovery_register(31, rec._h, minex.ctypes.data_as(_lib._I)) == 0
    assert lib.getPartStressStateSize(32) == -1
    nd, ns = lib.getPartDeformationStateSize(31), lib.getPartStressStateSize(31)
    assert nd == 3 * part.sam.nnod + 4 and ns == 4 + sum(3 + n for n in part.nstrp() if n > 0)
    q = np.ascontiguousarray(Q[:, 2])
    assert lib.fsr_recovery_update(31, 12, 0.375, 0.005, q.ctypes.data_as(_lib._D)) == 0
    d, s = np.zeros(nd), np.zeros(ns)
    assert lib.savePartDeformationState(31, d.ctypes.data_as(_lib._D), nd) and lib.savePartStressState(31, s.ctypes.data_as(_lib._D), ns)
    assert not lib.savePartStressState(31, s.ctypes.data_as(_lib._D), ns - 1)
    assert list(d[:4]) == [12.0, 0.375, 0.005, 31.0] and list(s[:4]) == [12.0, 0.375, 0.005, 31.0]
    sv = oracle.expand(b, q)
    ref = oracle.calc_stresses(b, sv)
    want = np.stack([sv[part.sam.madof[:-1] - 1 + k] for k in range(3)], 1)
    want[2] = 0.0
    assert np.abs(d[4:].reshape(-1, 3) - want).max() <= TOL * np.abs(sv).max()
    k, scale = 4, np.abs(ref["resmat"][:, 0]).max()
    for e in range(part.sam.nel):
        n = int(part.nstrp()[e])
        if n == 0:
            continue
        assert (s[k], s[k + 2]) == (e + 1, n)
        p0 = b["ptoff"][e]
        assert np.abs(s[k + 3:k + 3 + n] - ref["resmat"][p0:p0 + n, 0]).max() <= TOL * scale
        k += 3 + n
    assert k == ns
    # the loop over the parts of a mechanism in one call (stressRecoveryModule.f90:1021-1061): a second part, both queued
    # before either is waited for; every part keeps its own state, and the running envelope has taken the steps along
    part2 = tet10_block(2, 2, 2, ngen=3, seed=10, n_beams=3, curved="surface")
    b2 = oracle.bind_part(part2)
    rec2 = StressRecovery(part2)
    assert lib.fsr_recovery_register(44, rec2._h, None) == 0
    q1, q2 = np.ascontiguousarray(Q[:, 1]), np.ascontiguousarray(reduced_history(part2.sam.ndim, 2, seed=8)[:, 1])
    ids = np.array([31, 44], np.int32)
    qs = (_lib._D * 2)(q1.ctypes.data_as(_lib._D), q2.ctypes.data_as(_lib._D))
    assert lib.fsr_recovery_update_parts(2, ids.ctypes.data_as(_lib._I), 13, 0.38, 0.005, qs) == 0
    ns2 = lib.getPartStressStateSize(44)
    s2 = np.zeros(ns2)
    assert lib.savePartStressState(31, s.ctypes.data_as(_lib._D), ns) and lib.savePartStressState(44, s2.ctypes.data_as(_lib._D), ns2)
    assert list(s[:4]) == [13.0, 0.38, 0.005, 31.0] and list(s2[:4]) == [13.0, 0.38, 0.005, 44.0]
    for bb, qq, ss, pp in ((b, q1, s, part), (b2, q2, s2, part2)):
        r = oracle.calc_stresses(bb, oracle.expand(bb, qq))
        vm = np.concatenate([ss[k + 3:k + 3 + int(ss[k + 2])] for k in _vms_offsets(ss)])
        assert np.abs(vm - r["resmat"][:, 0]).max() <= TOL * np.abs(r["resmat"][:, 0]).max()
    mx, _ = rec.envelope()
    want_mx = np.maximum(ref["resmat"][:, 0], oracle.calc_stresses(b, oracle.expand(b, q1))["resmat"][:, 0])
    assert np.abs(mx - want_mx).max() <= TOL * np.abs(want_mx).max()
    assert lib.fsr_recovery_unregister(31) == 0 and lib.fsr_recovery_unregister(44) == 0
    rec.close(); rec2.close()


def _vms_offsets(s):
    k = 4
    while k < len(s):
        yield k
        k += 3 + int(s[k + 2])


def test_linear_solids_hex8_tet4_wedg6(oracle):
    """types 44, 45, 46: gradient-form von Mises kernel + dense full-result operator against the oracle"""
    part = linsolid_block(3, 2, 2, ngen=5, seed=8, shuffle_eq=True)
    assert {44, 45, 46} <= set(int(t) for t in part.sam.melcon)
    _check_part(oracle, part, nsteps=70, seed=3)
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 2, seed=5)
    rec = StressRecovery(part)
    full = rec.calc_stresses(Q[:, 1])
    ref = oracle.calc_stresses(b, oracle.expand(b, Q[:, 1]))
    for k, tol in (("stress", TOL), ("strain", TOL), ("resmat", 1e-9)):
        for c in range(ref[k].shape[1]):
            sc = np.abs(ref[k][:, c % 4::4] if k == "resmat" else ref[k]).max()
            assert np.abs(full[k][:, c] - ref[k][:, c]).max() <= tol * sc, (k, c)
    rec.close()


@pytest.mark.parametrize("form", [1, 2])
def test_linear_solids_gauss_point_forms(oracle, form):
    """-stressForm 1 (HEX8 volume average, WEDG6 mid-plane scheme) and 2 (extrapolation from the Gauss points)"""
    import ctypes as C
    from oracle_bind import _dp
    part = linsolid_block(2, 2, 2, ngen=4, seed=9, kinds=(44, 46))
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 9, seed=2)
    rec = StressRecovery(part, stress_form=form)
    vm_g = rec.recover(Q)
    X = part.elm.xyz
    off = rec.result_point_offsets()
    for e in range(part.sam.nel):
        t = int(part.sam.melcon[e])
        nodes = part.sam.mmnpc[part.sam.mpmnpc[e] - 1: part.sam.mpmnpc[e + 1] - 1] - 1
        nn = len(nodes)
        for s in (0, 8):
            sv = oracle.expand(b, Q[:, s])
            v = np.ascontiguousarray(np.stack([sv[3 * n: 3 * n + 3] for n in nodes]).ravel())
            xg, yg, zg = (np.ascontiguousarray(X[nodes, k]) for k in range(3))
            sig, eps = np.zeros(6 * nn), np.zeros(6 * nn)
            fn = oracle.lib.orc_str44 if t == 44 else oracle.lib.orc_str46
            assert fn(_dp(xg), _dp(yg), _dp(zg), C.c_double(part.elm.emod[e]), C.c_double(part.elm.rny[e]), form, _dp(v), _dp(sig), _dp(eps)) == 0
            vm_o = np.array([oracle.von_mises(sig[6 * p: 6 * p + 6]) for p in range(nn)])
            assert rel_err(vm_g[s, off[e]:off[e] + nn], vm_o) <= TOL, (t, e, s)
    rec.close()


@pytest.mark.parametrize("form", [0, 2])
def test_wedg15_block(oracle, form):
    """type 42: two-block gradient kernel + dense operator; -stressForm 2 = 3 x 2 Gauss points extrapolated (REAL*4 abscissa)"""
    import ctypes as C
    from oracle_bind import _dp
    part = wedg15_block(2, 2, 1, ngen=5, seed=10, shuffle_eq=True)
    b = oracle.bind_part(part)
    if form == 0:
        _check_part(oracle, part, nsteps=40, seed=4)
        Q = reduced_history(part.sam.ndim, 2, seed=5)
        rec = StressRecovery(part)
        full = rec.calc_stresses(Q[:, 1])
        ref = oracle.calc_stresses(b, oracle.expand(b, Q[:, 1]))
        for k, tol in (("stress", TOL), ("strain", TOL), ("resmat", 1e-9)):
            assert np.abs(full[k] - ref[k]).max() <= tol * np.abs(ref[k]).max() * (1 if k != "resmat" else 1), k
        rec.close()
        return
    Q = reduced_history(part.sam.ndim, 9, seed=2)
    rec = StressRecovery(part, stress_form=form)
    vm_g = rec.recover(Q)
    off = rec.result_point_offsets()
    X = part.elm.xyz
    for e in (0, 3, part.sam.nel - 1):
        nodes = part.sam.mmnpc[part.sam.mpmnpc[e] - 1: part.sam.mpmnpc[e + 1] - 1] - 1
        for s in (0, 8):
            sv = oracle.expand(b, Q[:, s])
            v = np.ascontiguousarray(np.stack([sv[3 * n: 3 * n + 3] for n in nodes]).ravel())
            xg, yg, zg = (np.ascontiguousarray(X[nodes, k]) for k in range(3))
            sig, eps = np.zeros(90), np.zeros(90)
            assert oracle.lib.orc_str42(_dp(xg), _dp(yg), _dp(zg), C.c_double(part.elm.emod[e]), C.c_double(part.elm.rny[e]), form,
                                        _dp(v), _dp(sig), _dp(eps)) == 0
            vm_o = np.array([oracle.von_mises(sig[6 * p: 6 * p + 6]) for p in range(15)])
            assert rel_err(vm_g[s, off[e]:off[e] + 15], vm_o) <= TOL, (e, s)
    rec.close()


@pytest.mark.parametrize("kinds", [(31, 32), (32,), (31,)])
def test_thick_shells(oracle, kinds):
    """types 31 / 32 (STR31 / STR32): dense stress + strain operators folded from the SCTS32 / SCQS32 stress matrices;
    the oracle side of this pair is pinned by the reference's testThickShell.pf cases (tests/test_thickshell_cpu.py)"""
    part = thickshell_panel(4, 3, ngen=5, seed=12, kinds=kinds, shuffle_eq=True)
    _check_part(oracle, part, nsteps=70, seed=6)


def test_thick_shells_with_normals_along_global_x(oracle):
    """panel turned so that its normals lie along X: LNCS30 cannot put x' in the z'-X plane and falls through to the
    z'-Y rule (scts.f:486-505), for the node systems and the sampling points alike"""
    rot = np.array([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])
    part = thickshell_panel(3, 3, ngen=4, seed=15, rotation=rot, curvature=(1e-5, 2e-5, 0.0))
    _check_part(oracle, part, nsteps=12, seed=7)


def test_thick_shell_with_bad_midside_node_gets_huge(oracle):
    """CHQA30: a mid-side node at 1/5 of its edge fails the element (hugeVal), the others are unaffected"""
    part = thickshell_panel(3, 2, ngen=3, seed=16, kinds=(32,))
    n = part.sam.mmnpc[part.sam.mpmnpc[2] - 1: part.sam.mpmnpc[3] - 1] - 1
    X = part.elm.xyz
    X[n[1]] = X[n[0]] + 0.2 * (X[n[2]] - X[n[0]])
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 6, seed=2)
    vm_o, _, _ = oracle.recover_history(b, Q)
    rec = StressRecovery(part)
    assert rec.n_failed >= 1
    vm_g = rec.recover(Q)
    bad = vm_o >= 1e300
    assert bad.any() and not bad.all() and np.array_equal(bad, vm_g >= 1e300)
    assert rel_err(vm_g[~bad], vm_o[~bad]) <= TOL
    rec.close()


def test_legacy_shells_default_formulations(oracle):
    """types 21 / 22 (part reduced without the ANDES formulation): with the default -fftStressForm 1 / -ffqStressForm 2 the
    reference's STR21 / STR22 run the statements of STR23 / STR24; any other formulation gives those elements no results"""
    part = plate_part(6, 5, ngen=4, seed=17, tri_fraction=0.4, warp=0.02)
    ref_vm = _check_part(oracle, part, nsteps=20, seed=8)
    legacy = plate_part(6, 5, ngen=4, seed=17, tri_fraction=0.4, warp=0.02)
    legacy.sam.melcon = legacy.sam.melcon - 2
    assert set(np.unique(legacy.sam.melcon)) == {21, 22}
    vm = _check_part(oracle, legacy, nsteps=20, seed=8)
    assert np.array_equal(vm, ref_vm)
    rec = StressRecovery(legacy, ffq_stress_form=0)      # FFQ quads drop out (STR22b is not built), the FFT triangles stay
    assert rec.npts == 6 * int((legacy.sam.melcon == 21).sum())
    rec.close()
    # -ffqStressForm 1: STR22a with one Gauss point, the centroid strain at every node
    oracle.lib.orc_set_ffq_stress_form(1)
    try:
        b = oracle.bind_part(legacy)
        Q = reduced_history(legacy.sam.ndim, 9, seed=9)
        vm_o, mx_o, mn_o = oracle.recover_history(b, Q)
        rec = StressRecovery(legacy, ffq_stress_form=1)
        vm_g = rec.recover(Q)
        assert rel_err(vm_g, vm_o) <= TOL
        rec2 = StressRecovery(legacy)
        assert rel_err(vm_g, rec2.recover(Q)) > 1e-3         # and it is a different formulation than the 2 x 2 one
        rec2.close()
        full = rec.calc_stresses(Q[:, 2])
        ref = oracle.calc_stresses(b, oracle.expand(b, Q[:, 2]))
        for key in ("stress", "strain", "sres"):
            assert rel_err(full[key], ref[key]) <= TOL, key
        rec.close()
    finally:
        oracle.lib.orc_set_ffq_stress_form(2)


def test_legacy_fft_shell_private_formulations(oracle):
    """type 21 with -fftStressForm 0 / 2: STR21 on FTS31 / FTS32 (Bergan / Felippa membrane triangle of tmrf.f + TEBA bending), the
    reference's statements including its two oddities (plane stress matrix where the membrane rigidity is expected, HH columns added
    in their own order); von Mises history, stresses, strains and stress resultants against the oracle's restatement"""
    legacy = plate_part(6, 5, ngen=4, seed=23, tri_fraction=1.0, warp=0.02)
    legacy.sam.melcon = legacy.sam.melcon - 2
    assert set(np.unique(legacy.sam.melcon)) == {21}
    Q = reduced_history(legacy.sam.ndim, 11, seed=10)
    rec1 = StressRecovery(legacy)
    vm_default = rec1.recover(Q)
    rec1.close()
    for form in (0, 2):
        oracle.lib.orc_set_fft_stress_form(form)
        try:
            b = oracle.bind_part(legacy)
            vm_o, mx_o, mn_o = oracle.recover_history(b, Q)
            rec = StressRecovery(legacy, fft_stress_form=form)
            assert rec.npts == 6 * legacy.sam.nel
            vm_g = rec.recover(Q)
            assert rel_err(vm_g, vm_o) <= TOL
            assert rel_err(vm_g, vm_default) > 1e-3        # a different formulation than FTSA31 / FTSA32
            full = rec.calc_stresses(Q[:, 3])
            ref = oracle.calc_stresses(b, oracle.expand(b, Q[:, 3]))
            for key in ("stress", "strain", "sres"):
                assert rel_err(full[key], ref[key]) <= TOL, key
            rec.close()
        finally:
            oracle.lib.orc_set_fft_stress_form(1)
    # next to ANDES triangles in one part the private formulation is not served: those FFT shells get no results
    mixed = plate_part(6, 5, ngen=4, seed=23, tri_fraction=1.0, warp=0.02)
    mixed.sam.melcon[::2] = 21
    rec = StressRecovery(mixed, fft_stress_form=0)
    assert rec.npts == 6 * int((mixed.sam.melcon == 23).sum())
    rec.close()


def test_edge_sizes(oracle):
    """ragged / empty inputs: one step, zero steps, no component modes, a part whose elements are all outside the selection"""
    part = plate_part(5, 4, ngen=0, seed=18, tri_fraction=0.5)            # no generalized DOFs: E matrix absent
    assert part.sam.ngen == 0
    _check_part(oracle, part, nsteps=3, seed=2)
    part = plate_part(5, 4, ngen=3, seed=19)
    b = oracle.bind_part(part)
    rec = StressRecovery(part)
    Q = reduced_history(part.sam.ndim, 1, seed=3)
    vm_o, mx_o, mn_o = oracle.recover_history(b, Q)
    vm_g = rec.recover(Q)                                                  # a single step
    assert vm_g.shape == (1, rec.npts) and rel_err(vm_g, vm_o) <= TOL
    rec.reset_envelope()
    vm0 = rec.recover(Q[:, :0])                                            # no step at all: nothing happens
    assert vm0.shape == (0, rec.npts)
    mx, mn = rec.envelope()
    assert (mx == 0.0).all() and (mn > 1e300).all()                        # untouched: max from 0, min from hugeVal
    rec.close()
    part.elm.elmid = -np.abs(part.elm.elmid)                               # every element outside the -group selection
    rec = StressRecovery(part)
    assert rec.npts == 0
    assert rec.recover(Q).shape == (1, 0)
    U = rec.calc_int_displacements(Q)                                      # the expansion still works
    assert rel_err(U[0], oracle.expand(oracle.bind_part(part), Q[:, 0])) <= TOL
    rec.close()


def test_expand_rows(oracle):
    """fsr_expand_rows: the expansion of selected DOFs only (rosette nodes) equals those entries of the full expansion"""
    import ctypes as C
    from oracle_bind import _dp, _ip
    part = plate_part(6, 6, ngen=4, seed=20, tri_fraction=0.2)
    b = oracle.bind_part(part)
    rec = StressRecovery(part, step_tile=64)
    Q = np.asfortranarray(reduced_history(part.sam.ndim, 150, seed=4))          # three device tiles, the last one ragged
    rows = np.array([0, 5, 17, part.sam.ndof - 1, 17, 40], np.int32)
    out = np.zeros((150, len(rows)))
    rc = rec._lib.fsr_expand_rows(rec._h, _dp(Q), Q.shape[0], 150, _ip(rows), len(rows), _dp(out))
    assert rc == 0
    for s_ in (0, 63, 64, 149):
        sv = oracle.expand(b, Q[:, s_])
        assert np.abs(out[s_] - sv[rows]).max() <= TOL * np.abs(sv).max()
    bad = np.array([part.sam.ndof], np.int32)
    assert rec._lib.fsr_expand_rows(rec._h, _dp(Q), Q.shape[0], 1, _ip(bad), 1, _dp(out)) < 0
    rec.close()


def _impose_field(part, field):
    """Makes the nodal field `field` [nnod, 6] the expansion of the part: B = 0, one component mode whose shape is the
    field at the internal DOFs, the external DOFs carry the field themselves.  Returns q = [finit; 1]."""
    sam = part.sam
    f = np.asarray(field, np.float64).ravel()
    eq2dof = np.zeros(sam.neq + 1, np.int64)
    nz = np.nonzero(sam.meqn > 0)[0]
    eq2dof[sam.meqn[nz]] = nz
    part.B = np.zeros((sam.ndof1, sam.ndof2), order="F")
    part.E = np.asfortranarray(f[eq2dof[sam.meqn1]].reshape(-1, 1))
    sam.ngen = 1
    rec = StressRecovery(part)
    probe = np.zeros((sam.ndim, sam.ndof2), order="F")
    probe[:sam.ndof2] = np.eye(sam.ndof2)
    U = rec.calc_int_displacements(probe)              # row j: unit external DOF j of finit
    ext = U.argmax(1)
    assert np.array_equal(np.sort(ext), np.nonzero(sam.msc == 2)[0]) and np.allclose(U.max(1), 1.0)
    q = np.concatenate([f[ext], [1.0]])
    Uf = rec.calc_int_displacements(q.reshape(-1, 1))[0]
    assert np.abs(Uf - f).max() <= 1e-15 * np.abs(f).max()
    return rec, q


def test_closed_form_twist_and_shear_on_the_device():
    """not via the oracle: a plate of ANDES quads, turned and moved in space, under (a) pure twist w = kappa x y and (b)
    in-plane shear -- the device's von Mises is sqrt(3) G t kappa resp. sqrt(3) G gamma at EVERY result point; the full
    result set gives the principal stresses +-G t kappa (+-G gamma), max shear G t kappa, zero trace"""
    from scipy.spatial.transform import Rotation
    t, E, nu = 0.012, 2.1e11, 0.3
    G = E / (2 * (1 + nu))
    R = Rotation.from_rotvec([0.4, -0.7, 0.2]).as_matrix()
    for case in ("twist", "shear"):
        part = plate_part(6, 5, ngen=1, seed=41, jitter=0.25, thickness=t, emod=E, rny=nu, with_recovery=False)
        X0 = part.elm.xyz.copy()
        kap, gam = 3.0e-2, 1.0e-3
        fld = np.zeros((part.sam.nnod, 6))
        if case == "twist":
            fld[:, 2] = kap * X0[:, 0] * X0[:, 1]
            fld[:, 3], fld[:, 4] = kap * X0[:, 0], -kap * X0[:, 1]
            want = np.sqrt(3.0) * G * t * kap
        else:
            fld[:, 0], fld[:, 1] = 0.5 * gam * X0[:, 1], 0.5 * gam * X0[:, 0]
            want = np.sqrt(3.0) * G * gam
        part.elm.xyz = X0 @ R.T + np.array([2.0, -1.0, 0.5])
        fld = np.hstack([fld[:, :3] @ R.T, fld[:, 3:] @ R.T])
        rec, q = _impose_field(part, fld)
        vm = rec.recover(np.tile(q.reshape(-1, 1), (1, 9)))
        assert np.abs(vm - want).max() <= 1e-9 * want, (case, np.abs(vm - want).max() / want)
        # the output axes (global X projected on the turned plate) are not the plate's own: compare invariants.
        # Pure shear tau: principal stresses +tau / -tau, max shear tau, trace 0
        full = rec.calc_stresses(q)
        tau = want / np.sqrt(3.0)
        assert np.abs(full["resmat"][:, 1] - tau).max() <= 1e-8 * tau and np.abs(full["resmat"][:, 2] + tau).max() <= 1e-8 * tau
        assert np.abs(full["resmat"][:, 3] - tau).max() <= 1e-8 * tau
        assert np.abs(full["stress"][:, 0] + full["stress"][:, 1]).max() <= 1e-9 * tau
        rec.close()
