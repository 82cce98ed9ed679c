"""Thick shells (types 31, 32): the reference's own known-answer test, re-expressed against the oracle.

src/vpmStress/vpmStressTests/testThickShell.pf is the one unit-level golden the reference holds for an element stress
routine: eight calls of ElStress on a unit T6 / Q8 (geometry, E = 2.1e11, nu = 0.3, t = 0.01 from vpmStressTests/ffl.f90:27-96)
with prescribed nodal displacements, asserting the strain components at every result point to 1.0e-15.  The cases below
carry the same displacements, the same expectations and the same tolerance; ElStress's tensorial-shear conversion
(elStressModule.f90:244-253) is applied the way the oracle's element loop does it."""
import ctypes as C
import os
import numpy as np
import pytest

from oracle_bind import Oracle, _dp

E, NU, THK = 2.1e11, 0.3, 0.01
TOL = 1.0e-15


def geometry(ieltyp):
    """ffl_getcoor of vpmStressTests/ffl.f90:27-47 for iel = 1 (T6) and iel = 2 (Q8)."""
    if ieltyp == 31:
        x, y = np.zeros(6), np.zeros(6)
        x[1] = 1.0; y[2] = 1.0; x[3:5] = 0.5; y[4:6] = 0.5
    else:
        x, y = np.zeros(8), np.zeros(8)
        x[1] = 0.5; x[2:5] = 1.0; x[5] = 0.5; y[3] = 0.5; y[4:7] = 1.0; y[7] = 0.5
    return x, y, np.zeros_like(x)


def el_stress(o, ieltyp, V):
    """ElStress for one thick shell: V (6, nenod) Fortran order.  Returns Strain (nstrp, 6) as ElStress leaves it."""
    nenod = 6 if ieltyp == 31 else 8
    nstrp = 2 * nenod
    x, y, z = geometry(ieltyp)
    thk = np.full(nenod, THK)
    ev = np.ascontiguousarray(V.T.reshape(-1))          # EV(6*(n-1)+d)
    sig, eps = np.zeros(6 * nstrp), np.zeros(6 * nstrp)
    f = o.lib.orc_str31 if ieltyp == 31 else o.lib.orc_str32
    f.restype = C.c_int
    rc = f(_dp(x), _dp(y), _dp(z), C.c_double(E), C.c_double(NU), _dp(thk), _dp(ev), _dp(sig), _dp(eps))
    assert rc == 0
    eps = eps.reshape(nstrp, 6).copy()
    eps[:, 3:] *= 0.5
    return sig.reshape(nstrp, 6), eps


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def check(eps, exx, eyy, exy):
    for i in range(eps.shape[0]):
        assert abs(eps[i, 0] - exx(i + 1)) <= TOL, f"eps_xx strain in point #{i + 1}"
        assert abs(eps[i, 1] - eyy(i + 1)) <= TOL, f"eps_yy strain in point #{i + 1}"
        assert abs(eps[i, 3] - exy(i + 1)) <= TOL, f"eps_xy strain in point #{i + 1}"


def const(v):
    return lambda i: v


def test_q8_uniform_stretch_x(oracle):          # testThickShell.pf:55-75
    V = np.zeros((6, 8))
    V[0, 2:5] = 0.01; V[0, 1] = 0.005; V[0, 5] = 0.005
    check(el_stress(oracle, 32, V)[1], const(0.01), const(0.0), const(0.0))


def test_q8_uniform_stretch_y(oracle):          # :77-97
    V = np.zeros((6, 8))
    V[1, 4:7] = 0.01; V[1, 3] = 0.005; V[1, 7] = 0.005
    check(el_stress(oracle, 32, V)[1], const(0.0), const(0.01), const(0.0))


def test_q8_constant_shear(oracle):             # :99-125
    V = np.zeros((6, 8))
    V[0, 4:7] = 0.01; V[0, 3] = 0.005; V[0, 7] = 0.005
    V[1, 2:5] = 0.01; V[1, 1] = 0.005; V[1, 5] = 0.005
    check(el_stress(oracle, 32, V)[1], const(0.0), const(0.0), const(0.01))


def test_q8_linear_stretch_x(oracle):           # :127-154
    V = np.zeros((6, 8))
    V[0, 2:5] = 0.04; V[0, 1] = 0.01; V[0, 5] = 0.01

    def exx(i):
        j = (i - 1) % 8 + 1
        return 0.08 if 3 <= j <= 5 else 0.04 if j in (2, 6) else 0.0
    check(el_stress(oracle, 32, V)[1], exx, const(0.0), const(0.0))


def test_t6_uniform_stretch_x(oracle):          # :169-186
    V = np.zeros((6, 6))
    V[0, 1] = 0.01; V[0, 3] = 0.005; V[0, 4] = 0.005
    check(el_stress(oracle, 31, V)[1], const(0.01), const(0.0), const(0.0))


def test_t6_uniform_stretch_y(oracle):          # :188-205
    V = np.zeros((6, 6))
    V[1, 2] = 0.01; V[1, 4] = 0.005; V[1, 5] = 0.005
    check(el_stress(oracle, 31, V)[1], const(0.0), const(0.01), const(0.0))


def test_t6_constant_shear(oracle):             # :207-228
    V = np.zeros((6, 6))
    V[0, 2] = 0.01; V[0, 4] = 0.005; V[0, 5] = 0.005
    V[1, 1] = 0.01; V[1, 3] = 0.005; V[1, 4] = 0.005
    check(el_stress(oracle, 31, V)[1], const(0.0), const(0.0), const(0.01))


def test_t6_linear_stretch_y(oracle):           # :230-256
    V = np.zeros((6, 6))
    V[1, 2] = 0.04; V[1, 4] = 0.01; V[1, 5] = 0.01

    def eyy(i):
        j = (i - 1) % 6 + 1
        return 0.08 if j == 3 else 0.04 if j >= 5 else 0.0
    check(el_stress(oracle, 31, V)[1], const(0.0), eyy, const(0.0))


def test_rotate3d_matches_the_reference_library(oracle):
    """orc_rotate3d against FFaTensorTransforms::rotate3D compiled unmodified from the reference (oracle/_ref)."""
    ref = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libfedem_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    L = C.CDLL(ref)
    f = getattr(L, "_ZN19FFaTensorTransforms8rotate3DEPKdS1_Pd")
    rng = np.random.default_rng(31)
    for _ in range(50):
        S = rng.standard_normal(6)
        Q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        rot = np.ascontiguousarray(Q.T.reshape(-1))       # column-major: rot[0:3] = first column
        a, b = np.zeros(6), np.zeros(6)
        f(_dp(S), _dp(rot), _dp(b))
        oracle.lib.orc_rotate3d(_dp(S), _dp(rot), _dp(a))
        assert np.array_equal(a, b)


def test_thick_shell_physics_on_a_curved_patch(oracle):
    """Rigid-body translation gives zero stress on a curved, tilted Q8 / T6; a degenerate mid-side node is reported."""
    rng = np.random.default_rng(32)
    for ieltyp in (31, 32):
        nenod = 6 if ieltyp == 31 else 8
        x, y, z = geometry(ieltyp)
        z = 0.15 * x * x + 0.1 * y * y + 0.05 * x * y       # shallow paraboloid
        Q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        P = np.stack([x, y, z], 1) @ Q.T
        x, y, z = (np.ascontiguousarray(P[:, k]) for k in range(3))
        thk = np.full(nenod, 0.02)
        V = np.zeros((nenod, 6)); V[:, :3] = rng.standard_normal(3)
        sig, eps = np.zeros(12 * nenod), np.zeros(12 * nenod)
        f = oracle.lib.orc_str31 if ieltyp == 31 else oracle.lib.orc_str32
        assert f(_dp(x), _dp(y), _dp(z), C.c_double(E), C.c_double(NU), _dp(thk), _dp(V.reshape(-1)), _dp(sig), _dp(eps)) == 0
        assert np.abs(sig).max() < 1e-4 * E * 1e-9
        # mid-side node moved to 1/5 of its edge: CHQT30 / CHQA30 reject the element
        k = 3 if ieltyp == 31 else 1
        a, b = (0, 1) if ieltyp == 31 else (0, 2)
        for c in (x, y, z):
            c[k] = c[a] + 0.2 * (c[b] - c[a])
        assert f(_dp(x), _dp(y), _dp(z), C.c_double(E), C.c_double(NU), _dp(thk), _dp(V.reshape(-1)), _dp(sig), _dp(eps)) == 1
