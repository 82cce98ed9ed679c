#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference's OWN C++ (oracle/_ref/libfedem_ref.so, which
oracle/Makefile compiles unmodified from /root/reference): tensor invariants
(FFaTensorTransforms.C, FFaMath.C) and fatigue (FFpFatigue.C, FFpCycle.C, FFpSNCurve.C).
Run in the build container (the reference sources do not exist on the GPU box):

    make -C oracle ref && python tools/make_golden.py

The fixtures pin the oracle (tests/test_oracle_cpu.py) and the CUDA kernels (tests/test_gpu_*.py)
to outputs of the reference itself."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_bind  # noqa: E402

CURVE = np.array([15.117, 17.146, 4.0, 5.0])  # loga1, loga2, m1 (gagemain.C defaults), m2 (FFpSNCurve.H)


def fatigue_series(rng):
    """Adversarial + realistic scalar histories."""
    out = []
    n = 400
    t = np.arange(n)
    out.append(("narrow_band", 150 + 120 * np.sin(0.31 * t) * np.cos(0.013 * t) + rng.normal(0, 8, n), 25.0))
    out.append(("random_walk", np.cumsum(rng.normal(0, 30, n)), 25.0))
    out.append(("white", rng.uniform(0, 300, n), 25.0))
    out.append(("white_gate0", rng.uniform(-1, 1, 257), 0.0))
    out.append(("integers_plateaus", rng.integers(0, 6, n).astype(float) * 20.0, 25.0))
    out.append(("integers_equal_ranges", np.tile([0.0, 100.0, 0.0, 100.0, 50.0, 100.0, 0.0], 30), 25.0))
    out.append(("monotone_up", np.linspace(0, 500, 100), 25.0))
    out.append(("monotone_down", np.linspace(500, 0, 100), 25.0))
    out.append(("below_gate", 100 + rng.uniform(-10, 10, 300), 25.0))
    out.append(("constant", np.full(50, 42.0), 25.0))
    out.append(("two_points", np.array([0.0, 100.0]), 25.0))
    out.append(("three_points", np.array([0.0, 100.0, 20.0]), 25.0))
    out.append(("one_point", np.array([7.0]), 25.0))
    out.append(("sawtooth_growing", np.array([(-1) ** i * (10 + 3 * i) for i in range(120)], float), 25.0))
    out.append(("sawtooth_shrinking", np.array([(-1) ** i * (400 - 3 * i) for i in range(120)], float), 25.0))
    out.append(("negative_only", -200 + 150 * np.sin(0.7 * t) + rng.normal(0, 20, n), 25.0))
    out.append(("big_gate", rng.uniform(0, 300, n), 200.0))
    out.append(("long", 150 + 100 * np.sin(0.2 * np.arange(5000)) + rng.normal(0, 40, 5000), 25.0))
    for k in range(12):
        m = int(rng.integers(5, 600))
        x = rng.normal(0, 1, m).cumsum() * rng.uniform(5, 60) + rng.normal(0, rng.uniform(0, 30), m)
        if k % 3 == 0:
            x = np.round(x / 10.0) * 10.0  # ties
        out.append((f"mixed{k}", x, float(rng.choice([0.0, 5.0, 25.0, 60.0]))))
    return out


def main():
    ref = oracle_bind.Reference()
    assert ref.available, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20261017)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    # ---- invariants ------------------------------------------------------------------------
    S2 = [rng.normal(0, 1e8, 3) for _ in range(200)]
    S2 += [np.array([1.0, 1.0, 0.0]), np.zeros(3), np.array([5e7, -5e7, 0.0]), np.array([3.0, 3.0, 1e-9]),
           np.array([2e8, 2e8, 0.0]), np.array([-1e-3, 4e-3, 2e-3])]
    S3 = [rng.normal(0, 1e8, 6) for _ in range(200)]
    S3 += [np.array([1.0, 1.0, 1.0, 0, 0, 0]), np.zeros(6), np.array([3e8, 2e8, 1e8, 0, 0, 0]),
           np.array([1e8, 1e8, 5e7, 0, 0, 0]), np.array([1.0, 2.0, 3.0, 1e-12, 0, 0]),
           np.array([5e7, 5e7, 5e7, 1e7, 1e7, 1e7]), np.array([1e-4, -2e-4, 3e-4, 5e-5, -1e-5, 2e-5])]
    S3 += [rng.normal(0, 1, 6) * np.array([1, 1, 1, 1e-9, 1e-9, 1e-9]) for _ in range(20)]
    S2 = np.array(S2); S3 = np.array(S3)
    vm2 = np.array([ref.von_mises(s) for s in S2]); vm3 = np.array([ref.von_mises(s) for s in S3])
    p2 = [ref.principal(s) for s in S2]; p3 = [ref.principal(s) for s in S3]
    np.savez(os.path.join(gold, "invariants_ref.npz"), S2=S2, S3=S3, vm2=vm2, vm3=vm3,
             ok2=np.array([o for o, _ in p2]), P2=np.array([p for _, p in p2]),
             ok3=np.array([o for o, _ in p3]), P3=np.array([p for _, p in p3]))

    # ---- fatigue -----------------------------------------------------------------------------
    names, gates, data, doff = [], [], [], [0]
    turns, toff, cyc, coff, dmg, ncyc, bins, rf_ok = [], [0], [], [0], [], [], [], []
    BIN, NB = 10.0, 64
    for name, x, gate in fatigue_series(rng):
        x = np.ascontiguousarray(x, np.float64)
        names.append(name); gates.append(gate); data.append(x); doff.append(doff[-1] + len(x))
        tp = ref.pvx(x, gate)
        turns.append(tp); toff.append(toff[-1] + len(tp))
        c = ref.rainflow(tp, gate)
        ok = c is not None
        c = c if ok else np.zeros((0, 2))
        cyc.append(c); coff.append(coff[-1] + len(c))
        d, n = ref.get_damage(x, gate, CURVE)
        dmg.append(d); ncyc.append(n); rf_ok.append(ok)  # ffp_getdamage ignores the closing failure
        b = np.zeros(NB, np.int32)
        for k in range(NB):
            v = ref.num_cycles(k * BIN, (k + 1) * BIN)
            b[k] = v  # -1 once the bin lies beyond the largest range (or no cycles at all)
        bins.append(b)
    np.savez(os.path.join(gold, "fatigue_ref.npz"), names=np.array(names), gates=np.array(gates),
             data=np.concatenate(data), doff=np.array(doff), turns=np.concatenate(turns), toff=np.array(toff),
             cycles=np.concatenate(cyc), coff=np.array(coff), damage=np.array(dmg), ncycles=np.array(ncyc),
             bins=np.array(bins), rf_ok=np.array(rf_ok), curve=CURVE, bin_size=BIN)
    print("wrote", os.listdir(gold))


if __name__ == "__main__":
    main()
