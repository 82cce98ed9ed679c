"""GPU parity of K3 (PVX + rainflow + damage + cycle histogram) through the C ABI against golden
vectors produced by the reference's own C++ (tests/golden/fatigue_ref.npz) and against the oracle on
seeded random histories.  Bar: identical cycle counts and histogram bins; damage <= 1e-10 relative."""
import os
import numpy as np
import pytest

from fedem_solvers_b200 import fatigue, FatigueCounter

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CURVE = [15.117, 17.146, 4.0, 5.0]


def _gold():
    g = np.load(os.path.join(GOLD, "fatigue_ref.npz"))
    for i, name in enumerate(g["names"]):
        yield dict(name=str(name), gate=float(g["gates"][i]), data=g["data"][g["doff"][i]:g["doff"][i + 1]],
                   damage=float(g["damage"][i]), ncycles=int(g["ncycles"][i]), ok=bool(g["rf_ok"][i]),
                   bins=g["bins"][i], bin_size=float(g["bin_size"]))


def test_fatigue_golden_one_by_one():
    for c in _gold():
        d, n, b = fatigue(c["data"][None, :], c["gate"], CURVE, c["bin_size"], len(c["bins"]))
        assert n[0] == c["ncycles"], c["name"]
        assert abs(d[0] - c["damage"]) <= 1e-10 * max(c["damage"], 1e-300), c["name"]
        assert np.array_equal(b[0], c["bins"]), c["name"]


def test_fatigue_golden_batched_ragged():
    """all golden series in ONE launch: ragged lengths padded by repeating the last sample (a repeated
    sample never creates a turning point: delta = 0 is ignored by the PVX gate test only when the
    reference would ignore it too, so pad only series that are compared on their own length)."""
    cases = [c for c in _gold() if c["gate"] == 25.0 and len(c["data"]) == 400]
    assert len(cases) >= 4
    H = np.stack([c["data"] for c in cases])
    d, n, b = fatigue(H, 25.0, CURVE, 10.0, 64)
    for i, c in enumerate(cases):
        assert n[i] == c["ncycles"] and np.array_equal(b[i], c["bins"]), c["name"]
        assert abs(d[i] - c["damage"]) <= 1e-10 * max(c["damage"], 1e-300)


def _random_histories(ng, ns, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(ns)
    H = np.empty((ng, ns))
    for g in range(ng):
        kind = g % 5
        if kind == 0:
            H[g] = 150 + 100 * np.sin(rng.uniform(0.05, 0.6) * t + rng.uniform(0, 6)) * np.cos(0.01 * t) + rng.normal(0, 15, ns)
        elif kind == 1:
            H[g] = rng.normal(0, 1, ns).cumsum() * 20
        elif kind == 2:
            H[g] = np.round(rng.normal(0, 1, ns).cumsum() * 2) * 10     # ties / plateaus
        elif kind == 3:
            H[g] = 100 + rng.uniform(-10, 10, ns)                        # all below the gate
        else:
            H[g] = rng.uniform(0, 300, ns)
    H[:, 0] = H[:, 1]  # histories that start on a plateau: duplicated first turning point in the reference
    return H


def _oracle_all(oracle, H, gate, bin_size, nbins):
    ng = H.shape[0]
    dmg = np.zeros(ng); ncyc = np.zeros(ng, np.int32); bins = np.zeros((ng, nbins), np.int32); ok = np.ones(ng, bool)
    for g in range(ng):
        tp = oracle.pvx(H[g], gate)
        cyc = oracle.rainflow(tp, gate)
        if cyc is None:
            ok[g] = False
            continue
        ncyc[g] = len(cyc)
        dmg[g] = oracle.damage(cyc, CURVE) if len(cyc) else 0.0
        r = np.abs(cyc[:, 0] - cyc[:, 1]) if len(cyc) else np.zeros(0)
        for k in range(nbins):
            lo, hi = k * bin_size, (k + 1) * bin_size
            bins[g, k] = -1 if (len(r) == 0 or lo > r.max()) else int(((r >= lo) & (r < hi)).sum())
    return dmg, ncyc, bins, ok


@pytest.mark.parametrize("ng,ns", [(1, 1), (3, 2), (37, 33), (300, 1000), (130, 4099)])
def test_fatigue_vs_oracle(oracle, ng, ns):
    H = _random_histories(ng, max(ns, 2), 7)[:, :ns]
    d, n, b = fatigue(H, 25.0, CURVE, 10.0, 48)
    do, no, bo, ok = _oracle_all(oracle, H, 25.0, 10.0, 48)
    assert np.array_equal(n[ok], no[ok])
    assert np.array_equal(b[ok], bo[ok])
    assert np.all(np.abs(d[ok] - do[ok]) <= 1e-10 * np.maximum(do[ok], 1e-300))


def test_streaming_tiles_both_layouts(oracle):
    """tile-by-tile feeding (step-major device tiles, as the rosette kernel writes them) equals the
    one-shot result; first turning points that lie in later tiles are found by the locate pass."""
    import torch
    ng, ns, tile = 500, 3000, 256
    H = _random_histories(ng, ns, 3)
    H[5, :1500] = 100.0          # first turning point far into the history
    H[6, :] = 50.0               # never any turning point
    do, no, bo, ok = _oracle_all(oracle, H, 25.0, 10.0, 40)
    Hd = torch.from_numpy(H).cuda()
    for layout in (FatigueCounter.GAGE_MAJOR, FatigueCounter.STEP_MAJOR):
        fc = FatigueCounter(ng, 25.0, CURVE, 10.0, 40, stack_cap=ns + 8)
        s = torch.cuda.current_stream().cuda_stream
        tiles = []
        for t0 in range(0, ns, tile):
            blk = Hd[:, t0:t0 + tile]
            blk = blk.contiguous() if layout == 0 else blk.t().contiguous()
            tiles.append((t0, blk))
        pend = ng
        for t0, blk in tiles:   # locate pass stops as soon as every gage has its first turning point
            ld = blk.shape[1]
            pend = fc.locate(blk.data_ptr(), ld, layout, t0, blk.shape[1] if layout == 0 else blk.shape[0], s)
            if pend == 0:
                break
        assert pend == 101      # gage 6 and the 100 all-below-gate histories never turn
        for t0, blk in tiles:
            fc.feed(blk.data_ptr(), blk.shape[1], layout, t0, blk.shape[1] if layout == 0 else blk.shape[0], s)
        torch.cuda.synchronize()
        r = fc.finish()
        assert np.array_equal(r["ncycles"][ok], no[ok])
        assert np.array_equal(r["bins"][ok], bo[ok])
        assert np.all(np.abs(r["damage"][ok] - do[ok]) <= 1e-10 * np.maximum(do[ok], 1e-300))
        assert np.all((r["status"] == 1) == ~ok)
        assert r["ncycles"][6] == 0 and r["damage"][6] == 0.0
        fc.close()


def test_stack_overflow_is_reported():
    """a growing sawtooth closes no cycle: the residue is the whole history; with a small stack the
    gage is flagged (status 2, results -1) instead of silently miscounted"""
    x = np.array([(-1) ** i * (10 + 3 * i) for i in range(400)], float)[None, :]
    fc = FatigueCounter(1, 5.0, CURVE, stack_cap=64)
    import torch
    xd = torch.from_numpy(x).cuda()
    fc.locate(xd.data_ptr(), 400, 0, 0, 400)
    fc.feed(xd.data_ptr(), 400, 0, 0, 400)
    r = fc.finish()
    assert r["status"][0] == 2 and r["ncycles"][0] == -1 and r["damage"][0] == -1.0
    fc.close()
