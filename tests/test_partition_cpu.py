"""CPU tests of the multi-GPU host logic: element-block sub-parts reproduce the parent part's
recovery exactly (checked with the oracle), work planning balances load, and the world_size-2
gloo path (broadcast of Q, gather of envelopes) returns the parent's result-point order."""
import os
import sys
import numpy as np
import pytest

from fedem_solvers_b200.model import plate_part, tet10_block, reduced_history, synthetic_recovery
from fedem_solvers_b200.partition import split_elements, sub_part, plan_work, element_costs


@pytest.mark.parametrize("maker", ["plate", "tets"])
def test_sub_parts_reproduce_parent(oracle, maker):
    if maker == "plate":
        part = plate_part(9, 8, ngen=4, seed=3, tri_fraction=0.3, shuffle_eq=True, n_fixed=3, n_constraints=4, warp=0.02)
    else:
        part = tet10_block(3, 2, 2, ngen=4, seed=4, shuffle_eq=True, n_beams=5)
    b = oracle.bind_part(part)
    Q = reduced_history(part.sam.ndim, 5, seed=1)
    vm, mx, mn = oracle.recover_history(b, Q)
    ranges = split_elements(part, 3)
    assert ranges[0][0] == 0 and ranges[-1][1] == part.sam.nel
    assert all(a[1] == c[0] for a, c in zip(ranges, ranges[1:]))
    got = []
    for e0, e1 in ranges:
        sp = sub_part(part, e0, e1)
        assert sp.part.sam.ndof2 == part.sam.ndof2 and sp.part.sam.nnod <= part.sam.nnod
        bs = oracle.bind_part(sp.part)
        vms, _, _ = oracle.recover_history(bs, Q)
        assert vms.shape[1] == sp.npts
        got.append((sp.pt0, vms))
        # bit-identical: same rows of B/E, same arithmetic order per DOF
        assert np.array_equal(vms, vm[:, sp.pt0:sp.pt0 + sp.npts])
    assert sum(g[1].shape[1] for g in got) == vm.shape[1]


def test_block_matrices_can_be_generated_locally():
    """a rank generates only ITS rows of the synthetic [B|E] (bench at 1M elements x 8 ranks)"""
    part = plate_part(12, 10, ngen=5, seed=2, n_ext=8)
    lo, hi = part.elm.xyz.min(0), part.elm.xyz.max(0)
    e0, e1 = split_elements(part, 4)[2]
    sp = sub_part(part, e0, e1)
    light = plate_part(12, 10, ngen=5, seed=2, n_ext=8, with_recovery=False)
    sp2 = sub_part(light, e0, e1, with_matrices=False)
    B, E = synthetic_recovery(sp2.part, bbox=(lo, hi))
    assert np.array_equal(B, sp.part.B) and np.array_equal(E, sp.part.E)


def test_split_balances_cost():
    part = plate_part(40, 30, ngen=2, seed=1, tri_fraction=0.5, with_recovery=False)
    cost = element_costs(part.sam.melcon)
    for n in (2, 4, 8):
        loads = [cost[a:b].sum() for a, b in split_elements(part, n)]
        assert max(loads) <= 1.02 * (cost.sum() / n) + cost.max()


def test_plan_work_config4():
    """six parts of {2M, 1M, 500k, 250k, 100k, 50k} elements on 8 GPUs: divisible-load packing"""
    from fedem_solvers_b200.partition import cost_fraction_to_elements
    costs = np.array([2e6, 1e6, 5e5, 2.5e5, 1e5, 5e4]) * 1164.0
    items, loads = plan_work(costs, 8)
    assert np.allclose(loads, costs.sum() / 8, rtol=1e-9)
    assert sum(len(i) for i in items) <= 6 + 8 - 1
    cover = {ip: [] for ip in range(6)}
    for its in items:
        for ip, f0, f1 in its:
            cover[ip].append((f0, f1))
    for ip, iv in cover.items():       # every part covered exactly once
        iv.sort()
        assert abs(iv[0][0]) < 1e-12 and abs(iv[-1][1] - 1) < 1e-12
        assert all(abs(a[1] - b[0]) < 1e-12 for a, b in zip(iv, iv[1:]))
    # fractions -> element ranges tile the part
    part = plate_part(20, 10, ngen=1, seed=1, tri_fraction=0.3, with_recovery=False)
    cuts = [cost_fraction_to_elements(part, a, b) for a, b in ((0, 0.3), (0.3, 0.85), (0.85, 1))]
    assert cuts[0][0] == 0 and cuts[-1][1] == part.sam.nel and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    import oracle_bind
    from fedem_solvers_b200.distributed import ShardedRecovery
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = oracle_bind.Oracle()
        part = plate_part(8, 7, ngen=3, seed=5, tri_fraction=0.4, n_constraints=2)

        class CpuStandIn:   # the oracle plays the device (TEST INFRASTRUCTURE; the product uses StressRecovery)
            def __init__(self, p):
                self.b = o.bind_part(p)

            def run(self, Q):
                return o.recover_history(self.b, Q)

        sh = ShardedRecovery(part, rank, world, CpuStandIn)
        # rank 0 owns the history and broadcasts it tile by tile
        ndim, ns, tile = part.sam.ndim, 12, 5
        Q = torch.zeros((ns, ndim), dtype=torch.float64)
        if rank == 0:
            Q[:] = torch.from_numpy(reduced_history(ndim, ns, seed=2).T.copy())
        mx = np.zeros(sh.block.npts); mn = np.full(sh.block.npts, np.finfo(float).max)
        for t0 in range(0, ns, tile):
            qt = Q[t0:t0 + tile].contiguous()
            dist.broadcast(qt, src=0)
            _, a, b_ = sh.rec.run(qt.numpy().T)
            mx = np.maximum(mx, a); mn = np.minimum(mn, b_)
        gmx, gmn = sh.gather_envelope(dist, mx, mn)
        if rank == 0:
            _, rmx, rmn = o.recover_history(o.bind_part(part), Q.numpy().T)
            q.put((bool(np.array_equal(gmx, rmx)), bool(np.array_equal(gmn, rmn)), len(gmx)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_broadcast_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] and res[1] and res[2] > 0
