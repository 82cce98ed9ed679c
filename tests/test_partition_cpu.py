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


def test_quads_next_to_other_element_types_cost_the_global_row_path():
    """a quadrilateral that shares a node with a triangle is charged the six-global-rows cost (it leaves the in-plane path);
    an all-quadrilateral plate and a solid part are charged by type alone"""
    from fedem_solvers_b200.partition import ELEMENT_COST, QUAD_GLOBAL_ROWS_COST
    part = plate_part(12, 10, ngen=2, seed=2, tri_fraction=0.3, with_recovery=False)
    s = part.sam
    c = element_costs(s.melcon, s.mpmnpc, s.mmnpc)
    tri_nodes = set()
    for e in np.nonzero(s.melcon == 23)[0]:
        tri_nodes.update(s.mmnpc[s.mpmnpc[e] - 1:s.mpmnpc[e + 1] - 1].tolist())
    nq_global = 0
    for e in np.nonzero(s.melcon == 24)[0]:
        touches = bool(tri_nodes.intersection(s.mmnpc[s.mpmnpc[e] - 1:s.mpmnpc[e + 1] - 1].tolist()))
        assert c[e] == (QUAD_GLOBAL_ROWS_COST if touches else ELEMENT_COST[24])
        nq_global += touches
    assert 0 < nq_global and np.all(c[s.melcon == 23] == ELEMENT_COST[23])
    plain = plate_part(6, 5, ngen=2, seed=2, with_recovery=False).sam
    assert np.all(element_costs(plain.melcon, plain.mpmnpc, plain.mmnpc) == ELEMENT_COST[24])
    assert np.array_equal(element_costs(s.melcon), np.where(s.melcon == 24, ELEMENT_COST[24], ELEMENT_COST[23]))
    # the K1 share scales with the reduced dimension of the part (quoted at n_red = 98)
    from fedem_solvers_b200.partition import ELEMENT_K1, ELEMENT_K2
    half = element_costs(s.melcon, n_red=49)
    assert np.allclose(half, np.where(s.melcon == 24, ELEMENT_K2[24] + 0.5 * ELEMENT_K1[24], ELEMENT_K2[23] + 0.5 * ELEMENT_K1[23]), rtol=1e-15)
    assert all(abs(ELEMENT_COST[t] - (ELEMENT_K1[t] + ELEMENT_K2[t])) == 0 for t in ELEMENT_COST)


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
    from fedem_solvers_b200.partition import ShardedRecovery
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


# ---- the native cut (libfedem_b200.so, sharded.cu) against the Python one above ---------------------------------
def _native_blockdef(part, e0, e1):
    import ctypes as C
    from fedem_solvers_b200 import _lib
    from fedem_solvers_b200.recovery import c_part_structs, c_options
    lib = _lib.load_library()
    keep = []
    sam, elm = c_part_structs(part, keep)
    opt = c_options()
    h = C.c_void_p()
    _lib.check(lib.fsr_blockdef_create(C.byref(h), C.byref(sam), C.byref(elm), C.byref(opt), e0, e1), "fsr_blockdef_create")
    s, e = lib.fsr_blockdef_sam(h).contents, lib.fsr_blockdef_elm(h).contents
    info = np.zeros(10, np.int32)
    _lib.check(lib.fsr_blockdef_info(h, info.ctypes.data_as(_lib._I), None, None), "fsr_blockdef_info")
    rows, nodes = np.zeros(max(info[5], 1), np.int32), np.zeros(max(info[4], 1), np.int32)
    lib.fsr_blockdef_info(h, None, rows.ctypes.data_as(_lib._I), nodes.ctypes.data_as(_lib._I))

    def arr(p, n):
        return np.array(p[:n]) if n > 0 else np.zeros(0)
    out = dict(nnod=s.nnod, nel=s.nel, ndof=s.ndof, ndof1=s.ndof1, ndof2=s.ndof2, ngen=s.ngen, neq=s.neq, nceq=s.nceq,
               madof=arr(s.madof, s.nnod + 1), msc=arr(s.msc, s.ndof), mpmnpc=arr(s.mpmnpc, s.nel + 1), mmnpc=arr(s.mmnpc, s.nmmnpc),
               melcon=arr(s.melcon, s.nel), mpmceq=arr(s.mpmceq, s.nceq + 1), mmceq=arr(s.mmceq, s.nmmceq), ttcc=arr(s.ttcc, s.nmmceq),
               meqn=arr(s.meqn, s.ndof), meqn1=arr(s.meqn1, s.ndof1), meqn2=arr(s.meqn2, s.ndof2),
               xyz=arr(e.xyz, 3 * s.nnod).reshape(-1, 3), emod=arr(e.emod, s.nel), rny=arr(e.rny, s.nel), thk=arr(e.thk, s.nel),
               elmid=arr(e.elmid, s.nel) if e.elmid else None,
               beam=arr(e.beam, 32 * s.nel).reshape(-1, 32) if e.beam else None,
               info=info.copy(), rows1=rows[:info[5]].copy(), nodes=nodes[:info[4]].copy())
    lib.fsr_blockdef_destroy(h)
    return out


@pytest.mark.parametrize("maker", ["plate", "tets"])
def test_native_blocks_equal_python_sub_parts(maker):
    """fsr_split_elements / fsr_blockdef_create (what fsr_part_create_block and fsr_group_create cut with) give exactly the
    arrays of partition.split_elements / sub_part, which test_sub_parts_reproduce_parent checks against the oracle"""
    from fedem_solvers_b200 import split_elements as native_split
    if maker == "plate":
        part = plate_part(9, 8, ngen=4, seed=3, tri_fraction=0.3, shuffle_eq=True, n_fixed=3, n_constraints=4, warp=0.02)
    else:
        part = tet10_block(3, 2, 2, ngen=4, seed=4, shuffle_eq=True, n_beams=5)
    for nb in (1, 2, 3, 5):
        assert native_split(part, nb) == split_elements(part, nb)
    for e0, e1 in split_elements(part, 3) + [(0, 0), (part.sam.nel, part.sam.nel), (2, 3)]:
        sp = sub_part(part, e0, e1)
        nb = _native_blockdef(part, e0, e1)
        s = sp.part.sam
        for k in ("nnod", "nel", "ndof", "ndof1", "ndof2", "ngen", "neq", "nceq"):
            assert nb[k] == getattr(s, k), (k, nb[k], getattr(s, k))
        for k in ("madof", "msc", "mpmnpc", "mmnpc", "melcon", "mpmceq", "mmceq", "ttcc", "meqn", "meqn1", "meqn2"):
            assert np.array_equal(nb[k], np.asarray(getattr(s, k))[:len(nb[k])]) and len(nb[k]) == len(getattr(s, k)), k
        assert np.array_equal(nb["xyz"], sp.part.elm.xyz) and np.array_equal(nb["emod"], sp.part.elm.emod)
        assert np.array_equal(nb["thk"], sp.part.elm.thk) and np.array_equal(nb["rny"], sp.part.elm.rny)
        if sp.part.elm.elmid is not None and e1 > e0:
            assert np.array_equal(nb["elmid"], sp.part.elm.elmid)
        if sp.part.elm.beam is not None and e1 > e0:
            assert np.array_equal(nb["beam"], sp.part.elm.beam)
        assert np.array_equal(nb["rows1"], sp.rows1) and np.array_equal(nb["nodes"], sp.nodes)
        assert tuple(nb["info"][:4]) == (e0, e1, sp.pt0, sp.npts) and nb["info"][6] == part.sam.ndof1
        assert nb["info"][7] == int(part.nstrp().sum()) and nb["info"][8] == s.ndof and nb["info"][9] == part.sam.nel


def test_native_split_skips_inactive_elements():
    from fedem_solvers_b200 import split_elements as native_split
    part = plate_part(12, 10, ngen=2, seed=5, tri_fraction=0.4, with_recovery=False)
    part.elm.elmid = np.arange(1, part.sam.nel + 1, dtype=np.int32)
    part.elm.elmid[:40] *= -1
    for nb in (2, 4, 7):
        assert native_split(part, nb) == split_elements(part, nb)
