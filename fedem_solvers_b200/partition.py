"""Element-block sharding of the recovery work across the GPUs of one box.

The reference has no parallelism: one serial element loop per FE part and one process per part
(src/vpmStress/stressRoutines.f90:169, stress.f90:126-128).  Elements are independent given the
nodal displacements and result points are per element, so a part is cut into contiguous element
blocks of equal cost; every block becomes a self-contained sub-part (own SamType arrays, the rows
of B and E of the nodes it touches, seam nodes duplicated) that one GPU recovers without talking
to the others.  The external DOFs (and therefore the reduced history Q) are common to all blocks.
Several parts of a mechanism (config 4) are packed onto the GPUs with longest-processing-time-first
bin packing of (part, block) work items."""
from dataclasses import dataclass
import numpy as np

from .model import SamData, ElementData, PartModel

I32 = np.int32
F64 = np.float64

# relative cost of one element.step = the K1 rows it brings (nodal DOFs x n_red, shared with the neighbours) + its K2 kernel, in
# measured picoseconds per element.step on B200 (profiles/R4_bench.json, R4_bench_configs.json, the per-piece times of
# R5_bench_c4_n8.json).  The K1 share is quoted at n_red = 98 and scales with the reduced dimension of the part: quad 26 + 24
# (flat regions on in-plane rows; 39 + 38 on six global rows), TET10 23 + 31 and HEX20 63 + 149 (step-lane kernels), triangles
# 20 + 37; linear solids 4.2 + 44.4 measured on config 4's mix at n_red = 44; the other types scaled by their DMMA counts.  With the scaling the six
# parts of config 4 (n_red 34 ... 88) come out within 5 % of their measured times (the small ones within 11 %).  The same table
# lives in csrc/sharded.cu (fsr_split_elements).
ELEMENT_K1 = {24: 26.0, 22: 26.0, 23: 20.0, 21: 20.0, 41: 23.0, 42: 70.0, 43: 63.0, 44: 14.0, 45: 8.0, 46: 12.0, 11: 1.0, 31: 56.0, 32: 78.0}
ELEMENT_K2 = {24: 24.0, 22: 24.0, 23: 37.0, 21: 37.0, 41: 31.0, 42: 170.0, 43: 149.0, 44: 67.0, 45: 36.0, 46: 57.0, 11: 2.0, 31: 160.0, 32: 230.0}
ELEMENT_COST = {t: ELEMENT_K1[t] + ELEMENT_K2[t] for t in ELEMENT_K1}   # at n_red = 98
NSTRP = {11: 0, 21: 6, 23: 6, 22: 8, 24: 8, 31: 12, 32: 16, 41: 10, 42: 15, 43: 20, 44: 8, 45: 4, 46: 6}
N_RED_REF = 98.0

QUAD_GLOBAL_ROWS_K1, QUAD_GLOBAL_ROWS_K2 = 39.0, 38.0   # a quadrilateral on six global rows per node: see element_costs
QUAD_GLOBAL_ROWS_COST = QUAD_GLOBAL_ROWS_K1 + QUAD_GLOBAL_ROWS_K2


def element_costs(melcon, mpmnpc=None, mmnpc=None, n_red=None):
    """Cost of every element by its type code; n_red (the part's reduced dimension, SAM ndim) scales the K1 share.  With the
    connectivity (SAM mpmnpc / mmnpc, 1-based) a quadrilateral that shares a node with any other element type (triangles,
    beams, solids) is charged the six-global-rows cost: the in-plane form does not pay where other elements read the global
    rows of the same nodes anyway (config 4's mixed plate, profiles/R5_bench_mixed_plate.json).  Fold lines inside an
    all-quadrilateral mesh are not looked for."""
    melcon = np.asarray(melcon)
    f = 1.0 if n_red is None else float(n_red) / N_RED_REF
    c = np.zeros(len(melcon), F64)
    for t in ELEMENT_K1:
        c[melcon == t] = ELEMENT_K2[t] + ELEMENT_K1[t] * f
    if mpmnpc is not None and mmnpc is not None:
        isq = (melcon == 24) | (melcon == 22)
        if isq.any() and not isq.all():
            mp = np.asarray(mpmnpc, np.int64) - 1
            nodes = np.asarray(mmnpc, np.int64) - 1
            owner = np.repeat(np.arange(len(melcon)), np.diff(mp))
            mixed = np.zeros(int(nodes.max()) + 1, bool)
            mixed[nodes[~isq[owner]]] = True
            touched = np.zeros(len(melcon), bool)
            np.logical_or.at(touched, owner, mixed[nodes])
            c[isq & touched] = QUAD_GLOBAL_ROWS_K2 + QUAD_GLOBAL_ROWS_K1 * f
    return c


def split_elements(part, nblocks):
    """Contiguous (SAM order) element ranges [(e0, e1), ...] of equal summed cost."""
    cost = element_costs(part.sam.melcon, part.sam.mpmnpc, part.sam.mmnpc, part.sam.ndim)
    if part.elm.elmid is not None:
        cost = np.where(part.elm.elmid < 1, 0.0, cost)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    total = cum[-1]
    cuts = [0]
    for b in range(1, nblocks):
        cuts.append(int(np.searchsorted(cum, total * b / nblocks, side="left")))
    cuts.append(part.sam.nel)
    cuts = np.maximum.accumulate(np.asarray(cuts))
    return [(int(cuts[b]), int(cuts[b + 1])) for b in range(nblocks)]


@dataclass
class SubPart:
    part: PartModel          # the self-contained block
    e0: int                  # first element (0-based, SAM order of the parent)
    e1: int                  # one past the last element
    pt0: int                 # first result point of the block in the parent's result-point order
    npts: int                # result points of the block
    nodes: np.ndarray        # parent node numbers (1-based) of the block's nodes
    rows1: np.ndarray        # rows of the parent's B / E (0-based) that the block keeps


def sub_part(part, e0, e1, with_matrices=True):
    """The element block [e0, e1) of `part` as a PartModel of its own.  Keeps: the block's nodes, all
    external nodes (so that ndof2 and Q stay those of the parent) and the master nodes of every
    constraint equation that a kept DOF depends on."""
    s, el = part.sam, part.elm
    ip0, ip1 = s.mpmnpc[e0] - 1, s.mpmnpc[e1] - 1
    keep = np.zeros(s.nnod + 1, bool)
    keep[s.mmnpc[ip0:ip1]] = True
    node_of_dof = np.repeat(np.arange(1, s.nnod + 1), np.diff(s.madof))
    keep[node_of_dof[s.msc == 2]] = True
    if s.nceq > 0:  # closure over constraint masters
        while True:
            kd = keep[node_of_dof]
            dep = np.nonzero(kd & (s.meqn < 0))[0]
            added = False
            for d in dep:
                ic = -int(s.meqn[d])
                for ip in range(s.mpmceq[ic - 1] + 1, s.mpmceq[ic]):
                    m = int(s.mmceq[ip - 1])
                    if 0 < m <= s.ndof and not keep[node_of_dof[m - 1]]:
                        keep[node_of_dof[m - 1]] = True
                        added = True
            if not added:
                break
    nodes = np.nonzero(keep)[0]                      # 1-based parent node numbers, ascending
    newnode = np.zeros(s.nnod + 1, np.int64)
    newnode[nodes] = np.arange(1, len(nodes) + 1)
    ndpn = np.diff(s.madof)[nodes - 1]
    madof = np.concatenate([[1], 1 + np.cumsum(ndpn)]).astype(I32)
    kd = keep[node_of_dof]                           # kept DOFs of the parent
    old_dofs = np.nonzero(kd)[0]
    newdof = np.zeros(s.ndof + 1, np.int64)          # 1-based old -> 1-based new
    newdof[old_dofs + 1] = np.arange(1, len(old_dofs) + 1)
    msc = s.msc[old_dofs].astype(I32)
    meqn_old = s.meqn[old_dofs]
    eq_kept = np.unique(meqn_old[meqn_old > 0])
    neweq = np.zeros(s.neq + 1, np.int64)
    neweq[eq_kept] = np.arange(1, len(eq_kept) + 1)
    meqn = np.zeros(len(old_dofs), I32)
    pos = meqn_old > 0
    meqn[pos] = neweq[meqn_old[pos]]
    # constraint equations
    mpmceq, mmceq, ttcc = [1], [], []
    ceq_kept = np.unique(-meqn_old[meqn_old < 0])
    newceq = {int(c): i + 1 for i, c in enumerate(ceq_kept)}
    for c in ceq_kept:
        for ip in range(s.mpmceq[c - 1], s.mpmceq[c]):        # incl. the leading (dependent, c0) entry
            m = int(s.mmceq[ip - 1])
            mmceq.append(int(newdof[m]) if 0 < m <= s.ndof else 0)
            ttcc.append(float(s.ttcc[ip - 1]))
        mpmceq.append(len(mmceq) + 1)
    neg = meqn_old < 0
    meqn[neg] = [-newceq[int(-v)] for v in meqn_old[neg]]
    k1 = neweq[s.meqn1] > 0
    rows1 = np.nonzero(k1)[0]
    meqn1 = neweq[s.meqn1[rows1]].astype(I32)
    meqn2 = neweq[s.meqn2].astype(I32)
    assert np.all(meqn2 > 0)
    # elements
    nel = e1 - e0
    mpmnpc = (s.mpmnpc[e0:e1 + 1] - s.mpmnpc[e0] + 1).astype(I32)
    mmnpc = newnode[s.mmnpc[ip0:ip1]].astype(I32)
    sam = SamData(nnod=len(nodes), nel=nel, ndof=len(old_dofs), ndof1=len(rows1), ndof2=s.ndof2, ngen=s.ngen,
                  neq=len(eq_kept), nceq=len(ceq_kept), madof=madof, msc=msc, mpmnpc=mpmnpc, mmnpc=mmnpc,
                  melcon=s.melcon[e0:e1].astype(I32), meqn=meqn, meqn1=meqn1, meqn2=meqn2,
                  mpmceq=np.asarray(mpmceq, I32), mmceq=np.asarray(mmceq, I32), ttcc=np.asarray(ttcc, F64),
                  minex=(s.minex[nodes - 1] if s.minex is not None else None))
    elm = ElementData(xyz=np.ascontiguousarray(el.xyz[nodes - 1]), emod=el.emod[e0:e1].copy(), rny=el.rny[e0:e1].copy(),
                      thk=el.thk[e0:e1].copy(), elmid=(el.elmid[e0:e1].copy() if el.elmid is not None else None),
                      beam=(el.beam[e0:e1].copy() if el.beam is not None else None))
    sp = PartModel(sam=sam, elm=elm, name=f"{part.name}[{e0}:{e1}]", recovery_seed=getattr(part, "recovery_seed", 0))
    if with_matrices and part.B is not None:
        sp.B = np.asfortranarray(part.B[rows1, :])
    if with_matrices and part.E is not None:
        sp.E = np.asfortranarray(part.E[rows1, :])
    nstrp = part.nstrp()
    off = np.concatenate([[0], np.cumsum(nstrp)])
    return SubPart(part=sp, e0=e0, e1=e1, pt0=int(off[e0]), npts=int(off[e1] - off[e0]), nodes=nodes, rows1=rows1)


def plan_work(part_costs, nranks):
    """Packing of several parts of a mechanism onto `nranks` GPUs (config 4).  Element blocks can be cut
    anywhere, so the work is divisible: the parts are laid end to end (largest first) and the line is
    cut into `nranks` equal shares.  Every rank gets a contiguous span = pieces of one or more parts;
    at most nparts + nranks - 1 pieces exist in total and the load is balanced to one element.
    Returns (items, loads): items[r] = list of (part index, f0, f1) with [f0, f1) the fraction of that
    part's summed element cost the rank takes (turn into elements with cost_fraction_to_elements)."""
    part_costs = np.asarray(part_costs, F64)
    order = np.argsort(-part_costs, kind="stable")
    total = part_costs.sum()
    share = total / nranks if nranks > 0 else 0.0
    items = [[] for _ in range(nranks)]
    loads = np.zeros(nranks)
    r, room = 0, share
    for ip in order:
        c, done = part_costs[ip], 0.0
        while c - done > 1e-12 * max(total, 1.0):
            take = min(c - done, room) if r < nranks - 1 else c - done
            items[r].append((int(ip), done / c, (done + take) / c))
            loads[r] += take
            done += take
            room -= take
            if room <= 1e-12 * max(total, 1.0) and r < nranks - 1:
                r, room = r + 1, share
    return items, loads


def cost_fraction_to_elements(part, f0, f1):
    """Element range [e0, e1) of `part` covering the cost fractions [f0, f1) (see plan_work)."""
    cost = element_costs(part.sam.melcon, part.sam.mpmnpc, part.sam.mmnpc, part.sam.ndim)
    if part.elm.elmid is not None:
        cost = np.where(part.elm.elmid < 1, 0.0, cost)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    e0 = int(np.searchsorted(cum, f0 * cum[-1], side="left")) if f0 > 0 else 0
    e1 = int(np.searchsorted(cum, f1 * cum[-1], side="left")) if f1 < 1 else part.sam.nel
    return min(e0, part.sam.nel), min(max(e1, e0), part.sam.nel)


# ---- one process per GPU (torch.distributed: NCCL over NVLink / NVSwitch on the box, gloo on CPU for the tests) -------------
# The only data that crosses GPUs: the small reduced history Q (n_red x steps, broadcast from rank 0 per tile of steps -- 392 KB
# for config 2) and, once at the end, the per-block von Mises envelopes (gathered to rank 0 into the parent part's result-point
# order).  No collective sits between K1 and K2: every block holds the B / E rows of all nodes its elements touch.  bench.py and
# tools/bench_c4.py do the same through the library's own communicator (fsr_comm_*); this class is the host-side logic the
# world-2 gloo test runs with the oracle standing in for the device.
class ShardedRecovery:
    """recover_cls(part) must offer recover_dev / recover / envelope like StressRecovery; tests pass a
    CPU stand-in built on the oracle to check the sharding logic under gloo."""

    def __init__(self, part, rank, world, recover_factory, with_matrices=True):
        self.rank, self.world = rank, world
        self.ranges = split_elements(part, world)
        e0, e1 = self.ranges[rank]
        self.block = sub_part(part, e0, e1, with_matrices=with_matrices)
        nstrp = part.nstrp()
        off = np.concatenate([[0], np.cumsum(nstrp)])
        self.pt_ranges = [(int(off[a]), int(off[b])) for a, b in self.ranges]
        self.npts_total = int(off[-1])
        self.rec = recover_factory(self.block.part)

    def gather_envelope(self, dist, local_max, local_min, device=None):
        """Per-block envelopes -> rank 0, concatenated in the parent's result-point order (blocks are
        contiguous element ranges, so concatenation IS the parent order).  Returns (max, min) on rank 0."""
        import torch
        counts = [b - a for a, b in self.pt_ranges]
        mine = torch.stack([torch.as_tensor(local_max, dtype=torch.float64), torch.as_tensor(local_min, dtype=torch.float64)])
        if device is not None:
            mine = mine.to(device)
        if self.world == 1:
            return mine[0].cpu().numpy(), mine[1].cpu().numpy()
        # gather wants equal shapes on every rank: pad the blocks to the largest one
        width = max(counts)
        padded = torch.zeros((2, width), dtype=torch.float64, device=mine.device)
        padded[:, :mine.shape[1]] = mine
        bufs = [torch.empty((2, width), dtype=torch.float64, device=mine.device) for _ in counts] if self.rank == 0 else None
        dist.gather(padded, bufs, dst=0)
        if self.rank != 0:
            return None, None
        full = torch.cat([b[:, :c] for b, c in zip(bufs, counts)], 1).cpu().numpy()
        return full[0], full[1]
