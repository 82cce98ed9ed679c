"""Host-side model containers (the data the reference keeps in SamType and in the FE-model
singleton) and seeded synthetic part generators for the BASELINE.json configurations.

Index arrays are 1-based exactly as the reference stores them in the .fsm file
(src/vpmStress/samStressModule.f90:273-316), so the same arrays can be handed to the C ABI, to
the Fortran driver and to the CPU oracle unchanged."""
from dataclasses import dataclass, field
import numpy as np

I32 = np.int32
F64 = np.float64


@dataclass
class SamData:
    """SamType subset (src/vpmCommon/samModule.f90:27-66) as read by initiateSAM."""
    nnod: int
    nel: int
    ndof: int
    ndof1: int
    ndof2: int
    ngen: int
    neq: int
    nceq: int
    madof: np.ndarray
    msc: np.ndarray
    mpmnpc: np.ndarray
    mmnpc: np.ndarray
    melcon: np.ndarray
    meqn: np.ndarray
    meqn1: np.ndarray
    meqn2: np.ndarray
    mpmceq: np.ndarray = field(default_factory=lambda: np.ones(1, I32))
    mmceq: np.ndarray = field(default_factory=lambda: np.zeros(0, I32))
    ttcc: np.ndarray = field(default_factory=lambda: np.zeros(0, F64))
    minex: np.ndarray = None

    @property
    def ndim(self):
        return self.ndof2 + self.ngen

    def dof_pos_in2(self):
        """invertEqPartition, src/vpmCommon/samModule.f90:960-970."""
        pos = {int(eq): j + 1 for j, eq in enumerate(self.meqn2)}
        return np.array([pos[int(self.meqn[d])] for d in np.nonzero(self.msc == 2)[0]], I32)


@dataclass
class ElementData:
    """What ffl_getcoor/getmat/getthick/getbeamsection/getpinflags/getelmid deliver
    (fedem-foundation/src/FFlLib/FFlLinkHandler_F.C:699-1193)."""
    xyz: np.ndarray          # [nnod, 3]
    emod: np.ndarray         # [nel]
    rny: np.ndarray          # [nel]
    thk: np.ndarray          # [nel]
    elmid: np.ndarray = None # [nel] external ids (<1: skipped)
    beam: np.ndarray = None  # [nel, 32] beam data (type 11)


@dataclass
class PartModel:
    sam: SamData
    elm: ElementData
    B: np.ndarray = None     # [ndof1, ndof2] column-major (Fortran order)
    E: np.ndarray = None     # [ndof1, ngen]  column-major
    name: str = ""
    recovery_seed: int = 0   # seed of the synthetic [B|E] fields (synthetic_recovery)

    def nstrp(self):
        tab = {11: 0, 21: 6, 23: 6, 22: 8, 24: 8, 31: 12, 32: 16, 41: 10, 42: 15, 43: 20, 44: 8, 45: 4, 46: 6}
        n = np.array([tab.get(int(t), 0) for t in self.sam.melcon], I32)
        if self.elm.elmid is not None:
            n[self.elm.elmid < 1] = 0
        return n


# ------------------------------------------------------------------------------------------
# SAM bookkeeping for a synthetic part
# ------------------------------------------------------------------------------------------
def _build_sam(nnod, ndof_per_node, conn_list, types, ext_nodes, fixed_dofs=(), constraints=(),
               rng=None, shuffle_eq=False):
    """conn_list: list of 1-based node arrays per element; ext_nodes: 1-based external nodes;
    fixed_dofs: 0-based DOF indices that are suppressed; constraints: list of
    (dependent_dof0, [(master_dof0, coeff), ...])."""
    ndpn = np.full(nnod, ndof_per_node, I32) if np.isscalar(ndof_per_node) else np.asarray(ndof_per_node, I32)
    madof = np.concatenate([[1], 1 + np.cumsum(ndpn)]).astype(I32)
    ndof = int(madof[-1] - 1)
    msc = np.ones(ndof, I32)
    for n in ext_nodes:
        msc[madof[n - 1] - 1: madof[n] - 1] = 2
    dep = {d: terms for d, terms in constraints}
    for d in fixed_dofs:
        msc[d] = 0
    for d in dep:
        msc[d] = 0  # dependent DOFs carry status 0 in msc; meqn holds -iceq
    free = np.nonzero(msc > 0)[0]
    neq = len(free)
    eqno = np.arange(1, neq + 1)
    if shuffle_eq and rng is not None:
        eqno = rng.permutation(eqno)
    meqn = np.zeros(ndof, I32)
    meqn[free] = eqno
    mpmceq = [1]
    mmceq, ttcc = [], []
    for ic, (d, terms) in enumerate(dep.items(), start=1):
        meqn[d] = -ic
        mmceq.append(d + 1)       # first entry: the dependent DOF itself with c0
        ttcc.append(0.0)
        for m, c in terms:
            mmceq.append(m + 1)
            ttcc.append(c)
        mpmceq.append(len(mmceq) + 1)
    int_dofs = np.nonzero(msc == 1)[0]
    ext_dofs = np.nonzero(msc == 2)[0]
    meqn1 = meqn[int_dofs].copy()
    meqn2 = meqn[ext_dofs].copy()
    if shuffle_eq and rng is not None:
        meqn1 = rng.permutation(meqn1)
        meqn2 = rng.permutation(meqn2)
    if isinstance(conn_list, np.ndarray) and conn_list.ndim == 2:  # uniform connectivity, fast path
        nper = conn_list.shape[1]
        mpmnpc = (1 + nper * np.arange(conn_list.shape[0] + 1, dtype=np.int64)).astype(I32)
        mmnpc = np.ascontiguousarray(conn_list, I32).ravel()
    else:
        mpmnpc = np.concatenate([[1], 1 + np.cumsum([len(c) for c in conn_list])]).astype(I32)
        mmnpc = np.concatenate(conn_list).astype(I32) if len(conn_list) else np.zeros(0, I32)
    return SamData(nnod=nnod, nel=len(conn_list), ndof=ndof, ndof1=len(int_dofs), ndof2=len(ext_dofs),
                   ngen=0, neq=neq, nceq=len(dep), madof=madof, msc=msc, mpmnpc=mpmnpc, mmnpc=mmnpc,
                   melcon=np.asarray(types, I32), meqn=meqn, meqn1=meqn1.astype(I32),
                   meqn2=meqn2.astype(I32), mpmceq=np.asarray(mpmceq, I32),
                   mmceq=np.asarray(mmceq, I32), ttcc=np.asarray(ttcc, F64),
                   minex=np.arange(1, nnod + 1, dtype=I32))


def _smooth_recovery_matrices(sam, xyz, ngen, rng, amp=1.0, bbox=None, rows=None):
    """Synthetic [B|E]: smooth low-order cosine fields over the part, one per reduced DOF, with
    a random per-component scale -- cheap to generate at 6M rows, and every internal DOF row is a
    distinct, well-scaled linear combination of the reduced DOFs.  A row depends
    only on its node's coordinates, its DOF component and the per-column parameters drawn from `rng`,
    so an element block of a part (partition.sub_part) can generate exactly its own rows when given
    the whole part's bounding box; `rows` (0-based) restricts the result to those rows of the part's
    matrices (what fsr_block_rows reports for a native element block)."""
    int_dofs = np.nonzero(sam.msc == 1)[0]
    # internal row k corresponds to equation meqn1[k]; map equation -> dof
    eq2dof = np.zeros(sam.neq + 1, np.int64)
    nz = np.nonzero(sam.meqn > 0)[0]
    eq2dof[sam.meqn[nz]] = nz
    rows_dof = eq2dof[sam.meqn1 if rows is None else sam.meqn1[np.asarray(rows, np.int64)]]   # dof of each B row
    nrow = len(rows_dof)
    node_of_dof = np.searchsorted(sam.madof, rows_dof + 1, side="right") - 1
    comp = rows_dof - (sam.madof[node_of_dof] - 1)
    lo, hi = (xyz.min(0), xyz.max(0)) if bbox is None else (np.asarray(bbox[0], F64), np.asarray(bbox[1], F64))
    span = np.where(hi - lo > 0, hi - lo, 1.0)
    u = (xyz[node_of_dof] - lo) / span                 # normalised coordinates in [0,1]
    ncol = sam.ndof2 + ngen
    M = np.empty((nrow, ncol), F64, order="F")
    cscale = np.array([1.0, 1.0, 1.0, 0.5, 0.5, 0.5])
    # cos(f*u + ph) = cos(f*u) cos(ph) - sin(f*u) sin(ph): tabulate the four harmonics per axis once
    ctab = [[np.cos(m * np.pi * u[:, ax]) for m in range(4)] for ax in range(3)]
    stab = [[np.sin(m * np.pi * u[:, ax]) for m in range(4)] for ax in range(3)]
    fs = rng.integers(0, 4, (ncol, 3))
    phs = rng.uniform(0, 2 * np.pi, (ncol, 3))
    amps = rng.normal(0.0, 1.0, (ncol, 6)) * cscale

    def fill(j):
        col = M[:, j]
        np.take(amp * amps[j], comp, out=col)
        tmp = np.empty(nrow, F64)
        for ax in range(3):
            np.multiply(ctab[ax][fs[j, ax]], np.cos(phs[j, ax]), out=tmp)
            tmp -= stab[ax][fs[j, ax]] * np.sin(phs[j, ax])
            col *= tmp

    if nrow * ncol > 5_000_000:
        from concurrent.futures import ThreadPoolExecutor
        import os
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            list(ex.map(fill, range(ncol)))
    else:
        for j in range(ncol):
            fill(j)
    B = np.asfortranarray(M[:, :sam.ndof2])
    E = np.asfortranarray(M[:, sam.ndof2:])
    return B, E


def reduced_history(ndim, nsteps, seed, amp=1.0e-3, dt=1.0e-3):
    """Band-limited synthetic reduced history Q [ndim x nsteps] (Fortran order): sum of 8
    sinusoids per reduced DOF, amplitude ~amp."""
    rng = np.random.default_rng(seed)
    t = np.arange(nsteps) * dt
    Q = np.zeros((ndim, nsteps), F64, order="F")
    for _ in range(8):
        w = rng.uniform(2.0, 60.0, ndim) * 2 * np.pi
        ph = rng.uniform(0, 2 * np.pi, ndim)
        a = rng.normal(0.0, amp / 3.0, ndim)
        Q += a[:, None] * np.sin(w[:, None] * t[None, :] + ph[:, None])
    return Q


# ------------------------------------------------------------------------------------------
# part generators
# ------------------------------------------------------------------------------------------
def plate_part(nx, ny, ngen=10, seed=1, tri_fraction=0.0, n_ext=4, jitter=0.02, lx=1.0, ly=1.0,
               thickness=0.01, emod=2.1e11, rny=0.3, warp=0.0, shuffle_eq=False, n_fixed=0,
               n_constraints=0, with_recovery=True):
    """Flat (or slightly warped) plate of nx x ny ANDES quads (type 24), optionally a fraction of
    the cells split into two ANDES triangles (type 23).  n_ext corner/edge nodes are external
    (6 DOFs each).  Config C1: plate_part(70, 70); config C2: plate_part(1000, 1000, ngen=50, n_ext=8)."""
    rng = np.random.default_rng(seed)
    nnx, nny = nx + 1, ny + 1
    nnod = nnx * nny
    gx, gy = np.meshgrid(np.linspace(0, lx, nnx), np.linspace(0, ly, nny), indexing="xy")
    xyz = np.zeros((nnod, 3), F64)
    xyz[:, 0] = gx.ravel()
    xyz[:, 1] = gy.ravel()
    hx, hy = lx / nx, ly / ny
    interior = np.ones((nny, nnx), bool)
    interior[0, :] = interior[-1, :] = interior[:, 0] = interior[:, -1] = False
    m = interior.ravel()
    xyz[m, 0] += rng.uniform(-jitter, jitter, m.sum()) * hx
    xyz[m, 1] += rng.uniform(-jitter, jitter, m.sum()) * hy
    if warp:
        xyz[:, 2] = warp * np.sin(np.pi * xyz[:, 0] / lx) * np.sin(np.pi * xyz[:, 1] / ly)
    # cells
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    n1 = (iy * nnx + ix + 1).ravel()
    quads = np.stack([n1, n1 + 1, n1 + 1 + nnx, n1 + nnx], 1).astype(I32)
    ncell = nx * ny
    split = np.zeros(ncell, bool)
    if tri_fraction > 0:
        split = rng.random(ncell) < tri_fraction
    conn, types = [], []
    if not split.any():
        conn = quads
        types = np.full(ncell, 24, I32)
    else:
        types = []
        for c in range(ncell):
            q = quads[c]
            if split[c]:
                conn.append(q[[0, 1, 2]]); types.append(23)
                conn.append(q[[0, 2, 3]]); types.append(23)
            else:
                conn.append(q); types.append(24)
        types = np.asarray(types, I32)
    corners = [1, nnx, nnod, nnod - nnx + 1]
    extra = [nnx // 2 + 1, nnod - nnx // 2, (nny // 2) * nnx + 1, (nny // 2 + 1) * nnx]
    ext_nodes = (corners + extra)[:n_ext]
    ext_set = set(ext_nodes)
    fixed, cons = [], []
    if n_fixed or n_constraints:
        cand = [n for n in rng.permutation(np.arange(1, nnod + 1)) if n not in ext_set]
        for n in cand[:n_fixed]:
            fixed.append(6 * (n - 1) + int(rng.integers(0, 6)))
        used = set(fixed)
        for n in cand[n_fixed:n_fixed + n_constraints]:
            d = 6 * (n - 1) + int(rng.integers(0, 6))
            masters = []
            for mnode in rng.choice(cand[n_fixed + n_constraints:], 3, replace=False):
                md = 6 * (int(mnode) - 1) + int(rng.integers(0, 6))
                if md not in used:
                    masters.append((md, float(rng.normal())))
            cons.append((d, masters))
            used.add(d)
    sam = _build_sam(nnod, 6, conn, types, ext_nodes, fixed, cons, rng, shuffle_eq)
    sam.ngen = ngen
    nel = sam.nel
    elm = ElementData(xyz=xyz, emod=np.full(nel, emod, F64), rny=np.full(nel, rny, F64),
                      thk=np.full(nel, thickness, F64), elmid=np.arange(1, nel + 1, dtype=I32))
    part = PartModel(sam=sam, elm=elm, name=f"plate{nx}x{ny}")
    part.recovery_seed = seed
    if with_recovery:
        part.B, part.E = synthetic_recovery(part)
    return part


_TET_SPLIT = [(0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (1, 2, 3, 7), (1, 6, 7, 5), (1, 2, 7, 6)]


def beam_record(x1, x2, zdir, emod=2.1e11, gmod=8.0e10, area=1.0e-4, iy=2.0e-9, iz=1.0e-9, it=2.5e-9,
                kxy=0.85, kxz=0.8, sy=0.0, sz=0.0, efflen=0.0, phi=0.0, ecc1=(0, 0, 0), ecc2=(0, 0, 0),
                rho=7850.0, pin_a=0, pin_b=0):
    """The 32 doubles a type-11 element carries over the C ABI (fsr_elmdata.beam): X(1:5), Y(1:5), Z(1:5)
    as ffl_getcoor delivers them for BEAM2 (FFlLinkHandler_F.C:766-812: ends incl. eccentricity, point on
    the local Z axis, the two nodes), BSEC(1:14) of ffl_getbeamsection (:1131-1188; the shear factors
    stored inverted), pin flags."""
    x1, x2, zdir = (np.asarray(v, F64) for v in (x1, x2, zdir))
    e1, e2 = np.asarray(ecc1, F64), np.asarray(ecc2, F64)
    P = np.stack([x1 + e1, x2 + e2, x1 + zdir + e1, x1, x2])      # [5, 3]
    ixx = iy + iz
    bsec = [rho, emod, gmod, area, iy, iz, it, ixx if ixx > 0 else it, 1.0 / kxy if kxy > 0 else 0.0,
            1.0 / kxz if kxz > 0 else 0.0, sy, sz, efflen, phi]
    return np.concatenate([P[:, 0], P[:, 1], P[:, 2], bsec, [pin_a, pin_b, 0.0]])


def tet10_block(nx, ny, nz, ngen=10, seed=3, n_ext=4, jitter=0.05, emod=2.1e11, rny=0.3,
                shuffle_eq=False, with_recovery=True, n_beams=0, curved="all"):
    """Structured block of nx*ny*nz hexahedral cells, each split into 6 ten-node tetrahedra
    (type 41), FEDEM node order: corners 1,3,5,10, mid-edges 2,4,6,7,8,9 (itet.f label 300).
    Mid-edge nodes get a small jitter so edges are curved (non-constant Jacobians): curved="all" every
    mid-edge node (the worst case for the kernels), "surface" only those on the outer faces of the block
    (what a mesher produces: mid-side nodes leave the chord only where they are snapped to curved CAD
    surfaces; interior elements stay straight-sided), "none" no node.  n_beams BEAM2
    stiffeners (type 11) run along random cell edges; their nodes carry 6 DOFs like in a real model
    (the solids use the first three), some with eccentricities, shear-centre offsets and a rotated
    principal axis so that every branch of BEAM31 is exercised (config C3)."""
    rng = np.random.default_rng(seed)
    NX, NY, NZ = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    # nodes of the doubled grid that are used by the 6-tet split: all vertices, edge mids, face-diagonal
    # mids and the cell-diagonal mid -- numbered in the order the elements first touch them
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([iz.ravel(order="F"), iy.ravel(order="F"), ix.ravel(order="F")], 1)  # cz, cy, cx loops
    order = np.lexsort((cells[:, 2], cells[:, 1], cells[:, 0]))
    cells = cells[order]
    dv = np.array([(dx, dy, dz) for dz in (0, 2) for dy in (0, 2) for dx in (0, 2)])
    hexsel = [0, 1, 3, 2, 4, 5, 7, 6]
    tets = []
    for t in _TET_SPLIT:
        a, b, c, d = (dv[hexsel[q]] for q in t)
        if np.dot(np.cross(b - a, c - a), d - a) < 0:
            b, c = c, b
        mid = lambda p, q: (p + q) // 2
        tets.append(np.stack([a, mid(a, b), b, mid(b, c), c, mid(c, a), mid(a, d), mid(b, d), mid(c, d), d]))
    tets = np.stack(tets)                                            # [6, 10, 3] offsets in the doubled grid
    base = 2 * cells[:, [2, 1, 0]]                                   # [ncell, 3] (x, y, z)
    gidx = base[:, None, None, :] + tets[None]                       # [ncell, 6, 10, 3]
    key = (gidx[..., 2].astype(np.int64) * NY + gidx[..., 1]) * NX + gidx[..., 0]
    flat = key.reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    rank = np.empty(len(uniq), np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    node = rank[np.searchsorted(uniq, flat)] + 1                     # numbering by first appearance
    conn = node.reshape(-1, 10).astype(I32)
    nnod = len(uniq)
    kk = uniq[np.argsort(first, kind="stable")]
    gi = np.stack([kk % NX, (kk // NX) % NY, kk // (NX * NY)], 1)
    xyz = gi / 2.0
    odd = (gi % 2).sum(1) > 0
    shift = rng.uniform(-jitter, jitter, (int(odd.sum()), 3)) * 0.5
    if curved == "surface":
        on_surface = ((gi == 0) | (gi == np.array([NX - 1, NY - 1, NZ - 1]))).any(1)
        shift[~on_surface[odd]] = 0.0
    elif curved == "none":
        shift[:] = 0.0
    else:
        assert curved == "all", curved
    xyz[odd] += shift
    node_of = {int(k): i + 1 for i, k in enumerate(kk)}

    def nid(i, j, k):
        return node_of[(k * NY + j) * NX + i]

    cand = [(0, 0, 0), (NX - 1, 0, 0), (0, NY - 1, 0), (NX - 1, NY - 1, NZ - 1), (0, 0, NZ - 1),
            (NX - 1, NY - 1, 0), (0, NY - 1, NZ - 1), (NX - 1, 0, NZ - 1)]
    ext_nodes = [nid(*c) for c in cand[:n_ext]]
    types = np.full(len(conn), 41, I32)
    ndpn = np.full(nnod, 3, I32)
    beams = []
    conn_list = conn
    if n_beams > 0:
        conn_list = [c for c in conn]
        for b in range(n_beams):
            ax = int(rng.integers(0, 3))
            c0 = [int(rng.integers(0, n + (0 if a == ax else 1))) for a, n in enumerate((nx, ny, nz))]
            p0 = [2 * v for v in c0]
            p1 = list(p0); p1[ax] += 2
            n1, n2 = nid(*p0), nid(*p1)
            ndpn[n1 - 1] = ndpn[n2 - 1] = 6
            zdir = np.roll(np.array([0.0, 0.3, 1.0]), ax) + rng.normal(0, 0.1, 3)
            kw = {}
            if b % 3 == 1:
                kw = dict(ecc1=rng.normal(0, 0.02, 3), ecc2=rng.normal(0, 0.02, 3), sy=0.004, sz=-0.003)
            if b % 3 == 2:
                kw = dict(phi=25.0, sy=0.002, sz=0.001, efflen=0.9)
            beams.append((len(conn_list), beam_record(xyz[n1 - 1], xyz[n2 - 1], zdir, **kw)))
            conn_list.append(np.array([n1, n2], I32))
        types = np.concatenate([types, np.full(n_beams, 11, I32)])
    sam = _build_sam(nnod, ndpn if n_beams > 0 else 3, conn_list, types, ext_nodes, rng=rng, shuffle_eq=shuffle_eq)
    sam.ngen = ngen
    nel = sam.nel
    beam = None
    if beams:
        beam = np.zeros((nel, 32), F64)
        for e, rec in beams:
            beam[e] = rec
    elm = ElementData(xyz=xyz, emod=np.full(nel, emod, F64), rny=np.full(nel, rny, F64),
                      thk=np.zeros(nel, F64), elmid=np.arange(1, nel + 1, dtype=I32), beam=beam)
    part = PartModel(sam=sam, elm=elm, name=f"tet10_{nx}x{ny}x{nz}" + (f"+{n_beams}beams" if n_beams else ""))
    part.recovery_seed = seed
    if with_recovery:
        part.B, part.E = synthetic_recovery(part)
    return part


# FEDEM HEX20 node order (DN2031, src/Femlib/ihex.f:2441-2500): offsets on the doubled grid
_HEX20 = [(0, 0, 0), (1, 0, 0), (2, 0, 0), (2, 1, 0), (2, 2, 0), (1, 2, 0), (0, 2, 0), (0, 1, 0),
          (0, 0, 1), (2, 0, 1), (2, 2, 1), (0, 2, 1),
          (0, 0, 2), (1, 0, 2), (2, 0, 2), (2, 1, 2), (2, 2, 2), (1, 2, 2), (0, 2, 2), (0, 1, 2)]


def hex20_block(nx, ny, nz, ngen=8, seed=6, n_ext=4, jitter=0.04, emod=2.1e11, rny=0.3, shuffle_eq=False,
                with_recovery=True):
    """Structured block of nx*ny*nz 20-node hexahedra (type 43), jittered nodes (curved edges)."""
    rng = np.random.default_rng(seed)
    NX, NY, NZ = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    off = np.asarray(_HEX20)
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = 2 * np.stack([cx.ravel(), cy.ravel(), cz.ravel()], 1)
    gidx = base[:, None, :] + off[None]
    key = (gidx[..., 2].astype(np.int64) * NY + gidx[..., 1]) * NX + gidx[..., 0]
    flat = key.reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    rank = np.empty(len(uniq), np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    conn = (rank[np.searchsorted(uniq, flat)] + 1).reshape(-1, 20).astype(I32)
    kk = uniq[np.argsort(first, kind="stable")]
    gi = np.stack([kk % NX, (kk // NX) % NY, kk // (NX * NY)], 1)
    xyz = gi / 2.0 + rng.uniform(-jitter, jitter, (len(kk), 3)) * 0.5
    node_of = {int(k): i + 1 for i, k in enumerate(kk)}
    cand = [(0, 0, 0), (NX - 1, 0, 0), (0, NY - 1, 0), (NX - 1, NY - 1, NZ - 1), (0, 0, NZ - 1),
            (NX - 1, NY - 1, 0), (0, NY - 1, NZ - 1), (NX - 1, 0, NZ - 1)]
    ext_nodes = [node_of[(c[2] * NY + c[1]) * NX + c[0]] for c in cand[:n_ext]]
    sam = _build_sam(len(kk), 3, conn, np.full(len(conn), 43, I32), ext_nodes, rng=rng, shuffle_eq=shuffle_eq)
    sam.ngen = ngen
    nel = sam.nel
    elm = ElementData(xyz=xyz, emod=np.full(nel, emod, F64), rny=np.full(nel, rny, F64), thk=np.zeros(nel, F64),
                      elmid=np.arange(1, nel + 1, dtype=I32))
    part = PartModel(sam=sam, elm=elm, name=f"hex20_{nx}x{ny}x{nz}")
    part.recovery_seed = seed
    if with_recovery:
        part.B, part.E = synthetic_recovery(part)
    return part


# WEDG15 node order (DN1531, src/Femlib/ipri.f:2867-2971): bottom triangle corner, mid, corner, mid, corner, mid; the three
# mid-height corner nodes; the top triangle.  Offsets on the doubled grid for the two wedges of a cell.
_W15_TRI = [[(0, 0), (1, 0), (2, 0), (2, 1), (2, 2), (1, 1)], [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]]
_WEDG15 = [[(x, y, 0) for x, y in t] + [(t[0][0], t[0][1], 1), (t[2][0], t[2][1], 1), (t[4][0], t[4][1], 1)] +
           [(x, y, 2) for x, y in t] for t in _W15_TRI]


def wedg15_block(nx, ny, nz, ngen=6, seed=10, n_ext=4, jitter=0.04, emod=2.1e11, rny=0.3, shuffle_eq=False, with_recovery=True):
    """Structured block of nx*ny*nz cells, each split into two 15-node wedges (type 42), jittered nodes."""
    rng = np.random.default_rng(seed)
    NX, NY, NZ = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = 2 * np.stack([cx.ravel(), cy.ravel(), cz.ravel()], 1)
    off = np.asarray(_WEDG15)                                   # [2, 15, 3]
    gidx = (base[:, None, None, :] + off[None]).reshape(-1, 15, 3)
    key = (gidx[..., 2].astype(np.int64) * NY + gidx[..., 1]) * NX + gidx[..., 0]
    flat = key.reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    conn = (rank[np.searchsorted(uniq, flat)] + 1).reshape(-1, 15).astype(I32)
    kk = uniq[order]
    gi = np.stack([kk % NX, (kk // NX) % NY, kk // (NX * NY)], 1)
    xyz = gi / 2.0 + rng.uniform(-jitter, jitter, (len(kk), 3)) * 0.5
    node_of = {int(k): i + 1 for i, k in enumerate(kk)}
    cand = [(0, 0, 0), (NX - 1, 0, 0), (0, NY - 1, 0), (NX - 1, NY - 1, NZ - 1), (0, 0, NZ - 1), (NX - 1, NY - 1, 0),
            (0, NY - 1, NZ - 1), (NX - 1, 0, NZ - 1)]
    ext_nodes = [node_of[(c[2] * NY + c[1]) * NX + c[0]] for c in cand[:n_ext]]
    sam = _build_sam(len(kk), 3, conn, np.full(len(conn), 42, I32), ext_nodes, rng=rng, shuffle_eq=shuffle_eq)
    sam.ngen = ngen
    nel = sam.nel
    elm = ElementData(xyz=xyz, emod=np.full(nel, emod, F64), rny=np.full(nel, rny, F64), thk=np.zeros(nel, F64),
                      elmid=np.arange(1, nel + 1, dtype=I32))
    part = PartModel(sam=sam, elm=elm, name=f"wedg15_{nx}x{ny}x{nz}")
    part.recovery_seed = seed
    if with_recovery:
        part.B, part.E = synthetic_recovery(part)
    return part


def linsolid_block(nx, ny, nz, ngen=6, seed=8, n_ext=4, jitter=0.08, emod=2.1e11, rny=0.3, shuffle_eq=False,
                   kinds=(44, 45, 46), with_recovery=True):
    """Structured block of nx*ny*nz cells filled, cell by cell in turn, with the linear solids of `kinds`:
    one 8-node hexahedron (type 44, node order of HEXA32: bottom face counter-clockwise, then the top face),
    six 4-node tetrahedra (type 45, positive volume as cstetbmat demands) or two 6-node wedges (type 46, bottom
    triangle then top triangle).  Jittered nodes, so no element is a parallelepiped."""
    rng = np.random.default_rng(seed)
    NX, NY, NZ = nx + 1, ny + 1, nz + 1
    nid = lambda i, j, k: 1 + i + NX * (j + NY * k)
    gi = np.stack(np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij"), -1).reshape(-1, 3)
    xyz = np.zeros((NX * NY * NZ, 3))
    for i, j, k in gi:
        xyz[nid(i, j, k) - 1] = (i, j, k)
    xyz += rng.uniform(-jitter, jitter, xyz.shape)
    conn, types = [], []
    cell = 0
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)]
                t = kinds[cell % len(kinds)]
                cell += 1
                if t == 44:
                    conn.append(np.array(c, I32)); types.append(44)
                elif t == 46:
                    for tri in ((0, 1, 2), (0, 2, 3)):
                        conn.append(np.array([c[q] for q in tri] + [c[q + 4] for q in tri], I32)); types.append(46)
                else:   # six tetrahedra around the cell diagonal 0-6
                    for a, b in ((1, 2), (2, 3), (3, 7), (7, 4), (4, 5), (5, 1)):
                        tet = [c[0], c[a], c[b], c[6]]
                        X = xyz[np.array(tet) - 1]
                        if np.dot(np.cross(X[1] - X[0], X[2] - X[0]), X[3] - X[0]) < 0:
                            tet[1], tet[2] = tet[2], tet[1]
                        conn.append(np.array(tet, I32)); types.append(45)
    cand = [(0, 0, 0), (nx, 0, 0), (0, ny, 0), (nx, ny, nz), (0, 0, nz), (nx, ny, 0), (0, ny, nz), (nx, 0, nz)]
    ext_nodes = [nid(*c) for c in cand[:n_ext]]
    sam = _build_sam(len(xyz), 3, conn, np.asarray(types, I32), ext_nodes, rng=rng, shuffle_eq=shuffle_eq)
    sam.ngen = ngen
    nel = sam.nel
    elm = ElementData(xyz=xyz, emod=np.full(nel, emod, F64), rny=np.full(nel, rny, F64), thk=np.zeros(nel, F64),
                      elmid=np.arange(1, nel + 1, dtype=I32))
    part = PartModel(sam=sam, elm=elm, name=f"linsolid_{nx}x{ny}x{nz}")
    part.recovery_seed = seed
    if with_recovery:
        part.B, part.E = synthetic_recovery(part)
    return part


# thick shells on the doubled (u, v) grid of a cell: QUAD8 counter-clockwise from a corner (corner, mid-side, ...), the
# cell centre unused; two TRI6 per cell in SAM order (3 corners, then the mid-sides 1-2, 2-3, 3-1), the cell centre is the
# mid-side node of the shared diagonal
_Q8 = [(0, 0), (1, 0), (2, 0), (2, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
_T6 = [[(0, 0), (2, 0), (2, 2), (1, 0), (2, 1), (1, 1)], [(0, 0), (2, 2), (0, 2), (1, 1), (1, 2), (0, 1)]]


def thickshell_panel(nx, ny, ngen=6, seed=12, n_ext=4, jitter=0.1, emod=2.1e11, rny=0.3, thk=0.02, kinds=(31, 32),
                     curvature=(0.15, 0.10, 0.05), rotation=None, shuffle_eq=False, with_recovery=True):
    """Doubly curved panel of nx*ny cells meshed with 8-noded (type 32) and/or 6-noded (type 31) thick shells, 6 DOFs per
    node; cells alternate between the requested kinds.  z = a u^2 + b v^2 + c u v over the unit-cell grid, nodes jittered in
    (u, v) so that mid-side nodes stay well inside CHQA30's 1:3 bound; `rotation` (3x3) turns the whole panel, e.g. to put
    shell normals along the global X axis (LNCS30's fallback branch)."""
    rng = np.random.default_rng(seed)
    NX, NY = 2 * nx + 1, 2 * ny + 1
    conn_list, types = [], []
    for j in range(ny):
        for i in range(nx):
            kind = kinds[(i + j) % len(kinds)]
            pats = [_Q8] if kind == 32 else _T6
            for pat in pats:
                conn_list.append([(2 * j + dv) * NX + 2 * i + du for du, dv in pat])
                types.append(kind)
    flat = np.concatenate([np.asarray(c, np.int64) for c in conn_list])
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    lookup = rank[np.searchsorted(uniq, flat)] + 1
    conn, k = [], 0
    for c in conn_list:
        conn.append(lookup[k:k + len(c)].astype(I32))
        k += len(c)
    kk = uniq[order]
    gu, gv = (kk % NX).astype(F64), (kk // NX).astype(F64)
    corner = ((kk % NX) % 2 == 0) & ((kk // NX) % 2 == 0)
    u = gu / 2.0 + np.where(corner, rng.uniform(-jitter, jitter, len(kk)), 0.0)
    v = gv / 2.0 + np.where(corner, rng.uniform(-jitter, jitter, len(kk)), 0.0)
    # mid-side / centre nodes: between their (jittered) neighbours, then shifted along the edge by up to +-10 % of it
    cu, cv = {int(q): a for q, a in zip(kk, u)}, {int(q): a for q, a in zip(kk, v)}
    for n, q in enumerate(kk):
        if corner[n]:
            continue
        a, b = int(q % NX), int(q // NX)
        if a % 2 == 1 and b % 2 == 0:
            ends = [b * NX + a - 1, b * NX + a + 1]
        elif a % 2 == 0 and b % 2 == 1:
            ends = [(b - 1) * NX + a, (b + 1) * NX + a]
        else:
            ends = [(b - 1) * NX + a - 1, (b + 1) * NX + a + 1]
        w = 0.5 + rng.uniform(-0.1, 0.1)
        u[n] = (1 - w) * cu[ends[0]] + w * cu[ends[1]]
        v[n] = (1 - w) * cv[ends[0]] + w * cv[ends[1]]
    ca, cb, cc = curvature
    xyz = np.stack([u, v, ca * u * u / max(nx, 1) + cb * v * v / max(ny, 1) + cc * u * v / max(nx, ny, 1)], 1)
    if rotation is not None:
        xyz = xyz @ np.asarray(rotation, F64).T
    node_of = {int(q): i + 1 for i, q in enumerate(kk)}
    cand = [(0, 0), (NX - 1, 0), (0, NY - 1), (NX - 1, NY - 1), (NX // 2 // 2 * 2, 0), (0, NY // 2 // 2 * 2)]
    ext_nodes = []
    for c in cand:
        n = node_of.get(c[1] * NX + c[0])
        if n is not None and n not in ext_nodes and len(ext_nodes) < n_ext:
            ext_nodes.append(n)
    sam = _build_sam(len(kk), 6, conn, np.asarray(types, I32), ext_nodes, rng=rng, shuffle_eq=shuffle_eq)
    sam.ngen = ngen
    nel = sam.nel
    elm = ElementData(xyz=np.ascontiguousarray(xyz), emod=np.full(nel, emod, F64), rny=np.full(nel, rny, F64),
                      thk=np.full(nel, thk, F64), elmid=np.arange(1, nel + 1, dtype=I32))
    part = PartModel(sam=sam, elm=elm, name=f"thickshell_{nx}x{ny}")
    part.recovery_seed = seed
    if with_recovery:
        part.B, part.E = synthetic_recovery(part)
    return part


def synthetic_recovery(part, bbox=None, rows=None):
    """(B, E) of a synthetic part (or of an element block of it, given the whole part's bbox; or only `rows` of them)."""
    rng = np.random.default_rng([int(getattr(part, "recovery_seed", 0)), 7719])
    return _smooth_recovery_matrices(part.sam, part.elm.xyz, part.sam.ngen, rng, bbox=bbox, rows=rows)


# ------------------------------------------------------------------------------------------
# strain rosettes on a shell part (config C5)
# ------------------------------------------------------------------------------------------
def rosettes_on_part(part, nros, seed=5, rtype="TRIPLE_GAGE_45", zero_init_fraction=0.0, top_surface=True):
    """nros rosettes on randomly chosen shell elements (types 23/24) of `part`, placed like the
    reference's strain coats on the element surface (zPos = +-t/2), rosette X axis at a random
    in-plane angle, explicit position matrix as in the .fsi input format."""
    from .gage import Rosette
    rng = np.random.default_rng(seed)
    sam, elm = part.sam, part.elm
    shells = np.nonzero((sam.melcon == 24) | (sam.melcon == 23))[0]
    pick = rng.choice(shells, nros, replace=len(shells) < nros)
    out = []
    for k, e in enumerate(pick):
        nodes = sam.mmnpc[sam.mpmnpc[e] - 1: sam.mpmnpc[e + 1] - 1]
        X = elm.xyz[nodes - 1]
        if len(nodes) == 4:
            n = np.cross(X[2] - X[0], X[3] - X[1])
        else:
            n = np.cross(X[1] - X[0], X[2] - X[0])
        n = n / np.linalg.norm(n)
        ex = X[1] - X[0]
        ex = ex - n * (ex @ n)
        ex /= np.linalg.norm(ex)
        ey = np.cross(n, ex)
        ang = rng.uniform(0, 2 * np.pi)
        xr = np.cos(ang) * ex + np.sin(ang) * ey
        yr = np.cross(n, xr)
        rpos = np.stack([xr, yr, n, X.mean(0)], 1)
        t = float(elm.thk[e])
        out.append(Rosette(id=k + 1, nodes=[int(v) for v in nodes], rpos=rpos, type=rtype,
                           zpos=(0.5 * t if top_surface else -0.5 * t), emod=float(elm.emod[e]),
                           nu=float(elm.rny[e]), zero_init=bool(rng.random() < zero_init_fraction)))
    return out
