"""FE part files (.ftl): reader binding and a writer for the synthetic parts.

The reader is csrc/io_ftl.cu (see include/fedem_b200.h, "FE part file"): it replaces the per-element
ffl_getcoor / ffl_getmat / ffl_getthick / ffl_getbeamsection / ffl_getpinflags / ffl_getelmid calls of
fedem_stress (fedem-foundation/src/FFlLib/FFlLinkHandler_F.C:699-1193) by one pass over the file.  The
writer emits the records the reference's FFlFedemWriter would (NODE, element records with {PMAT} {PTHICK}
{PBEAMSECTION} {PORIENT} {PBEAMECCENT} {PBEAMPIN} {PEFFLENGTH} references, attribute records, GROUP) so
that generated models can be fed to the stress driver through the same door as real ones."""
import ctypes as C
import os
import numpy as np

from . import _lib
from ._lib import check
from .model import ElementData

F64 = np.float64
I32 = np.int32
NAMES = {11: "BEAM2", 21: "TRI3", 23: "TRI3", 22: "QUAD4", 24: "QUAD4", 31: "TRI6", 32: "QUAD8", 41: "TET10",
         42: "WEDG15", 43: "HEX20", 44: "HEX8", 45: "TET4", 46: "WEDG6"}


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class FtlPart:
    """One parsed and resolved .ftl file (ffl_init)."""

    def __init__(self, path, groups=""):
        self.lib = _lib.load_library()
        self.h = C.c_void_p()
        check(self.lib.fsr_ftl_open(C.byref(self.h), os.fsencode(path)), "fsr_ftl_open")
        self.ignored_groups = 0
        if groups:
            self.ignored_groups = check(self.lib.fsr_ftl_activate_groups(self.h, groups.encode()), "fsr_ftl_activate_groups")

    def close(self):
        if self.h:
            self.lib.fsr_ftl_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sizes(self):
        """ffl_getsize: dict of the twelve size parameters + 'nael' (elements with the calculation flag on)."""
        sz = np.zeros(12, I32)
        nael = check(self.lib.fsr_ftl_sizes(self.h, _ip(sz)), "fsr_ftl_sizes")
        keys = ("nnod", "nel", "ndof", "nmnpc", "nmat", "nxnod", "npbeam", "nrgd", "nrbar", "nwavgm", "nprop", "ncons")
        d = {k: int(v) for k, v in zip(keys, sz)}
        d["nael"] = nael
        return d

    def nodes(self):
        """ffl_getnodes: (madof, minex, mnode, msc, xyz[nnod, 3])."""
        s = self.sizes()
        madof, minex, mnode = np.zeros(s["nnod"] + 1, I32), np.zeros(s["nnod"], I32), np.zeros(s["nnod"], I32)
        msc, xyz = np.zeros(s["ndof"], I32), np.zeros((s["nnod"], 3), F64)
        check(self.lib.fsr_ftl_get_nodes(self.h, _ip(madof), _ip(minex), _ip(mnode), _ip(msc), _dp(xyz)), "fsr_ftl_get_nodes")
        return madof, minex, mnode, msc, xyz

    def topology(self, use_andes=False):
        """ffl_gettopol: (melcon, mpmnpc, mmnpc)."""
        s = self.sizes()
        melcon, mpmnpc, mmnpc = np.zeros(s["nel"], I32), np.zeros(s["nel"] + 1, I32), np.zeros(max(s["nmnpc"], 1), I32)
        check(self.lib.fsr_ftl_get_topology(self.h, int(use_andes), _ip(melcon), _ip(mpmnpc), _ip(mmnpc)),
              "fsr_ftl_get_topology")
        return melcon, mpmnpc, mmnpc[:mpmnpc[-1] - 1]

    def element_data(self):
        """(ElementData, rho[nel], status[nel]) in SAM element order."""
        s = self.sizes()
        nel = s["nel"]
        emod, rny, rho, thk = (np.zeros(nel, F64) for _ in range(4))
        elmid, status, beam = np.zeros(nel, I32), np.zeros(nel, I32), np.zeros((nel, 32), F64)
        check(self.lib.fsr_ftl_get_elmdata(self.h, _dp(emod), _dp(rny), _dp(rho), _dp(thk), _ip(elmid), _dp(beam),
                                           _ip(status)), "fsr_ftl_get_elmdata")
        xyz = self.nodes()[4]
        return ElementData(xyz=xyz, emod=emod, rny=rny, thk=thk, elmid=elmid, beam=beam), rho, status

    def ext2int(self, ident, node=True):
        return int(self.lib.fsr_ftl_ext2int(self.h, int(node), int(ident)))

    def strain_coats(self):
        """ffl_getnostrc + ffl_getstraincoat for every strain coat element: list of dicts."""
        n = check(self.lib.fsr_ftl_num_strain_coats(self.h), "fsr_ftl_num_strain_coats")
        arr = (_lib.FsrStrainCoat * max(n, 1))()
        check(self.lib.fsr_ftl_get_strain_coats(self.h, arr, n), "fsr_ftl_get_strain_coats")
        out = []
        for c in arr[:n]:
            out.append(dict(id=c.id, nodes=list(c.nodes[:c.nnod]), npts=c.npts, elm_id=c.elm_id, mat_id=list(c.mat_id[:c.npts]),
                            res_set=list(c.res_set[:c.npts]), sn_curve=[tuple(c.sn_curve[k]) for k in range(c.npts)],
                            emod=list(c.emod[:c.npts]), nu=list(c.nu[:c.npts]), zpos=list(c.zpos[:c.npts]), scf=list(c.scf[:c.npts])))
        return out


def _fmt(v):
    return repr(float(v))


def write_ftl(path, part, groups=None, rho=7850.0, comments=True, strain_coats=None):
    """Writes `part` (model.PartModel) as an .ftl file.  groups: {id: [external element ids]}.
    strain_coats: list of dicts {id, elm (0-based element index of the shell underneath), sets: [(name, height or None)],
    fatigue: (snStd, snIdx, scf) or None}: one STRCT3 / STRCQ4 element on the shell's nodes with a PSTRC per result set
    (height given: PHEIGHT, else PTHICKREF +-0.5 to the shell's PTHICK) as FFlFedemWriter / the strain coat creator write them.
    Node ids = minex, element ids = elmid (or 1..nel); external nodes (all DOFs status 2) get status 1,
    nodes with suppressed DOFs the negative bit mask of FFlNode::isFixed."""
    sam, elm = part.sam, part.elm
    minex = sam.minex if sam.minex is not None else np.arange(1, sam.nnod + 1)
    ids = np.abs(elm.elmid) if elm.elmid is not None else np.arange(1, sam.nel + 1)
    mats, thks, secs, oris, eccs, pins, effs = {}, {}, {}, {}, {}, {}, {}

    def key(d, k):
        return d.setdefault(k, len(d) + 1)

    out = ["FTLVERSION{7 ASCII}"]
    if comments:
        out += ["#", "# Nodal coordinates", "#"]
    for n in range(sam.nnod):
        sc = sam.msc[sam.madof[n] - 1: sam.madof[n + 1] - 1]
        if (sc == 2).all():
            st = 1
        else:   # a constraint-equation DOF is not a node property; only plain suppressed DOFs are written
            st = -int(sum(1 << i for i, c in enumerate(sc) if c == 0 and sam.meqn[sam.madof[n] - 1 + i] == 0))
        x = elm.xyz[n]
        out.append(f"NODE{{{int(minex[n])} {st} {_fmt(x[0])} {_fmt(x[1])} {_fmt(x[2])}}}")
    if comments:
        out += ["#", "# Element definitions", "#"]
    for e in range(sam.nel):
        t = int(sam.melcon[e])
        nodes = [int(minex[k - 1]) for k in sam.mmnpc[sam.mpmnpc[e] - 1: sam.mpmnpc[e + 1] - 1]]
        if t == 31:   # the file lists a TRI6 around its perimeter; ffl_gettopol moves the mid-side nodes last
            nodes = [nodes[0], nodes[3], nodes[1], nodes[4], nodes[2], nodes[5]]
        refs = []
        if t == 11:
            b = elm.beam[e]
            X, Y, Z, sec = b[0:5], b[5:10], b[10:15], b[15:29]
            P = np.stack([X, Y, Z], 1)
            e1, e2, zdir = P[0] - P[3], P[1] - P[4], P[2] - P[0]
            m = key(mats, (sec[1], sec[2], 0.3, sec[0]))
            kxy = 1.0 / sec[8] if sec[8] > 0 else 0.0
            kxz = 1.0 / sec[9] if sec[9] > 0 else 0.0
            s = key(secs, (sec[3], sec[4], sec[5], sec[6], kxy, kxz, sec[10], sec[11], sec[13]))
            refs += [f"{{PMAT {m}}}", f"{{PBEAMSECTION {s}}}", f"{{PORIENT {key(oris, tuple(zdir))}}}"]
            if np.abs(e1).max() > 0 or np.abs(e2).max() > 0:
                refs.append(f"{{PBEAMECCENT {key(eccs, tuple(e1) + tuple(e2))}}}")
            if b[29] > 0 or b[30] > 0:
                refs.append(f"{{PBEAMPIN {key(pins, (int(b[29]), int(b[30])))}}}")
            if sec[12] != 0.0:
                refs.append(f"{{PEFFLENGTH {key(effs, (sec[12],))}}}")
        else:
            nu = float(elm.rny[e])
            m = key(mats, (float(elm.emod[e]), float(elm.emod[e]) / (2 * (1 + nu)), nu, rho))
            refs.append(f"{{PMAT {m}}}")
            if t in (21, 22, 23, 24, 31, 32):
                refs.insert(0, f"{{PTHICK {key(thks, (float(elm.thk[e]),))}}}")
        out.append(f"{NAMES[t]}{{{int(ids[e])} {' '.join(map(str, nodes))} {' '.join(refs)}}}")
    coat_lines, pstrc, prefs, pheights, pfats = [], [], {}, {}, {}
    for sc in strain_coats or []:
        e = int(sc["elm"])
        t = int(sam.melcon[e])
        nodes = [int(minex[k - 1]) for k in sam.mmnpc[sam.mpmnpc[e] - 1: sam.mpmnpc[e + 1] - 1]]
        nu = float(elm.rny[e])
        m = key(mats, (float(elm.emod[e]), float(elm.emod[e]) / (2 * (1 + nu)), nu, rho))
        refs = []
        for name, height in sc["sets"]:
            if height is not None:
                sub = f"{{PHEIGHT {key(pheights, (float(height),))}}}"
            else:
                fac = {"Bottom": -0.5, "Mid": 0.0, "Top": 0.5}[name]
                sub = f"{{PTHICKREF {key(prefs, (fac, key(thks, (float(elm.thk[e]),))))}}}"
            pstrc.append((len(pstrc) + 1, name, m, sub))
            refs.append(f"{{PSTRC {len(pstrc)}}}")
        if sc.get("fatigue") is not None:
            refs.append(f"{{PFATIGUE {key(pfats, tuple(sc['fatigue']))}}}")
        refs.append(f"{{FE {int(ids[e])}}}")
        coat_lines.append(f"{'STRCT3' if len(nodes) == 3 else 'STRCQ4'}{{{int(sc['id'])} {' '.join(map(str, nodes))} {' '.join(refs)}}}")
    if coat_lines:
        if comments:
            out += ["#", "# Strain coat elements", "#"]
        out += coat_lines
        for i, name, m, sub in pstrc:
            out.append(f"PSTRC{{{i} \"{name}\" {{PMAT {m}}} {sub}}}")
        for (fac, th), i in prefs.items():
            out.append(f"PTHICKREF{{{i} {_fmt(fac)} {{PTHICK {th}}}}}")
        for (h,), i in pheights.items():
            out.append(f"PHEIGHT{{{i} {_fmt(h)}}}")
        for (a, b, scf), i in pfats.items():
            out.append(f"PFATIGUE{{{i} {int(a)} {int(b)} {_fmt(scf)}}}")
    for title, name, table in (("Material properties", "PMAT", mats), ("Shell thicknesses", "PTHICK", thks),
                               ("Beam cross sections", "PBEAMSECTION", secs), ("Orientation vectors", "PORIENT", oris),
                               ("Beam eccentricities", "PBEAMECCENT", eccs), ("Beam pin flags", "PBEAMPIN", pins),
                               ("Effective beam lengths", "PEFFLENGTH", effs)):
        if not table:
            continue
        if comments:
            out += ["#", f"# {title}", "#"]
        for k, i in table.items():
            vals = " ".join(str(v) if isinstance(v, int) else _fmt(v) for v in k)
            out.append(f"{name}{{{i} {vals}}}")
    if groups:
        if comments:
            out += ["#", "# Element groups", "#"]
        for gid, els in groups.items():
            out.append(f"GROUP{{{gid} {' '.join(str(int(e)) for e in els)} {{NAME \"group {gid}\"}}}}")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
