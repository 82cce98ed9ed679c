#!/usr/bin/env python
"""bench_cli.py -- the drop-in executable on a large part: bin/fedem_stress on the files a reducer + solver run leaves
behind (.ftl, _SAM.fsm, _B.fmx, _E.fmx, fedem_solver.fsi, th_p_1.frs) -> stress results database (.frs).

Workload (default): 500 x 500 ANDES quads (250,000 elements, 1.5 M DOF), 8 triads (48 external DOFs) + 50 component
modes, 2,000 time steps, `-vmStress` (float file, the reference's default) and, as a second line, every measure
(`-vmStress -maxPStress ... -maxSStrain -stress -strain -SR`).  Reported: wall time of the whole program and where it
went (the executable's own split: input files / CUDA context / part on device / time loop = history read | device |
device-to-host | file), element.steps per second of wall time, and a parity check of stored values against the oracle
(TEST INFRASTRUCTURE use of oracle/: the checker, never the thing measured).

Prints one JSON line per run."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALL = ["-vmStress", "-maxPStress", "-minPStress", "-maxSStress", "-vmStrain", "-maxPStrain", "-minPStrain", "-maxSStrain",
       "-stress", "-strain", "-SR"]


def make_case(d, nx, ny, nsteps, ngen, n_ext, seed=2):
    from test_frs_cpu import _write_solver_file
    from fedem_solvers_b200.files import save_part
    from fedem_solvers_b200.fsi import SolverPart, write_fsi
    from fedem_solvers_b200.ftl import write_ftl
    from fedem_solvers_b200.model import plate_part
    t0 = time.perf_counter()
    part = plate_part(nx, ny, ngen=ngen, n_ext=n_ext, seed=seed)
    base = 40
    write_ftl(os.path.join(d, "plate.ftl"), part, comments=False)
    save_part(os.path.join(d, "plate"), part, checksum=7, part_id=base)
    rng = np.random.default_rng(seed)
    triads, tr_undef, sup, tri, gen = _write_solver_file(os.path.join(d, "th_p_1.frs"), rng, nsteps, n_ext, ngen, dt=0.001, step0=1,
                                                         sup_base=base)
    sp = SolverPart(base_id=base, user_id=1, descr="plate", ngen=ngen, sup_pos=sup[0], gravity=np.zeros(3), model_file="",
                    triad_base_id=np.array([t[0] for t in triads]), triad_user_id=np.array([t[1] for t in triads]),
                    ndofs=np.full(n_ext, 6), first_dof=np.zeros(n_ext, int), tr_undef=tr_undef, triad_ur=tri[0], gen_first_dof=0)
    write_fsi(os.path.join(d, "fedem_solver.fsi"), [sp])
    inputs = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    return part, base, (sup, tri, gen), inputs, time.perf_counter() - t0


def parse_log(text):
    out = {}
    m = re.search(r"Wall time ([\d.]+) s: input files ([\d.]+) \(CUDA context ([\d.]+) beside it\), part on device \+ B/E ([\d.]+), time loop ([\d.]+)", text)
    if m:
        out.update(program_s=float(m[1]), input_files_s=float(m[2]), cuda_context_s=float(m[3]), part_setup_s=float(m[4]),
                   time_loop_s=float(m[5]))
    m = re.search(r"Time loop: history read ([\d.]+) s \| device \(K1 \+ record kernels\) ([\d.]+) s \| device-to-host ([\d.]+) s \| file ([\d.]+) s \(([\d.]+) MB in (\d+) tiles\)", text)
    if m:
        out.update(history_read_s=float(m[1]), device_s=float(m[2]), d2h_s=float(m[3]), file_s=float(m[4]), file_mb=float(m[5]),
                   tiles=int(m[6]))
    return out


def run(exe, d, opts, rdbfile, stotm=1.0e9, env=None):
    cmd = [exe, "-cwd", d, "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs",
           "-rdbfile", rdbfile, "-statm", "0", "-stotm", repr(stotm), "-tinc", "0"] + opts
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
    return wall, parse_log(r.stdout)


def check_parity(d, part, base, hist, out_path, nsteps, nchk_steps=3, nelem=200, seed=7):
    """stored float values of `nelem` random elements at the first steps against the oracle"""
    import oracle_bind
    from test_frs_cpu import _build_finit_numpy
    from fedem_solvers_b200.frs import FrsReader
    from fedem_solvers_b200.fsi import read_fsi
    sup, tri, gen = hist
    o = oracle_bind.Oracle()
    b = o.bind_part(part)
    n_ext, ngen = tri.shape[1], gen.shape[1]
    tr = read_fsi(os.path.join(d, "fedem_solver.fsi"), base).tr_undef
    Q = _build_finit_numpy(sup[:nchk_steps], tri[:nchk_steps], tr, np.full(n_ext, 6), 1 + 6 * np.arange(n_ext), gen[:nchk_steps], 6 * n_ext + 1,
                           6 * n_ext + ngen)
    vm_o, _, _ = o.recover_history(b, Q, want_history=True, nthreads=os.cpu_count() or 1)
    rd = FrsReader(out_path)
    assert rd.nsteps == nsteps, (rd.nsteps, nsteps)
    rng = np.random.default_rng(seed)
    worst = 0.0
    for e in rng.choice(part.sam.nel, min(nelem, part.sam.nel), replace=False):
        p0 = b["ptoff"][e]
        for side, k in (("Top", 0), ("Bottom", 4)):
            h = rd.find(f"Elements|{int(part.elm.elmid[e])}|QUAD4|Element nodes|{side}|1|Von Mises stress", "Part", base)
            got = rd.read(h, 0, nchk_steps)[:, 0]
            want = vm_o[:, p0 + k]
            worst = max(worst, float(np.abs(got - want).max() / np.abs(want).max()))
    rd.close()
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=500)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--ngen", type=int, default=50)
    ap.add_argument("--next", type=int, default=8, help="external nodes (triads), 6 DOFs each")
    ap.add_argument("--dir", default=None, help="case directory (default: a temporary one, removed afterwards)")
    ap.add_argument("--all-steps", type=int, default=200, help="time steps of the every-measure run (its file is 14x larger per step)")
    ap.add_argument("--shm", action="store_true", help="also write the results database to /dev/shm (no disk in the way)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--writer-ab", action="store_true", help="every -vmStress run twice: pwritev helpers and FSR_RDB_MMAP=1 (mapped file)")
    ap.add_argument("--skip-all", action="store_true", help="no every-measure run")
    args = ap.parse_args()
    exe = os.path.join(ROOT, "fedem_solvers_b200", "bin", "fedem_stress")
    d = args.dir or tempfile.mkdtemp(prefix="bench_cli_")
    os.makedirs(d, exist_ok=True)
    try:
        part, base, hist, in_bytes, t_make = make_case(d, args.nx, args.ny, args.steps, args.ngen, args.next)
        nel = part.sam.nel
        common = {"metric": "element_timestep_stress_evals_per_sec", "unit": "element*steps/s", "n_gpus": 1, "dtype": "f64 compute, f32 file",
                  "elements": nel, "ndof": int(part.sam.ndof), "n_red": int(part.sam.ndim), "input_mb": in_bytes / 1e6, "case_setup_s": t_make}
        run(exe, d, ["-vmStress"], "warm.frs", stotm=0.0105)     # first start: page cache, CUDA module load
        targets = [("disk", os.path.join(d, "plate.frs"))]
        if args.shm and os.path.isdir("/dev/shm"):
            targets.append(("shm", "/dev/shm/bench_cli_plate.frs"))
        variants = [("pwritev", {"FSR_RDB_MMAP": "0"}), ("mmap", {"FSR_RDB_MMAP": "1"})] if args.writer_ab else [("default", {})]
        for where, rdbfile, (wname, wenv) in [(a, b, v) for a, b in targets for v in variants]:
            wall, split = run(exe, d, ["-vmStress"], rdbfile, env=wenv)
            out = rdbfile.replace(".frs", "_1.frs")
            line = dict(common, config="cli-vmStress", results_database=where, writer=wname, steps=args.steps, value=nel * args.steps / wall, seconds_wall=wall,
                        split=split, file_mb=os.path.getsize(out) / 1e6,
                        workload=f"bin/fedem_stress -vmStress: {args.nx}x{args.ny} ANDES quads, n_red={part.sam.ndim}, {args.steps} steps -> .frs on {where}")
            if not args.no_parity and where == "disk" and wname != "pwritev":
                line["parity_max_rel_vs_oracle_float_file"] = check_parity(d, part, base, hist, out, args.steps)
            print(json.dumps(line), flush=True)
            os.remove(out)
        # every measure + tensors + stress resultants (J2: the full-output path), fewer steps
        ns = min(args.all_steps, args.steps)
        if args.skip_all:
            return
        wall, split = run(exe, d, ALL, os.path.join(d, "plate_all.frs"), stotm=0.001 * (ns - 1) + 0.0005)
        out = os.path.join(d, "plate_all_1.frs")
        print(json.dumps(dict(common, config="cli-all-measures", results_database="disk", steps=ns, value=nel * ns / wall, seconds_wall=wall,
                              split=split, file_mb=os.path.getsize(out) / 1e6,
                              workload=f"bin/fedem_stress {' '.join(ALL)}: {args.nx}x{args.ny} ANDES quads, {ns} steps -> .frs on disk")),
              flush=True)
    finally:
        if not args.dir:
            shutil.rmtree(d, ignore_errors=True)
        for f in ("/dev/shm/bench_cli_plate_1.frs",):
            if os.path.exists(f):
                os.remove(f)


if __name__ == "__main__":
    main()
