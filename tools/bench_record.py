#!/usr/bin/env python
"""bench_record.py -- the full-output path of the results database on the device ("von Mises /
max-principal / max-shear evaluation, reported as achieved HBM GB/s").

One part (default 500 x 500 ANDES quads, n_red = 98) through fsr_rdb_create / fsr_rdb_write_steps with the file on
/dev/shm, once per output selection.  The library times its own tiles with CUDA events on the stream the kernels run on
(fsr_rdb_flush): K1 (H2D + expansion) and the record kernels separately from the PCIe copy and the file.  Reported per
selection: element.steps/s of the record kernels alone and with K1, their algorithmic bytes (8 B x element DOFs read +
bytes of the record written, per element.step) against the measured HBM peak, and the pipeline stages.

  vm      -vmStress                         (tuned von Mises kernel writing float records)
  all     every measure + tensors + SR      (record_points_dmma_kernel: DMMA stress rows -> shared memory -> invariants)
  all64   the same with -double

Prints one JSON line per selection."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=500)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--part", default="plate", choices=["plate", "tets"])
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--only", default="", help="comma list of selections (vm, all, all64)")
    args = ap.parse_args()
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import plate_part, tet10_block, reduced_history
    from fedem_solvers_b200.rdb import StressRdb, out_mask
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    lib = load_library()
    if args.part == "plate":
        part = plate_part(args.nx, args.nx, ngen=50, n_ext=8, seed=2)
        nedof, nstrp, ncmp, nsr = 24, 8, 3, 24
    else:
        n = args.nx
        part = tet10_block(n, n, n, ngen=50, seed=3, n_ext=16, curved="surface")
        nedof, nstrp, ncmp, nsr = 30, 10, 6, 0
    rec = StressRecovery(part, device=0, step_tile=256)
    nel, ndim = part.sam.nel, part.sam.ndim
    Q = reduced_history(ndim, args.steps, seed=2)
    stepno, tm = np.arange(1, args.steps + 1), 1e-3 * np.arange(args.steps)
    every = dict(vmStress=True, maxPStress=True, minPStress=True, maxSStress=True, vmStrain=True, maxPStrain=True, minPStrain=True,
                 maxSStrain=True, stress=True, strain=True, SR=True)
    sels = {"vm": (out_mask(vmStress=True), False, nstrp * 1), "all": (out_mask(**every), False, nsr + nstrp * (2 * ncmp + 8)),
            "all64": (out_mask(**every), True, nsr + nstrp * (2 * ncmp + 8))}
    for name, (mask, dbl, nval) in sels.items():
        if args.only and name not in args.only.split(","):
            continue
        path = os.path.join(args.dir, f"bench_record_{name}.frs")
        lib.fsr_kernel_launches(1)
        with StressRdb(rec, path, mask, double=dbl, rdbinc=0, base_id=1, user_id=1, descr="bench", elmid=part.elm.elmid, minex=part.sam.minex) as rdb:
            rdb.write_steps(Q[:, :64], stepno[:64], tm[:64])          # warm-up (first launches, page faults of the buffers)
            t0 = rdb.flush()                                            # cumulative: subtracted below
            t_w0 = time.perf_counter()
            rdb.write_steps(Q, stepno, tm)
            t1 = rdb.flush()
            wall = time.perf_counter() - t_w0
        os.remove(path)
        rec_ms = (t1["compute_ms"] - t1["k1_ms"]) - (t0["compute_ms"] - t0["k1_ms"])
        k1_ms = t1["k1_ms"] - t0["k1_ms"]
        vb = 8 if dbl else 4
        alg = (8.0 * nedof + vb * nval) * nel * args.steps
        print(json.dumps({
            "config": f"record-{name}", "metric": "element_timestep_stress_evals_per_sec", "unit": "element*steps/s", "n_gpus": 1,
            "value_record_kernels": nel * args.steps / (rec_ms * 1e-3), "value_with_k1": nel * args.steps / ((rec_ms + k1_ms) * 1e-3),
            "value_wall_incl_pcie_and_file": nel * args.steps / wall,
            "workload": f"{part.name}: {nel} elements, n_red={ndim}, {args.steps} steps, {nval} {'double' if dbl else 'float'} values per element.step "
                        f"into step records ({name}), file on {args.dir}",
            "roofline": {"kernel": "k2_shell_vm_kernel<..., float, record>" if name == "vm" and args.part == "plate" else "record_points_dmma_kernel",
                         "bound": "hbm", "achieved": alg / (rec_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg / (rec_ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_element_step": 8.0 * nedof + vb * nval,
                         "ms_record_kernels": rec_ms, "ms_k1": k1_ms},
            "pipeline_ms": {"device": t1["compute_ms"] - t0["compute_ms"], "d2h": t1["d2h_ms"] - t0["d2h_ms"], "file": t1["disk_ms"] - t0["disk_ms"],
                            "wall": wall * 1e3, "tiles": t1["tiles"] - t0["tiles"], "file_mb": (t1["bytes"] - t0["bytes"]) / 1e6},
            "gpu_launches": int(lib.fsr_kernel_launches(0))}), flush=True)
    rec.close()


if __name__ == "__main__":
    main()
