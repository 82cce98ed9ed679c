#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configurations that are NOT the headline line of bench.py, measured on one
B200 with the same rules (CUDA events on the launching stream, >= 3 warm-up steps, inputs larger than L2):

  c3  mixed 10-node-tet solid + beam part (default 2,000,000 tets + 2 % beams), von Mises + envelope
      -> element.time-step evaluations/s, roofline of the TET10 kernel (320 algorithmic bytes per element.step)
  c5  strain-gage rosette recovery with rainflow counting and damage (default 100,000 rosettes, 100,000 steps)
      -> gage-point.time-step evaluations/s; the rosette GEMM against the FP64 peak, the streaming
         PVX + rainflow kernel in samples/s and GB/s of history consumed (8 B per sample)
  c1  the reference's own CPU-sized case (70 x 70 quads, n_red = 34, 1,000 steps) through the host API

One JSON line per configuration on stdout.  These are profile lines (profiles/), not the driver's bench contract."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HBM_PEAK = 6546.2
try:
    HBM_PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
DGEMM_PEAK = 35.45   # TFLOP/s, profiles/r01_fp64_peaks.txt


def c3(args):
    import torch
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import tet10_block, reduced_history
    lib = load_library()
    n = round((args.elements / 6) ** (1 / 3))
    t0 = time.time()
    part = tet10_block(n, n, n, ngen=50, seed=3, n_ext=16, n_beams=max(1, int(0.02 * 6 * n ** 3)), curved=args.curved)
    tile, steps, warm = args.tile, args.steps, 3
    rec = StressRecovery(part, device=0, step_tile=((tile + 63) // 64) * 64)
    setup = time.time() - t0
    nel, ndim, npts = part.sam.nel, part.sam.ndim, rec.npts
    ntet = int((part.sam.melcon == 41).sum())
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * (steps + warm), seed=3).T)).to(dev)
    for i in range(warm):
        rec.recover_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    torch.cuda.synchronize()
    rec.reset_envelope(); rec.timing_reset(); lib.fsr_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rec.recover_dev(Q[(warm + i) * tile:(warm + i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tm = rec.last_timing()
    k2 = tm["k2_ms"] / max(tm["tiles"], 1)
    k1 = tm["k1_ms"] / max(tm["tiles"], 1)
    # envelope only: 240 B read per TET10 element.step (the 80 B of von Mises stay in registers -> envelope)
    alg = 240.0 * ntet * tile
    mx, mn = rec.envelope()
    return {"config": "C3", "curved_elements": args.curved, "metric": "element_timestep_stress_evals_per_sec", "value": nel * tile * steps / (ms * 1e-3),
            "unit": "element*steps/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "dtype": "f64",
            "workload": f"{n}x{n}x{n} cells -> {ntet} TET10 + {nel - ntet} BEAM2, {part.sam.ndof} DOF, n_red={ndim}, {tile} time "
                        f"steps per step, von Mises envelope (no per-step history kept)",
            "roofline": {"kernel": "k2_tet10_vm_kernel (dense 60x30)" if os.environ.get("FSR_TET10_DENSE") else "k2_tet10_affine_vm_kernel (lane = corner x step)" if os.environ.get("FSR_TET10_STEPLANE") == "0" else "k2_tet10_steplane_vm_kernel (lane = step)", "bound": "hbm", "achieved": alg / (k2 * 1e-3) / 1e9, "peak": HBM_PEAK,
                         "unit": "GB/s", "frac": alg / (k2 * 1e-3) / 1e9 / HBM_PEAK, "ms_per_launch": k2,
                         "algorithmic_bytes_per_launch": alg,
                         "dmma_tflops_issued": (4096.0 if os.environ.get("FSR_TET10_DENSE") else 2304.0) * ntet * tile / (k2 * 1e-3) / 1e12},
            "k1": {"ms_per_launch": k1, "tflops": 2.0 * part.sam.ndof * ndim * tile / (k1 * 1e-3) / 1e12, "peak": DGEMM_PEAK},
            "gpu_launches": int(lib.fsr_kernel_launches(0)), "setup_s": setup, "max_von_mises": float(mx.max())}


def chex(args):
    """HEX20 block (type 43, named by the north_star): von Mises envelope."""
    import torch
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import hex20_block, reduced_history
    lib = load_library()
    n = round(args.hex_elements ** (1 / 3))
    part = hex20_block(n, n, n, ngen=50, seed=6, n_ext=16)
    tile, steps, warm = args.tile, args.steps, 3
    rec = StressRecovery(part, device=0, step_tile=((tile + 63) // 64) * 64)
    nel, ndim = part.sam.nel, part.sam.ndim
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * (steps + warm), seed=3).T)).to(dev)
    for i in range(warm):
        rec.recover_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    torch.cuda.synchronize()
    rec.reset_envelope(); rec.timing_reset(); lib.fsr_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rec.recover_dev(Q[(warm + i) * tile:(warm + i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tm = rec.last_timing()
    k2, k1 = tm["k2_ms"] / max(tm["tiles"], 1), tm["k1_ms"] / max(tm["tiles"], 1)
    alg = 480.0 * nel * tile
    mx, mn = rec.envelope()
    return {"config": "HEX20", "metric": "element_timestep_stress_evals_per_sec", "value": nel * tile * steps / (ms * 1e-3),
            "unit": "element*steps/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "dtype": "f64",
            "workload": f"{n}x{n}x{n} HEX20 ({nel} elements, {part.sam.ndof} DOF), n_red={ndim}, {tile} time steps per step, "
                        "von Mises envelope",
            "roofline": {"kernel": "k2_solid_smem_vm_kernel<20>" if os.environ.get("FSR_HEX20_DENSE") else "k2_bigsolid_grad_vm_kernel (DMMA gradient form)" if os.environ.get("FSR_HEX20_STEPLANE") == "0" else "k2_hex20_steplane_vm_kernel (natural derivatives at the nodes, lane = step)",
                         "bound": "hbm", "achieved": alg / (k2 * 1e-3) / 1e9, "peak": HBM_PEAK, "unit": "GB/s",
                         "frac": alg / (k2 * 1e-3) / 1e9 / HBM_PEAK, "ms_per_launch": k2, "algorithmic_bytes_per_launch": alg},
            "k1": {"ms_per_launch": k1, "tflops": 2.0 * part.sam.ndof * ndim * tile / (k1 * 1e-3) / 1e12, "peak": DGEMM_PEAK},
            "gpu_launches": int(lib.fsr_kernel_launches(0)), "max_von_mises": float(mx.max())}


def cthick(args):
    """Thick-shell panel (types 32 QUAD8 and 31 TRI6): dense per-element operators, k2_dense6_vm_kernel."""
    import torch
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import thickshell_panel, reduced_history
    lib = load_library()
    n = max(2, round((args.hex_elements / 1.5) ** 0.5))      # cells alternate QUAD8 / 2 x TRI6: 1.5 elements per cell
    part = thickshell_panel(n, n, ngen=50, seed=12, n_ext=4)
    tile, steps, warm = args.tile, args.steps, 3
    rec = StressRecovery(part, device=0, step_tile=((tile + 63) // 64) * 64)
    nel, ndim = part.sam.nel, part.sam.ndim
    nq, nt = int((part.sam.melcon == 32).sum()), int((part.sam.melcon == 31).sum())
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * (steps + warm), seed=3).T)).to(dev)
    for i in range(warm):
        rec.recover_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    torch.cuda.synchronize()
    rec.reset_envelope(); rec.timing_reset(); lib.fsr_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rec.recover_dev(Q[(warm + i) * tile:(warm + i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tm = rec.last_timing()
    k2, k1 = tm["k2_ms"] / max(tm["tiles"], 1), tm["k1_ms"] / max(tm["tiles"], 1)
    alg = (512.0 * nq + 384.0 * nt) * tile
    mx, mn = rec.envelope()
    return {"config": "THICK", "metric": "element_timestep_stress_evals_per_sec", "value": nel * tile * steps / (ms * 1e-3),
            "unit": "element*steps/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "dtype": "f64",
            "workload": f"{n}x{n} cell panel, {nq} QUAD8 + {nt} TRI6 thick shells ({part.sam.ndof} DOF), n_red={ndim}, {tile} time steps "
                        "per step, von Mises envelope",
            "roofline": {"kernel": "k2_dense6_vm_kernel<16,48> + <12,36>", "bound": "hbm", "achieved": alg / (k2 * 1e-3) / 1e9,
                         "peak": HBM_PEAK, "unit": "GB/s", "frac": alg / (k2 * 1e-3) / 1e9 / HBM_PEAK, "ms_per_launch_pair": k2,
                         "algorithmic_bytes_per_launch_pair": alg},
            "k1": {"ms_per_launch": k1, "tflops": 2.0 * part.sam.ndof * ndim * tile / (k1 * 1e-3) / 1e12, "peak": DGEMM_PEAK},
            "gpu_launches": int(lib.fsr_kernel_launches(0)), "max_von_mises": float(mx.max())}


def ctri(args):
    """All-triangle ANDES plate (type 23): k2_shell_vm_kernel<5>, 192 algorithmic bytes per element.step."""
    import torch
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import plate_part, reduced_history
    lib = load_library()
    n = max(2, round((args.elements / 2) ** 0.5))
    part = plate_part(n, n, ngen=50, n_ext=8, seed=2, tri_fraction=1.0)
    tile, steps, warm = args.tile, args.steps, 3
    rec = StressRecovery(part, device=0, step_tile=((tile + 63) // 64) * 64)
    nel, ndim = part.sam.nel, part.sam.ndim
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * (steps + warm), seed=3).T)).to(dev)
    for i in range(warm):
        rec.recover_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    torch.cuda.synchronize()
    rec.reset_envelope(); rec.timing_reset(); lib.fsr_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rec.recover_dev(Q[(warm + i) * tile:(warm + i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tm = rec.last_timing()
    k2, k1 = tm["k2_ms"] / max(tm["tiles"], 1), tm["k1_ms"] / max(tm["tiles"], 1)
    alg = 192.0 * nel * tile
    return {"config": "TRI", "metric": "element_timestep_stress_evals_per_sec", "value": nel * tile * steps / (ms * 1e-3),
            "unit": "element*steps/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "dtype": "f64",
            "workload": f"{n}x{n} cells -> {nel} ANDES triangles ({part.sam.ndof} DOF), n_red={ndim}, {tile} time steps per step, von Mises envelope",
            "roofline": {"kernel": "k2_shell_vm_kernel<5>", "bound": "hbm", "achieved": alg / (k2 * 1e-3) / 1e9, "peak": HBM_PEAK, "unit": "GB/s",
                         "frac": alg / (k2 * 1e-3) / 1e9 / HBM_PEAK, "ms_per_launch": k2, "algorithmic_bytes_per_launch": alg},
            "k1": {"ms_per_launch": k1, "tflops": 2.0 * part.sam.ndof * ndim * tile / (k1 * 1e-3) / 1e12, "peak": DGEMM_PEAK},
            "gpu_launches": int(lib.fsr_kernel_launches(0))}


def cmixed(args):
    """Config 4's mixed plate (half of the cells split into triangles, flat): quadrilaterals scattered among triangles; the
    library's cost rule decides between the in-plane rows and the global rows (FSR_QUAD_PLANAR=2 forces the in-plane form)."""
    import torch
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import plate_part, reduced_history
    lib = load_library()
    part = plate_part(572, 572, ngen=30, n_ext=6, seed=42, tri_fraction=0.5)
    tile, steps, warm = args.tile, args.steps, 3
    rec = StressRecovery(part, device=0, step_tile=((tile + 63) // 64) * 64)
    nel, ndim = part.sam.nel, part.sam.ndim
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * (steps + warm), seed=3).T)).to(torch.device("cuda", 0))
    for i in range(warm):
        rec.recover_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    torch.cuda.synchronize()
    rec.reset_envelope(); rec.timing_reset(); lib.fsr_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rec.recover_dev(Q[(warm + i) * tile:(warm + i + 1) * tile].data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tm = rec.last_timing()
    k2, k1 = tm["k2_ms"] / max(tm["tiles"], 1), tm["k1_ms"] / max(tm["tiles"], 1)
    return {"config": "MIXED", "metric": "element_timestep_stress_evals_per_sec", "value": nel * tile * steps / (ms * 1e-3),
            "unit": "element*steps/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "dtype": "f64",
            "workload": f"572x572 cells -> {nel} ANDES triangles + quadrilaterals ({part.sam.ndof} DOF), n_red={ndim}, {tile} time steps per step, von Mises envelope",
            "ps_per_element_step": 1e9 * (ms / steps) / (nel * tile), "k1_ms": k1, "k2_ms": k2, "vm_path": rec.vm_path_info(),
            "quad_planar_env": os.environ.get("FSR_QUAD_PLANAR", ""), "gpu_launches": int(lib.fsr_kernel_launches(0))}


def c5(args):
    import torch
    from fedem_solvers_b200 import StressRecovery, StrainGages, load_library
    from fedem_solvers_b200.model import plate_part, reduced_history, rosettes_on_part
    lib = load_library()
    part = plate_part(200, 200, ngen=50, n_ext=8, seed=5)
    rec = StressRecovery(part, device=0, step_tile=512)
    ros = rosettes_on_part(part, args.gages, seed=5)
    t0 = time.time()
    g = StrainGages(rec, ros)
    setup = time.time() - t0
    rec.close()
    ndim, tile = part.sam.ndim, args.tile
    ntiles = args.nsteps // tile
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    # a band-limited random history with enough amplitude to produce cycles above the gate
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * ntiles, seed=5, amp=2.0e-3).T)).to(dev)
    g.fatigue_begin(to_mpa=1.0e-6, gate=5.0, bin_size=10.0, nbins=8)
    for i in range(min(3, ntiles)):   # warm-up on throw-away state
        g.fatigue_feed_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, i * tile, tile, 1, stream.cuda_stream)
    g.fatigue_end()
    torch.cuda.synchronize()
    lib.fsr_kernel_launches(1)
    g.fatigue_begin(to_mpa=1.0e-6, gate=5.0, bin_size=10.0, nbins=8)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    pending, i = 1, 0
    while pending and i < ntiles:   # locate pass (normally one tile)
        pending = g.fatigue_feed_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, i * tile, tile, 0, stream.cuda_stream, want_pending=True)
        i += 1
    e1.record()
    for i in range(ntiles):
        g.fatigue_feed_dev(Q[i * tile:(i + 1) * tile].data_ptr(), ndim, i * tile, tile, 1, stream.cuda_stream)
    res = g.fatigue_end()
    e2.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e2)
    nseries = 4 * args.gages
    samples = float(nseries) * tile * ntiles
    return {"config": "C5", "metric": "gage_point_timestep_evals_per_sec", "value": args.gages * tile * ntiles / (ms * 1e-3),
            "unit": "rosette*steps/s", "n_gpus": 1, "ms_total": ms, "ms_locate_pass": e0.elapsed_time(e1), "dtype": "f64",
            "workload": f"{args.gages} TRIPLE_GAGE_45 rosettes x {tile * ntiles} time steps (tiles of {tile}), n_red={ndim}: "
                        "Bcart GEMM -> Mohr circle / leg stresses -> PVX + rainflow + Miner damage + 8-bin cycle histogram, "
                        f"{nseries} series",
            "rainflow": {"samples_per_s": samples / (ms * 1e-3), "history_gbs": 8.0 * samples / (ms * 1e-3) / 1e9,
                         "hbm_frac_if_materialised": 8.0 * samples / (ms * 1e-3) / 1e9 / HBM_PEAK},
            "gemm_tflops_equiv": 2.0 * 3 * args.gages * ndim * tile * ntiles / (ms * 1e-3) / 1e12,
            "gpu_launches": int(lib.fsr_kernel_launches(0)), "setup_s": setup,
            "cycles_total": int(res["ncycles"][res["ncycles"] > 0].sum()), "series_with_status": int((res["status"] != 0).sum()),
            "damage_max": float(res["damage"].max())}


def ccoat(args):
    """Strain coat recovery summary (envelopes + angle bins + biaxiality) of args.gages result points, device-resident history."""
    import ctypes as C
    import torch
    from fedem_solvers_b200 import StressRecovery, StrainGages, load_library
    from fedem_solvers_b200._lib import check
    from fedem_solvers_b200.model import plate_part, reduced_history, rosettes_on_part
    lib = load_library()
    part = plate_part(200, 200, ngen=50, n_ext=8, seed=5)
    rec = StressRecovery(part, device=0, step_tile=512)
    ros = rosettes_on_part(part, args.gages, seed=5)
    g = StrainGages(rec, ros)
    rec.close()
    ndim, tile = part.sam.ndim, args.tile
    ntiles = max(4, min(args.nsteps, 20480) // tile)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    Q = torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * ntiles, seed=5, amp=2.0e-3).T)).to(dev)
    check(lib.fsr_coat_begin(g._h, 541, 10.0e6), "fsr_coat_begin")
    for i in range(3):
        check(lib.fsr_coat_feed_dev(g._h, C.c_void_p(Q[i * tile:(i + 1) * tile].data_ptr()), ndim, tile, C.c_void_p(stream.cuda_stream)), "feed")
    torch.cuda.synchronize()
    check(lib.fsr_coat_begin(g._h, 541, 10.0e6), "fsr_coat_begin")
    lib.fsr_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ntiles):
        check(lib.fsr_coat_feed_dev(g._h, C.c_void_p(Q[i * tile:(i + 1) * tile].data_ptr()), ndim, tile, C.c_void_p(stream.cuda_stream)), "feed")
    env, summ, nb = np.zeros((8, args.gages)), np.zeros((6, args.gages)), np.zeros(args.gages, np.int32)
    # the work runs on the handle's own stream (torch's default stream handle is 0): fsr_coat_end synchronises it, so the
    # events bracket launch -> results on the host
    check(lib.fsr_coat_end(g._h, env.ctypes.data_as(C.POINTER(C.c_double)), summ.ctypes.data_as(C.POINTER(C.c_double)),
                           nb.ctypes.data_as(C.POINTER(C.c_int))), "fsr_coat_end")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"config": "COAT", "metric": "coat_point_timestep_evals_per_sec", "value": args.gages * tile * ntiles / (ms * 1e-3),
            "unit": "point*steps/s", "n_gpus": 1, "ms_total": ms, "ms_per_tile": ms / ntiles, "dtype": "f64",
            "workload": f"{args.gages} strain coat result points x {tile * ntiles} time steps (tiles of {tile}), n_red={ndim}: Bcart GEMM -> "
                        "Mohr circle -> running envelopes + 540 angle bins per point + biaxiality sums",
            "state_bytes": int(args.gages) * 540 * 36, "gpu_launches": int(lib.fsr_kernel_launches(0)),
            "mean_angle_spread_deg": float(summ[3].mean()), "mean_gated_steps": float(nb.mean())}


def c1(args):
    import torch
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import plate_part, reduced_history
    lib = load_library()
    part = plate_part(70, 70, ngen=10, n_ext=4, seed=1)
    rec = StressRecovery(part, device=0)
    Q = reduced_history(part.sam.ndim, 1000, seed=1)
    rec.recover(Q[:, :64], want_history=False)
    torch.cuda.synchronize()
    rec.reset_envelope(); lib.fsr_kernel_launches(1)
    t0 = time.perf_counter()
    vm = rec.recover(Q, want_history=True)     # host in, host out: the whole 1,000-step von Mises history
    mx, mn = rec.envelope()
    dt = time.perf_counter() - t0
    return {"config": "C1", "metric": "element_timestep_stress_evals_per_sec", "value": part.sam.nel * 1000 / dt,
            "unit": "element*steps/s", "n_gpus": 1, "seconds": dt, "dtype": "f64",
            "workload": f"70x70 ANDES quads ({part.sam.nel} elements, {part.sam.ndof} DOF), n_red={part.sam.ndim}, 1000 steps, "
                        "host Q in -> full von Mises history + envelope back on the host (fsr_recover + fsr_get_envelope)",
            "h2d_bytes": int(Q.nbytes), "d2h_bytes": int(vm.nbytes + mx.nbytes + mn.nbytes),
            "gpu_launches": int(lib.fsr_kernel_launches(0))}


def c1cli(args):
    """Config 1 through the drop-in executable: the files a reducer + solver run leaves behind (70 x 70 ANDES quads, 4 triads,
    10 component modes, 1,000 time steps) -> bin/fedem_stress -vmStress (float .frs, the reference's default) -> wall time of
    the whole program incl. reading the inputs, CUDA context creation and writing the 157 MB results database; next to it the
    CPU restatement of the reference's time loop (oracle, single thread like fedem_stress) on the same history."""
    import subprocess
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind
    from test_frs_cpu import _write_solver_file, _build_finit_numpy
    from fedem_solvers_b200.files import save_part
    from fedem_solvers_b200.fsi import SolverPart, write_fsi, read_fsi
    from fedem_solvers_b200.ftl import write_ftl
    from fedem_solvers_b200.frs import FrsReader
    from fedem_solvers_b200.model import plate_part
    part = plate_part(70, 70, ngen=10, n_ext=4, seed=1)
    d = tempfile.mkdtemp(prefix="c1cli_")
    rng = np.random.default_rng(1)
    nsteps, base = 1000, 40
    write_ftl(os.path.join(d, "plate.ftl"), part)
    save_part(os.path.join(d, "plate"), part, checksum=7, part_id=base)
    triads, tr_undef, sup, tri, gen = _write_solver_file(os.path.join(d, "th_p_1.frs"), rng, nsteps, 4, 10, dt=0.001, step0=1, sup_base=base)
    sp = SolverPart(base_id=base, user_id=1, descr="plate", ngen=10, sup_pos=sup[0], gravity=np.zeros(3), model_file="",
                    triad_base_id=np.array([t[0] for t in triads]), triad_user_id=np.array([t[1] for t in triads]), ndofs=np.full(4, 6),
                    first_dof=np.zeros(4, int), tr_undef=tr_undef, triad_ur=tri[0], gen_first_dof=0)
    write_fsi(os.path.join(d, "fedem_solver.fsi"), [sp])
    exe = os.path.join(ROOT, "fedem_solvers_b200", "bin", "fedem_stress")
    cmd = [exe, "-cwd", d, "-linkfile", "plate.ftl", "-samfile", "plate_SAM.fsm", "-fsifile", "fedem_solver.fsi", "-frsfile", "th_p_1.frs",
           "-vmStress", "-statm", "0", "-stotm", "10", "-tinc", "0"]
    subprocess.run(cmd, capture_output=True, text=True)          # first run: CUDA module load, page cache
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout
    out = os.path.join(d, "plate_1.frs")
    rd = FrsReader(out)
    assert rd.nsteps == nsteps
    # the CPU restatement on the same reduced history
    o = oracle_bind.Oracle()
    b = o.bind_part(part)
    tr = read_fsi(os.path.join(d, "fedem_solver.fsi"), base).tr_undef
    Q = _build_finit_numpy(sup, tri, tr, np.full(4, 6), 1 + 6 * np.arange(4), gen, 25, 34)
    ns_cpu = 100
    t0 = time.perf_counter()
    vm_o, _, _ = o.recover_history(b, Q[:, :ns_cpu], want_history=True, nthreads=1)
    cpu = time.perf_counter() - t0
    h = rd.find(f"Elements|{int(part.elm.elmid[17])}|QUAD4|Element nodes|Top|1|Von Mises stress", "Part", base)
    got = rd.read(h)[:ns_cpu, 0]
    p0 = b["ptoff"][17]
    err = float(np.abs(got - vm_o[:, p0]).max() / np.abs(vm_o).max())
    return {"config": "C1-cli", "metric": "element_timestep_stress_evals_per_sec", "value": part.sam.nel * nsteps / wall, "unit": "element*steps/s",
            "n_gpus": 1, "seconds_wall": wall, "dtype": "f64 compute, f32 file",
            "workload": f"bin/fedem_stress -vmStress on generated reducer/solver files: {part.sam.nel} ANDES quads, n_red=34, {nsteps} steps; "
                        f"results database {os.path.getsize(out) / 1e6:.0f} MB",
            "cpu_port": {"value": part.sam.nel * ns_cpu / cpu, "unit": "element*steps/s", "cores": 1, "seconds_for_1000_steps": cpu * nsteps / ns_cpu,
                         "sample": f"oracle restatement of the reference time loop, {ns_cpu} of the {nsteps} steps, compute only (no file output)"},
            "max_rel_diff_vs_oracle_float_file": err}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c3", "c5"])
    ap.add_argument("--elements", type=int, default=2_000_000, help="c3: number of TET10 elements")
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--hex-elements", type=int, default=250_000)
    ap.add_argument("--gages", type=int, default=100_000)
    ap.add_argument("--nsteps", type=int, default=100_000)
    ap.add_argument("--curved", default="all", choices=["all", "surface", "none"], help="c3: which TET10 mid-edge nodes leave the chord")
    args = ap.parse_args()
    for c in args.configs:
        print(json.dumps({"c1": c1, "c1cli": c1cli, "c3": c3, "c5": c5, "hex20": chex, "thick": cthick, "coat": ccoat, "tri": ctri, "mixed": cmixed}[c](args)), flush=True)


if __name__ == "__main__":
    main()
