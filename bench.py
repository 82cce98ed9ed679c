#!/usr/bin/env python
"""bench.py -- headline benchmark of the stress-recovery hot path (BASELINE.json metric).

Metric   : element.time-step stress evaluations per second (one evaluation = all result points of
           one element at one time step, von Mises requested, envelope kept).
Workload : BASELINE.json configs[1] -- synthetic 1000x1000 ANDES-quad plate (1,000,000 elements,
           6.01 M DOF), 48 external DOFs + 50 component modes (n_red = 98), 10,000 time steps,
           full-field von Mises + envelope.  One bench "step" = one pass of the hot path
           (Q pack -> K1 DMMA expansion -> K2 element kernel with fused envelope) over one batch
           of --tile time steps (default 500), so the default --steps 20 covers the 10,000-step
           history once.  At N > 1 GPUs ONE part of N x 1,000,000 elements (a 1000 x 1000N plate) is
           cut into N element blocks by the library (fsr_split_elements / fsr_part_create_block;
           weak scaling: the per-GPU block stays at the named size); rank r recovers block r.  The
           reduced history is broadcast from rank 0 every step and the per-block envelopes are
           gathered to rank 0 at the end, both with the library's own NCCL communicator
           (fsr_comm_*), both inside the timed region.
Legs     : `value`  inputs resident in HBM, the full von Mises history of the tile written to HBM.
           `e2e`    through the host API with page-locked host buffers: per step the H2D copy of
                    the step's Q tile and the D2H read-back of the step's result (the running
                    envelopes, 2 x 8 B per result point) are inside the timed region, the read-back
                    of step i overlapping the compute of step i+1 (fsr_recover_async /
                    fsr_get_envelope_async); the per-step history stays on the device in this leg.
Checks   : the timed part is compared with the CPU oracle (full field, first steps of the last
           tile + the history of sampled elements) -> `parity`; the run fails above 1e-10.
Secondary: at N = 1 also config 3 with every element curved, config 5 (rosettes + rainflow, the
           reference's own compiled fatigue code as CPU arm) and the HEX20 kernel -> `secondary`; at
           every N the STRONG scaling of the fixed config-3 part (TET10 + beams) -> `strong_c3`.
Arms     : default           this repo's CUDA path through the C ABI (libfedem_b200.so)
           --impl reference  the reference's CPU algorithm (oracle/ restatement; the reference's
                             Fortran cannot be compiled in this image) on all host threads, on a
                             bounded sample of the same workload.
Prints ONE JSON line on rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "element_timestep_stress_evals_per_sec"
UNIT = "element*steps/s"
NRED_EXT_NODES = 8      # 48 external DOFs
NGEN = 50               # component modes
# algorithmic bytes per quad element.step: 24 nodal DOFs read (192 B) + 8 von Mises values written (64 B);
# in the in-plane form of flat regions an element reads 16 values (u, v, theta1, theta2 of its four nodes): 128 + 64 B
QUAD_BYTES = {"dense": 256, "flat": 256, "inplane": 192}
# DMMA flop per element.step: 18 DMMA.8x8x4 per 8 steps (24x24 operator) / 12 (membrane | bending blocks) / 8 (in-plane blocks)
QUAD_DMMA_FLOP = {"dense": 1152, "flat": 768, "inplane": 512}
DGEMM_PEAK = 35.45      # TFLOP/s, cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_fp64_peaks.txt)
DMMA_PEAK = 37.1        # TFLOP/s, DMMA issue peak measured by tools/microbench/fp64_peaks.cu (same file)
PARITY_TOL = 1.0e-10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=1000, help="plate is nx x nx quads per rank")
    ap.add_argument("--tile", type=int, default=500, help="time steps per bench step")
    ap.add_argument("--cpu-sample-nx", type=int, default=160)
    ap.add_argument("--cpu-sample-steps", type=int, default=0,
                    help="time steps of the CPU sample (0 = 128 for cpu_baseline ~10 s on one thread, 48 per reference-arm step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip config 3 / config 5 / strong-scaling blocks")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--c3-elements", type=int, default=2_000_000)
    ap.add_argument("--elem-order", type=int, default=0, help="0 = Morton order (default), 1 = SAM order")
    return ap.parse_args()


def workload_text(nx, tile):
    return (f"C2: {nx}x{nx} ANDES-quad plate block per GPU ({nx * nx} elements), n_red=48+50, {tile} time steps per bench step, "
            "full-field von Mises + envelope")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        # "under load": the upper half of the samples (idle samples before/after are dropped)
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(sm_load)) if sm_load else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class CpuSample:
    """The CPU restatement of the reference loop (per step: column-AXPY expansion, then every element
    rebuilt from its coordinates) on a bounded sample of the workload: an nx x nx sub-plate with the
    same reduced dimension (48 + 50)."""

    def __init__(self, nx, nsteps_total):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_bind
        from fedem_solvers_b200.model import plate_part, reduced_history
        self.o = oracle_bind.Oracle()
        self.part = plate_part(nx, nx, ngen=NGEN, n_ext=NRED_EXT_NODES, seed=2)
        self.b = self.o.bind_part(self.part)
        self.Q = reduced_history(self.part.sam.ndim, nsteps_total, seed=2)
        self.nel = self.part.sam.nel

    def run(self, s0, ns, nthreads):
        t0 = time.perf_counter()
        self.o.recover_history(self.b, self.Q[:, s0:s0 + ns], want_history=False, nthreads=nthreads)
        return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nx, ns = args.cpu_sample_nx, args.cpu_sample_steps or 48
    cs = CpuSample(nx, ns * (args.steps + args.warmup))
    for i in range(args.warmup):
        cs.run(i * ns, ns, cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cs.run((args.warmup + i) * ns, ns, cores)
    wall = time.perf_counter() - t0
    value = cs.nel * ns * args.steps / wall
    sample = (f"{nx}x{nx}-quad sub-plate ({cs.nel} elements, n_red=98) x {ns} time steps per bench step, "
              f"oracle C restatement of the reference loop, OpenMP over elements / DOF rows on {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_text(args.nx, args.tile)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
def oracle_parity(part, Q2, vm_rows, sample_hist=None):
    """max relative difference of the device results against the CPU oracle (TEST INFRASTRUCTURE used as the checker):
    vm_rows [k, npts] = von Mises of the k steps Q2 [ndim, k] at every result point of `part`; sample_hist = (Qall, elements,
    vm_hist_of_their_points) for the per-element history check.  Per value, relative to max(|oracle value|, 1e-6 of the field max)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind
    o = oracle_bind.Oracle()
    b = o.bind_part(part)
    vm_o, _, _ = o.recover_history(b, np.asfortranarray(Q2), want_history=True, nthreads=os.cpu_count() or 1)
    floor = 1.0e-6 * np.abs(vm_o).max()
    err = float((np.abs(vm_rows - vm_o) / np.maximum(np.abs(vm_o), floor)).max())
    return err, int(vm_o.size)


def secondary_hex20():
    """the 20-node hexahedron the north_star names beside the TET10: 250,000 elements x 256 steps per step, von Mises envelope"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs
    r = bench_configs.chex(argparse.Namespace(hex_elements=250_000, tile=256, steps=10))
    return {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "workload", "roofline", "k1", "gpu_launches")}


def secondary_c5(lib, peaks):
    """config 5 at its named size: 100,000 rosettes x 100,000 steps (tiles of 256), rainflow + damage"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs
    a = argparse.Namespace(gages=100_000, nsteps=99_840, tile=256)
    r = bench_configs.c5(a)
    out = {k: r[k] for k in ("metric", "value", "unit", "ms_total", "workload", "rainflow", "gpu_launches", "cycles_total")}
    # CPU baseline: the reference's OWN compiled C++ (FFpFatigue.C, FFpCycle.C, FFpSNCurve.C from oracle/_ref) on series
    # shaped like the device ones, one thread like fedem_gage
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_bind
        ref = oracle_bind.Reference()
        if not ref.available:
            raise RuntimeError("oracle/_ref/libfedem_ref.so is missing (built from /root/reference by __graft_entry__.build())")
        rng = np.random.default_rng(5)
        nser, ns = 64, 99_840
        t = np.arange(ns) * 1.0e-3
        series = [np.ascontiguousarray(sum(rng.normal(0, 30.0) * np.sin(2 * np.pi * rng.uniform(2, 60) * t + rng.uniform(0, 6.28))
                                           for _ in range(8))) for _ in range(nser)]
        t0 = time.perf_counter()
        ncyc = 0
        for x in series:
            ref.get_damage(x, 5.0, [15.117, 17.146, 4.0, 5.0])
            ncyc += max(ref.num_cycles(0.0, 1.0e30), 0)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nser * ns / dt, "unit": "samples/s", "cores": 1, "kind": "reference",
                               "sample": f"{nser} narrow-band series x {ns} samples through the reference's own compiled FFpFatigue (PVX + "
                                         f"rainflow + Miner sum + cycle count, {ncyc} cycles), {dt:.1f} s on one thread like fedem_gage; "
                                         "compare with rainflow.samples_per_s"}
    except Exception as e:  # the reference objects are built from /root/reference in the build container and travel with the repo
        out["cpu_baseline"] = {"unavailable": str(e)[:200]}
    torch.cuda.empty_cache()
    return out


def c3_block(args, rank, world, local_rank, comm, dev, lib, max_over_ranks, sync_all, curved="surface"):
    """config 3: ONE fixed part (1.97 M TET10 + 2 % beams) cut into `world` element blocks -> strong scaling.
    curved = which mid-edge nodes leave the chord of their edge: "surface" (default: the outer faces of the block, as a
    mesher leaves it -- interior elements are straight-sided and take the constant-Jacobian kernel) or "all" (every
    element curved: the worst case, all on the general DMMA kernel)."""
    import torch
    from fedem_solvers_b200 import StressRecovery, split_elements
    from fedem_solvers_b200.model import tet10_block, reduced_history, synthetic_recovery
    n = round((args.c3_elements / 6) ** (1 / 3))
    t0 = time.time()
    whole = tet10_block(n, n, n, ngen=50, seed=3, n_ext=16, n_beams=max(1, int(0.02 * 6 * n ** 3)), with_recovery=False, curved=curved)
    cuts = split_elements(whole, world)
    tile, steps, warm = 256, 10, 3
    rec = StressRecovery(whole, device=local_rank, step_tile=tile, block=cuts[rank] if world > 1 else None)
    rows, _ = rec.block_rows()
    B, E = synthetic_recovery(whole, rows=rows if world > 1 else None)
    rec.open_B_and_E_matrices(B, E)
    del B, E
    setup = time.time() - t0
    nel, ndim = whole.sam.nel, whole.sam.ndim
    ntet = int((whole.sam.melcon == 41).sum())
    ntet_blk = int((whole.sam.melcon[cuts[rank][0]:cuts[rank][1]] == 41).sum())
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    Q = torch.empty((tile * (steps + warm), ndim), dtype=torch.float64, device=dev)
    if rank == 0:
        Q.copy_(torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, tile * (steps + warm), seed=3).T)))
    nstrp = whole.nstrp()
    off = np.concatenate([[0], np.cumsum(nstrp)])
    pt0 = [int(off[a]) for a, _ in cuts]
    npts = [int(off[b] - off[a]) for a, b in cuts]
    env = torch.empty((2, int(off[-1])), dtype=torch.float64, device=dev) if rank == 0 else None

    def step(i):
        q = Q[i * tile:(i + 1) * tile]
        if world > 1:
            comm.broadcast(q.data_ptr(), q.numel(), 0, stream.cuda_stream)
        rec.recover_dev(q.data_ptr(), ndim, tile, None, 0, stream.cuda_stream)

    def gather():
        if world > 1:
            comm.gather_envelope(rec, pt0, npts, env[0].data_ptr() if rank == 0 else None, env[1].data_ptr() if rank == 0 else None, 0,
                                 stream.cuda_stream)

    for i in range(warm):
        step(i)
    gather()
    sync_all()
    rec.reset_envelope(); rec.timing_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(steps):
        step(warm + i)
    gather()
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    tm = rec.last_timing()
    k2 = tm["k2_ms"] / max(tm["tiles"], 1)
    k1 = tm["k1_ms"] / max(tm["tiles"], 1)
    k2_max, k1_max = max_over_ranks(k2), max_over_ranks(k1)
    fam = rec.family_counts().get("tet10", (0, 0, 0))
    rec.close()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    hbm = float(peaks_json().get("hbm_gbs", 6650.0))
    alg = 240.0 * ntet_blk * tile     # envelope only: 240 B of displacements read per TET10 element.step
    return {"config": "C3", "curved_elements": curved, "tet10_on_rank0": {"elements": fam[0], "straight_sided_kernel": fam[1], "general_kernel": fam[2]},
            "metric": METRIC, "value": nel * tile * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "strong",
            "steps": steps, "warmup": warm, "ms_per_step": ms / steps,
            "workload": f"{n}x{n}x{n} cells -> {ntet} TET10 + {nel - ntet} BEAM2 ({whole.sam.ndof} DOF), n_red={ndim}, ONE part cut into "
                        f"{world} element block(s), {tile} time steps per step, von Mises envelope; NCCL broadcast of Q per step + "
                        "gather of the envelopes inside the timed region",
            "roofline": {"kernel": "k2_tet10_steplane_vm_kernel, straight-sided + curved launches (rank 0 block)", "bound": "hbm",
                         "bound_note": "FP64 pipe in fact; reported against HBM as the north_star asks",
                         "achieved": alg / (k2 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": alg / (k2 * 1e-3) / 1e9 / hbm,
                         "ms_per_launch": k2, "algorithmic_bytes_per_launch": alg},
            "k1_ms": k1, "k2_ms_slowest_rank": k2_max, "k1_ms_slowest_rank": k1_max, "setup_s": setup}


def c3_cpu_baseline():
    """the CPU restatement on a small TET10 + beam block, one thread"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind
    from fedem_solvers_b200.model import tet10_block, reduced_history
    part = tet10_block(12, 12, 12, ngen=50, seed=3, n_ext=16, n_beams=200)
    o = oracle_bind.Oracle()
    b = o.bind_part(part)
    ns = 24
    Q = reduced_history(part.sam.ndim, ns + 1, seed=3)
    o.recover_history(b, Q[:, :1], want_history=False, nthreads=1)
    t0 = time.perf_counter()
    o.recover_history(b, Q[:, 1:], want_history=False, nthreads=1)
    dt = time.perf_counter() - t0
    return {"value": part.sam.nel * ns / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"12x12x12 cells ({part.sam.nel} TET10 + beams, n_red={part.sam.ndim}) x {ns} steps, {dt:.1f} s, oracle C restatement of "
                      "the reference loop (ITET32 per element and step), single thread"}


_PEAKS = None


def peaks_json():
    global _PEAKS
    if _PEAKS is None:
        try:
            _PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            _PEAKS = {}
    return _PEAKS


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from fedem_solvers_b200 import StressRecovery, Comm, load_library, split_elements
    from fedem_solvers_b200.model import plate_part, reduced_history, synthetic_recovery

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load_library()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    comm = None
    if world > 1:
        # the library's own NCCL communicator carries the data path; torch.distributed only passes the 128-byte id around
        def exchange(ident):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if ident is not None:
                t.copy_(torch.frombuffer(bytearray(ident), dtype=torch.uint8))
            dist.broadcast(t, src=0)
            return bytes(t.cpu().numpy().tobytes())
        comm = Comm(rank, world, local_rank, exchange)

    tile = args.tile
    nsteps_total = tile * (args.steps + args.warmup)
    # ---- setup (untimed): the rank's element block, recovery matrices, reduced history ----
    if world == 1:
        part = plate_part(args.nx, args.nx, ngen=NGEN, n_ext=NRED_EXT_NODES, seed=2)
        rec = StressRecovery(part, device=local_rank, step_tile=((tile + 63) // 64) * 64, elem_order=args.elem_order)
        cuts, whole = [(0, part.sam.nel)], part
    else:
        # one part, `world` element blocks, cut natively; the rank generates only its own rows of the synthetic [B|E]
        whole = plate_part(args.nx, args.nx * world, ngen=NGEN, n_ext=NRED_EXT_NODES, seed=2, ly=float(world), with_recovery=False)
        cuts = split_elements(whole, world)
        rec = StressRecovery(whole, device=local_rank, step_tile=((tile + 63) // 64) * 64, elem_order=args.elem_order, block=cuts[rank])
        rows, _ = rec.block_rows()
        B, E = synthetic_recovery(whole, rows=rows)
        rec.open_B_and_E_matrices(B, E)
        part = None
    nel = cuts[rank][1] - cuts[rank][0]
    ndim, npts, ndof_blk = rec.ndim, rec.npts, rec.ndof
    fam = rec.family_counts()
    n_flat = fam.get("quad", (0, 0, 0))[1]
    path = rec.vm_path_info()
    quad_path = "inplane" if path["quads_inplane"] == nel else ("flat" if n_flat == nel else "dense")
    k1_rows = path["inplane_rows"] + 128 * path["global_row_tiles"] if path["inplane_rows"] else ndof_blk
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)
    nstrp = whole.nstrp()
    off = np.concatenate([[0], np.cumsum(nstrp)])
    pt0 = [int(off[a]) for a, _ in cuts]
    npts_r = [int(off[b] - off[a]) for a, b in cuts]

    Q_host = torch.empty((nsteps_total, ndim), dtype=torch.float64).pin_memory()
    Q_np = reduced_history(ndim, nsteps_total, seed=2) if (rank == 0 or not args.no_parity) else None
    if rank == 0:
        Q_host.numpy()[:] = Q_np.T
    Q_dev = torch.empty((nsteps_total, ndim), dtype=torch.float64, device=dev)
    vm_tile = torch.empty((tile, npts), dtype=torch.float64, device=dev)
    env_host = torch.empty((2, 2, npts), dtype=torch.float64).pin_memory()
    env_root = torch.empty((2, int(off[-1])), dtype=torch.float64, device=dev) if (world > 1 and rank == 0) else None

    # ---------------- leg 1: device-resident inputs (`value`) ----------------
    if rank == 0:
        Q_dev.copy_(Q_host, non_blocking=True)
    if world > 1:
        comm.broadcast(Q_dev.data_ptr(), Q_dev.numel(), 0, stream.cuda_stream)
    torch.cuda.synchronize()

    def device_step(i):
        q = Q_dev[i * tile:(i + 1) * tile]
        if world > 1:
            comm.broadcast(q.data_ptr(), q.numel(), 0, stream.cuda_stream)  # the small reduced history goes to every element block
        rec.recover_dev(q.data_ptr(), ndim, tile, vm_tile.data_ptr(), npts, stream.cuda_stream)

    def gather():
        if world > 1:
            comm.gather_envelope(rec, pt0, npts_r, env_root[0].data_ptr() if rank == 0 else None,
                                 env_root[1].data_ptr() if rank == 0 else None, 0, stream.cuda_stream)

    for i in range(args.warmup):
        device_step(i)
    gather()   # brings up NCCL's point-to-point channels before the timed region
    sync_all()
    rec.reset_envelope()
    rec.timing_reset()
    lib.fsr_kernel_launches(1)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        device_step(args.warmup + i)
    gather()   # per-block envelopes gathered to rank 0 over NVLink
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.fsr_kernel_launches(0))
    clk = clocks.stop() if rank == 0 else None
    tm = rec.last_timing()
    value = world * nel * tile * args.steps / (ms * 1e-3)
    k2_ms = tm["k2_ms"] / max(tm["tiles"], 1)
    k1_ms = tm["k1_ms"] / max(tm["tiles"], 1)

    # ---------------- parity of the timed part against the CPU oracle ----------------
    parity = None
    if not args.no_parity:
        nchk = 2
        last = args.warmup + args.steps - 1
        Qchk = Q_np[:, last * tile:last * tile + nchk]
        if world == 1:
            blk_part = part
        else:
            from fedem_solvers_b200.partition import sub_part    # the Python cut: an independent check of the native one
            sp = sub_part(whole, cuts[rank][0], cuts[rank][1], with_matrices=False)
            blk_part = sp.part
            blk_part.B, blk_part.E = B, E
        err, nval = oracle_parity(blk_part, Qchk, vm_tile[:nchk].cpu().numpy())
        # envelope property on the whole timed history: max >= every value of the last tile, min <= (size-independent check)
        mx_d, mn_d = rec.envelope_dev_ptrs()
        env_now = torch.empty((2, npts), dtype=torch.float64, device=dev)
        rec.copy_envelope_dev(env_now[0].data_ptr(), env_now[1].data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        env_ok = bool((env_now[0] >= vm_tile.max(0).values).all().item() and (env_now[1] <= vm_tile.min(0).values).all().item())
        err = max_over_ranks(err)
        env_ok = max_over_ranks(0.0 if env_ok else 1.0) == 0.0
        parity = {"max_rel_vs_oracle": err, "values_compared_per_rank": nval, "tolerance": PARITY_TOL,
                  "what": f"von Mises at every result point of the timed part, first {nchk} steps of the last timed tile, per value "
                          "relative to max(|oracle|, 1e-6 field max); envelope >= / <= every value of the last tile",
                  "envelope_consistent": env_ok}
        if world > 1:
            del B, E
    del vm_tile
    torch.cuda.empty_cache()

    # ---------------- leg 2: end to end through the host API (`e2e`) ----------------
    qh = Q_host.numpy()

    def host_step(i):
        buf = env_host[i & 1]
        if world > 1:
            # rank 0 owns the history; the other ranks receive the tile over NCCL
            q = Q_dev[i * tile:(i + 1) * tile]
            if rank == 0:
                q.copy_(Q_host[i * tile:(i + 1) * tile], non_blocking=True)
            comm.broadcast(q.data_ptr(), q.numel(), 0, stream.cuda_stream)
            rec.recover_dev(q.data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
        else:
            rec.recover_async(qh[i * tile:(i + 1) * tile].T)                 # H2D of the step's Q (page-locked) inside
        rec.envelope_async(buf[0].numpy(), buf[1].numpy())                   # D2H of the step's result, overlapping the next step

    host_step(0)
    rec.synchronize()
    sync_all()
    rec.reset_envelope()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.perf_counter()
    h0.record()
    for i in range(args.steps):
        host_step(args.warmup + i)
    h1.record()
    rec.synchronize()          # the last read-back has landed in host memory
    wall_e2e = (time.perf_counter() - t_wall0) * 1e3
    sync_all()
    ms_e2e = max_over_ranks(max(h0.elapsed_time(h1), wall_e2e))
    e2e_value = world * nel * tile * args.steps / (ms_e2e * 1e-3)
    e2e_env_max = float(env_host[(args.warmup + args.steps - 1) & 1][0].max())

    # ---------------- secondary configurations ----------------
    secondary, strong = {}, None
    rec.close()
    del Q_dev
    torch.cuda.empty_cache()
    if not args.no_secondary:
        strong = c3_block(args, rank, world, local_rank, comm, dev, lib, max_over_ranks, sync_all)
        if world == 1:
            if not args.no_cpu_baseline:
                strong["cpu_baseline"] = c3_cpu_baseline()
            worst = c3_block(args, rank, world, local_rank, comm, dev, lib, max_over_ranks, sync_all, curved="all")
            secondary["C3_all_elements_curved"] = {k: worst[k] for k in ("value", "unit", "ms_per_step", "curved_elements", "tet10_on_rank0", "roofline",
                                                                          "k1_ms", "workload")}
            secondary["C5"] = secondary_c5(lib, peaks_json())
            secondary["HEX20"] = secondary_hex20()

    if rank != 0:
        if comm:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = peaks_json()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    k2_name = {"inplane": "k2_quad_planar_vm_kernel", "flat": "k2_quad_flat_vm_kernel", "dense": "k2_shell_vm_kernel<6>"}[quad_path]
    traffic, k1_traffic = None, None   # measured DRAM bytes per launch, from the committed ncu captures of this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        t = tj[k2_name]
        if t["nx"] == args.nx and t["tile"] == tile:
            traffic = {"bytes_per_launch": t["dram_bytes_per_launch"], "source": t["source"]}
        t = tj["k1_expand_kernel"]
        if t["nx"] == args.nx and t["tile"] == tile:
            k1_traffic = {"bytes_per_launch": t["dram_bytes_per_launch"], "source": t["source"]}
    except Exception:
        pass
    achieved = QUAD_BYTES[quad_path] * nel * tile / (k2_ms * 1e-3) / 1e9
    k1_flops = 2.0 * k1_rows * ndim * tile
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.nx, tile),
                   "elements_per_gpu": nel, "ndof_per_gpu": int(ndof_blk), "n_red": ndim,
                   "time_steps_per_step": tile, "time_steps_timed": tile * args.steps, "parallelism": f"element-block x{world}",
                   "l2": "inputs larger than L2 (U tile %.1f GB, vm tile %.1f GB per step)" %
                         (ndof_blk * tile * 8 / 1e9, npts * tile * 8 / 1e9)},
        "k2": {"kernel": k2_name + " (ANDES quad von Mises + envelope" +
                         {"inplane": ", flat region: in-plane rows (u, v, theta1, theta2 per node) and two 12 x 8 operator blocks)",
                          "flat": ", membrane / bending split of flat elements)", "dense": ")"}[quad_path],
               "bound": "hbm",
               "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
               "peak_source": peak_src, "traffic": traffic["bytes_per_launch"] if traffic else None,
               "traffic_source": traffic["source"] if traffic else None, "ms_per_launch": k2_ms,
               "algorithmic_bytes_per_launch": QUAD_BYTES[quad_path] * nel * tile,
               "algorithmic_bytes_per_element_step": QUAD_BYTES[quad_path],
               # what the counters say limits this kernel: `frac` counts every node's displacements once per element that
               # reads them (256 B per element.step, so it can pass 1 when L2 serves the node sharing); the DRAM
               # really moved and the FP64 pipe (DMMA + scalar FP64 share it) are below
               "dram_frac": (traffic["bytes_per_launch"] / (k2_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
               "fp64_pipe_frac": QUAD_DMMA_FLOP[quad_path] * nel * tile / (k2_ms * 1e-3) / 1e12 / DMMA_PEAK,
               "fp64_pipe_note": "DMMA flops issued / DMMA issue peak (37.1 TFLOP/s measured); the von Mises epilogue adds scalar "
                                 "FP64 on the same pipe (ncu: tensor + fp64 pipe active, profiles/)"},
        "k1": {"kernel": "k1_expand_kernel (DMMA.8x8x4, bulk-TMA staged)", "bound": "tensor", "ms_per_launch": k1_ms,
               "achieved": k1_flops / (k1_ms * 1e-3) / 1e12, "peak": DGEMM_PEAK, "unit": "TFLOP/s",
               "frac": k1_flops / (k1_ms * 1e-3) / 1e12 / DGEMM_PEAK,
               "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_fp64_peaks.txt); MEASURED_PEAKS.json holds no FP64 figure",
               "traffic": k1_traffic["bytes_per_launch"] if k1_traffic else None,
               "traffic_source": k1_traffic["source"] if k1_traffic else None,
               "rows_expanded": int(k1_rows), "rows_note": "nodal DOFs of the block" if quad_path != "inplane" else
                                "4 in-plane rows per node (the rotation into the plane is folded into the recovery operator once) "
                                f"instead of the {int(ndof_blk)} nodal DOFs; FSR_QUAD_PLANAR=0 gives the 6-row path",
               "algorithmic_flops_per_launch": k1_flops, "dmma_issue_peak_frac": k1_flops / (k1_ms * 1e-3) / 1e12 / DMMA_PEAK},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ndim * tile * 8),
                "d2h_bytes_per_step": int(2 * npts * 8), "ms_per_step": ms_e2e / args.steps,
                "api": "fsr_recover_async(page-locked host Q tile) + fsr_get_envelope_async(page-locked host) per step, fsr_synchronize at the "
                       "end; the result read back per step is the running von Mises ENVELOPE (max, min per result point), the "
                       "per-step history stays on the device in this leg",
                "last_envelope_max": e2e_env_max},
        "gpu_launches": launches, "clocks": clk, "parity": parity,
        "element_paths": {k: {"elements": v[0], "fast_path": v[1], "general": v[2]} for k, v in fam.items()},
        "vm_path": path,
    }
    # `roofline` = the kernel that takes the larger share of the step (the two are within a few percent of each other)
    dom = "k1" if k1_ms >= k2_ms else "k2"
    line["roofline"] = dict(line[dom], share_of_step=(k1_ms if dom == "k1" else k2_ms) / (k1_ms + k2_ms),
                            other_kernel={"k1": "k2", "k2": "k1"}[dom])
    if strong:
        line["strong_c3"] = strong
    if secondary:
        line["secondary"] = secondary
    if world == 1 and not args.no_cpu_baseline:
        ncs = args.cpu_sample_steps or 128
        cs = CpuSample(args.cpu_sample_nx, ncs + 1)
        cs.run(0, 1, 1)
        dt = cs.run(1, ncs, 1)
        v, n_el = cs.nel * ncs / dt, cs.nel
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{args.cpu_sample_nx}x{args.cpu_sample_nx}-quad sub-plate ({n_el} elements, n_red=98) x "
                      f"{ncs} time steps, {dt:.1f} s, single thread like the serial reference "
                      "(oracle C restatement; the reference's Fortran cannot be built in this image)"}
    print(json.dumps(line), flush=True)
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    if parity and (parity["max_rel_vs_oracle"] > PARITY_TOL or not parity["envelope_consistent"]):
        print(f"PARITY FAILURE: {parity}", file=sys.stderr, flush=True)
        sys.exit(3)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks"}))
        sys.exit(2)
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
