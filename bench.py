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
           cut into N element blocks by fedem_solvers_b200.partition (weak scaling: the per-GPU block
           stays at the named size); rank r recovers block r.  The reduced history is broadcast
           from rank 0 with NCCL every step and the per-block envelopes are gathered to rank 0 at
           the end, both inside the timed region.

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
QUAD_BYTES = 256        # algorithmic bytes per quad element.step: 192 read + 64 written


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
    ap.add_argument("--elem-order", type=int, default=0, help="0 = Morton order (default), 1 = SAM order")
    return ap.parse_args()


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
            "config": {"workload": "C2: 1M-element ANDES-quad plate, n_red=48+50, von Mises + envelope "
                                   "(bounded CPU sample, see cpu_baseline.sample)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from fedem_solvers_b200 import StressRecovery, load_library
    from fedem_solvers_b200.model import plate_part, reduced_history

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load_library()

    tile = args.tile
    nsteps_total = tile * (args.steps + args.warmup)
    # ---- setup (untimed): the rank's element block, recovery matrices, reduced history ----
    if world == 1:
        part = plate_part(args.nx, args.nx, ngen=NGEN, n_ext=NRED_EXT_NODES, seed=2)
    else:
        # one part, `world` element blocks: the rank builds the (matrix-free) part, cuts its block and
        # generates only its own rows of the synthetic [B|E]
        from fedem_solvers_b200.partition import split_elements, sub_part
        from fedem_solvers_b200.model import synthetic_recovery
        whole = plate_part(args.nx, args.nx * world, ngen=NGEN, n_ext=NRED_EXT_NODES, seed=2, ly=float(world),
                           with_recovery=False)
        bbox = (whole.elm.xyz.min(0), whole.elm.xyz.max(0))
        e0, e1 = split_elements(whole, world)[rank]
        part = sub_part(whole, e0, e1, with_matrices=False).part
        del whole
        part.B, part.E = synthetic_recovery(part, bbox=bbox)
    nel, ndim = part.sam.nel, part.sam.ndim
    rec = StressRecovery(part, device=local_rank, step_tile=((tile + 63) // 64) * 64, elem_order=args.elem_order)
    npts = rec.npts
    part.B = part.E = None  # host copies no longer needed
    stream = torch.cuda.current_stream()
    rec.set_stream(stream.cuda_stream)

    Q_host = torch.empty((nsteps_total, ndim), dtype=torch.float64).pin_memory()
    if rank == 0:
        Q_host.numpy()[:] = reduced_history(ndim, nsteps_total, seed=2).T
    Q_dev = torch.empty((nsteps_total, ndim), dtype=torch.float64, device=dev)
    vm_tile = torch.empty((tile, npts), dtype=torch.float64, device=dev)
    env_host = torch.empty((2, npts), dtype=torch.float64).pin_memory()
    gather_buf = [torch.empty((2, npts), dtype=torch.float64, device=dev) for _ in range(world)] \
        if (world > 1 and rank == 0) else None

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

    # ---------------- leg 1: device-resident inputs (`value`) ----------------
    if rank == 0:
        Q_dev.copy_(Q_host, non_blocking=True)
    if world > 1:
        dist.broadcast(Q_dev, src=0)
    torch.cuda.synchronize()

    def device_step(i):
        q = Q_dev[i * tile:(i + 1) * tile]
        if world > 1:
            dist.broadcast(q, src=0)  # the small reduced history goes to every element block
        rec.recover_dev(q.data_ptr(), ndim, tile, vm_tile.data_ptr(), npts, stream.cuda_stream)

    for i in range(args.warmup):
        device_step(i)
    if world > 1:  # bring up NCCL's point-to-point channels (gather = send/recv) before the timed region
        env = torch.empty((2, npts), dtype=torch.float64, device=dev)
        rec.copy_envelope_dev(env[0].data_ptr(), env[1].data_ptr(), stream.cuda_stream)
        dist.gather(env, gather_buf, dst=0)
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
    if world > 1:  # per-part envelopes gathered to rank 0 over NVLink
        rec.copy_envelope_dev(env[0].data_ptr(), env[1].data_ptr(), stream.cuda_stream)
        dist.gather(env, gather_buf, dst=0)
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.fsr_kernel_launches(0))
    clk = clocks.stop() if rank == 0 else None
    tm = rec.last_timing()
    value = world * nel * tile * args.steps / (ms * 1e-3)
    k2_ms = tm["k2_ms"] / max(tm["tiles"], 1)
    k1_ms = tm["k1_ms"] / max(tm["tiles"], 1)

    # ---------------- leg 2: end to end through the host API (`e2e`) ----------------
    qh = Q_host.numpy()
    mx_h, mn_h = env_host[0].numpy(), env_host[1].numpy()

    def host_step(i):
        if world > 1:
            # rank 0 owns the history; the other ranks receive the tile over NCCL, then use the host API
            q = Q_dev[i * tile:(i + 1) * tile]
            if rank == 0:
                q.copy_(Q_host[i * tile:(i + 1) * tile], non_blocking=True)
            dist.broadcast(q, src=0)
            rec.recover_dev(q.data_ptr(), ndim, tile, None, 0, stream.cuda_stream)
        else:
            rec.recover(qh[i * tile:(i + 1) * tile].T, want_history=False)  # H2D of the step's Q inside
        rec.envelope(mx_h, mn_h)                                           # D2H of the step's result

    host_step(0)
    sync_all()
    rec.reset_envelope()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    h0.record()
    for i in range(args.steps):
        host_step(args.warmup + i)
    h1.record()
    sync_all()
    ms_e2e = max_over_ranks(h0.elapsed_time(h1))
    e2e_value = world * nel * tile * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    traffic = None   # measured DRAM bytes per launch of the dominant kernel, from the committed ncu capture of this workload
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k2_shell_vm_kernel<6>"]
        if t["nx"] == args.nx and t["tile"] == tile:
            traffic = {"bytes_per_launch": t["dram_bytes_per_launch"], "source": t["source"]}
    except Exception:
        pass
    achieved = QUAD_BYTES * nel * tile / (k2_ms * 1e-3) / 1e9
    k1_flops = 2.0 * part.sam.ndof * ndim * tile
    dgemm_peak = 35.45  # TFLOP/s, cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_fp64_peaks.txt)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: {args.nx}x{args.nx} ANDES-quad plate block per GPU ({nel} elements, "
                               f"{part.sam.ndof} DOF), n_red=48+50, {tile} time steps per bench step "
                               f"({tile * args.steps} steps timed), full-field von Mises + envelope",
                   "elements_per_gpu": nel, "ndof_per_gpu": int(part.sam.ndof), "n_red": ndim,
                   "time_steps_per_step": tile, "parallelism": f"element-block x{world}",
                   "l2": "inputs larger than L2 (U tile %.1f GB, vm tile %.1f GB per step)" %
                         (part.sam.ndof * tile * 8 / 1e9, npts * tile * 8 / 1e9)},
        "roofline": {"kernel": "k2_shell_vm_kernel<6> (ANDES quad von Mises + envelope)", "bound": "hbm",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_source": peak_src, "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None, "ms_per_launch": k2_ms,
                     "algorithmic_bytes_per_launch": QUAD_BYTES * nel * tile},
        "k1": {"kernel": "k1_expand_kernel (DMMA.8x8x4)", "bound": "fp64 tensor", "ms_per_launch": k1_ms,
               "achieved": k1_flops / (k1_ms * 1e-3) / 1e12, "peak": dgemm_peak, "unit": "TFLOP/s",
               "frac": k1_flops / (k1_ms * 1e-3) / 1e12 / dgemm_peak,
               "peak_source": "cuBLAS DGEMM 8192^3 measured (profiles/r01_fp64_peaks.txt)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ndim * tile * 8),
                "d2h_bytes_per_step": int(2 * npts * 8), "ms_per_step": ms_e2e / args.steps,
                "api": "fsr_recover(host Q tile) + fsr_get_envelope(host)"},
        "gpu_launches": launches, "clocks": clk,
    }
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
    if world > 1:
        dist.destroy_process_group()


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
