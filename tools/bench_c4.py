#!/usr/bin/env python
"""bench_c4.py -- BASELINE.json config 4: a mechanism with 6 flexible superelements of mixed size and element type,
per-part stress recovery load-balanced across the GPUs of one box.

  python tools/bench_c4.py                                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_c4.py

The reference runs one fedem_stress process per part (the parts are independent).  Here the parts are laid end to end by
their measured element cost and the line is cut into N equal shares (partition.plan_work: element blocks can be cut
anywhere, so the load is divisible); a rank recovers its pieces one after the other, every piece a self-contained
element block (partition.sub_part) with the B/E rows of its own nodes.  Per step tile rank 0 broadcasts the reduced
histories of all parts (NCCL), at the end the per-piece von Mises envelopes are gathered to rank 0 (NCCL gather), both
inside the timed region.  Prints one JSON line on rank 0: whole-job element.steps/s, the per-rank device times (load
balance) and the plan."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_parts(scale, which="c4"):
    """Six parts of mixed size / type; `scale` multiplies the element counts (1.0 ~ 1.9 M elements in total).
    which = "c3": BASELINE config 3 instead, ONE mixed TET10 + beam part of ~2 M elements (element-sharded over the ranks)."""
    from fedem_solvers_b200.model import plate_part, tet10_block, hex20_block, linsolid_block
    if which == "c3":
        n = max(2, int(round((2_000_000 * scale / 6) ** (1.0 / 3.0))))
        return [tet10_block(n, n, n, ngen=50, seed=3, n_ext=16, n_beams=max(1, int(0.02 * 6 * n ** 3)), with_recovery=False)]
    # scale = 1: the part sizes SURVEY 8(d) names for config 4, {2 M, 1 M, 500 k, 250 k, 100 k, 50 k} elements
    s = scale ** 0.5
    c = scale ** (1.0 / 3.0)
    n = lambda v, f: max(2, int(round(v * f)))
    kw = dict(with_recovery=False)      # every rank generates only the B / E rows of its own element blocks
    return [
        plate_part(n(2000, s), n(1000, s), ngen=40, n_ext=8, seed=41, lx=2.0, **kw),                              # 2.00 M ANDES quads
        tet10_block(n(55, c), n(55, c), n(55, c), ngen=30, seed=43, n_ext=8, n_beams=2000, curved="surface", **kw),   # 1.00 M TET10 + beams
        plate_part(n(572, s), n(572, s), ngen=30, n_ext=6, seed=42, tri_fraction=0.5, **kw),                      # 0.50 M quads + triangles
        hex20_block(n(63, c), n(63, c), n(63, c), ngen=20, seed=44, n_ext=8, **kw),                               # 0.25 M HEX20
        linsolid_block(n(32, c), n(32, c), n(32, c), ngen=20, seed=45, n_ext=8, **kw),                            # 0.10 M HEX8 / TET4 / WEDG6
        plate_part(n(224, s), n(224, s), ngen=10, n_ext=4, seed=46, **kw),                                        # 0.05 M quads
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--parts", default="c4", choices=["c4", "c3"], help="c4: six mixed parts; c3: one 2 M TET10 + beam part")
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from fedem_solvers_b200 import StressRecovery, Comm, load_library
    from fedem_solvers_b200.model import reduced_history, synthetic_recovery
    from fedem_solvers_b200.partition import element_costs, plan_work, cost_fraction_to_elements

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load_library()
    comm = None
    if world > 1:   # the library's own NCCL communicator carries Q and the envelopes; torch.distributed passes the id around
        def exchange(ident):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if ident is not None:
                t.copy_(torch.frombuffer(bytearray(ident), dtype=torch.uint8))
            dist.broadcast(t, src=0)
            return bytes(t.cpu().numpy().tobytes())
        comm = Comm(rank, world, local_rank, exchange)
    t0 = time.time()
    parts = build_parts(args.scale, args.parts)          # every rank builds the (seeded, identical) parts and keeps only its pieces
    costs = []
    for p in parts:
        c = element_costs(p.sam.melcon, p.sam.mpmnpc, p.sam.mmnpc, p.sam.ndim)
        costs.append(float(c.sum()))
    items, loads = plan_work(costs, world)
    tile, steps, warm = args.tile, args.steps, args.warmup
    pieces = []
    for ip, f0, f1 in items[rank]:
        e0, e1 = cost_fraction_to_elements(parts[ip], f0, f1)
        if e1 <= e0:
            continue
        # native element block (fsr_part_create_block) + only its rows of the synthetic [B | E]
        rec = StressRecovery(parts[ip], device=local_rank, step_tile=((tile + 63) // 64) * 64, block=(e0, e1))
        rows, _ = rec.block_rows()
        B, E = synthetic_recovery(parts[ip], rows=rows)
        rec.open_B_and_E_matrices(B, E)
        del B, E
        pieces.append(dict(part=ip, e0=e0, e1=e1, rec=rec, nel=e1 - e0, npts=rec.npts))
    nel_total = sum(p.sam.nel for p in parts)
    part_sizes = [int(p.sam.nel) for p in parts]
    ndims = [p.sam.ndim for p in parts]
    del parts
    stream = torch.cuda.current_stream()
    for pc in pieces:
        pc["rec"].set_stream(stream.cuda_stream)
        pc["env"] = torch.empty((2, pc["npts"]), dtype=torch.float64, device=dev)
    # reduced histories of all parts, concatenated: [steps, sum ndim]; rank 0 owns them
    off = np.concatenate([[0], np.cumsum(ndims)])
    nq = tile * (steps + warm)
    Q = torch.empty((nq, int(off[-1])), dtype=torch.float64, device=dev)
    if rank == 0:
        Qh = np.concatenate([reduced_history(nd, nq, seed=50 + i).T for i, nd in enumerate(ndims)], axis=1)
        Q.copy_(torch.from_numpy(np.ascontiguousarray(Qh)))
    setup = time.time() - t0

    step_ev = []   # (before the broadcast, after it, after the pieces) of every timed step

    def step(i, timed=False):
        q = Q[i * tile:(i + 1) * tile]
        if timed:
            step_ev.append([torch.cuda.Event(enable_timing=True) for _ in range(3)])
            step_ev[-1][0].record()
        if world > 1:
            comm.broadcast(q.data_ptr(), q.numel(), 0, stream.cuda_stream)
        if timed:
            step_ev[-1][1].record()
        for pc in pieces:
            ip = pc["part"]
            qp = q[:, int(off[ip]):int(off[ip + 1])]      # [tile, ndim] view with row stride sum(ndim)
            pc["rec"].recover_dev(qp.data_ptr(), int(off[-1]), tile, None, 0, stream.cuda_stream)
        if timed:
            step_ev[-1][2].record()

    state = {}

    def gather():
        for pc in pieces:
            pc["rec"].copy_envelope_dev(pc["env"][0].data_ptr(), pc["env"][1].data_ptr(), stream.cuda_stream)
        if world > 1:
            mine = torch.cat([pc["env"] for pc in pieces], 1) if pieces else torch.empty((2, 0), dtype=torch.float64, device=dev)
            # NCCL gather wants equal shapes: pad every rank's envelope block to the largest one
            if "gather_pad" not in state:
                counts = [None] * world
                dist.all_gather_object(counts, int(mine.shape[1]))
                state["gather_pad"] = max(counts)
                state["gather_bufs"] = [torch.empty((2, max(counts)), dtype=torch.float64, device=dev) for _ in counts] if rank == 0 else None
                state["gather_mine"] = torch.zeros((2, max(counts)), dtype=torch.float64, device=dev)
            state["gather_mine"][:, :mine.shape[1]].copy_(mine)
            dist.gather(state["gather_mine"], state["gather_bufs"], dst=0)

    for i in range(warm):
        step(i)
    gather()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for pc in pieces:
        pc["rec"].reset_envelope()
        pc["rec"].timing_reset()
    lib.fsr_kernel_launches(1)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    # the envelope resets above are synchronous host-to-device copies whose length depends on the rank's pieces: every rank
    # starts its clock only when all of them are through (without this the ranks with small pieces spent the head start waiting
    # at the first broadcast -- 34 ms of 270 at N = 2, read as an 8 % "load imbalance" in the lines before R4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for i in range(steps):
        step(warm + i, timed=True)
    e1.record()
    gather()
    e2.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    my_compute, my_total = e0.elapsed_time(e1), e0.elapsed_time(e2)
    t = torch.tensor([my_compute, my_total], dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allt, t)
    else:
        allt = [t]
    # one more (untimed) step with an event between the pieces: the stream time each piece really takes, gaps included
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(pieces) + 2)]
    qx = Q[(warm + steps - 1) * tile:(warm + steps) * tile]
    evs[0].record()
    if world > 1:
        comm.broadcast(qx.data_ptr(), qx.numel(), 0, stream.cuda_stream)
    evs[1].record()
    for k, pc in enumerate(pieces):
        ip = pc["part"]
        pc["rec"].recover_dev(qx[:, int(off[ip]):int(off[ip + 1])].data_ptr(), int(off[-1]), tile, None, 0, stream.cuda_stream)
        evs[k + 2].record()
    torch.cuda.synchronize()
    bracket = [evs[k + 1].elapsed_time(evs[k + 2]) for k in range(len(pieces))]
    bcast_ms = evs[0].elapsed_time(evs[1])
    # the library's own per-piece K1 / K2 times (CUDA events on the part's stream): what the cost table is calibrated with
    mine_t = []
    for pc in pieces:
        lt = pc["rec"].last_timing()
        nt = max(1, int(lt["tiles"]))   # sums over the timed tiles of the ring
        mine_t.append(dict(part=pc["part"], elements=int(pc["nel"]), stream_ms=round(bracket[len(mine_t)], 4), bcast_ms=round(bcast_ms, 4), k1_ms=round(float(lt["k1_ms"]) / nt, 4), k2_ms=round(float(lt["k2_ms"]) / nt, 4),
                           ps_per_element_step=round(1e9 * float(lt["k1_ms"] + lt["k2_ms"]) / nt / max(1, pc["nel"] * tile), 2)))
    mine_t.append(dict(steps_bcast_ms=[round(e[0].elapsed_time(e[1]), 3) for e in step_ev],
                       steps_pieces_ms=[round(e[1].elapsed_time(e[2]), 3) for e in step_ev]))
    piece_t = [None] * world
    if world > 1:
        dist.all_gather_object(piece_t, mine_t)
    else:
        piece_t = [mine_t]
    if rank == 0:
        comp = [float(x[0]) for x in allt]
        tot = max(float(x[1]) for x in allt)
        print(json.dumps({
            "config": "C4" if args.parts == "c4" else "C3-sharded", "metric": "element_timestep_stress_evals_per_sec", "value": nel_total * tile * steps / (tot * 1e-3),
            "unit": "element*steps/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": tot / steps, "dtype": "f64",
            "scaling": "strong",
            "part_elements": part_sizes,
            "workload": (f"6 parts, {nel_total} elements in total (ANDES quads/triangles, TET10 + beams, HEX20, HEX8/TET4/WEDG6), "
                         if args.parts == "c4" else f"one part, {nel_total} elements (TET10 + 2 % beams), cut into element blocks, ") +
                        f"{tile} time steps per step, von Mises envelopes gathered to rank 0",
            "rank_compute_ms_per_step": [c / steps for c in comp],
            "pieces_per_rank": piece_t,
            "load_imbalance": max(comp) / (sum(comp) / len(comp)) - 1.0,
            "planned_load_share": [float(l / sum(loads)) for l in loads],
            "plan": [[(ip, round(f0, 4), round(f1, 4)) for ip, f0, f1 in it] for it in items],
            "gpu_launches_rank0": int(lib.fsr_kernel_launches(0)), "setup_s": setup}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
