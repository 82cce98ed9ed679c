"""GPU tests of the native element-block / multi-GPU layer (csrc/sharded.cu): blocks reproduce the unsharded part bit for
bit, the in-process group (NCCL broadcast of Q, NCCL gather of the envelopes) and the one-process-per-GPU communicator
return the parent's result-point order.  The multi-device cases skip on a one-GPU box."""
import os
import subprocess
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from fedem_solvers_b200 import StressRecovery, GroupRecovery, split_elements  # noqa: E402
from fedem_solvers_b200.model import plate_part, tet10_block, reduced_history  # noqa: E402


def _parts():
    return [plate_part(13, 11, ngen=6, seed=3, tri_fraction=0.3, shuffle_eq=True, n_fixed=3, n_constraints=4, warp=0.02),
            tet10_block(4, 3, 2, ngen=5, seed=4, shuffle_eq=True, n_beams=7)]


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ip", [0, 1])
def test_blocks_on_one_device_reproduce_the_unsharded_part_bit_for_bit(ip):
    part = _parts()[ip]
    Q = reduced_history(part.sam.ndim, 77, seed=2)
    whole = StressRecovery(part, step_tile=64)
    vm = whole.recover(Q)
    mx, mn = whole.envelope()
    U = whole.calc_int_displacements(Q[:, :3])
    for nb in (2, 3):
        cuts = split_elements(part, nb)
        assert cuts[0][0] == 0 and cuts[-1][1] == part.sam.nel
        npts = 0
        for e0, e1 in cuts:
            blk = StressRecovery(part, step_tile=64, block=(e0, e1))      # takes the PARENT's B and E and picks its rows
            assert blk.is_block and blk.parent_npts == whole.npts and blk.ndim == whole.ndim
            got = blk.recover(Q)
            bmx, bmn = blk.envelope()
            assert np.array_equal(got, vm[:, blk.pt0:blk.pt0 + blk.npts])
            assert np.array_equal(bmx, mx[blk.pt0:blk.pt0 + blk.npts]) and np.array_equal(bmn, mn[blk.pt0:blk.pt0 + blk.npts])
            # the block's nodal displacements are the parent's for its nodes
            rows, nodes = blk.block_rows()
            Ub = blk.calc_int_displacements(Q[:, :3])
            madof = part.sam.madof
            dofs = np.concatenate([np.arange(madof[n - 1] - 1, madof[n] - 1) for n in nodes])
            assert np.array_equal(Ub, U[:, dofs])
            npts += blk.npts
            blk.close()
        assert npts == whole.npts
    whole.close()


def test_group_on_one_device_is_the_part():
    part = _parts()[0]
    Q = reduced_history(part.sam.ndim, 40, seed=5)
    whole = StressRecovery(part)
    vm = whole.recover(Q)
    mx, mn = whole.envelope()
    grp = GroupRecovery(part, devices=[0])
    assert grp.nblocks == 1 and grp.npts == whole.npts
    assert np.array_equal(grp.recover(Q), vm)
    gmx, gmn = grp.envelope()
    assert np.array_equal(gmx, mx) and np.array_equal(gmn, mn)
    grp.close(); whole.close()


@pytest.mark.parametrize("ip", [0, 1])
def test_group_on_all_gpus_reproduces_the_unsharded_part(ip):
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU: the NCCL group needs two devices")
    part = _parts()[ip]
    Q = reduced_history(part.sam.ndim, 130, seed=6)
    whole = StressRecovery(part, step_tile=64)
    vm = whole.recover(Q)
    mx, mn = whole.envelope()
    grp = GroupRecovery(part, devices=None, step_tile=64)
    assert grp.nblocks == n
    got = grp.recover(Q)                      # Q: host -> device 0 -> ncclBroadcast; history merged in the parent's order
    assert np.array_equal(got, vm)
    gmx, gmn = grp.envelope()                 # ncclSend / ncclRecv gather
    assert np.array_equal(gmx, mx) and np.array_equal(gmn, mn)
    grp.reset_envelope()
    grp.recover(Q[:, :50], want_history=False)
    grp.recover(Q[:, 50:], want_history=False)   # envelopes accumulate over calls
    gmx, gmn = grp.envelope()
    assert np.array_equal(gmx, mx) and np.array_equal(gmn, mn)
    t = grp.last_timing()
    assert t["tiles"] >= 1 and 0 < t["balance"] <= 1
    grp.close(); whole.close()


_WORKER = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1])
rank, world, idfile = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
import torch
from fedem_solvers_b200 import StressRecovery, Comm, split_elements
from fedem_solvers_b200.model import plate_part, reduced_history
part = plate_part(13, 11, ngen=6, seed=3, tri_fraction=0.3, shuffle_eq=True, n_fixed=3, n_constraints=4, warp=0.02)
def exchange(ident):
    if ident is not None:
        open(idfile + ".tmp", "wb").write(ident); os.replace(idfile + ".tmp", idfile)
        return ident
    for _ in range(600):
        if os.path.exists(idfile):
            return open(idfile, "rb").read()
        time.sleep(0.1)
    raise RuntimeError("no id")
torch.cuda.set_device(rank)
comm = Comm(rank, world, rank, exchange)
cuts = split_elements(part, world)
blk = StressRecovery(part, device=rank, step_tile=64, block=cuts[rank])
pt0, npts = [], []
off = np.concatenate([[0], np.cumsum(part.nstrp())])
for e0, e1 in cuts:
    pt0.append(int(off[e0])); npts.append(int(off[e1] - off[e0]))
assert pt0[rank] == blk.pt0 and npts[rank] == blk.npts
ndim, ns = part.sam.ndim, 90
Qd = torch.zeros((ns, ndim), dtype=torch.float64, device=f"cuda:{rank}")
if rank == 0:
    Qd.copy_(torch.from_numpy(np.ascontiguousarray(reduced_history(ndim, ns, seed=6).T)))
s = torch.cuda.current_stream()
blk.set_stream(s.cuda_stream)
comm.broadcast(Qd.data_ptr(), Qd.numel(), 0, s.cuda_stream)
blk.recover_dev(Qd.data_ptr(), ndim, ns, None, 0, s.cuda_stream)
env = torch.zeros((2, blk.parent_npts), dtype=torch.float64, device=f"cuda:{rank}") if rank == 0 else None
comm.gather_envelope(blk, pt0, npts, env[0].data_ptr() if rank == 0 else None, env[1].data_ptr() if rank == 0 else None, 0, s.cuda_stream)
torch.cuda.synchronize()
if rank == 0:
    whole = StressRecovery(part, device=0, step_tile=64)
    whole.recover(reduced_history(ndim, ns, seed=6), want_history=False)
    mx, mn = whole.envelope()
    e = env.cpu().numpy()
    assert np.array_equal(e[0], mx) and np.array_equal(e[1], mn)
    print("COMM OK")
comm.close()
'''


def test_one_process_per_gpu_communicator(tmp_path):
    if _ngpu() < 2:
        pytest.skip("one GPU")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "COMM OK" in outs[0]


def test_group_results_database_equals_the_single_device_file(tmp_path, monkeypatch):
    """every GPU fills the record slots of its element block; the merged file is byte-identical to the one-device file"""
    from fedem_solvers_b200.rdb import StressRdb, out_mask
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU")
    monkeypatch.setenv("FSR_RDB_TILE", "16")     # several tiles per call: exercises the double buffering
    for ip, mask in ((0, out_mask(vmStress=True, SR=True, stress=True)), (1, out_mask(vmStress=True, maxPStress=True, strain=True, SR=True))):
        part = _parts()[ip]
        ns = 70
        Q = reduced_history(part.sam.ndim, ns, seed=8)
        stepno, time = np.arange(1, ns + 1), 0.01 * np.arange(ns)
        files = []
        for rec in (StressRecovery(part, step_tile=64), GroupRecovery(part, devices=None, step_tile=64)):
            path = str(tmp_path / f"p{ip}_{len(files)}.frs")
            with StressRdb(rec, path, mask, double=(ip == 1), rdbinc=0, base_id=21, user_id=3, descr=part.name, elmid=part.elm.elmid,
                           minex=part.sam.minex) as rdb:
                rdb.write_steps(Q[:, :33], stepno[:33], time[:33])
                rdb.write_steps(Q[:, 33:], stepno[33:], time[33:])
                t = rdb.flush()
                assert t["tiles"] >= 4 and t["bytes"] == ns * rdb.step_bytes
            raw = open(path, "rb").read()
            files.append(raw[raw.find(b"\nDATA:"):])
            rec.close()
        assert files[0] == files[1]
