"""One process per GPU (torch.distributed, NCCL over NVLink / NVSwitch; gloo on CPU for tests):
element-block sharded recovery of one FE part, or of several parts of a mechanism.

The only data that crosses GPUs: the small reduced history Q (n_red x steps, broadcast from rank 0
per tile of steps -- 392 KB for config 2) and, once at the end, the per-block von Mises envelopes
(gathered to rank 0 into the parent part's result-point order).  No collective sits between K1 and
K2: every block holds the B/E rows of all nodes its elements touch."""
import numpy as np

from .partition import split_elements, sub_part

F64 = np.float64


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
