"""Process-group plumbing for the N-replica benchmark (one process per GPU, independent replicas: there is
no collective on the data path — SURVEY.md §8e).  torch.distributed is used only to line the ranks up
(barrier) and to take the max / sum of their timings.  NCCL on GPU boxes, gloo in the CPU tests."""
from __future__ import annotations

import os


class Group:
    def __init__(self, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = f"cuda:{self.local_rank}"
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group("gloo")
            self.dist = dist

    def barrier(self):
        if self.dist is None:
            return
        import torch
        if self.device != "cpu":
            self.dist.barrier(device_ids=[self.local_rank])
            torch.cuda.synchronize()
        else:
            self.dist.barrier()

    def _reduce(self, x: float, op) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else float(x)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else float(x)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None


def aggregate_throughput(group: Group, units_this_rank: float, seconds_this_rank: float) -> float:
    """Whole-job value of a weak-scaling replica run: units of ALL ranks / max-over-ranks time."""
    return group.sum(units_this_rank) / group.max(seconds_this_rank)
