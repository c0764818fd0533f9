"""FedAvg: sample-weighted parameter average over the flat arena.

The reference has no federated code at all; this implements the specification
authored in SURVEY.md §8.2 / §8e (McMahan et al. 2017):

    w^{r+1} = sum_k (n_k / sum_j n_j) * w_k        over every floating-point entry of the state

Data path per round, per GPU: each local client's trained weights are folded into an fp32
accumulator with one HBM sweep (``fv_fedavg_accum``: 12 B/param, 8 for the first client), then ONE
``ncclAllReduce(sum)`` over the flat buffer (345 MB for ViT-B, 1.22 GB for ViT-L) makes every rank
hold w^{r+1}. Clients shard across GPUs round-robin (client k -> rank k mod G); there is no other
collective on the path.

Reduction order: within a rank, clients are folded in ascending id with one rounded multiply and
one rounded add each (bit-identical to oracle/fedavg.py on a single GPU); across ranks the order is
NCCL's (sum of <= 8 pre-scaled terms), within the 1e-6 gate.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .arena import FlatArena


def dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def clients_of_rank(num_clients: int, rank: int, world: int) -> List[int]:
    """Client k trains on GPU k mod G; several clients on one GPU run back to back."""
    return [k for k in range(num_clients) if k % world == rank]


def client_weight(n_k: int, n_total: int) -> float:
    """fp32(n_k / sum n) — the scalar every side multiplies by (see oracle/fedavg.py)."""
    return float(torch.tensor(n_k / float(n_total), dtype=torch.float32).item())


class FedAvgAggregator:
    """Round protocol over one flat parameter buffer. ``fold`` is the HBM sweep
    (``ops.fedavg_accum`` — the CUDA kernel); it is a parameter only so the protocol and the
    collective can be exercised by the CPU/gloo tests with a stand-in."""

    def __init__(self, model: nn.Module, arena, fold=None) -> None:
        self.model = model
        self.arena = arena
        self._fold = fold if fold is not None else ops.fedavg_accum
        self.global_flat = arena.params.clone()
        self.acc = torch.zeros_like(arena.params)
        self._folded = 0
        # floating-point buffers outside the arena (BatchNorm running stats of the metadata MLP)
        self._fbufs = [(n, b) for n, b in model.named_buffers() if b.is_floating_point()]
        self._ibufs = [(n, b) for n, b in model.named_buffers() if not b.is_floating_point()]
        self._gbuf = [b.detach().clone() for _, b in self._fbufs]
        self._accbuf = [torch.zeros_like(b, dtype=torch.float32) for _, b in self._fbufs]
        self._ibuf0: Optional[List[torch.Tensor]] = None

    # -- round protocol -----------------------------------------------------------------------
    @torch.no_grad()
    def begin_round(self) -> None:
        """Snapshot w^r: every client of this round starts from it."""
        self.global_flat.copy_(self.arena.params)
        for g, (_, b) in zip(self._gbuf, self._fbufs):
            g.copy_(b)
        self._folded = 0
        self._ibuf0 = None

    @torch.no_grad()
    def load_global(self) -> None:
        self.arena.params.copy_(self.global_flat)
        for g, (_, b) in zip(self._gbuf, self._fbufs):
            b.copy_(g)
        if self.arena.lp is not None:
            self.arena.refresh_lp(force=True)

    @torch.no_grad()
    def fold(self, n_k: int, n_total: int, client_id: int = 0) -> None:
        """acc += (n_k / n_total) * w_k for the client whose weights are in the arena now."""
        w = client_weight(n_k, n_total)
        first = self._folded == 0
        self._fold(self.acc, self.arena.params, w, first)
        for a, (_, b) in zip(self._accbuf, self._fbufs):
            t = b.to(torch.float32) * w
            if first:
                a.copy_(t)
            else:
                a.add_(t)
        if client_id == 0:
            self._ibuf0 = [b.detach().clone() for _, b in self._ibufs]
        self._folded += 1

    @torch.no_grad()
    def finish(self) -> None:
        """Cross-GPU sum (one NCCL allreduce of the flat buffer) and install w^{r+1} everywhere."""
        rank, world = dist_info()
        if self._folded == 0:
            self.acc.zero_()
            for a in self._accbuf:
                a.zero_()
        if world > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
            for a in self._accbuf:
                dist.all_reduce(a, op=dist.ReduceOp.SUM)
            # integer buffers (num_batches_tracked) come from client 0 == rank 0's first client
            for i, (_, b) in enumerate(self._ibufs):
                src = self._ibuf0[i] if (rank == 0 and self._ibuf0 is not None) else b.detach().clone()
                dist.broadcast(src, src=0)
                b.copy_(src)
        elif self._ibuf0 is not None:
            for (_, b), s in zip(self._ibufs, self._ibuf0):
                b.copy_(s)
        self.arena.params.copy_(self.acc)
        for a, (_, b) in zip(self._accbuf, self._fbufs):
            b.copy_(a.to(b.dtype))
        if self.arena.lp is not None:
            self.arena.refresh_lp(force=True)


@torch.no_grad()
def broadcast_initial(arena: FlatArena, model: nn.Module) -> None:
    """Round-0 broadcast of rank 0's initial weights."""
    _, world = dist_info()
    if world > 1:
        dist.broadcast(arena.params, src=0)
        for b in model.buffers():
            dist.broadcast(b, src=0)
    if arena.lp is not None:
        arena.refresh_lp(force=True)


def fedavg_state_dicts(states: Sequence[Dict[str, torch.Tensor]], n_k: Sequence[int],
                       device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
    """Convenience entry point for callers that hold K ``state_dict()``s (e.g. checkpoints of
    clients trained elsewhere): flattens each onto the GPU, folds with the kernel, un-flattens."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    keys = [k for k, v in states[0].items() if v.is_floating_point()]
    sizes = [states[0][k].numel() for k in keys]
    total = sum(sizes)
    pad = (total + 3) // 4 * 4
    acc = torch.zeros(pad, device=device, dtype=torch.float32)
    buf = torch.zeros(pad, device=device, dtype=torch.float32)
    n_total = sum(n_k)
    for i, (sd, n) in enumerate(zip(states, n_k)):
        off = 0
        for k, sz in zip(keys, sizes):
            buf[off:off + sz].copy_(sd[k].reshape(-1))
            off += sz
        ops.fedavg_accum(acc, buf, client_weight(n, n_total), i == 0)
    out: Dict[str, torch.Tensor] = {}
    off = 0
    for k, sz in zip(keys, sizes):
        ref = states[0][k]
        out[k] = acc[off:off + sz].view(ref.shape).to(ref.dtype).clone()
        off += sz
    for k, v in states[0].items():
        if not v.is_floating_point():
            out[k] = v.clone()
    return out
