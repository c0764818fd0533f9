"""FedAvg: sample-weighted parameter average over the flat arena.

The reference has no federated code at all; this implements the specification
authored in SURVEY.md §8.2 / §8e (McMahan et al. 2017):

    w^{r+1} = sum_k (n_k / sum_j n_j) * w_k        over every floating-point entry of the state

Data path per round, per GPU: each local client's trained weights are folded into an fp32
accumulator with one HBM sweep (``fv_fedavg_accum``: 12 B/param, 8 for the first client); the LAST
local client is folded straight into the parameter arena (``fv_fedavg_fold_into``), ONE
``ncclAllReduce(sum)`` runs on that arena in place (345 MB for ViT-B, 1.22 GB for ViT-L) and the bf16
shadow is re-cast — no accumulator -> arena copy, and with one client per GPU (BASELINE config 2) no
snapshot of the global weights and no accumulator at all. There is no other collective on the path.

Client placement: equal shards go round-robin (client k -> rank k mod G). Unequal shards
(BASELINE config 4: 16 non-IID clients of 512 ... 2048 samples on 8 GPUs) are placed by
longest-processing-time-first on n_k — a round lasts as long as its most loaded GPU — see
``assign_clients``.

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


def assign_clients(sizes: Sequence[int], world: int, policy: str = "auto") -> List[List[int]]:
    """Clients of every rank. Equal shard sizes: client k -> rank k mod G. Unequal
    sizes: longest-processing-time-first — clients in descending n_k (ties: lower id first), each to
    the rank with the least samples so far (ties: lowest rank) — because the round ends when the most
    loaded GPU does. Deterministic, identical on every rank; within a rank clients run and are
    folded in ascending id, so the single-GPU fold order (and its bit-exactness) is unchanged.
    ``policy`` (config key ``federated.placement``): "auto" (the above), "lpt", or "round_robin"."""
    k = len(sizes)
    if policy not in ("auto", "lpt", "round_robin"):
        raise ValueError(f"unknown federated.placement {policy!r} (auto | lpt | round_robin)")
    if world <= 1:
        return [list(range(k))]
    if policy == "round_robin" or (policy == "auto" and len(set(int(s) for s in sizes)) <= 1):
        return [[c for c in range(k) if c % world == r] for r in range(world)]
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for c in sorted(range(k), key=lambda c: (-int(sizes[c]), c)):
        r = min(range(world), key=lambda r: (load[r], r))
        out[r].append(c)
        load[r] += int(sizes[c])
    return [sorted(cs) for cs in out]


def clients_of_rank(num_clients: int, rank: int, world: int, sizes: Optional[Sequence[int]] = None) -> List[int]:
    """Clients that train on GPU ``rank`` (several on one GPU run back to back): k mod G, or the
    load-balanced placement of ``assign_clients`` when unequal shard sizes are given."""
    if sizes is not None:
        if len(sizes) != num_clients:
            raise ValueError("sizes must have num_clients entries")
        return assign_clients(sizes, world)[rank]
    return [k for k in range(num_clients) if k % world == rank]


def client_weight(n_k: int, n_total: int) -> float:
    """fp32(n_k / sum n) — the scalar every side multiplies by (see oracle/fedavg.py)."""
    return float(torch.tensor(n_k / float(n_total), dtype=torch.float32).item())


class FedAvgAggregator:
    """Round protocol over one flat parameter buffer. ``fold`` is the HBM sweep
    (``ops.fedavg_accum`` — the CUDA kernel); it is a parameter only so the protocol and the
    collective can be exercised by the CPU/gloo tests with a stand-in.

        agg.begin_round(n_local)               # n_local = clients this rank trains this round
        for c in my_clients:
            agg.load_global()                  # restart from w^r (free for the first client)
            ... local epochs ...
            agg.fold(n_c, n_total, c, last=c == my_clients[-1])
        agg.finish(root=rank_of_client_0)      # one allreduce, in place on the parameter arena
    """

    def __init__(self, model: nn.Module, arena, fold=None) -> None:
        self.model = model
        self.arena = arena
        self._custom_fold = fold is not None
        self._fold = fold if fold is not None else ops.fedavg_accum
        # allocated on first need: with one client per GPU neither exists (2 x 345 MB at ViT-B)
        self.global_flat: Optional[torch.Tensor] = None
        self.acc: Optional[torch.Tensor] = None
        self._folded = 0
        self._loaded = 0
        self._have_snapshot = False
        self._in_params = False   # the weighted sum of this rank's clients already sits in arena.params
        self._lp_written = False  # ... and its bf16 copy in arena.lp
        # floating-point buffers outside the arena (BatchNorm running stats of the metadata MLP)
        self._fbufs = [(n, b) for n, b in model.named_buffers() if b.is_floating_point()]
        self._ibufs = [(n, b) for n, b in model.named_buffers() if not b.is_floating_point()]
        self._gbuf = [b.detach().clone() for _, b in self._fbufs]
        self._accbuf = [torch.zeros_like(b, dtype=torch.float32) for _, b in self._fbufs]
        self._ibuf0: Optional[List[torch.Tensor]] = None

    # -- round protocol -----------------------------------------------------------------------
    @torch.no_grad()
    def begin_round(self, n_local: Optional[int] = None) -> None:
        """Start of round r; the arena holds w^r. A snapshot of it is only taken when a second local
        client will have to restart from it (``n_local`` unknown or > 1)."""
        self._have_snapshot = n_local is None or n_local > 1
        if self._have_snapshot:
            if self.global_flat is None:
                self.global_flat = torch.empty_like(self.arena.params)
            self.global_flat.copy_(self.arena.params)
        for g, (_, b) in zip(self._gbuf, self._fbufs):
            g.copy_(b)
        self._folded = 0
        self._loaded = 0
        self._in_params = False
        self._lp_written = False
        self._ibuf0 = None

    @torch.no_grad()
    def load_global(self) -> None:
        """Reset the arena to w^r for the next local client. The first client of a round finds w^r
        already there (nothing has trained since ``begin_round``): no copy, no re-cast."""
        first = self._loaded == 0 and self._folded == 0
        self._loaded += 1
        for g, (_, b) in zip(self._gbuf, self._fbufs):
            b.copy_(g)
        if first:
            return
        if not self._have_snapshot:
            raise RuntimeError("FedAvgAggregator: begin_round(n_local=1) took no snapshot of the global "
                               "weights, but a second client asked for them")
        self.arena.params.copy_(self.global_flat)
        if self.arena.lp is not None:
            self.arena.refresh_lp(force=True)

    @torch.no_grad()
    def fold(self, n_k: int, n_total: int, client_id: int = 0, last: bool = False) -> None:
        """acc += (n_k / n_total) * w_k for the client whose weights are in the arena now. ``last``
        (this rank's final client of the round): the sum is written into the parameter arena itself,
        where ``finish`` reduces it in place."""
        w = client_weight(n_k, n_total)
        first = self._folded == 0
        if self._in_params:
            raise RuntimeError("FedAvgAggregator.fold: called after the round's last fold")
        p = self.arena.params
        if last and not self._custom_fold:
            world = dist_info()[1]
            lp = self.arena.lp if world == 1 else None  # across ranks the bf16 copy follows the allreduce
            ops.fedavg_fold_into(None if first else self.acc, p, w, p, lp)
            self._in_params = True
            self._lp_written = lp is not None
        else:
            if self.acc is None:
                self.acc = torch.empty_like(p)
            self._fold(self.acc, p, w, first)
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
    def finish(self, root: int = 0) -> None:
        """Cross-GPU sum (one NCCL allreduce, in place on the parameter arena) and install of w^{r+1}
        everywhere. ``root`` = the rank that trained client 0 (integer buffers come from it)."""
        rank, world = dist_info()
        p = self.arena.params
        if self._folded == 0:  # a rank without clients contributes zeros
            p.zero_()
            for a in self._accbuf:
                a.zero_()
        elif not self._in_params:
            p.copy_(self.acc)
        if world > 1:
            dist.all_reduce(p, op=dist.ReduceOp.SUM)
            for a in self._accbuf:
                dist.all_reduce(a, op=dist.ReduceOp.SUM)
            # integer buffers (num_batches_tracked) come from client 0
            for i, (_, b) in enumerate(self._ibufs):
                src = self._ibuf0[i] if (rank == root and self._ibuf0 is not None) else b.detach().clone()
                dist.broadcast(src, src=root)
                b.copy_(src)
        elif self._ibuf0 is not None:
            for (_, b), s in zip(self._ibufs, self._ibuf0):
                b.copy_(s)
        for a, (_, b) in zip(self._accbuf, self._fbufs):
            b.copy_(a.to(b.dtype))
        if self.arena.lp is not None:
            if self._lp_written and world == 1:
                self.arena.mark_lp_fresh()
            else:
                self.arena.refresh_lp(force=True)
        self._in_params = False


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
