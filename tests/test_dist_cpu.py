"""The N>1 FedAvg path on CPU: world_size-2 gloo processes run the aggregator's round protocol
(partition -> local folds -> one allreduce -> install) and must reproduce the sequential oracle.
The HBM fold kernel is replaced by the oracle's fold here (no GPU in this container); the protocol,
the weights and the collective are the product's."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


class _Arena:
    def __init__(self, n):
        self.params = torch.zeros(n)
        self.lp = None


def _cpu_fold(acc, w, weight, init):
    term = w * torch.tensor(weight, dtype=torch.float32)
    if init:
        acc.copy_(term)
    else:
        acc.add_(term)


def _worker(rank, world, port, n_clients, tmp, lpt=False):
    sys.path.insert(0, str(ROOT))
    import fedvit_b200  # noqa: F401
    from fedvit_b200 import fedavg

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 4099
        sizes = [64 * (1 + (k % 3)) for k in range(n_clients)]
        if lpt:  # unequal shards placed by longest-processing-time-first; client 0 may land on any rank
            sizes = [64 * (1 + (7 * k + 3) % 5) for k in range(n_clients)]
        place = fedavg.assign_clients(sizes, world) if lpt else [fedavg.clients_of_rank(n_clients, r, world) for r in range(world)]
        root = next(r for r, cs in enumerate(place) if 0 in cs)
        model = torch.nn.BatchNorm1d(4)  # float buffers (running stats) + an integer buffer
        arena = _Arena(n)
        g0 = torch.Generator().manual_seed(7)
        arena.params.copy_(torch.randn(n, generator=g0) if rank == 0 else torch.zeros(n))
        fedavg.broadcast_initial(arena, model)
        agg = fedavg.FedAvgAggregator(model, arena, fold=_cpu_fold)
        agg.begin_round(len(place[rank]))
        for c in place[rank]:
            agg.load_global()
            g = torch.Generator().manual_seed(100 + c)
            arena.params.add_(torch.randn(n, generator=g) * 0.1)  # "local training" of client c
            model.running_mean.fill_(float(c))
            model.num_batches_tracked.fill_(c + 5)
            agg.fold(sizes[c], sum(sizes), client_id=c, last=c == place[rank][-1])
        agg.finish(root=root)
        torch.save({"params": arena.params.clone(), "rm": model.running_mean.clone(),
                    "nbt": model.num_batches_tracked.clone(), "sizes": sizes}, f"{tmp}/r{rank}.pt")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clients,lpt", [(2, False), (5, False), (5, True)])
def test_fedavg_round_two_ranks_gloo(tmp_path, n_clients, lpt):
    port = 29500 + os.getpid() % 2000 + n_clients + (7 if lpt else 0)
    mp.spawn(_worker, args=(2, port, n_clients, str(tmp_path), lpt), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["params"], r1["params"])  # every rank holds the same w^{r+1}
    sys.path.insert(0, str(ROOT))
    from oracle import fedavg as ofed

    n = r0["params"].numel()
    base = torch.randn(n, generator=torch.Generator().manual_seed(7))
    clients = [base + torch.randn(n, generator=torch.Generator().manual_seed(100 + c)) * 0.1 for c in range(n_clients)]
    want = ofed.fedavg_flat(clients, r0["sizes"])
    rel = float((r0["params"] - want).norm() / want.norm())
    assert rel < 1e-6, rel  # north_star gate for the aggregate
    ws = ofed.client_weights(r0["sizes"])
    assert torch.allclose(r0["rm"], torch.full((4,), sum(w * c for c, w in enumerate(ws))), rtol=1e-6)
    assert int(r0["nbt"]) == 5 and int(r1["nbt"]) == 5  # integer buffers come from client 0
