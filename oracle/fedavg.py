"""FedAvg — the authored specification. There is no reference counterpart
(the reference has no federated code); parity for the aggregate is therefore
UNPINNED by the reference and pinned only by this restatement of McMahan et al. 2017:

    w_global = sum_k (n_k / sum_j n_j) * w_k

TEST INFRASTRUCTURE (see oracle/__init__.py).
Fixed reduction order: clients k = 0..K-1 sequentially, fp32, one rounded multiply and one rounded
add per client (no FMA): acc_0 = fl(c_0*w_0); acc_k = fl(acc_{k-1} + fl(c_k*w_k)), c_k = fl32(n_k/sum n).
Floating-point entries of the state (parameters and float buffers such as BatchNorm running
stats) are averaged; integer buffers (num_batches_tracked) are taken from client 0.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch


def client_weights(n_k: Sequence[int]) -> List[float]:
    tot = float(sum(n_k))
    return [float(np.float32(n / tot)) for n in n_k]


def fedavg_flat(flats: Sequence[torch.Tensor], n_k: Sequence[int]) -> torch.Tensor:
    """Sequential fp32 weighted sum of flat fp32 vectors, in client order."""
    cs = client_weights(n_k)
    acc = None
    for w, c in zip(flats, cs):
        term = w.to(torch.float32) * torch.tensor(c, dtype=torch.float32, device=w.device)
        acc = term if acc is None else acc + term
    return acc


def fedavg_state_dicts(states: Sequence[Dict[str, torch.Tensor]], n_k: Sequence[int]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for key, ref in states[0].items():
        if ref.is_floating_point():
            out[key] = fedavg_flat([s[key] for s in states], n_k).to(ref.dtype)
        else:
            out[key] = ref.clone()
    return out


def fedavg_numpy(flats: Sequence[np.ndarray], n_k: Sequence[int]) -> np.ndarray:
    """Same order of operations in numpy float32 (independent of torch), for cross-checking."""
    cs = client_weights(n_k)
    acc = None
    for w, c in zip(flats, cs):
        term = w.astype(np.float32) * np.float32(c)
        acc = term if acc is None else (acc + term).astype(np.float32)
    return acc
