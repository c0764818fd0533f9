"""Fused optimiser step over the flat arena: global-norm clip + AdamW (per-range lr / weight decay
for the LLRD groups) + optional EMA + bf16 weight re-cast in ONE kernel sweep.

Replaces the reference's four separate passes per step (train.py:156-162): ``scaler.unscale_`` /
``clip_grad_norm_`` (utils.py:192-193), ``torch.optim.AdamW.step`` over L+3 param groups
(train.py:253-261) and the Python loop of ``EMA.update`` (utils.py:76-83).

``FusedAdamW`` is a ``torch.optim.Optimizer``: ``param_groups`` (and therefore the reference's
``WarmupCosineScheduler``, which rewrites ``group["lr"]`` each epoch), ``state_dict`` /
``load_state_dict`` and ``zero_grad`` behave as usual. Update rule and operation order follow
``torch.optim.AdamW`` (betas (0.9, 0.999), eps 1e-8, decoupled decay, bias correction).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .arena import FlatArena


def grad_sumsq(arena: FlatArena, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum of squares of every gradient in the arena (device scalar)."""
    if out is None:
        out = torch.zeros(1, device=arena.device, dtype=torch.float32)
    ops.sumsq(arena.grads, out, False)
    return out


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, arena: Optional[FlatArena] = None,
                 fuse_zero_grad: bool = True) -> None:
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if arena is None:
            arenas = {id(p._fv_arena[0]): p._fv_arena[0] for g in self.param_groups for p in g["params"]
                      if getattr(p, "_fv_arena", None) is not None}
            if len(arenas) != 1:
                raise RuntimeError("FusedAdamW needs its parameters in one FlatArena — build "
                                   "FlatArena(model) before the optimiser")
            arena = next(iter(arenas.values()))
        for g in self.param_groups:
            for p in g["params"]:
                if not arena.owns(p):
                    raise RuntimeError("parameter outside the arena (was the model moved after FlatArena?)")
            if tuple(g["betas"]) != tuple(betas) or g["eps"] != eps:
                raise ValueError("per-group betas/eps are not supported (the reference uses one setting)")
        self.arena = arena
        self.exp_avg = torch.zeros_like(arena.params)
        self.exp_avg_sq = torch.zeros_like(arena.params)
        self.step_count = 0
        self.ema = None           # set by utils.EMA.attach(optimizer): shadow updated in the same sweep
        self._pending_clip = None  # (sumsq device scalar, max_norm) from utils.clip_grad_norm
        self._seg_key = None
        self._seg = None
        # CUDA-graph mode (graphs.GraphedTrainStep): the bias corrections of the current step live in a
        # device tensor that graph_tick() refreshes before every replay
        self._bias_corr = None      # device fp32 [2]: bias corrections of the current step
        self._step_dev = None       # device int64 [1]: the step counter the captured tick kernel advances
        self._tick_pending = False   # graph_tick() ran and no step has consumed it yet
        # the sweep is the gradient buffer's last reader: it writes zeros back, and the zero_grad() that
        # follows every step in the reference loop (train.py:160) finds the buffer clean — no memset pass
        self.fuse_zero_grad = bool(fuse_zero_grad)

    # -- CUDA-graph support ---------------------------------------------------------------------------
    def enable_graph_mode(self) -> None:
        """From now on ``step()`` launches the sweep variant that reads the step-dependent scalars from
        device memory. ``graph_tick()`` advances the step (count, bias corrections, lr table): a graph
        replay needs it called first (GraphedTrainStep does); an eager ``step()`` calls it itself unless
        the caller already has. The per-group lr / weight-decay table is updated in place when a scheduler
        changes it, so its device addresses stay valid inside a captured graph."""
        if self._bias_corr is None:
            self._bias_corr = torch.ones(2, device=self.arena.device, dtype=torch.float32)
            self._step_dev = torch.full((1,), self.step_count, device=self.arena.device, dtype=torch.int64)
            self._seg_key, self._seg = None, None  # rebuilt per param group: fixed length from here on
            self._segments()

    def graph_tick(self) -> None:
        """Host-side bookkeeping of the next optimisation step: the step count mirrored for checkpoints and
        the lr table if a scheduler moved it. The step-dependent numbers the sweep needs never leave the
        device: a one-thread kernel captured in the graph right before the sweep (``fv_adamw_tick``)
        increments a device counter and derives the bias corrections from it, so a host that runs any
        number of replays ahead of the GPU cannot hand a replay another step's corrections (the pinned
        ring this replaces was overwritten after 8 un-synchronised steps)."""
        if self._bias_corr is None:
            raise RuntimeError("graph_tick() needs enable_graph_mode()")
        self.step_count += 1
        self._segments()
        self._tick_pending = True

    def _set_device_step(self, step: int) -> None:
        if self._step_dev is not None:
            self._step_dev.fill_(int(step))

    # -- hooks used by utils.clip_grad_norm / utils.EMA -------------------------------------------
    def defer_clip(self, sumsq: torch.Tensor, max_norm: float) -> None:
        """Fold the clip coefficient into the next step's gradient read instead of rescaling the
        gradient buffer in a pass of its own."""
        self._pending_clip = (sumsq, float(max_norm))

    def _segments(self):
        key = tuple((float(g["lr"]), float(g["weight_decay"])) for g in self.param_groups)
        if key != self._seg_key:
            ends, lrs, wds = self.arena.segments(self.param_groups, by_group=self._bias_corr is not None)
            dev = self.arena.device
            new = (torch.tensor(ends, device=dev, dtype=torch.int64),
                   torch.tensor(lrs, device=dev, dtype=torch.float32),
                   torch.tensor(wds, device=dev, dtype=torch.float32))
            if self._bias_corr is not None and self._seg is not None:
                if self._seg[0].numel() != new[0].numel():
                    raise RuntimeError("FusedAdamW (graph mode): the param-group layout changed after "
                                       "enable_graph_mode() — build a new GraphedTrainStep")
                for old_t, new_t in zip(self._seg, new):  # same device addresses for the captured launch
                    old_t.copy_(new_t)
            else:
                self._seg = new
            self._seg_key = key
        return self._seg

    def _gather_foreign_grads(self) -> None:
        """A gradient that autograd allocated outside the arena (p.grad was None at backward time)
        is copied into its slot so the sweep sees it."""
        a = self.arena
        for g in self.param_groups:
            for p in g["params"]:
                slot = a.grad_view(p)
                if p.grad is None:
                    slot.zero_()
                elif p.grad.data_ptr() != slot.data_ptr():
                    slot.copy_(p.grad)
                    p.grad = slot

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._gather_foreign_grads()
        seg_end, seg_lr, seg_wd = self._segments()
        b1, b2 = self.defaults["betas"]
        sumsq, max_norm = self._pending_clip if self._pending_clip is not None else (None, 0.0)
        self._pending_clip = None
        ema = self.ema
        a = self.arena
        if self._bias_corr is not None:  # graph mode: graph_tick() advances the step
            if not torch.cuda.is_current_stream_capturing():
                if not self._tick_pending:
                    self.graph_tick()
                    seg_end, seg_lr, seg_wd = self._seg
                self._tick_pending = False
            ops.adamw_tick(self._step_dev, self._bias_corr, b1, b2)
            ops.adamw_flat_dev(a.params, a.grads, self.exp_avg, self.exp_avg_sq, seg_end, seg_lr, seg_wd,
                               sumsq, max_norm, b1, b2, self.defaults["eps"], self._bias_corr,
                               ema.flat if ema is not None else None, ema.decay if ema is not None else 0.0, a.lp,
                               self.fuse_zero_grad)
        else:
            self.step_count += 1
            ops.adamw_flat(a.params, a.grads, self.exp_avg, self.exp_avg_sq, seg_end, seg_lr, seg_wd,
                           sumsq, max_norm, b1, b2, self.defaults["eps"], self.step_count,
                           ema.flat if ema is not None else None, ema.decay if ema is not None else 0.0, a.lp,
                           self.fuse_zero_grad)
        if self.fuse_zero_grad:
            a.grads_clean = True
        if a.lp is not None:
            a.mark_lp_fresh()
        if ema is not None:
            ema._fused_updates += 1
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        """One memset of the flat gradient buffer. Gradients stay attached as arena views even
        for ``set_to_none=True`` (the reference passes it, train.py:160): the kernels accumulate
        into those views, which is what makes the single-sweep step possible. Right after a ``step()`` the
        buffer is already zero (the sweep wrote the zeros behind its last read) and nothing is launched.
        A clip coefficient deferred by ``utils.clip_grad_norm`` that no ``step()`` consumed is dropped
        here: it belongs to the gradients being discarded."""
        self._pending_clip = None
        self.arena.zero_grads()

    # -- checkpoint interchange: moments are exposed per parameter like torch.optim.AdamW ----------
    def state_dict(self):
        sd = super().state_dict()
        a = self.arena
        idx = 0
        state = {}
        for g in self.param_groups:
            for p in g["params"]:
                n = p._fv_arena[1]
                state[idx] = {"step": torch.tensor(float(self.step_count)),
                              "exp_avg": a.view(self.exp_avg, n).clone(),
                              "exp_avg_sq": a.view(self.exp_avg_sq, n).clone()}
                idx += 1
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        shell = dict(state_dict)
        shell["state"] = {}
        super().load_state_dict(shell)
        a = self.arena
        idx = 0
        for g in self.param_groups:
            for p in g["params"]:
                s = state.get(idx, state.get(str(idx)))
                if s is not None:
                    n = p._fv_arena[1]
                    a.view(self.exp_avg, n).copy_(s["exp_avg"])
                    a.view(self.exp_avg_sq, n).copy_(s["exp_avg_sq"])
                    self.step_count = int(float(s["step"]))
                idx += 1
        self._seg_key = None
        self._set_device_step(self.step_count)

    def reset_state(self) -> None:
        """Fresh moments and step count — the per-round client reset of canonical FedAvg."""
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0
        self._set_device_step(0)
