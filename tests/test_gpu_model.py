"""End-to-end parity on a real B200, through the reference-facing API (build_model / forward /
build_loss / train_one_epoch / FedAvg round), against (a) the fixtures written by the reference's
own code and (b) the CPU oracle on seeded inputs. Gates (north_star): logits and gradients within
1e-4 relative in fp32 and 2e-2 in bf16; FedAvg aggregate within 1e-6 (bit-exact on one GPU);
identical label predictions on a fixed eval batch."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

import fedvit_b200  # noqa: F401
from conftest import micro_config, rel_err, state_from_golden
from fedvit_b200 import data, fedavg, losses, model, optim, train, utils
from fedvit_b200.arena import FlatArena
from oracle import asl, fedavg as ofed, isic, step

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
# Parameters AFTER Adam steps are compared more loosely than logits / gradients: Adam divides by
# sqrt(v), so gradient entries that are mathematically zero (the key bias of every attention block:
# softmax is shift-invariant) are pure rounding noise that gets normalised to +-lr steps — any two
# correct fp32 implementations disagree there by O(lr). The 1e-4 gate applies to logits/gradients.
POST_ADAM_TOL = 2e-3


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _grad_errs(ours: torch.nn.Module, want: dict):
    """max over parameters of ||g - g_ref|| / max(||g_ref||, 1e-3 * global norm)."""
    total = float(np.sqrt(sum(float(np.sum(np.square(v.astype(np.float64)))) for v in want.values())))
    worst, who = 0.0, None
    for n, p in ours.named_parameters():
        ref = torch.from_numpy(np.asarray(want[n])).double()
        err = float((p.grad.detach().double().cpu() - ref).norm() / max(float(ref.norm()), 1e-3 * total))
        if err > worst:
            worst, who = err, n
    return worst, who


def _fixture_model(g, masked=False):
    m = model.build_model(micro_config(masked)).to(DEV)
    m.load_state_dict(state_from_golden(g))
    return m.train()


@pytest.mark.parametrize("masked", [False, True])
def test_fp32_forward_backward_matches_reference_fixture(golden_rgb, golden_masked, masked):
    g = golden_masked if masked else golden_rgb
    m = _fixture_model(g, masked)
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    logits = m(x)["logits"]
    logits.retain_grad()
    loss = losses.build_loss(micro_config())(logits, y)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 1e-4
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-4)
    assert rel_err(logits.grad, torch.from_numpy(g["dlogits"])) < 1e-4
    worst, who = _grad_errs(m, {k[5:]: v for k, v in g.items() if k.startswith("grad/")})
    assert worst < 1e-4, (who, worst)
    assert torch.equal(logits.argmax(1).cpu(), torch.from_numpy(g["logits"]).argmax(1))


def test_cls_only_last_block_is_exact(golden_rgb):
    """model.cls_only_last_block (opt-in): the last block's proj / LN2 / MLP on the cls rows only. Same
    logits and parameter gradients as the reference fixture in fp32, and as the dense schedule in bf16
    with stochastic depth on (same masks through the same seed)."""
    g = golden_rgb
    m = _fixture_model(g)
    m.backbone.cls_only_last_block = True
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    logits = m(x)["logits"]
    loss = losses.build_loss(micro_config())(logits, y)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 1e-4
    worst, who = _grad_errs(m, {k[5:]: v for k, v in g.items() if k.startswith("grad/")})
    assert worst < 1e-4, (who, worst)
    grads = {}
    for flag in (False, True):
        mm = _fixture_model(g)
        mm.backbone.cls_only_last_block = flag
        for i, blk in enumerate(mm.backbone.blocks):
            blk.drop_path_rate = 0.3 * (i + 1) / len(mm.backbone.blocks)
        torch.manual_seed(7)
        with torch.amp.autocast("cuda", dtype=torch.bfloat16):
            lg = mm(x)["logits"]
            ls = losses.build_loss(micro_config())(lg, y)
        ls.backward()
        grads[flag] = (lg.detach().float(), {n: p.grad.detach().clone() for n, p in mm.named_parameters() if p.grad is not None})
    assert rel_err(grads[True][0], grads[False][0]) < 1e-5
    for n, gd in grads[False][1].items():
        assert rel_err(grads[True][1][n], gd) < 2e-3, n  # same kernels on a row subset: only summation order moves


def test_two_fused_steps_match_reference_fixture(golden_rgb):
    """clip(1.0) + AdamW over the reference's LLRD groups + EMA, two steps — the fixture was written
    by torch.optim.AdamW / utils.clip_grad_norm / utils.EMA of the reference (make_golden.py)."""
    g = golden_rgb
    m = _fixture_model(g)
    arena = FlatArena(m)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=arena)
    ema = utils.EMA(m, decay=0.9).attach(opt)
    crit = losses.build_loss(micro_config())
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    for i in range(2):
        opt.zero_grad(set_to_none=True)
        crit(m(x)["logits"], y).backward()
        norm = utils.clip_grad_norm(m.parameters(), 1.0, optimizer=opt)
        if i == 0:
            assert float(norm) == pytest.approx(float(g["grad_norm"]), rel=1e-4)
        opt.step()
        ema.update()
    for n, p in m.named_parameters():
        assert rel_err(p, torch.from_numpy(g[f"after2/{n}"])) < POST_ADAM_TOL, n
        assert rel_err(ema.shadow[n], torch.from_numpy(g[f"ema2/{n}"])) < POST_ADAM_TOL, n
    assert torch.equal(m.backbone.cls_token.cpu(), torch.from_numpy(g["state/backbone.cls_token"]))  # never stepped
    m.eval()
    with torch.no_grad():
        ev = m(x)["logits"]
    assert rel_err(ev, torch.from_numpy(g["eval_logits_after2"])) < POST_ADAM_TOL
    assert torch.equal(ev.argmax(1).cpu(), torch.from_numpy(g["eval_logits_after2"]).argmax(1))
    # EMA swap-in / restore round trip (reference train.py:289-295)
    before = arena.params.clone()
    ema.apply_shadow()
    assert rel_err(m.backbone.norm.weight, torch.from_numpy(g["ema2/backbone.norm.weight"])) < POST_ADAM_TOL
    ema.restore()
    assert torch.equal(arena.params, before)


def _tiny_pair(seed=0, masked=False, batch=8):
    cfg = {"model": {"backbone": "vit_tiny_patch16_224", "num_classes": 7, "image_size": 224, "pretrained": False,
                     "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}},
           "data": {"use_segmentation_mask": masked}}
    torch.manual_seed(seed)
    ora = isic.model_from_config(cfg).train()
    with torch.no_grad():
        for p in ora.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    ours = model.build_model(cfg)
    ours.load_state_dict(ora.state_dict())
    ours = ours.to(DEV).train()
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(batch, 4 if masked else 3, 224, 224, generator=g)
    y = torch.randint(0, 7, (batch,), generator=g)
    return cfg, ora, ours, x, y


def _oracle_grads(ora, x, y):
    ora.zero_grad(set_to_none=True)
    logits = ora(x)["logits"]
    loss = asl.asymmetric_focal_loss(logits, y)
    loss.backward()
    return logits.detach(), float(loss), {n: p.grad.numpy() for n, p in ora.named_parameters()}


def test_vit_tiny_fp32_vs_oracle():
    """BASELINE config 1 shapes (ViT-Tiny/16 224, 7 classes), fp32 arithmetic."""
    _, ora, ours, x, y = _tiny_pair()
    want_logits, want_loss, want_grads = _oracle_grads(ora, x, y)
    logits = ours(x.to(DEV))["logits"]
    loss = losses.AsymmetricFocalLoss()(logits, y.to(DEV))
    loss.backward()
    assert rel_err(logits, want_logits) < 1e-4
    assert float(loss) == pytest.approx(want_loss, rel=1e-4)
    worst, who = _grad_errs(ours, want_grads)
    assert worst < 1e-4, (who, worst)


@pytest.mark.parametrize("masked", [False, True])
def test_vit_tiny_bf16_vs_oracle(masked):
    """bf16 tensor-core path under autocast vs the fp32 oracle: 2e-2 gate, identical predictions."""
    _, ora, ours, x, y = _tiny_pair(seed=1, masked=masked)
    want_logits, want_loss, want_grads = _oracle_grads(ora, x, y)
    FlatArena(ours)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        logits = ours(x.to(DEV))["logits"]
        loss = losses.AsymmetricFocalLoss()(logits, y.to(DEV))
    loss.backward()
    assert rel_err(logits, want_logits) < 2e-2
    assert float(loss) == pytest.approx(want_loss, rel=2e-2)
    worst, who = _grad_errs(ours, want_grads)
    assert worst < 2e-2, (who, worst)
    ours.eval(), ora.eval()
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
        pred = ours(x.to(DEV))["logits"].argmax(1).cpu()
    with torch.no_grad():
        assert torch.equal(pred, ora(x)["logits"].argmax(1))


def test_gradient_accumulation_and_frozen_backbone():
    _, ora, ours, x, y = _tiny_pair(seed=2, batch=4)
    crit = losses.AsymmetricFocalLoss()
    xd, yd = x.to(DEV), y.to(DEV)
    crit(ours(xd)["logits"], yd).backward()
    once = {n: p.grad.clone() for n, p in ours.named_parameters()}
    crit(ours(xd)["logits"], yd).backward()  # second micro-batch accumulates (train.py:151-155)
    for n, p in ours.named_parameters():
        assert rel_err(p.grad, 2 * once[n]) < 1e-5, n
    ours.zero_grad(set_to_none=True)
    ours.freeze_backbone()
    crit(ours(xd)["logits"], yd).backward()
    assert all(p.grad is None for p in ours.backbone.parameters())
    assert all(p.grad is not None for p in ours.classifier.parameters())


def test_train_one_epoch_matches_oracle_local_epoch(golden_rgb):
    cfg = micro_config()
    m = _fixture_model(golden_rgb)
    ora = isic.model_from_config(cfg)
    ora.load_state_dict(state_from_golden(golden_rgb))
    loader = data.SyntheticClientLoader(0, 24, 6, 32, channels=3, num_classes=7, pin=False)
    batches = [{k: v.clone() for k, v in b.items()} for b in loader]
    arena = FlatArena(m)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=arena)
    ema = utils.EMA(m, decay=0.9).attach(opt)
    got = train.train_one_epoch(m, loader, losses.build_loss(cfg), opt, None, None, ema, DEV, cfg, 1, None)
    oopt = torch.optim.AdamW(isic.llrd_groups(ora, 1e-3, 0.75, 1e-2), weight_decay=1e-2)
    oema = step.OracleEMA(ora, 0.9)
    want = step.local_epoch(ora, batches, asl.loss_from_config(cfg), oopt, grad_clip=1.0, ema=oema)
    assert got == pytest.approx(want, rel=POST_ADAM_TOL)  # later batches see already-updated weights
    for n, p in ora.named_parameters():
        assert rel_err(dict(m.named_parameters())[n], p) < POST_ADAM_TOL, n
        assert rel_err(ema.shadow[n], oema.shadow[n]) < POST_ADAM_TOL, n
    # gradient accumulation path: 2 micro-batches per step
    cfg2 = micro_config()
    cfg2["training"]["gradient_accumulation_steps"] = 2
    m2 = _fixture_model(golden_rgb)
    ora2 = isic.model_from_config(cfg)
    ora2.load_state_dict(state_from_golden(golden_rgb))
    opt2 = optim.FusedAdamW(model.get_layerwise_lr_groups(m2, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=FlatArena(m2))
    got2 = train.train_one_epoch(m2, loader, losses.build_loss(cfg), opt2, None, None, None, DEV, cfg2, 1, None)
    want2 = step.local_epoch(ora2, batches, asl.loss_from_config(cfg), torch.optim.AdamW(isic.llrd_groups(ora2, 1e-3, 0.75, 1e-2), weight_decay=1e-2),
                             grad_clip=1.0, accum_steps=2)
    assert got2 == pytest.approx(want2, rel=POST_ADAM_TOL)
    assert rel_err(m2.classifier[0].weight, ora2.classifier[0].weight) < POST_ADAM_TOL


def test_graphed_train_step_matches_eager_steps(golden_rgb):
    """graphs.GraphedTrainStep (zero_grad + forward + loss + backward + clip + fused AdamW + EMA as ONE
    CUDA graph, bias corrections read from device memory) against the same steps run eagerly: same
    parameters, EMA shadow and losses over 3 steps (construction rolls its warm-up steps back); a scheduler-style lr change between
    replays lands in the captured launch through the in-place lr table."""
    g = golden_rgb
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    x2 = x.flip(0).contiguous()
    from fedvit_b200 import graphs

    crit = losses.build_loss(micro_config())

    def fresh():
        m = _fixture_model(g)
        opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=FlatArena(m))
        ema = utils.EMA(m, decay=0.9).attach(opt)
        return m, opt, ema

    m1, o1, e1 = fresh()
    m2, o2, e2 = fresh()
    batches = [(x, y), (x2, y), (x, y)]
    losses_eager = []
    for i, (bx, by) in enumerate(batches):
        if i == 2:
            for grp in o1.param_groups:
                grp["lr"] *= 0.5
        o1.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda", dtype=torch.bfloat16):
            ls = crit(m1(bx)["logits"], by)
        ls.backward()
        utils.clip_grad_norm(m1.parameters(), 1.0, optimizer=o1)
        o1.step()
        e1.update()
        losses_eager.append(float(ls.detach()))
    before = o2.arena.params.clone()
    step = graphs.GraphedTrainStep(m2, crit, o2, x, y, grad_clip=1.0)  # warm-up steps are rolled back
    assert o2.step_count == 0 and torch.equal(o2.arena.params, before)
    got = []
    for i, (bx, by) in enumerate(batches):
        if i == 2:
            for grp in o2.param_groups:
                grp["lr"] *= 0.5
        got.append(float(step(bx, by)))
        e2.update()
    assert o2.step_count == o1.step_count == 3
    assert got == pytest.approx(losses_eager, rel=POST_ADAM_TOL)
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if n.endswith("attn.qkv.bias"):
            # the key-bias gradient is mathematically zero: what Adam steps on is the rounding noise of the
            # atomically summed column sums, different in any two runs — compare the q and v thirds
            d = p1.numel() // 3
            p1, p2 = torch.cat([p1[:d], p1[2 * d:]]), torch.cat([p2[:d], p2[2 * d:]])
        assert rel_err(p2, p1) < POST_ADAM_TOL, n  # same kernels; only the order of atomic adds may differ
    assert rel_err(e2.shadow["backbone.norm.weight"], e1.shadow["backbone.norm.weight"]) < POST_ADAM_TOL
    # parameters installed through PyTorch between steps (what a FedAvg round does) reach the captured forward
    with torch.no_grad():
        for mm in (m1, m2):
            for p in mm.parameters():
                p.mul_(0.5)
    o1.zero_grad(set_to_none=True)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        ls = crit(m1(x)["logits"], y)
    assert float(step(x, y)) == pytest.approx(float(ls.detach()), rel=1e-3)
    # an eager step after the replays keeps counting (graph mode ticks inside step())
    o2.zero_grad(set_to_none=True)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        crit(m2(x)["logits"], y).backward()
    o2.step()
    assert o2.step_count == 5


def test_train_one_epoch_with_cuda_graph_matches_eager_epoch(golden_rgb):
    """training.cuda_graph: the epoch loop replays the captured step; same epoch loss and parameters as the
    eager loop, twice in a row (the graph is cached on the model across epochs)."""
    loader = data.SyntheticClientLoader(0, 24, 6, 32, channels=3, num_classes=7, pin=False)
    out = {}
    for flag in (False, True):
        cfg = micro_config()
        cfg["training"]["use_amp"] = True
        cfg["training"]["cuda_graph"] = flag
        m = _fixture_model(golden_rgb)
        opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=FlatArena(m))
        crit = losses.build_loss(cfg)
        l0 = train.train_one_epoch(m, loader, crit, opt, None, None, None, DEV, cfg, 0, None)
        l1 = train.train_one_epoch(m, loader, crit, opt, None, None, None, DEV, cfg, 1, None)
        assert opt.step_count == 2 * len(loader)
        out[flag] = (l0, l1, {n: p.detach().clone() for n, p in m.named_parameters()})
    assert out[True][0] == pytest.approx(out[False][0], rel=POST_ADAM_TOL)
    assert out[True][1] == pytest.approx(out[False][1], rel=5 * POST_ADAM_TOL)
    for n, p in out[False][2].items():
        q = out[True][2][n]
        if n.endswith("attn.qkv.bias"):  # zero-gradient key bias: Adam steps on rounding noise (see above)
            d = p.numel() // 3
            p, q = torch.cat([p[:d], p[2 * d:]]), torch.cat([q[:d], q[2 * d:]])
        assert rel_err(q, p) < 5 * POST_ADAM_TOL, n


@pytest.mark.parametrize("lag", [1, 2, 5])
def test_per_step_loss_readback_matches_device_accumulation(golden_rgb, lag):
    """training.sync_loss_every_step (the reference reads loss.item() every step, train.py:164): the
    lagged pinned read-back returns the same epoch loss as the device-side accumulation, whatever the
    lag (also larger than the number of steps)."""
    loader = data.SyntheticClientLoader(0, 24, 6, 32, channels=3, num_classes=7, pin=False)
    got = []
    for sync in (False, True):
        cfg = micro_config()
        cfg["training"]["sync_loss_every_step"] = sync
        cfg["training"]["loss_read_lag"] = lag
        m = _fixture_model(golden_rgb)
        opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=FlatArena(m))
        got.append(train.train_one_epoch(m, loader, losses.build_loss(cfg), opt, None, None, None, DEV, cfg, 1, None))
    assert got[1] == pytest.approx(got[0], rel=1e-6)


def test_validate_reports_reference_metrics(golden_rgb):
    cfg = micro_config()
    m = _fixture_model(golden_rgb)
    loader = data.SyntheticClientLoader(5, 30, 6, 32, num_classes=7, pin=False)
    out = train.validate(m, loader, losses.build_loss(cfg), DEV, cfg)
    ora = isic.model_from_config(cfg).eval()
    ora.load_state_dict(state_from_golden(golden_rgb))
    ys, ps, ls = [], [], []
    with torch.no_grad():
        for b in loader:
            lg = ora(b["image"])["logits"]
            ls.append(float(asl.asymmetric_focal_loss(lg, b["label"])) * 6)
            ps.append(lg.argmax(1)), ys.append(b["label"])
    want = train.classification_metrics(torch.cat(ys).numpy(), torch.cat(ps).numpy(), 7)
    assert out["loss"] == pytest.approx(sum(ls) / 30, rel=1e-4)
    for k in ("accuracy", "balanced_accuracy", "macro_f1"):
        assert out[k] == pytest.approx(want[k])


def test_fedavg_round_single_gpu_matches_oracle(golden_rgb):
    """Config-1 shape of a round (2 clients, 1 local epoch) on the toy model: the whole round loop
    (restart from global, local epoch, weighted fold, install) vs the oracle doing the same on CPU;
    the aggregate step itself is bit-exact."""
    cfg = micro_config()
    cfg["federated"] = {"num_clients": 2, "rounds": 1, "local_epochs": 1, "samples_per_client": [24, 12]}
    # small lr: Adam turns rounding-level gradient differences into +-lr steps (see POST_ADAM_TOL),
    # which six steps at lr 1e-3 amplify chaotically; the protocol is what is under test here
    cfg["training"]["optimizer"] = {"lr": 2e-5, "weight_decay": 1e-2}
    torch.manual_seed(42)
    out = train.run_federated(cfg, device=DEV)
    ours = out["model"]
    # oracle: same initial weights (seeded build), same shards
    utils.seed_everything(42)
    init = model.build_model(cfg).state_dict()
    finals, sizes = [], [24, 12]
    for c in range(2):
        ora = isic.model_from_config(cfg)
        ora.load_state_dict(init)
        loader = data.SyntheticClientLoader(c, sizes[c], 6, 32, num_classes=7, pin=False)
        oopt = torch.optim.AdamW(isic.llrd_groups(ora, 2e-5, 0.75, 1e-2), weight_decay=1e-2)
        step.local_epoch(ora, list(loader), asl.loss_from_config(cfg), oopt, grad_clip=1.0)
        finals.append({k: v.clone() for k, v in ora.state_dict().items()})
    want = ofed.fedavg_state_dicts(finals, sizes)
    # several Adam steps from zero-initialised biases: per-tensor comparison relative to the
    # larger of the tensor and the global update, then the functional check on a probe batch
    upd = max(float((want[k].double() - init[k].double()).norm()) for k in want if want[k].is_floating_point())
    for k, v in ours.state_dict().items():
        err = float((v.detach().double().cpu() - want[k].double()).norm())
        assert err < 1e-2 * max(float(want[k].double().norm()), upd), k
    glob = isic.model_from_config(cfg).eval()
    glob.load_state_dict(want)
    probe = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    ours.eval()
    with torch.no_grad():
        assert rel_err(ours(probe.to(DEV))["logits"], glob(probe)["logits"]) < 2e-2  # Adam-noise amplified, see POST_ADAM_TOL
    assert out["rounds"][0]["images_per_s"] > 0
    # the aggregate alone: arena fold vs oracle on identical client weights -> bit exact
    arena = out["arena"]
    agg = fedavg.FedAvgAggregator(ours, arena)
    agg.begin_round()
    flats = []
    for c in range(3):
        agg.load_global()
        arena.params.add_(torch.randn(arena.numel, device=DEV, generator=torch.Generator(device="cuda").manual_seed(c)) * 0.01)
        flats.append(arena.params.detach().cpu().clone())
        agg.fold([5, 9, 2][c], 16, client_id=c)
    agg.finish()
    assert torch.equal(arena.params.cpu(), ofed.fedavg_flat(flats, [5, 9, 2]))
    got = fedavg.fedavg_state_dicts(finals, sizes, device=DEV)
    for k in want:
        assert torch.equal(got[k].cpu(), want[k]), k


def test_vit_base_full_size_properties():
    """BASELINE config 2 at full size (ViT-B/16, batch 256, bf16): finite logits / loss / grads,
    eval forward deterministic, a full fused step moves the weights, bf16 shadow tracks them."""
    cfg = {"model": {"backbone": "vit_base_patch16_224", "num_classes": 7, "image_size": 224, "pretrained": False,
                     "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}}}
    torch.manual_seed(0)
    m = model.build_model(cfg).to(DEV).train()
    assert model.count_parameters(m) == 86_195_975
    arena = FlatArena(m)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m), weight_decay=1e-5, arena=arena)
    x = torch.randn(256, 3, 224, 224, device=DEV)
    y = torch.randint(0, 7, (256,), device=DEV)
    before = arena.params.clone()
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        logits = m(x)["logits"]
        loss = losses.AsymmetricFocalLoss()(logits, y)
    loss.backward()
    assert torch.isfinite(logits).all() and torch.isfinite(loss)
    assert torch.isfinite(arena.grads).all() and float(arena.grads.abs().sum()) > 0
    norm = utils.clip_grad_norm(m.parameters(), 1.0, optimizer=opt)
    opt.step()
    assert torch.isfinite(norm) and torch.isfinite(arena.params).all()
    assert not torch.equal(arena.params, before)
    assert torch.equal(arena.lp, arena.params.bfloat16())
    assert torch.equal(m.backbone.pos_embed, arena.view(before, "backbone.pos_embed"))  # never stepped
    m.eval()
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
        a, b = m(x[:32])["logits"], m(x[:32])["logits"]
    assert torch.equal(a, b)


def test_vit_large_384_bf16_vs_oracle():
    """BASELINE config 4 architecture (ViT-Large/16 384 px: D=1024, L=24, H=16, N=577 — the ragged
    long-sequence attention path) at a small batch, bf16 vs the fp32 CPU oracle."""
    cfg = {"model": {"backbone": "vit_large_patch16_384", "num_classes": 7, "image_size": 384, "pretrained": False,
                     "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}}}
    torch.manual_seed(4)
    ora = isic.model_from_config(cfg).train()
    ours = model.build_model(cfg)
    ours.load_state_dict(ora.state_dict())
    ours = ours.to(DEV).train()
    assert model.count_parameters(ours) == 304_219_143
    g = torch.Generator().manual_seed(1004)
    x, y = torch.randn(2, 3, 384, 384, generator=g), torch.randint(0, 7, (2,), generator=g)
    want_logits, want_loss, want_grads = _oracle_grads(ora, x, y)
    FlatArena(ours)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        logits = ours(x.to(DEV))["logits"]
        loss = losses.AsymmetricFocalLoss()(logits, y.to(DEV))
    loss.backward()
    assert rel_err(logits, want_logits) < 2e-2
    assert float(loss) == pytest.approx(want_loss, rel=2e-2)
    worst, who = _grad_errs(ours, want_grads)
    assert worst < 2e-2, (who, worst)


def test_metadata_branch_and_eval_mode(golden_rgb):
    """metadata.enabled: true (the reference default) — BatchNorm MLP + concat head on top of the
    kernel backbone, train and eval mode, with and without the metadata tensor (zero-fill path)."""
    cfg = micro_config()
    cfg["model"]["metadata"] = {"enabled": True, "input_dim": 13, "hidden_dim": 32, "output_dim": 16, "dropout": 0.0}
    torch.manual_seed(9)
    ora = isic.model_from_config(cfg)
    ours = model.build_model(cfg)
    ours.load_state_dict(ora.state_dict())
    ours = ours.to(DEV)
    g = torch.Generator().manual_seed(3)
    x, meta = torch.randn(5, 3, 32, 32, generator=g), torch.rand(5, 13, generator=g)
    for mode in ("train", "eval"):
        getattr(ora, mode)(), getattr(ours, mode)()
        a = ours(x.to(DEV), metadata=meta.to(DEV))["logits"]
        b = ora(x, metadata=meta)["logits"]
        assert rel_err(a, b) < 1e-4, mode
    assert rel_err(ours(x.to(DEV))["logits"], ora(x)["logits"]) < 1e-4
    assert rel_err(ours.metadata_branch.net[1].running_mean, ora.metadata_branch.net[1].running_mean) < 1e-5


def test_graphed_eval_forward_matches_eager():
    """CUDA-graph replay of the eval forward (config 5's launch-bound small batches) is bit-identical
    to the eager launch sequence, and new inputs flow through the captured graph."""
    from fedvit_b200.graphs import GraphedForward

    _, ora, ours, x, y = _tiny_pair(seed=3, batch=4)
    FlatArena(ours)
    ours.eval()
    xd = x.to(DEV)
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
        eager = ours(xd)["logits"].clone()
    fwd = GraphedForward(ours, xd)
    assert torch.equal(fwd(xd), eager)
    x2 = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(77)).to(DEV)
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
        eager2 = ours(x2)["logits"].clone()
    assert torch.equal(fwd(x2), eager2)
    with pytest.raises(ValueError):
        fwd(x2[:2])


def test_config1_vit_tiny_fedavg_round_fp32_vs_oracle():
    """BASELINE configs[0] end to end: ViT-Tiny/16 224, 2 FedAvg clients, 1 local epoch, batch 16,
    64 samples per client, fp32 — the whole round on the GPU vs the CPU oracle doing the same
    (local epochs with torch AdamW over the LLRD groups, then the sequential fp32 FedAvg)."""
    cfg = {
        "seed": 42,
        "model": {"backbone": "vit_tiny_patch16_224", "num_classes": 7, "image_size": 224, "pretrained": False,
                  "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}},
        "data": {"use_segmentation_mask": False},
        "training": {"use_amp": False, "grad_clip": 1.0, "gradient_accumulation_steps": 1, "batch_size": 16,
                     "optimizer": {"lr": 2e-5, "weight_decay": 1e-5}, "llrd": {"enabled": True, "decay_rate": 0.75}},
        "augmentation": {"mixup": {"alpha": 0.0}, "cutmix": {"prob": 0.0}},
        "loss": {"asymmetric": {"gamma_neg": 4, "gamma_pos": 1, "clip": 0.05}},
        "federated": {"num_clients": 2, "rounds": 1, "local_epochs": 1, "samples_per_client": 64},
    }
    out = train.run_federated(cfg, device=DEV)
    ours = out["model"].eval()
    utils.seed_everything(42)
    init = model.build_model(cfg).state_dict()
    finals = []
    for c in range(2):
        ora = isic.model_from_config(cfg)
        ora.load_state_dict(init)
        loader = data.SyntheticClientLoader(c, 64, 16, 224, num_classes=7, pin=False)
        oopt = torch.optim.AdamW(isic.llrd_groups(ora, 2e-5, 0.75, 1e-5), weight_decay=1e-5)
        loss = step.local_epoch(ora, list(loader), asl.loss_from_config(cfg), oopt, grad_clip=1.0)
        finals.append({k: v.clone() for k, v in ora.state_dict().items()})
    want = ofed.fedavg_state_dicts(finals, [64, 64])
    glob = isic.model_from_config(cfg).eval()
    glob.load_state_dict(want)
    probe = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a, b = ours(probe.to(DEV))["logits"], glob(probe)["logits"]
    assert rel_err(a, b) < 2e-2  # Adam-noise amplified (POST_ADAM_TOL note); first-step parity is 1e-6
    assert torch.equal(a.argmax(1).cpu(), b.argmax(1))
    assert out["rounds"][0]["mean_client_loss"] == pytest.approx(loss, rel=0.2)


def test_federated_rounds_with_opt_in_modes_match_default():
    """run_federated (3 clients, 2 rounds, bf16, EMA, cosine schedule) with training.cuda_graph and
    model.cls_only_last_block on against the default schedule: same global model. Exercises the captured
    step across clients and rounds — optimiser reset, global weights installed through PyTorch, lr changed
    by the scheduler between rounds."""
    outs = {}
    for flag in (False, True):
        cfg = micro_config()
        cfg["model"]["cls_only_last_block"] = flag
        cfg["training"].update({"use_amp": True, "cuda_graph": flag, "batch_size": 6,
                                "scheduler": {"warmup_epochs": 1, "min_lr": 1e-5}, "ema": {"enabled": True, "decay": 0.9}})
        cfg["federated"] = {"num_clients": 3, "rounds": 2, "local_epochs": 1, "samples_per_client": 24}
        outs[flag] = train.run_federated(cfg, device=DEV)
    a, b = outs[True], outs[False]
    for r1, r2 in zip(a["rounds"], b["rounds"]):
        assert r1["mean_client_loss"] == pytest.approx(r2["mean_client_loss"], rel=1e-2)
    probe = torch.randn(6, 3, 32, 32, generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        la, lb = a["model"].eval()(probe)["logits"], b["model"].eval()(probe)["logits"]
    assert rel_err(la, lb) < 2e-2 and torch.equal(la.argmax(1), lb.argmax(1))


@pytest.mark.parametrize("amp", [False, True])
def test_stochastic_depth_matches_oracle_with_shared_masks(amp):
    """drop_path_rate > 0 (the reference's default is 0.4, config.yaml:32): per-sample keep masks,
    branch scaled by 1/keep_prob in the forward epilogue and in the backward operand. Both sides are
    fed the same masks; fp32 gate 1e-4, bf16 gate 2e-2."""
    from oracle.timm.models import vision_transformer as ovt

    cfg = micro_config()
    cfg["model"]["backbone"] = "vit_tiny_patch16_224"
    cfg["model"]["image_size"] = 224
    cfg["model"]["drop_path_rate"] = 0.4
    torch.manual_seed(11)
    ora = isic.model_from_config(cfg).train()
    ours = model.build_model(cfg)
    ours.load_state_dict(ora.state_dict())
    ours = ours.to(DEV).train()
    B = 6
    g = torch.Generator().manual_seed(21)
    x, y = torch.randn(B, 3, 224, 224, generator=g), torch.randint(0, 7, (B,), generator=g)
    rates = [b.drop_path_rate for b in ours.backbone.blocks]
    assert rates[0] == 0.0 and rates[-1] == pytest.approx(0.4)
    masks = {}
    for i, r in enumerate(rates):
        if r > 0:
            masks[i] = [torch.empty(B).bernoulli_(1 - r, generator=g) / (1 - r) for _ in range(2)]

    # oracle: replace every DropPath by a fixed-mask multiply
    for i, blk in enumerate(ora.backbone.blocks):
        if i in masks:
            for j, name in enumerate(("drop_path1", "drop_path2")):
                m = masks[i][j]

                class Fixed(torch.nn.Module):
                    def __init__(self, m):
                        super().__init__()
                        self.m = m

                    def forward(self, t):
                        return t * self.m.view(-1, 1, 1)

                setattr(blk, name, Fixed(m))
    # ours: the backbone asks for its factors block by block, branch by branch
    queue = []
    for i in range(len(rates)):
        queue += [masks[i][0], masks[i][1]] if i in masks else [None, None]
    it = iter(queue)
    ours.backbone._drop_path_scale = lambda rate, batch, device: (lambda m: None if m is None else m.to(device))(next(it))

    want_logits, want_loss, want_grads = _oracle_grads(ora, x, y)
    tol = 2e-2 if amp else 1e-4
    if amp:
        FlatArena(ours)
    with torch.amp.autocast("cuda", enabled=amp, dtype=torch.bfloat16):
        logits = ours(x.to(DEV))["logits"]
        loss = losses.AsymmetricFocalLoss()(logits, y.to(DEV))
    loss.backward()
    assert rel_err(logits, want_logits) < tol
    worst, who = _grad_errs(ours, want_grads)
    assert worst < tol, (who, worst)
    # eval mode: no masks are drawn, the forward is deterministic
    ours.eval()
    del ours.backbone._drop_path_scale
    with torch.no_grad():
        a, b = ours(x.to(DEV))["logits"], ours(x.to(DEV))["logits"]
    assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("lp", [False, True])
@pytest.mark.parametrize("batch,width,classes", [(8, 192, 7), (33, 896, 8)])
def test_native_classifier_head_matches_torch(lp, batch, width, classes):
    """head.py (scope row f2): Linear-GELU-Dropout-Linear forward and every gradient on the libfedvit
    GEMMs against the stock PyTorch modules (reference model.py:139-144,206)."""
    from fedvit_b200.head import classifier_head

    torch.manual_seed(7)
    ref = nn.Sequential(nn.Linear(width, 512), nn.GELU(), nn.Dropout(0.5), nn.Linear(512, classes)).to(DEV)
    ours = nn.Sequential(nn.Linear(width, 512), nn.GELU(), nn.Dropout(0.5), nn.Linear(512, classes)).to(DEV)
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(batch, width, device=DEV)
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    dl = torch.randn(batch, classes, device=DEV)
    ref.eval()  # dropout off on the comparator; ours gets training=False for the same effect
    ref(xr).backward(dl)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16, enabled=lp):
        out = classifier_head(xo, ours, training=False)
    out.backward(dl)
    tol = 2e-2 if lp else 1e-4
    assert rel_err(out, ref(x)) < tol
    assert rel_err(xo.grad, xr.grad) < tol
    for po, pr in zip(ours.parameters(), ref.parameters()):
        assert rel_err(po.grad, pr.grad) < tol
    # training mode: inverted dropout on the hidden layer — finite, and E[out] unchanged in scale
    with torch.amp.autocast("cuda", dtype=torch.bfloat16, enabled=lp):
        out_t = classifier_head(x.clone().requires_grad_(True), ours, training=True)
    out_t.sum().backward()
    assert torch.isfinite(out_t).all() and all(torch.isfinite(p.grad).all() for p in ours.parameters())


@pytest.mark.gpu
def test_train_one_epoch_with_mixup_cutmix_runs_reference_wiring():
    """augmentation.mixup / cutmix wired as in reference train.py:115-124,139-150: the epoch runs,
    the loss is finite and differs from the unmixed epoch on the same data and seed."""
    import numpy as np

    def epoch(mixup_alpha, cutmix_prob):
        cfg = micro_config()
        cfg["training"] = {"use_amp": True, "amp_dtype": "bf16", "grad_clip": 1.0, "gradient_accumulation_steps": 1}
        cfg["augmentation"] = {"mixup": {"alpha": mixup_alpha}, "cutmix": {"prob": cutmix_prob, "alpha": 1.0}}
        utils.seed_everything(5)
        np.random.seed(5)
        m = model.build_model(cfg).to(DEV)
        arena = FlatArena(m)
        opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=arena)
        g = torch.Generator().manual_seed(9)
        batches = [{"image": torch.randn(8, 3, 32, 32, generator=g), "label": torch.randint(0, 7, (8,), generator=g)}
                   for _ in range(3)]
        return train.train_one_epoch(m, batches, losses.build_loss(cfg), opt, None, None, None, DEV, cfg, 0, None)

    plain, mixed = epoch(0.0, 0.0), epoch(0.4, 0.5)
    assert np.isfinite(plain) and np.isfinite(mixed) and plain != mixed


@pytest.mark.gpu
def test_evaluate_and_tta_match_reference_semantics():
    """utils.evaluate / utils.evaluate_with_tta (reference utils.py:200-280): metrics agree with
    sklearn on the same predictions; the TTA logits are the mean over the views of the per-view logits."""
    from sklearn.metrics import balanced_accuracy_score, confusion_matrix, f1_score

    cfg = micro_config()
    utils.seed_everything(3)
    m = model.build_model(cfg).to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    batches = [{"image": torch.randn(8, 3, 32, 32, generator=g), "label": torch.randint(0, 7, (8,), generator=g)}
               for _ in range(3)]
    out = utils.evaluate(m, batches, DEV, use_metadata=False, use_amp=False)
    y, p = out["all_labels"], out["all_preds"]
    assert out["balanced_accuracy"] == pytest.approx(balanced_accuracy_score(y, p))
    assert out["macro_f1"] == pytest.approx(f1_score(y, p, average="macro", zero_division=0))
    assert np.array_equal(out["confusion_matrix"], confusion_matrix(y, p, labels=list(range(7))))
    with torch.no_grad():
        want_loss = sum(float(torch.nn.functional.cross_entropy(m(b["image"].to(DEV))["logits"], b["label"].to(DEV))) * 8
                        for b in batches) / 24
    assert out["loss"] == pytest.approx(want_loss, rel=1e-5)
    views = torch.randn(4, 5, 3, 32, 32, generator=g)
    preds, labels, logits = utils.evaluate_with_tta(m, [{"images": views, "label": torch.tensor([0, 1, 2, 3])}], DEV,
                                                    use_metadata=False, use_amp=False)
    with torch.no_grad():
        per_view = m(views.view(-1, 3, 32, 32).to(DEV))["logits"].view(4, 5, -1).mean(1)
    assert rel_err(torch.from_numpy(logits), per_view) < 1e-6 and preds == per_view.argmax(1).tolist() and labels == [0, 1, 2, 3]
