"""Round-2 parity holes (VERDICT r1 "what's weak"): the benchmarked configuration against the oracle on
the kernels it is benchmarked on, the ``torch.library`` route on a GPU, config 4's round loop
(Dirichlet labels, unequal shards, two clients per GPU), the un-synchronised graph replays the advisor
flagged, and the small kernels this round added (zero-grad fused sweep, device step tick, in-place
FedAvg install, cls-row gradient scatter)."""
import numpy as np
import pytest
import torch

import fedvit_b200  # noqa: F401
from conftest import micro_config, rel_err, state_from_golden
from fedvit_b200 import _lib, data, fedavg, graphs, losses, model, ops, optim, train, utils
from fedvit_b200.arena import FlatArena
from oracle import asl, fedavg as ofed, isic, step

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _grad_errs(ours: torch.nn.Module, want: dict):
    worst, who = 0.0, None
    for n, p in ours.named_parameters():
        w = torch.from_numpy(want[n])
        if float(w.double().norm()) < 1e-12:  # mathematically zero gradient (attention key bias)
            continue
        e = rel_err(p.grad, w)
        if e > worst:
            worst, who = e, n
    return worst, who


# ------------------------------------------------------------------------------------------------
# BASELINE configs[1] / [2] at the size where the benchmarked kernels run
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("masked", [False, True])
def test_vit_base_bf16_vs_oracle_runs_on_the_pair_gemm(masked):
    """ViT-B/16 224 (configs[1]; ``masked``: configs[2], 4-channel patch GEMM) at batch 40 — large enough
    that every forward / dgrad / wgrad GEMM of the blocks takes the CTA-pair kernel the bench line is
    measured on — logits, loss and EVERY parameter gradient against the fp32 CPU oracle (2e-2, the
    north_star bf16 gate), identical predictions; the per-family launch counters prove which kernel ran."""
    cfg = {"model": {"backbone": "vit_base_patch16_224", "num_classes": 7, "image_size": 224, "pretrained": False,
                     "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}},
           "data": {"use_segmentation_mask": masked}}
    torch.manual_seed(11)
    ora = isic.model_from_config(cfg).train()
    with torch.no_grad():
        for p in ora.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    ours = model.build_model(cfg)
    ours.load_state_dict(ora.state_dict())
    ours = ours.to(DEV).train()
    batch = 40
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(batch, 4 if masked else 3, 224, 224, generator=g)
    if masked:
        x[:, 3] = (torch.bernoulli(torch.full((batch, 224, 224), 0.3), generator=g) - 0.5) / 0.5
    y = torch.randint(0, 7, (batch,), generator=g)
    ora.zero_grad(set_to_none=True)
    want_logits = ora(x)["logits"]
    want_loss = asl.asymmetric_focal_loss(want_logits, y)
    want_loss.backward()
    want_grads = {n: p.grad.numpy() for n, p in ora.named_parameters()}

    FlatArena(ours)
    pair0, single0 = _lib.kernel_launches(_lib.KERNEL_GEMM_TC_PAIR), _lib.kernel_launches(_lib.KERNEL_GEMM_TC)
    at0 = _lib.kernel_launches(_lib.KERNEL_ATTN_FWD), _lib.kernel_launches(_lib.KERNEL_ATTN_BWD)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        logits = ours(x.to(DEV))["logits"]
        loss = losses.AsymmetricFocalLoss()(logits, y.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    pair = _lib.kernel_launches(_lib.KERNEL_GEMM_TC_PAIR) - pair0
    single = _lib.kernel_launches(_lib.KERNEL_GEMM_TC) - single0
    # 12 blocks x (4 forward + 4 dgrad + 4 wgrad) GEMMs on the pair kernel; patch embedding (+ its weight
    # gradient) and the 512-wide head stay on the single-CTA kernel
    assert pair >= 12 * 12, (pair, single)
    assert single <= 8, (pair, single)
    assert _lib.kernel_launches(_lib.KERNEL_ATTN_FWD) - at0[0] == 12
    assert _lib.kernel_launches(_lib.KERNEL_ATTN_BWD) - at0[1] == 12
    assert rel_err(logits, want_logits) < 2e-2
    assert float(loss) == pytest.approx(float(want_loss), rel=2e-2)
    worst, who = _grad_errs(ours, want_grads)
    assert worst < 2e-2, (who, worst)
    ours.eval(), ora.eval()
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
        pred = ours(x.to(DEV))["logits"].argmax(1).cpu()
    with torch.no_grad():
        ref = ora(x)["logits"]
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2e-2 * ref.abs().max()  # ties inside the bf16 tolerance may flip
    assert torch.equal(pred[clear], ref.argmax(1)[clear])


# ------------------------------------------------------------------------------------------------
# torch.library route on the GPU
# ------------------------------------------------------------------------------------------------
def _all_ops():
    return [v for v in vars(ops).values() if isinstance(v, ops._Op)]


def test_dispatcher_route_matches_direct_route(golden_rgb):
    """Every product call goes through ``torch.ops.fedvit.*`` (what FEDVIT_DISPATCH=1 selects) for one
    training step of the fixture model in both arithmetic modes: same logits bit for bit (the forward has
    no atomics), same gradients up to atomic summation order."""
    x, y = torch.from_numpy(golden_rgb["x"]).to(DEV), torch.from_numpy(golden_rgb["y"]).to(DEV)
    results = {}
    for via in (False, True):
        for op in _all_ops():
            op._via_dispatcher = via
        try:
            for amp in (False, True):
                m = model.build_model(micro_config())
                m.load_state_dict(state_from_golden(golden_rgb))
                m = m.to(DEV).train()
                arena = FlatArena(m)
                opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=arena)
                with torch.amp.autocast("cuda", enabled=amp, dtype=torch.bfloat16):
                    logits = m(x)["logits"]
                    loss = losses.AsymmetricFocalLoss()(logits, y)
                loss.backward()
                grads = arena.grads.clone()
                utils.clip_grad_norm(m.parameters(), 1.0, optimizer=opt)
                opt.step()
                results[(via, amp)] = (logits.detach().clone(), grads, arena.params.clone())
        finally:
            for op in _all_ops():
                op._via_dispatcher = False
    for amp in (False, True):
        a, b = results[(False, amp)], results[(True, amp)]
        assert torch.equal(a[0], b[0])
        assert rel_err(b[1], a[1]) < 1e-5
        assert rel_err(b[2], a[2]) < 1e-5


def test_registered_ops_pass_opcheck():
    """``torch.library.opcheck`` (schema / mutation annotations, FakeTensor shapes and dtypes, autograd
    registration) on functional and mutating ops of the namespace."""
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(37, 192, device=DEV, generator=g)
    gam, bet = torch.randn(192, device=DEV, generator=g), torch.randn(192, device=DEV, generator=g)
    torch.library.opcheck(torch.ops.fedvit.layernorm_fwd.default, (x, gam, bet, 1e-6, True), test_utils=checks)
    logits = torch.randn(9, 7, device=DEV, generator=g)
    tgt = torch.randint(0, 7, (9,), device=DEV, generator=g)
    torch.library.opcheck(torch.ops.fedvit.asl_loss.default, (logits, tgt, 4.0, 1.0, 0.05, 1e-8), test_utils=checks)
    qkv = torch.randn(2 * 197, 3 * 3 * 64, device=DEV, generator=g).bfloat16()
    torch.library.opcheck(torch.ops.fedvit.attention_fwd.default, (qkv, 2, 197, 3, 0.125), test_utils=checks)
    a = torch.randn(200, 64, device=DEV, generator=g).bfloat16()
    b = torch.randn(72, 64, device=DEV, generator=g).bfloat16()
    out = torch.empty(200, 72, device=DEV)
    torch.library.opcheck(torch.ops.fedvit.gemm.default, (a, b, None, out, None, 0, 0, 0, 1, 0), test_utils=checks)
    src = torch.randn(1024, device=DEV, generator=g)
    dst = torch.empty(1024, device=DEV, dtype=torch.bfloat16)
    torch.library.opcheck(torch.ops.fedvit.cast_bf16.default, (src, dst), test_utils=checks)
    # and the dispatcher route produces what the direct route does
    y1, m1, r1 = torch.ops.fedvit.layernorm_fwd(x, gam, bet, 1e-6, False)
    y2, m2, r2 = ops.layernorm_fwd(x, gam, bet, 1e-6, False)
    assert torch.equal(y1, y2) and torch.equal(m1, m2) and torch.equal(r1, r2)
    p = torch.randn(4096, device=DEV, generator=g)
    gr, m, v = torch.randn(4096, device=DEV, generator=g), torch.zeros(4096, device=DEV), torch.zeros(4096, device=DEV)
    seg = (torch.tensor([4096], device=DEV), torch.tensor([1e-3], device=DEV), torch.tensor([1e-2], device=DEV))
    p2, g2, m2_, v2 = p.clone(), gr.clone(), m.clone(), v.clone()
    torch.ops.fedvit.adamw_flat(p, gr, m, v, *seg, None, 0.0, 0.9, 0.999, 1e-8, 1, None, 0.0, None, True)
    ops.adamw_flat(p2, g2, m2_, v2, *seg, None, 0.0, 0.9, 0.999, 1e-8, 1, None, 0.0, None, True)
    assert torch.equal(p, p2) and float(gr.abs().sum()) == 0.0 and float(g2.abs().sum()) == 0.0


# ------------------------------------------------------------------------------------------------
# config 4's round loop: Dirichlet labels, unequal shards, two clients on one GPU
# ------------------------------------------------------------------------------------------------
def test_run_federated_dirichlet_unequal_two_clients_per_gpu_vs_oracle():
    """BASELINE configs[3] shape of a round at ViT-Tiny scale: ``partition: dirichlet`` (non-IID label
    skew), unequal ``samples_per_client`` (sample weighting), two clients sequentially on one GPU, bf16 —
    the whole round vs the CPU oracle doing the same (local epochs on the same skewed shards, then the
    sequential fp32 FedAvg)."""
    sizes = [48, 32]
    cfg = {
        "seed": 42,
        "model": {"backbone": "vit_tiny_patch16_224", "num_classes": 7, "image_size": 224, "pretrained": False,
                  "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}},
        "data": {"use_segmentation_mask": False},
        "training": {"use_amp": True, "amp_dtype": "bf16", "grad_clip": 1.0, "gradient_accumulation_steps": 1,
                     "batch_size": 16, "optimizer": {"lr": 2e-5, "weight_decay": 1e-5},
                     "llrd": {"enabled": True, "decay_rate": 0.75}},
        "augmentation": {"mixup": {"alpha": 0.0}, "cutmix": {"prob": 0.0}},
        "loss": {"asymmetric": {"gamma_neg": 4, "gamma_pos": 1, "clip": 0.05}},
        "federated": {"num_clients": 2, "rounds": 1, "local_epochs": 1, "samples_per_client": sizes,
                      "partition": "dirichlet", "dirichlet_alpha": 0.5},
    }
    out = train.run_federated(cfg, device=DEV)
    assert out["placement"] == [[0, 1]]
    ours = out["model"].eval()
    probs = data.client_label_probs(2, 7, "dirichlet", 0.5, 42)
    utils.seed_everything(42)
    init = model.build_model(cfg).state_dict()
    finals, seen_labels = [], []
    for c in range(2):
        ora = isic.model_from_config(cfg)
        ora.load_state_dict(init)
        loader = data.SyntheticClientLoader(c, sizes[c], 16, 224, num_classes=7, label_probs=probs[c], pin=False)
        seen_labels.append(torch.cat([b["label"] for b in loader]))
        oopt = torch.optim.AdamW(isic.llrd_groups(ora, 2e-5, 0.75, 1e-5), weight_decay=1e-5)
        step.local_epoch(ora, list(loader), asl.loss_from_config(cfg), oopt, grad_clip=1.0)
        finals.append({k: v.clone() for k, v in ora.state_dict().items()})
    # the shards really are skewed differently (non-IID) and weighted 48 : 32
    h0 = torch.bincount(seen_labels[0], minlength=7).float() / sizes[0]
    h1 = torch.bincount(seen_labels[1], minlength=7).float() / sizes[1]
    assert float((h0 - h1).abs().sum()) > 0.3
    want = ofed.fedavg_state_dicts(finals, sizes)
    glob = isic.model_from_config(cfg).eval()
    glob.load_state_dict(want)
    probe = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a, b = ours(probe.to(DEV))["logits"], glob(probe)["logits"]
    assert rel_err(a, b) < 2e-2
    # the global model is closer to the oracle's aggregate than to either client alone by a wide margin
    upd = max(float((want[k].double() - init[k].double()).norm()) for k in want if want[k].is_floating_point())
    for k, v in ours.state_dict().items():
        if v.is_floating_point():
            err = float((v.detach().double().cpu() - want[k].double()).norm())
            assert err < 2e-2 * max(float(want[k].double().norm()), upd), k
    assert len(out["rounds"][0]["rank_busy_ms"]) == 1


def test_lpt_placement_balances_unequal_shards():
    sizes = [512, 640, 768, 896, 1024, 1152, 1280, 1408, 1536, 1664, 1792, 1920, 2048, 512, 1024, 2048]
    place = fedavg.assign_clients(sizes, 8)
    assert sorted(sum(place, [])) == list(range(16)) and all(cs == sorted(cs) for cs in place)
    load = [sum(sizes[c] for c in cs) for cs in place]
    rr = [sum(sizes[c] for c in range(16) if c % 8 == r) for r in range(8)]
    assert max(load) <= 1.06 * sum(sizes) / 8 < max(rr)  # round-robin's tail is 1.5x the mean here


# ------------------------------------------------------------------------------------------------
# CUDA-graph replays far ahead of the device
# ------------------------------------------------------------------------------------------------
def test_forty_unsynchronised_graph_replays_match_eager(golden_rgb):
    """40 replays of the captured step enqueued without a single host sync (the host runs far ahead of
    the device) against 40 eager steps: the bias corrections come from the device-side step counter the
    captured tick kernel advances, so every replay sees its own step."""
    g = golden_rgb
    cfg = micro_config()
    crit = losses.build_loss(cfg)
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)

    def fresh():
        m = model.build_model(cfg)
        m.load_state_dict(state_from_golden(g))
        m = m.to(DEV).train()
        arena = FlatArena(m)
        o = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 2e-5, 0.75, 1e-2), weight_decay=1e-2, arena=arena)
        return m, o

    m1, o1 = fresh()
    for _ in range(40):
        o1.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda", dtype=torch.bfloat16):
            ls = crit(m1(x)["logits"], y)
        ls.backward()
        utils.clip_grad_norm(m1.parameters(), 1.0, optimizer=o1)
        o1.step()
    m2, o2 = fresh()
    start = o2.arena.params.clone()
    stepper = graphs.GraphedTrainStep(m2, crit, o2, x, y, grad_clip=1.0)
    torch.cuda.synchronize()
    for _ in range(40):
        stepper(x, y)  # no .item(), no synchronize: 40 graph launches queued back to back
    torch.cuda.synchronize()
    assert o2.step_count == 40 and int(o2._step_dev) == 40
    bc = o2._bias_corr.cpu()
    b1, b2 = float(np.float32(0.9)), float(np.float32(0.999))  # the betas cross the C ABI as fp32
    assert float(bc[0]) == pytest.approx(1 - b1 ** 40, rel=1e-6)
    assert float(bc[1]) == pytest.approx((1 - b2 ** 40) ** 0.5, rel=1e-6)
    moved = float((o1.arena.params - start).norm())
    assert float((o2.arena.params - o1.arena.params).norm()) < 0.05 * moved
    # a wrong early bias correction scales the first updates by up to 1.5x: the first moment shows it
    assert rel_err(o2.exp_avg, o1.exp_avg) < 2e-2


# ------------------------------------------------------------------------------------------------
# small kernels added this round
# ------------------------------------------------------------------------------------------------
def test_sweep_leaves_the_gradient_buffer_zeroed_and_zero_grad_is_free(golden_rgb):
    m = model.build_model(micro_config())
    m.load_state_dict(state_from_golden(golden_rgb))
    m = m.to(DEV).train()
    arena = FlatArena(m)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=arena)
    x, y = torch.from_numpy(golden_rgb["x"]).to(DEV), torch.from_numpy(golden_rgb["y"]).to(DEV)
    losses.AsymmetricFocalLoss()(m(x)["logits"], y).backward()
    assert not arena.grads_clean and float(arena.grads.abs().sum()) > 0
    # cls_token / pos_embed are in no optimiser group (lr < 0 range): their gradients are zeroed as well
    assert float(m.backbone.pos_embed.grad.abs().sum()) > 0
    utils.clip_grad_norm(m.parameters(), 1.0, optimizer=opt)
    opt.step()
    assert arena.grads_clean and float(arena.grads.abs().sum()) == 0.0
    n0 = _lib.launch_count()
    opt.zero_grad(set_to_none=True)
    assert _lib.launch_count() == n0 and m.backbone.pos_embed.grad is not None
    # a deferred clip that no step consumed dies with the gradients it was computed from
    losses.AsymmetricFocalLoss()(m(x)["logits"], y).backward()
    utils.clip_grad_norm(m.parameters(), 1e-3, optimizer=opt)
    assert opt._pending_clip is not None
    opt.zero_grad(set_to_none=True)
    assert opt._pending_clip is None and float(arena.grads.abs().sum()) == 0.0
    # opting out keeps the gradients readable after the step
    opt2 = optim.FusedAdamW(model.get_layerwise_lr_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2, arena=arena,
                            fuse_zero_grad=False)
    losses.AsymmetricFocalLoss()(m(x)["logits"], y).backward()
    opt2.step()
    assert float(arena.grads.abs().sum()) > 0


def test_cls_grad_rows_kernel():
    g = torch.Generator(device="cuda").manual_seed(3)
    B, N, D = 5, 197, 192
    dcls = torch.randn(B, D, device=DEV, generator=g)
    scale = torch.rand(B, device=DEV, generator=g) * 2
    dx = torch.full((B * N, D), 7.0, device=DEV)
    dy = torch.full((B * N, D), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.cls_grad_rows(dcls, scale, dx, dy, B, N, D)
    want = torch.zeros(B, N, D, device=DEV)
    want[:, 0] = dcls
    assert torch.equal(dx.view(B, N, D), want)
    want[:, 0] = dcls * scale[:, None]
    assert torch.equal(dy.view(B, N, D), want.bfloat16())
    dy32 = torch.empty((B * N, D), device=DEV)
    ops.cls_grad_rows(dcls, None, None, dy32, B, N, D)
    want[:, 0] = dcls
    assert torch.equal(dy32.view(B, N, D), want)


def test_fold_into_is_bit_exact_and_installs_in_place():
    g = torch.Generator(device="cuda").manual_seed(4)
    n = 64 * 1031
    ws = [torch.randn(n, device=DEV, generator=g) for _ in range(3)]
    n_k = [5, 9, 2]
    want = ofed.fedavg_flat([w.cpu() for w in ws], n_k)
    cw = [fedavg.client_weight(k, sum(n_k)) for k in n_k]
    acc = torch.empty(n, device=DEV)
    ops.fedavg_accum(acc, ws[0], cw[0], True)
    ops.fedavg_accum(acc, ws[1], cw[1], False)
    last = ws[2].clone()
    lp = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ops.fedavg_fold_into(acc, last, cw[2], last, lp)  # result lands in the client's own buffer
    assert torch.equal(last.cpu(), want) and torch.equal(lp, last.bfloat16())
    solo = ws[0].clone()
    ops.fedavg_fold_into(None, solo, cw[0], solo, None)
    assert torch.equal(solo.cpu(), ws[0].cpu() * torch.tensor(cw[0], dtype=torch.float32))


def test_aggregator_in_place_round_matches_oracle_and_skips_the_snapshot(golden_rgb):
    m = model.build_model(micro_config())
    m.load_state_dict(state_from_golden(golden_rgb))
    m = m.to(DEV)
    arena = FlatArena(m)
    agg = fedavg.FedAvgAggregator(m, arena)
    # one client on this rank: no snapshot, no accumulator, result written in place (+ bf16 shadow)
    agg.begin_round(1)
    agg.load_global()
    arena.params.add_(0.01)
    w = arena.params.detach().cpu().clone()
    agg.fold(7, 7, client_id=0, last=True)
    agg.finish()
    assert agg.global_flat is None and agg.acc is None
    assert torch.equal(arena.params.cpu(), ofed.fedavg_flat([w], [7]))
    assert torch.equal(arena.lp, arena.params.bfloat16())
    with pytest.raises(RuntimeError):
        agg.begin_round(1)
        agg.load_global()
        agg.fold(1, 2, client_id=0)
        agg.load_global()  # a second client without a snapshot
    # three clients: folds in ascending id, the last one in place
    agg.begin_round(3)
    flats = []
    for c in range(3):
        agg.load_global()
        arena.params.add_(torch.randn(arena.numel, device=DEV, generator=torch.Generator(device="cuda").manual_seed(c)) * 0.01)
        flats.append(arena.params.detach().cpu().clone())
        agg.fold([5, 9, 2][c], 16, client_id=c, last=c == 2)
    agg.finish()
    assert torch.equal(arena.params.cpu(), ofed.fedavg_flat(flats, [5, 9, 2]))
    assert torch.equal(arena.lp, arena.params.bfloat16())


def test_asl_gradient_edge_cases_match_autograd_conventions():
    """gamma_pos = 0 (the ASL paper's default) with a saturated softmax: torch's pow backward gives 0,
    not 0 * inf; labels outside [0, C) poison the loss instead of training silently."""
    logits = torch.tensor([[80.0, 0.0, -3.0], [0.5, 0.2, 0.1]], device=DEV)
    y = torch.tensor([0, 2], device=DEV)
    loss, dl = ops.asl_loss(logits, y, 4.0, 0.0, 0.05, 1e-8)
    ref_logits = logits.cpu().clone().requires_grad_(True)
    ref = asl.asymmetric_focal_loss(ref_logits, y.cpu(), gamma_neg=4.0, gamma_pos=0.0, clip=0.05, eps=1e-8)
    ref.backward()
    assert torch.isfinite(dl).all() and torch.isfinite(ref_logits.grad).all()
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    assert rel_err(dl, ref_logits.grad) < 1e-4
    bad, _ = ops.asl_loss(logits, torch.tensor([0, 3], device=DEV), 4.0, 1.0, 0.05, 1e-8)
    assert torch.isnan(bad)
    bad, _ = ops.ce_loss(logits, torch.tensor([-1, 1], device=DEV))
    assert torch.isnan(bad)


# ------------------------------------------------------------------------------------------------
# native metadata branch (scope row f2): Linear + BatchNorm1d + GELU (+ dropout) fused stages
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch", [2, 37, 256])
def test_metadata_branch_kernels_match_torch_modules(batch):
    """``MetadataBranch`` on the fused kernels against the same Sequential run by stock nn modules (the
    reference's model.py:27-60 arithmetic): embedding, every parameter gradient, running statistics and
    num_batches_tracked, in training mode (batch statistics) and eval mode (running statistics)."""
    from fedvit_b200.model import MetadataBranch

    torch.manual_seed(batch)
    ours = MetadataBranch(13, 256, 128, dropout=0.0).to(DEV)
    ref = torch.nn.Sequential(
        torch.nn.Linear(13, 256), torch.nn.BatchNorm1d(256), torch.nn.GELU(), torch.nn.Dropout(0.0),
        torch.nn.Linear(256, 128), torch.nn.BatchNorm1d(128), torch.nn.GELU()).to(DEV)
    with torch.no_grad():
        for p in ours.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    ref.load_state_dict(ours.net.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    for step_i in range(3):  # running statistics accumulate over steps
        x = torch.rand(batch, 13, device=DEV, generator=g)
        dy = torch.randn(batch, 128, device=DEV, generator=g)
        ours.train(), ref.train()
        ours.zero_grad(set_to_none=True), ref.zero_grad(set_to_none=True)
        a = ours(x)
        a.backward(dy)
        b = ref(x)
        b.backward(dy)
        assert rel_err(a, b) < 1e-5
        grads = dict(ours.net.named_parameters())
        for (n, p), (_, q) in zip(ours.net.named_parameters(), ref.named_parameters()):
            if n in ("0.bias", "4.bias"):
                # a bias in front of a training-mode BatchNorm has a mathematically zero gradient (the batch
                # mean absorbs it): both sides hold rounding noise only
                wmax = float(grads[n.replace("bias", "weight")].grad.abs().max())
                assert float(p.grad.abs().max()) < 1e-3 * wmax and float(q.grad.abs().max()) < 1e-3 * wmax, (step_i, n)
                continue
            assert rel_err(p.grad, q.grad) < 2e-4, (step_i, n)
        for (n, u), (_, v) in zip(ours.net.named_buffers(), ref.named_buffers()):
            if u.is_floating_point():
                assert rel_err(u, v) < 1e-5, (step_i, n)
            else:
                assert int(u) == int(v) == step_i + 1, n
    ours.eval(), ref.eval()
    x = torch.rand(batch, 13, device=DEV, generator=g)
    with torch.no_grad():
        assert rel_err(ours(x), ref(x)) < 1e-5
    xg = x.clone()
    ours.zero_grad(set_to_none=True), ref.zero_grad(set_to_none=True)
    ours(xg).sum().backward()   # eval-mode backward: running statistics, no batch correction terms
    ref(xg).sum().backward()
    for (n, p), (_, q) in zip(ours.net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 2e-4, n
    n0 = _lib.launch_count()
    with torch.no_grad():
        ours(x)
    assert _lib.launch_count() - n0 == 2  # the whole branch is two launches


def test_metadata_branch_dropout_and_errors():
    from fedvit_b200.model import MetadataBranch

    torch.manual_seed(0)
    m = MetadataBranch(13, 64, 32, dropout=0.5).to(DEV).train()
    x = torch.rand(64, 13, device=DEV)
    torch.manual_seed(5)
    a = m(x)
    torch.manual_seed(5)
    b = m(x)
    assert torch.equal(a, b)               # the mask comes from torch's generator: reproducible under a seed
    m.eval()
    with torch.no_grad():
        c = m(x)
    assert not torch.equal(a, c)
    m.train()
    with pytest.raises(ValueError):
        m(x[:1])                            # BatchNorm1d: more than one value per channel in training mode
    with pytest.raises(Exception):
        m(x.cpu())


def test_fedavg_round_with_metadata_branch_averages_running_stats_vs_oracle(golden_rgb):
    """``metadata.enabled: true`` (the reference default) through a whole FedAvg round on the GPU path:
    parameters AND BatchNorm running statistics are sample-weighted averages, num_batches_tracked comes
    from client 0 — against the oracle doing the same on CPU."""
    cfg = micro_config()
    cfg["model"]["metadata"] = {"enabled": True, "input_dim": 13, "hidden_dim": 32, "output_dim": 16, "dropout": 0.0}
    cfg["training"]["optimizer"] = {"lr": 2e-5, "weight_decay": 1e-2}
    sizes = [24, 12]
    cfg["federated"] = {"num_clients": 2, "rounds": 1, "local_epochs": 1, "samples_per_client": sizes}
    out = train.run_federated(cfg, device=DEV)
    ours = out["model"]
    utils.seed_everything(42)
    init = model.build_model(cfg).state_dict()
    finals = []
    for c in range(2):
        ora = isic.model_from_config(cfg)
        ora.load_state_dict(init)
        loader = data.SyntheticClientLoader(c, sizes[c], 6, 32, num_classes=7, metadata_dim=13, pin=False)
        oopt = torch.optim.AdamW(isic.llrd_groups(ora, 2e-5, 0.75, 1e-2), weight_decay=1e-2)
        step.local_epoch(ora, list(loader), asl.loss_from_config(cfg), oopt, grad_clip=1.0, use_meta=True)
        finals.append({k: v.clone() for k, v in ora.state_dict().items()})
    want = ofed.fedavg_state_dicts(finals, sizes)
    sd = ours.state_dict()
    for k in ("metadata_branch.net.1.running_mean", "metadata_branch.net.1.running_var",
              "metadata_branch.net.5.running_mean", "metadata_branch.net.5.running_var"):
        assert rel_err(sd[k], want[k]) < 1e-3, k
        assert not torch.allclose(sd[k].cpu(), init[k])  # they moved, and were averaged
    for k in ("metadata_branch.net.1.num_batches_tracked", "metadata_branch.net.5.num_batches_tracked"):
        assert int(sd[k]) == int(want[k]) == 4  # client 0's count: 24 samples / batch 6
    upd = max(float((want[k].double() - init[k].double()).norm()) for k in want if want[k].is_floating_point())
    for k, v in sd.items():
        if v.is_floating_point():
            err = float((v.detach().double().cpu() - want[k].double()).norm())
            assert err < 1e-2 * max(float(want[k].double().norm()), upd), k


# ------------------------------------------------------------------------------------------------
# second-generation attention kernels (opt-in: FEDVIT_ATTN_FWD=v2 / FEDVIT_ATTN_BWD=v2)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (3, 64, 1), (2, 65, 2), (1, 1, 1), (2, 256, 2), (4, 129, 2),
                                   (1, 16, 1), (3, 200, 1), (40, 197, 12)])
def test_second_generation_attention_kernels_vs_fp64(B, N, H, monkeypatch):
    """attn_tc_fwd2_kernel (two threads per query row) and attn_tc_bwd2_kernel (keys on lanes: P^T as the
    TMEM A operand of dV, dS^T through shared memory once) against an fp64 reference over ragged shapes, and
    against the first-generation kernels they can replace (same outputs up to bf16 rounding)."""
    import math

    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn(B * N, 3 * H * 64, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B * N, H * 64, device=DEV, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(64)
    q, k, v = (qkv.double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)[i].clone().requires_grad_(True) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * scale
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    o.backward(dout.double())
    ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3, H * 64)
    monkeypatch.setenv("FEDVIT_ATTN_FWD", "v1")
    out1, lse1 = ops.attention_fwd(qkv, B, N, H, scale)
    monkeypatch.setenv("FEDVIT_ATTN_FWD", "v2")
    out2, lse2 = ops.attention_fwd(qkv, B, N, H, scale)
    assert rel_err(out2, o.detach()) < 5e-3 and rel_err(lse2, torch.logsumexp(s.detach(), -1)) < 1e-5
    assert rel_err(out2, out1) < 1e-2 and torch.equal(ops.attention_fwd(qkv, B, N, H, scale)[0], out2)
    monkeypatch.setenv("FEDVIT_ATTN_BWD", "v2")
    d2 = ops.attention_bwd(qkv, out2, dout, lse2, B, N, H, scale).float().view(B * N, 3, H * 64)
    scale_ref = float(ref.norm()) + 1e-6 * float(dout.double().norm())  # N == 1: dq, dk are exactly 0
    for i, name in enumerate("qkv"):
        err = float((d2[:, i].double().cpu() - ref[:, i].cpu()).norm())
        assert err < 1e-2 * max(float(ref[:, i].norm()), 1e-3 * scale_ref), name
    again = ops.attention_bwd(qkv, out2, dout, lse2, B, N, H, scale).float().view(B * N, 3, H * 64)
    assert torch.equal(again, d2)  # no atomics anywhere on the path
    monkeypatch.setenv("FEDVIT_ATTN_BWD", "v1")
    d1 = ops.attention_bwd(qkv, out2, dout, lse2, B, N, H, scale).float().view(B * N, 3, H * 64)
    assert float((d2 - d1).norm()) < 1e-2 * max(float(d1.norm()), 1e-3 * scale_ref)


# ------------------------------------------------------------------------------------------------
# fourth-generation attention forward (default for N <= 208): one pass over the scores, shift from the first keys
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,H,spread,late", [(2, 197, 3, 1.0, 1.0), (3, 64, 1, 1.0, 1.0), (2, 65, 2, 1.0, 1.0),
                                                (1, 1, 1, 1.0, 1.0), (1, 16, 1, 1.0, 1.0), (2, 17, 2, 1.0, 1.0),
                                                (2, 33, 1, 1.0, 1.0), (4, 129, 2, 1.0, 1.0), (3, 200, 1, 1.0, 1.0),
                                                (2, 208, 2, 1.0, 1.0), (5, 150, 3, 1.0, 1.0), (40, 197, 12, 1.0, 1.0),
                                                (8, 197, 12, 4.0, 1.0), (4, 197, 3, 2.0, 40.0), (3, 150, 2, 3.0, 25.0),
                                                (2, 100, 2, 6.0, 8.0)])
def test_single_pass_attention_forward_vs_fp64(B, N, H, spread, late, monkeypatch):
    """attn_tc_fwd4_kernel against an fp64 softmax(QK^T)V over ragged shapes — including inputs whose later keys
    beat the first 64 by far more than 2^64 (``late``: keys from token 40 on scaled up), which send rows through
    the shift-raising slow path — and against the first-generation two-pass kernel; the backward consumes its
    LSE unchanged."""
    import math

    g = torch.Generator(device="cuda").manual_seed(1000 + N)
    qkv = torch.randn(B * N, 3 * H * 64, device=DEV, generator=g) * spread
    if late != 1.0:
        qkv.view(B, N, 3, H * 64)[:, 40:, 1] *= late
    qkv = qkv.bfloat16()
    scale = 1.0 / math.sqrt(64)
    q, k, v = (qkv.double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)[i] for i in range(3))
    s = (q @ k.transpose(-1, -2)) * scale
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    monkeypatch.setenv("FEDVIT_ATTN_FWD", "v4")
    n0 = _lib.kernel_launches(_lib.KERNEL_ATTN_FWD)
    out4, lse4 = ops.attention_fwd(qkv, B, N, H, scale)
    assert _lib.kernel_launches(_lib.KERNEL_ATTN_FWD) == n0 + 1
    assert torch.isfinite(out4.float()).all() and torch.isfinite(lse4).all()
    assert rel_err(out4, o) < 5e-3 and rel_err(lse4, torch.logsumexp(s, -1)) < 1e-5
    assert torch.equal(ops.attention_fwd(qkv, B, N, H, scale)[0], out4)  # deterministic
    monkeypatch.setenv("FEDVIT_ATTN_FWD", "v1")
    out1, lse1 = ops.attention_fwd(qkv, B, N, H, scale)
    assert rel_err(out4, out1) < 1e-2 and rel_err(lse4, lse1) < 1e-5
    # the backward recomputes P from the saved LSE: same gradients from either forward
    dout = torch.randn(B * N, H * 64, device=DEV, generator=g).bfloat16()
    d4 = ops.attention_bwd(qkv, out4, dout, lse4, B, N, H, scale).float()
    d1 = ops.attention_bwd(qkv, out1, dout, lse1, B, N, H, scale).float()
    assert float((d4 - d1).norm()) <= 1e-2 * float(d1.norm()) + 1e-6


def test_patchify_rows_leaves_a_zero_cls_slot():
    """fv_patchify_rows: the patch rows of fv_patchify with ``lead_rows`` zero rows in front of every image's
    patches (the layout the patch-embedding weight gradient contracts against the token-major stream gradient)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randn(3, 4, 32, 48, device=DEV, generator=g)
    for bf in (False, True):
        plain = ops.patchify(img, bf)
        for lead in (1, 2):
            full = ops.patchify(img, bf, lead).view(3, lead + 6, -1)
            assert torch.equal(full[:, lead:].reshape(18, -1), plain)
            assert not full[:, :lead].any()
    # the reference unfold: rows in (c, py, px) order
    ref = torch.nn.functional.unfold(img, 16, stride=16).transpose(1, 2).reshape(18, -1)
    assert torch.equal(ops.patchify(img, False), ref)
