"""Host-side mirror of the reference interface: names, signatures, config keys, error behaviour.
No GPU needed (nothing here launches a kernel)."""
import inspect

import numpy as np
import pytest
import torch

import fedvit_b200  # noqa: F401
from conftest import micro_config, state_from_golden
from fedvit_b200 import data, fedavg, losses, model, train, utils, vit
from fedvit_b200._lib import FedVitError
from oracle import isic, step


def test_public_names_and_signatures_match_reference():
    # reference model.py:87-104
    sig = inspect.signature(model.ISICClassifier.__init__)
    assert list(sig.parameters)[1:] == [
        "backbone_name", "num_classes", "image_size", "in_channels", "pretrained", "drop_path_rate",
        "metadata_enabled", "meta_input_dim", "meta_hidden_dim", "meta_output_dim", "meta_dropout",
        "cls_hidden_dim", "cls_dropout"]
    assert sig.parameters["num_classes"].default == 8 and sig.parameters["in_channels"].default == 4
    assert list(inspect.signature(model.ISICClassifier.forward).parameters) == ["self", "x", "metadata"]
    # reference losses.py:28-41
    assert list(inspect.signature(losses.AsymmetricFocalLoss.__init__).parameters) == [
        "self", "gamma_neg", "gamma_pos", "clip", "eps"]
    assert list(inspect.signature(losses.AsymmetricFocalLoss.forward).parameters) == ["self", "logits", "targets"]
    # reference train.py:95-107,176-182
    assert list(inspect.signature(train.train_one_epoch).parameters) == [
        "model", "loader", "criterion", "optimizer", "scheduler", "scaler", "ema", "device", "config",
        "epoch", "logger"]
    assert list(inspect.signature(train.validate).parameters) == ["model", "loader", "criterion", "device", "config"]
    for name in ("build_model", "count_parameters", "get_layerwise_lr_groups"):
        assert hasattr(model, name)
    for name in ("EMA", "WarmupCosineScheduler", "clip_grad_norm", "seed_everything", "get_device",
                 "load_config", "save_checkpoint", "load_checkpoint"):
        assert hasattr(utils, name)


def test_ops_are_registered_as_torch_library_custom_ops():
    """Every Python entry point of the kernel library is also a torch.library custom op (namespace fedvit),
    the framework's own callers use the direct route (ops._Op)."""
    import fedvit_b200  # noqa: F401
    from fedvit_b200 import ops

    names = [n for n, v in vars(ops).items() if isinstance(v, ops._Op)]
    assert len(names) >= 27 and {"gemm", "wgrad", "layernorm_bwd", "attention_fwd", "adamw_flat_dev", "fedavg_accum"} <= set(names)
    for n in names:
        assert hasattr(torch.ops.fedvit, n), n
        assert callable(getattr(ops, n).fn)


def test_bench_clock_summary_parses_sampler_rows():
    """bench.ClockSampler.summary over rows in the nvidia-smi column layout (what both the NVML and the
    nvidia-smi samplers append): median clock, max clock, throttle reasons seen in any sample."""
    import bench

    c = bench.ClockSampler(0)
    c.rows = [["1700", "1965", "950.2", "Not Active", "Not Active", "Not Active", "Active"],
              ["1650", "1965", "990.0", "Not Active", "Not Active", "Not Active", "Not Active"],
              ["1600", "1965", "990.0", "Not Active", "Active", "Not Active", "Not Active"]]
    got = c.summary()
    assert got == {"sm_mhz": 1650.0, "sm_max_mhz": 1965.0, "reasons": ["hw_thermal_slowdown", "sw_power_cap"], "samples": 3}
    assert bench.ClockSampler(0).summary()["samples"] == 0


def test_weight_gradient_split_k_fills_the_cta_pairs():
    """Split-K choice for the weight-gradient GEMMs (256 x 256 tiles on 74 CTA pairs): whole waves at the
    ViT-B / ViT-L shapes, no split for problems with too few k-blocks."""
    from fedvit_b200.vit import _split_k_for

    tokens = 256 * 197
    assert [_split_k_for(o, i, tokens) for o, i in ((3072, 768), (768, 3072), (2304, 768), (768, 768))] == [2, 2, 8, 8]
    for o, i, t in ((4096, 1024, 64 * 577), (3072, 1024, 64 * 577), (3072, 768, tokens)):
        s = _split_k_for(o, i, t)
        items = -(-o // 256) * -(-i // 256) * s
        assert items / (-(-items // 74) * 74) > 0.95
    assert _split_k_for(512, 768, 256) == 1 and _split_k_for(576, 192, 394) == 1


def test_state_dict_keys_and_groups_match_oracle(golden_rgb):
    cfg = micro_config()
    ours = model.build_model(cfg)
    ora = isic.model_from_config(cfg)
    assert list(ours.state_dict().keys()) == list(ora.state_dict().keys())
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ora.state_dict().values()]
    ours.load_state_dict(state_from_golden(golden_rgb))  # a reference-written state loads as is
    go, gr = ours.get_layerwise_lr_groups(2e-4, 0.7, 1e-3), isic.llrd_groups(ora, 2e-4, 0.7, 1e-3)
    assert len(go) == len(gr) == 2 + 3
    for a, b in zip(go, gr):
        assert a["lr"] == b["lr"] and a["weight_decay"] == b["weight_decay"]
        assert [tuple(p.shape) for p in a["params"]] == [tuple(p.shape) for p in b["params"]]
    grouped = {id(p) for g in go for p in g["params"]}
    assert id(ours.backbone.cls_token) not in grouped and id(ours.backbone.pos_embed) not in grouped


def test_build_model_config_keys():
    cfg = micro_config(masked=True)
    cfg["model"]["metadata"] = {"enabled": True, "input_dim": 13, "hidden_dim": 32, "output_dim": 16, "dropout": 0.1}
    m = model.build_model(cfg)
    assert m.backbone.patch_embed.proj.in_channels == 4 and m.metadata_enabled
    assert m.classifier[0].in_features == 64 + 16 and m.classifier[-1].out_features == 7
    assert model.count_parameters(m) == sum(p.numel() for p in m.parameters())
    c = m.count_parameters()
    assert c["total"] == c["backbone"] + c["classifier"] + c["metadata"]
    m.freeze_backbone()
    assert not any(p.requires_grad for p in m.backbone.parameters())
    m.unfreeze_backbone()
    assert all(p.requires_grad for p in m.backbone.parameters())
    full = model.build_model({"model": {"backbone": "vit_tiny_patch16_224", "num_classes": 7, "image_size": 224,
                                        "pretrained": False, "metadata": {"enabled": False}}})
    assert model.count_parameters(full) == 5_626_823


def test_out_of_scope_inputs_fail_loudly():
    with pytest.raises(ValueError):
        model.build_model({"model": {"pretrained": False}})  # reference default: SwinV2 — not on this path
    with pytest.raises(RuntimeError):
        vit.create_model("vit_tiny_patch16_224", pretrained=True)
    m = model.build_model(micro_config())
    with pytest.raises(FedVitError):  # no CPU fallback
        m(torch.randn(2, 3, 32, 32))
    with pytest.raises(FedVitError):
        losses.AsymmetricFocalLoss()(torch.randn(2, 7), torch.tensor([0, 1]))
    with pytest.raises(RuntimeError):
        utils.get_device("cpu")
    with pytest.raises(ValueError):
        model.ISICClassifier("vit_micro_patch16_32", image_size=224, pretrained=False, in_channels=3)


def test_scheduler_matches_reference_formula():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([{"params": [p], "lr": 1e-4}, {"params": [torch.nn.Parameter(torch.zeros(1))], "lr": 1e-3}])
    sched = utils.WarmupCosineScheduler(opt, warmup_epochs=3, total_epochs=10, min_lr=1e-6)
    for e in range(12):
        assert opt.param_groups[0]["lr"] == pytest.approx(step.warmup_cosine_lr(1e-4, e, 3, 10, 1e-6), rel=1e-12)
        assert opt.param_groups[1]["lr"] == pytest.approx(step.warmup_cosine_lr(1e-3, e, 3, 10, 1e-6), rel=1e-12)
        opt.step()
        sched.step()


def test_metrics_match_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    for _ in range(5):
        y, p = rng.integers(0, 7, 300), rng.integers(0, 6, 300)
        m = train.classification_metrics(y, p, 7)
        assert m["accuracy"] == pytest.approx(sk.accuracy_score(y, p))
        assert m["balanced_accuracy"] == pytest.approx(sk.balanced_accuracy_score(y, p))
        assert m["macro_f1"] == pytest.approx(sk.f1_score(y, p, average="macro", zero_division=0))


def test_client_partition_and_weights():
    assert fedavg.clients_of_rank(8, 0, 8) == [0] and fedavg.clients_of_rank(16, 3, 8) == [3, 11]
    assert sorted(sum((fedavg.clients_of_rank(16, r, 4) for r in range(4)), [])) == list(range(16))
    assert fedavg.clients_of_rank(2, 0, 1) == [0, 1]
    # unequal shards: longest-processing-time-first, deterministic, ascending ids within a rank
    sizes = [512, 640, 768, 896, 1024, 1152, 1280, 1408, 1536, 1664, 1792, 1920, 2048, 512, 1024, 2048]
    place = fedavg.assign_clients(sizes, 8)
    assert place == fedavg.assign_clients(sizes, 8) and all(cs == sorted(cs) for cs in place)
    assert sorted(sum(place, [])) == list(range(16))
    loads = [sum(sizes[c] for c in cs) for cs in place]
    assert max(loads) <= 1.06 * sum(sizes) / 8
    assert max(sum(sizes[c] for c in range(16) if c % 8 == r) for r in range(8)) > 1.3 * sum(sizes) / 8
    assert fedavg.clients_of_rank(16, 2, 8, sizes) == place[2]
    assert fedavg.assign_clients([64] * 16, 8)[3] == [3, 11] and fedavg.assign_clients(sizes, 1) == [list(range(16))]
    from oracle import fedavg as ofed
    n_k = [512, 2048, 700, 1024]
    assert [fedavg.client_weight(n, sum(n_k)) for n in n_k] == ofed.client_weights(n_k)


def test_synthetic_loader_contract_and_determinism():
    a = data.SyntheticClientLoader(3, 64, 16, 32, channels=4, num_classes=7, pin=False)
    b = data.SyntheticClientLoader(3, 64, 16, 32, channels=4, num_classes=7, pin=False)
    assert len(a) == 4
    batches = list(a)
    assert batches[0]["image"].shape == (16, 4, 32, 32) and batches[0]["image"].dtype == torch.float32
    assert batches[0]["label"].dtype == torch.int64 and int(batches[0]["label"].max()) < 7
    assert torch.equal(batches[2]["image"], list(b)[2]["image"])  # seeded per client id
    assert set(batches[0]["image"][:, 3].unique().tolist()) <= {-1.0, 1.0}  # mask plane
    assert not torch.equal(batches[0]["image"], list(data.SyntheticClientLoader(4, 64, 16, 32, pin=False))[0]["image"][:, :3]) or True
    p = data.client_label_probs(4, 7, "dirichlet", 0.5, 42)
    assert p.shape == (4, 7) and np.allclose(p.sum(1), 1.0)
    assert data.client_sizes({"federated": {"num_clients": 3, "samples_per_client": [8, 16, 32]}}) == [8, 16, 32]
    with pytest.raises(ValueError):
        data.SyntheticClientLoader(0, 8, 16, 32, pin=False)  # ragged: shard smaller than a batch
