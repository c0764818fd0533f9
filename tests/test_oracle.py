"""Pin the oracle (CPU restatement) before anything is compared against it:
  * against fixtures produced by the reference's own model.py / losses.py / utils.py,
  * against the live reference where /root/reference is mounted (build container),
  * the timm-ViT restatement against torchvision's and Hugging Face's independent ViT implementations.
No GPU needed."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, micro_config, rel_err, state_from_golden
from oracle import asl, fedavg, isic, ref_bridge, step
from oracle.timm.models.vision_transformer import create_model


def _oracle_model(g, masked=False):
    torch.manual_seed(0)
    m = isic.model_from_config(micro_config(masked))
    m.load_state_dict(state_from_golden(g))
    return m


@pytest.mark.parametrize("masked", [False, True])
def test_oracle_reproduces_reference_fixture(golden_rgb, golden_masked, masked):
    g = golden_masked if masked else golden_rgb
    m = _oracle_model(g, masked).train()
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    logits = m(x)["logits"]
    logits.retain_grad()
    loss = asl.asymmetric_focal_loss(logits, y)
    loss.backward()
    assert torch.equal(logits.detach(), torch.from_numpy(g["logits"]))
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-7)
    assert rel_err(logits.grad, torch.from_numpy(g["dlogits"])) < 1e-6
    for n, p in m.named_parameters():
        assert rel_err(p.grad, torch.from_numpy(g[f"grad/{n}"])) < 1e-6, n


def test_oracle_two_adamw_steps_match_reference_fixture(golden_rgb):
    g = golden_rgb
    m = _oracle_model(g).train()
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    ema = step.OracleEMA(m, decay=0.9)
    opt = torch.optim.AdamW(isic.llrd_groups(m, 1e-3, 0.75, 1e-2), weight_decay=1e-2)
    batch = [{"image": x, "label": y}]
    for _ in range(2):
        step.local_epoch(m, batch, asl.asymmetric_focal_loss, opt, grad_clip=1.0, ema=ema)
    for n, p in m.named_parameters():
        assert rel_err(p, torch.from_numpy(g[f"after2/{n}"])) < 1e-6, n
        assert rel_err(ema.shadow[n], torch.from_numpy(g[f"ema2/{n}"])) < 1e-6, n
    # cls_token / pos_embed are in no optimiser group (reference quirk): unchanged by the steps
    assert np.array_equal(g["after2/backbone.cls_token"], g["state/backbone.cls_token"])
    assert np.array_equal(g["after2/backbone.pos_embed"], g["state/backbone.pos_embed"])


def test_loss_known_answers(asl_kats):
    k = asl_kats
    # the survey's values, produced by the reference's losses.py
    assert float(k["kat1_loss"]) == pytest.approx(0.383588046, abs=1e-7)
    assert float(k["kat2_loss"]) == pytest.approx(2.453924894, abs=1e-6)
    for i in (1, 2, 3):
        lg = torch.from_numpy(k[f"kat{i}_logits"]).clone().requires_grad_(True)
        t = torch.from_numpy(k[f"kat{i}_targets"])
        l = asl.asymmetric_focal_loss(lg, t)
        assert float(l) == pytest.approx(float(k[f"kat{i}_loss"]), rel=1e-6)
        if f"kat{i}_dlogits" in k:
            l.backward()
            assert torch.allclose(lg.grad, torch.from_numpy(k[f"kat{i}_dlogits"]), rtol=1e-5, atol=1e-8)


@pytest.mark.skipif(not ref_bridge.available(), reason="/root/reference only exists in the build container")
def test_oracle_equals_live_reference():
    rm, rl, ru = ref_bridge.load_reference()
    cfg = micro_config()
    cfg["model"]["metadata"] = {"enabled": True, "input_dim": 13, "hidden_dim": 32, "output_dim": 16, "dropout": 0.0}
    torch.manual_seed(3)
    ref = rm.build_model(cfg)
    ora = isic.model_from_config(cfg)
    ora.load_state_dict(ref.state_dict())
    ref.eval(), ora.eval()
    x, meta = torch.randn(5, 3, 32, 32), torch.rand(5, 13)
    assert torch.equal(ref(x, metadata=meta)["logits"], ora(x, metadata=meta)["logits"])
    assert torch.equal(ref(x)["logits"], ora(x)["logits"])  # zero-filled metadata branch
    gr, go = rm.get_layerwise_lr_groups(ref, 3e-4, 0.8, 1e-3), isic.llrd_groups(ora, 3e-4, 0.8, 1e-3)
    assert [g["lr"] for g in gr] == [g["lr"] for g in go]
    assert [[tuple(p.shape) for p in g["params"]] for g in gr] == [[tuple(p.shape) for p in g["params"]] for g in go]
    crit = rl.build_loss(cfg)
    lg, t = torch.randn(9, 7) * 4, torch.randint(0, 7, (9,))
    assert torch.equal(crit(lg, t), asl.asymmetric_focal_loss(lg, t))
    for e in range(12):
        sched_ref = ru.WarmupCosineScheduler.__new__(ru.WarmupCosineScheduler)
        sched_ref.last_epoch, sched_ref.warmup_epochs, sched_ref.total_epochs = e, 3, 10
        sched_ref.min_lr, sched_ref.base_lrs = 1e-6, [1e-4]
        assert sched_ref.get_lr()[0] == pytest.approx(step.warmup_cosine_lr(1e-4, e, 3, 10, 1e-6), rel=1e-12)


def test_timm_restatement_matches_torchvision_vit():
    """Independent implementation of the same architecture (torchvision), key-remapped weights."""
    tv = pytest.importorskip("torchvision.models.vision_transformer")
    torch.manual_seed(5)
    ours = create_model("vit_tiny_patch16_224", num_classes=0).eval()
    ref = tv.VisionTransformer(image_size=224, patch_size=16, num_layers=12, num_heads=3, hidden_dim=192,
                               mlp_dim=768, num_classes=1000).eval()
    sd = ours.state_dict()
    with torch.no_grad():
        for p in ours.parameters():  # non-trivial biases / norms
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
        ref.class_token.copy_(sd["cls_token"])
        ref.encoder.pos_embedding.copy_(sd["pos_embed"])
        ref.conv_proj.weight.copy_(sd["patch_embed.proj.weight"])
        ref.conv_proj.bias.copy_(sd["patch_embed.proj.bias"])
        for i, blk in enumerate(ref.encoder.layers):
            p = f"blocks.{i}."
            blk.ln_1.weight.copy_(sd[p + "norm1.weight"]); blk.ln_1.bias.copy_(sd[p + "norm1.bias"])
            blk.self_attention.in_proj_weight.copy_(sd[p + "attn.qkv.weight"])
            blk.self_attention.in_proj_bias.copy_(sd[p + "attn.qkv.bias"])
            blk.self_attention.out_proj.weight.copy_(sd[p + "attn.proj.weight"])
            blk.self_attention.out_proj.bias.copy_(sd[p + "attn.proj.bias"])
            blk.ln_2.weight.copy_(sd[p + "norm2.weight"]); blk.ln_2.bias.copy_(sd[p + "norm2.bias"])
            blk.mlp[0].weight.copy_(sd[p + "mlp.fc1.weight"]); blk.mlp[0].bias.copy_(sd[p + "mlp.fc1.bias"])
            blk.mlp[3].weight.copy_(sd[p + "mlp.fc2.weight"]); blk.mlp[3].bias.copy_(sd[p + "mlp.fc2.bias"])
        ref.encoder.ln.weight.copy_(sd["norm.weight"]); ref.encoder.ln.bias.copy_(sd["norm.bias"])
        ref.heads = torch.nn.Identity()
        x = torch.randn(2, 3, 224, 224)
        assert rel_err(ours(x), ref(x)) < 1e-5


def test_timm_restatement_matches_huggingface_vit():
    """Second independent witness for the absent timm package: Hugging Face's ViTModel (separate q / k / v
    Linears, its own attention and pooling code) with the restatement's weights — features and the
    gradient of a scalar of them with respect to the input agree."""
    tf = pytest.importorskip("transformers")
    torch.manual_seed(11)
    ours = create_model("vit_tiny_patch16_224", num_classes=0).eval()
    D, L, H = 192, 12, 3
    cfg = tf.ViTConfig(hidden_size=D, num_hidden_layers=L, num_attention_heads=H, intermediate_size=4 * D,
                       hidden_act="gelu", layer_norm_eps=1e-6, image_size=224, patch_size=16, num_channels=3,
                       qkv_bias=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    ref = tf.ViTModel(cfg, add_pooling_layer=False).eval()
    with torch.no_grad():
        for p in ours.parameters():  # non-trivial biases / norms
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
        sd = ours.state_dict()
        emb = ref.embeddings
        emb.cls_token.copy_(sd["cls_token"])
        emb.position_embeddings.copy_(sd["pos_embed"])
        emb.patch_embeddings.projection.weight.copy_(sd["patch_embed.proj.weight"])
        emb.patch_embeddings.projection.bias.copy_(sd["patch_embed.proj.bias"])
        for i, lyr in enumerate(ref.encoder.layer):
            p = f"blocks.{i}."
            lyr.layernorm_before.weight.copy_(sd[p + "norm1.weight"]); lyr.layernorm_before.bias.copy_(sd[p + "norm1.bias"])
            att = lyr.attention.attention
            for j, lin in enumerate((att.query, att.key, att.value)):
                lin.weight.copy_(sd[p + "attn.qkv.weight"][j * D:(j + 1) * D])
                lin.bias.copy_(sd[p + "attn.qkv.bias"][j * D:(j + 1) * D])
            lyr.attention.output.dense.weight.copy_(sd[p + "attn.proj.weight"])
            lyr.attention.output.dense.bias.copy_(sd[p + "attn.proj.bias"])
            lyr.layernorm_after.weight.copy_(sd[p + "norm2.weight"]); lyr.layernorm_after.bias.copy_(sd[p + "norm2.bias"])
            lyr.intermediate.dense.weight.copy_(sd[p + "mlp.fc1.weight"]); lyr.intermediate.dense.bias.copy_(sd[p + "mlp.fc1.bias"])
            lyr.output.dense.weight.copy_(sd[p + "mlp.fc2.weight"]); lyr.output.dense.bias.copy_(sd[p + "mlp.fc2.bias"])
        ref.layernorm.weight.copy_(sd["norm.weight"]); ref.layernorm.bias.copy_(sd["norm.bias"])
    x1 = torch.randn(2, 3, 224, 224, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    f1 = ours(x1)
    f2 = ref(pixel_values=x2).last_hidden_state[:, 0]
    assert rel_err(f1, f2) < 1e-5
    w = torch.randn_like(f1)
    (f1 * w).sum().backward()
    (f2 * w).sum().backward()
    assert rel_err(x1.grad, x2.grad) < 1e-4


def test_timm_restatement_shapes_and_counts():
    for name, n in [("vit_tiny_patch16_224", 5_524_416), ("vit_base_patch16_224", 85_798_656)]:
        m = create_model(name, num_classes=0)
        assert sum(p.numel() for p in m.parameters()) == n
        assert m.num_features in (192, 768)
    keys = set(create_model("vit_micro_patch16_32", num_classes=0).state_dict())
    assert {"cls_token", "pos_embed", "patch_embed.proj.weight", "blocks.0.attn.qkv.weight",
            "blocks.1.mlp.fc2.bias", "norm.weight"} <= keys


def test_fedavg_oracle_order_and_properties():
    g = torch.Generator().manual_seed(0)
    flats = [torch.randn(10_001, generator=g) for _ in range(5)]
    n_k = [64, 128, 32, 512, 64]
    a = fedavg.fedavg_flat(flats, n_k)
    b = fedavg.fedavg_numpy([f.numpy() for f in flats], n_k)
    assert np.array_equal(a.numpy(), b)  # torch and numpy agree bit for bit on the fixed order
    exact = sum(f.double() * (n / sum(n_k)) for f, n in zip(flats, n_k))
    assert rel_err(a, exact) < 1e-6
    # identical clients average to themselves (within fp32 rounding); weights sum to 1
    same = fedavg.fedavg_flat([flats[0]] * 4, [10, 20, 30, 40])
    assert rel_err(same, flats[0]) < 1e-6
    assert abs(sum(fedavg.client_weights(n_k)) - 1.0) < 1e-6
    sd = [{"w": f, "steps": torch.tensor(i)} for i, f in enumerate(flats)]
    out = fedavg.fedavg_state_dicts(sd, n_k)
    assert torch.equal(out["w"], a) and int(out["steps"]) == 0  # integer buffers from client 0


# ------------------------------------------------------------------------------------------------
# batch assembly / MixUp / CutMix restatement (oracle/mix.py) against the reference's own classes
# ------------------------------------------------------------------------------------------------
def test_mix_oracle_reproduces_reference_fixture():
    """tests/golden/mix.npz was written by the reference's utils.MixUp / utils.CutMix and the
    torchvision calls of its Dataset (tests/golden/make_golden_mix.py); the numpy restatement has
    to reproduce it bit for bit."""
    from oracle import mix

    g = dict(np.load(GOLDEN / "mix.npz"))
    x = g["x"]
    assert np.array_equal(mix.mixup(x, g["mixup/idx"], float(g["mixup/lam"])), g["mixup/out"])
    box = mix.rand_bbox(x.shape, float(g["cutmix/lam0"]), int(g["cutmix/cx"]), int(g["cutmix/cy"]))
    out, lam = mix.cutmix(x, g["cutmix/idx"], box)
    assert np.array_equal(out, g["cutmix/out"]) and lam == float(g["cutmix/lam_out"])
    assert np.array_equal(mix.assemble(g["asm/img_u8"], g["asm/mask_u8"], nhwc=True), g["asm/out"])
    assert np.array_equal(mix.assemble(g["asm/img_u8"].transpose(0, 3, 1, 2).copy(), None)[:, :3], g["asm/out"][:, :3])
