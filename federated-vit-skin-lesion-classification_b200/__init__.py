"""fedvit_b200 — B200-native drop-in for the federated ViT client-training hot path.

Import name: ``fedvit_b200`` (the on-disk directory carries the reference repository's name and is
not a valid identifier; ``fedvit_b200.py`` at the repo root maps one onto the other).

Layout
    csrc/ + libfedvit.so   hand-written sm_100a CUDA kernels behind the C ABI in include/fedvit.h
    _lib.py, ops.py        ctypes loader and the ``torch.library`` custom ops (namespace ``fedvit``)
    timm_b200.py           ``create_model`` — the call model.py makes into timm, answered natively
    vit.py                 timm-compatible VisionTransformer on the custom ops (manual fwd/bwd)
    arena.py, optim.py     flat fp32 parameter arena, fused clip+AdamW(+EMA+bf16 shadow)
    model.py, losses.py,   the reference's module-level API for this path (same names, same
    utils.py, train.py     signatures, same config.yaml keys) + the FedAvg round loop
    fedavg.py              sample-weighted aggregate: local fold kernel + one NCCL allreduce
"""
__version__ = "0.1.0"
