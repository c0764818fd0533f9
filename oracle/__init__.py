"""oracle/ — CPU restatement of the reference's algorithm for the hot path. TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import anything from here, and only as the checker / the timed CPU baseline —
never on the product path (``fedvit_b200`` raises when its CUDA library is missing; it has no
route into this package).

What is restated, and from where
    timm/                     the un-vendored third-party backbone the reference instantiates at
                              model.py:112-117 (``timm>=0.9``, requirements.txt:3, unpinned, NOT
                              installed here): timm's public ``VisionTransformer`` algorithm,
                              with timm's state_dict keys. Importable as top-level ``timm`` by
                              putting this directory on ``sys.path`` so the reference's own
                              model.py runs unmodified (``ref_bridge.py``).
    isic.py                   model.py:27-60,67-207,228-270,302-324 (metadata MLP, head, forward,
                              LLRD groups, build_model)
    asl.py                    losses.py:41-67 (AsymmetricFocalLoss.forward)
    step.py                   train.py:95-168 (one optimisation step of train_one_epoch) and
                              utils.py:50-105,171-193 (EMA, cosine schedule, clip)
    fedavg.py                 no reference counterpart: the FedAvg spec of
                              SURVEY.md §8.2, sequential fp32 sum in client order

Pinning status: the reference ships no golden vectors, so the pin is
    (a) tests/golden/*.npz — outputs of the reference's OWN model.py / losses.py / utils.py,
        imported from /root/reference in the build container and driven through the timm shim
        (generator: tests/golden/make_golden.py), plus the survey's known-answer loss values;
    (b) a numerical cross-check of the timm restatement against torchvision's independent
        VisionTransformer with a key-remapped state_dict (tests/test_oracle.py).
The timm backbone itself could not be executed here (package absent) — that part of parity is
pinned on (b) only, and DESIGN.md says so.
"""
