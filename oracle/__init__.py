"""CPU oracle for the CSWin-SimAM-UNet hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker or the timed CPU baseline.  Nothing under
``cswin-simam-unet_b200/`` imports it; the product path fails loudly when ``libcsb200.so`` is missing.

What is restated, and how it is pinned
--------------------------------------
* ``ops.stripe_attention`` / ``ops.lepe`` — LePEAttention.forward of the reference
  (train_cswinunet_segmentation.py, "C:", lines 220-298).  PINNED: checked against the live reference
  module (imported through ``reference_shim``) by ``tests/golden/make_golden.py``; the resulting
  vectors are committed under ``tests/golden/`` and re-checked by ``tests/test_oracle.py``.
* ``models.cswin_unet_forward`` / ``models.unet_forward`` — CSWinTransformer.forward (C:489-688) and
  UNet.forward (train_unet_segmentation.py, "U:", lines 177-250).  PINNED the same way.
* ``ops.simam`` — PARITY UNPINNED BY THE REFERENCE: the checkout contains no SimAM code at all
  (SURVEY.md §0.2).  The restatement follows the public SimAM module (Yang et al., ICML 2021,
  ``simam_module.forward``); its analytic backward is cross-checked against autograd in fp64.

The reference ships no tests, fixtures or golden vectors of its own, so every vector
here was produced by running the reference's modules in the build container.
"""
