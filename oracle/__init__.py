"""CPU oracle for the conv-stack hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-torch fp32 (optionally fp64) CPU restatement of what the
reference (BAMresearch/automatic-sem-image-segmentation, Release 1.2.0) executes
through Keras 3.5 on its torch backend for the MultiRes-UNet / CycleGAN conv
stacks.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or the timed CPU baseline) -- never as the product.  The product package
``automatic-sem-image-segmentation_b200`` does not import anything from here and
fails loudly when its CUDA library is missing.

Parity status
-------------
* Keras / TensorFlow are not installed in the build container and the reference
  ships no tests, so op-level parity is **unpinned by the reference**
  (SURVEY.md section 8c).
* What *is* pinned: the UNet inference graph, weight layout and BN-inference
  reading are checked against the reference's shipped trained weights
  (``ImageJ Plugin/SEM_Particle_Segmentation_Models/TiO2_UNet_Masks_*.pb``) on
  the reference's ``Datasets/`` with the IoU definition of
  ``Archive/Other Scripts/Calculate_Scores.py:69-70``; the resulting mean IoU
  lands inside the band the reference publishes (README.md:55-57).  See
  ``oracle/make_golden.py`` and ``tests/golden/pb_known_answers.json``.
* Layer algebra is additionally cross-checked against naive numpy loops in
  ``tests/test_oracle_layers.py``.
"""
