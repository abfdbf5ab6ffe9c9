"""CPU oracle for the CLAS-FV full-video inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline), never as the thing shipped.

Every function restates one piece of the reference in plain PyTorch-CPU /
NumPy and cites the reference ``file:line`` it follows (paths relative to the
reference checkout).  Pinning status (see DESIGN.md "Oracle"):

* ``model_ref``  - pinned: compared against the unmodified reference class
  ``src/model/R2plus1D_18_MotionNet.py`` imported in the build container
  (``tests/test_oracle_vs_reference.py``) and against committed golden
  vectors produced by that class (``tests/golden/``, ``oracle/make_golden.py``).
* ``fuse_ref.divide_to_consecutive_clips`` - pinned the same way (the
  reference function runs unmodified once its unused imports are stubbed).
* ``fuse_ref.generate_2dmotion_field`` / ``warp`` - pinned against the
  reference function (its ``.cuda()`` calls patched to no-ops) in the build
  container, plus golden vectors.
* ``fuse_ref.segment_a_video_with_fusion`` - control flow pinned against the
  reference function run with a stub voter; the voter itself (LabelFusion
  ``fuse_images``; package absent, version unpinned by the reference) is
  restated as plain majority voting: **parity unpinned** for SIMPLE/STAPLE.
* ``fuse_ref.warp_fuse`` - the north-star warp-and-fuse operator does not
  exist in the reference; it is specified here as a composition of the pinned
  primitives (softmax + warp + sum).
"""
