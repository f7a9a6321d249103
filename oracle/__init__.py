"""CPU oracle for the FFN inference hot path — TEST INFRASTRUCTURE, not product code.

Everything under ``oracle/`` is a CPU restatement (numpy + torch-CPU conv3d) of the reference's
flood-fill inference path, written to check the CUDA engine in ``ffn_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product path never does: ``ffn_b200`` fails loudly when its CUDA library is
missing instead of falling back to this code.

Parity pinning status (see DESIGN.md "Oracle"):

* flood-fill logic (Canvas.update_at/segment_at/segment_all, FaceMaxMovementPolicy,
  get_scored_move_offsets, quantize_probability, PolicyGrid3d): PINNED — checked against outputs of
  the reference's own Python modules imported in the build container with third-party stubs
  (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); likewise the masked run (mask, seed_mask,
  shift mask via MovementRestrictor), the anisotropic geometry of configs[4]
  (``make_golden_masks.py``), Canvas.history / history_deleted (``make_golden_history.py``) and
  MovementRestrictor.is_valid_pos at every voxel (``make_golden_restrictor.py``).
* network arithmetic (TensorFlow Conv3D/BiasAdd/Relu via tf_slim, un-vendored and not
  installable offline): PARITY UNPINNED — restated with torch CPU ``conv3d`` from
  ffn/training/models/convstack_3d.py:26-56,83-95 and ffn/training/model.py:168-183; the only
  anchors are the shipped checkpoint's op attributes and the plausibility of the resulting
  segmentations (results/fib25/sample-training2.npz sanity ranges).
"""
