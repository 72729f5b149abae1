"""CPU oracle for the metric-learning hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithm
(johndpope/multimodal_similarity: src/utils.py:55-360, src/networks.py:797-870,
src/evaluate_late_fusion.py:115-116).  It exists to *check* the CUDA path.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``multimodal_similarity_b200/`` imports it; the product path has no CPU
fallback and fails loudly when the CUDA extension is missing.

Parity pin status
-----------------
* NumPy half (distances, ranking, AP, recall@K, precision@recall, evaluate,
  evaluate_simple): PINNED -- ``oracle/make_golden.py`` ran the unmodified
  reference functions (tensorflow stubbed by an empty module) in the build
  container and committed their outputs as ``tests/golden/*.npz``;
  ``tests/test_oracle_golden.py`` checks this restatement against them
  bit-for-bit.
* TensorFlow half (cdist_tf, batch_hard, lifted_loss): TensorFlow is not
  installable here and the reference has no tests or golden vectors, so
  TensorFlow's own kernels are never executed.  The restatement
  (``oracle/losses_torch.py``) follows the TF graph op for op and is pinned by
  (a) the hand-checked known-answer tests of SURVEY.md Appendix B and (b) the
  reference's OWN functions executed unmodified under a torch-backed stand-in
  for the ``tf.*`` calls they make (``oracle/tf_shim.py``,
  ``oracle/make_golden_losses.py`` -> ``tests/golden/loss_*.npz``,
  ``tests/test_oracle_losses_golden.py``).  That pins the reference's source
  as executed code; the stand-in's reading of TF semantics (even tie split of
  reduce_max/min gradients, reducers of empty tensors) is stated in its header.
* tf.contrib triplet_semihard_loss / lifted_struct_loss: third-party source
  absent from /root/reference -- PARITY UNPINNED.
"""
