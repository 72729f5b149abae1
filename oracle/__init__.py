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
* TensorFlow half (cdist_tf, batch_hard, lifted_loss): PARITY UNPINNED --
  TensorFlow is not installable here, the reference has no tests or golden
  vectors.  The restatement follows the TF graph op for op and is pinned only
  by the hand-checked known-answer tests of SURVEY.md Appendix B.
"""
