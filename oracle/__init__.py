"""CPU oracle for the structure-tokenization hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (NumPy fp64 /
torch-CPU fp32) of the reference algorithm
(xwang112358/protein-structure-tokenizer, mounted read-only at /root/reference
while this repo was developed).  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it, and only as the checker or the timed CPU baseline -- never as the product
path.  The product (``protein-structure-tokenizer_b200/``) must not import it.

Parity pinning status
---------------------
* Featuriser (frames, centroid, k-NN, 27-d edge features): PINNED.  The
  restatement in ``oracle/featurize.py`` was checked bit-for-bit against the
  reference's own NumPy functions (``compute_nearest_neighbors_graph``,
  ``make_transform_from_reference``) imported through stub modules on all 31
  bundled CASP14 structures; outputs of the reference functions are committed
  as golden vectors under ``tests/golden/`` (generator:
  ``tests/golden/make_golden.py``).
* Model forward (Haiku/JAX): PINNED AGAINST THE REFERENCE'S SOURCE, not against
  XLA.  ``tests/golden/make_golden_model.py`` executes the reference's own
  ``Vq3D.encode_and_quantize`` (files under /root/reference, unmodified) over
  NumPy stand-ins for jax / haiku (``tests/golden/refshim.py``: JAX's x64-off
  dtype rules, Haiku 0.0.10 module naming, layer_stack's creator/getter hooks)
  for four released configs on CASP14 structures and commits tokens, bounded
  latents, pre-projection embeddings and the Haiku parameter names under
  ``tests/golden/model_ref_*``.  ``oracle/model.py`` reproduces them to <= 4e-6
  with identical tokens (``tests/test_golden_model.py``).  What remains
  unpinned is XLA:CPU's fp32 arithmetic itself (summation order, last-ulp
  tanh/exp/sin/cos/pow): jax and haiku are not installable in this image.
"""
