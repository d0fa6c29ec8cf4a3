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
* Model forward (Haiku/JAX): PARITY UNPINNED.  The reference ships no tests
  or golden vectors for the model, and jax/haiku are not installable here, so
  the fp32 forward in ``oracle/model.py`` follows the reference source
  line-by-line (citations in each function) but has never been compared with
  a real JAX run.  Golden vectors for the model are outputs of this oracle.
"""
