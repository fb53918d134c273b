"""
oracle/ -- CPU restatement of geNomad's nn-classification hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import or execute anything in this directory, and only as the
checker or as the timed CPU baseline -- never as a fallback for the CUDA path.
Nothing under ``genomad_b200/`` imports it.

What is pinned and what is not
------------------------------
* Tokenizer / windowing / FASTA reader (``oracle/tokenizer.py``, NumPy; plus ``oracle/tokenizer_c.c``, a plain-C
  restatement of the tokenizer and the window rules built by ``oracle/build_c.py`` / ``__graft_entry__.build()``):
  PINNED.  They are checked against golden vectors produced by running the *real*
  reference code (``/root/reference/genomad/sequence.py`` under numba) in the build
  container; generator: ``tests/golden/make_golden.py``, vectors: ``tests/golden/*.npz|json``.
* IGLOO model (``oracle/igloo_model.py``): STRUCTURE PINNED, TensorFlow's arithmetic unpinned.
  TensorFlow, Keras and h5py are not installable in the build container (no network,
  not in the wheelhouse) and the reference ships no tests or golden vectors for the
  model.  The reference's own model definition (``genomad/neural_network/model.py`` and
  ``igloo.py``, imported by path) is executed on a NumPy stand-in for the TF / Keras calls it
  makes (``tests/golden/keras_shim.py``); the resulting vectors are committed
  (``tests/golden/reference_graph_golden.npz``, generator ``make_reference_graph_golden.py``)
  and the restatement matches them to 5e-14 in fp64, with the shipped weights and with
  synthetic O(1) IGLOO weights.  The restatement follows ``model.py:9-45`` and
  ``igloo.py:30-83,190-217`` plus the Keras defaults they rely on.  Two independent
  formulations (op-for-op "as written" and a closed form) are also checked against each
  other and against frozen fp64 vectors in ``tests/golden/model_golden.npz``.
"""
