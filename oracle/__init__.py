"""CPU oracle for the sparse-embedding + feature-interaction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline -- never as the implementation of a layer.

PARITY UNPINNED: the reference (PatrickHwang/Explicit-tf2-Recommendation)
ships no tests, golden vectors or recorded outputs for these layers, its
arithmetic lives in TensorFlow/Keras 2.8.0 (un-vendored, version read from
``*/output/saved_model.pb``) and TensorFlow is not installable in this image.
The oracle is therefore an op-for-op restatement of the reference ``call()``
bodies (file:line cited per function) pinned only by the hand-derivable
known-answer vectors in ``oracle/kats.py``.
"""
