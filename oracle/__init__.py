"""CPU oracle for the PCGCv1 per-cube compress/decompress hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pcgcv1_b200/`` (the product) may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / baseline.

Each function restates, on the CPU, what the reference computes and cites the reference
file:line it follows (paths relative to the reference checkout).

Pinning status (see DESIGN.md "Oracle"):
  * ``oracle.entropy`` / ``oracle.nets`` graph wiring / ``oracle.topk`` are pinned against the
    reference's OWN Python source executed in the build container through a NumPy-backed
    TensorFlow shim (``tests/golden/make_golden.py``); the outputs are committed under
    ``tests/golden/``.
  * ``oracle.coder`` (pmf_to_quantized_cdf, range coder) restates the published algorithm of
    ``tensorflow.contrib.coder`` (tensorflow-gpu==1.13.1), which is NOT in the reference tree
    and cannot be run offline: **parity unpinned** for the coder byte stream and the
    quantised CDF normaliser.  Guaranteed instead: exact encode→decode round trips, valid
    CDF rows, coded size within a small constant of the estimated bits.
"""
