"""CPU oracle for the affinity U-Net watershed path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``iterseg_b200`` (the product) imports this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute it, and there only as the
checker / the CPU baseline -- never as the thing shipped.

What is pinned and what is not
------------------------------
The reference (AbigailMcGovern/iterseg) holds no tests, golden vectors or
known-answer fixtures for this path (its ``_tests`` package is empty), so the
pinning is done against outputs of the reference's own modules executed in the
build container (``oracle/ref_harness.py`` imports them *verbatim* from
``/root/reference``; ``scripts/make_golden.py`` writes ``tests/golden/``):

* ``flood.c`` / ``flood.py``      -- pinned: bit-identical to the reference numba
  kernel ``raveled_affinity_watershed`` (watershed.py:95-159) and its
  ``py_func`` on every golden scene.
* ``unet_ref.py``                 -- pinned: bit-identical to the reference
  ``UNet.forward`` (unet.py:284-364) on the golden chunk.
* ``chunks.py``                   -- pinned against ``make_chunks`` /
  ``process_chunks`` (predict.py:38-96).
* ``post.py`` (seeds, mask, CCL)  -- PARITY UNPINNED for the scikit-image calls:
  scikit-image is not installed here and cannot be installed, so
  ``skimage_shim.py`` restates its published algorithms on scipy/numpy
  (SURVEY.md Appendix B).  The reference's own control flow around those calls
  (watershed.py:165-251) *is* executed verbatim on top of the shim when the
  golden vectors are made.
"""
