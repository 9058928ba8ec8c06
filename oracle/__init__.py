"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of the reference's refinement-loop algorithm (zhixuan-lin/IODINE,
``lib/modeling/iodine.py``) plus a loader that runs the unmodified reference when
``/root/reference`` is present (build container only; never on the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from this package, and only as the
checker or the timed CPU baseline -- never as a fallback for ``iodine_b200``.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so the restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF, run in the build container with injected noise, and the resulting
vectors are committed under ``tests/golden/`` together with the generating script
(``oracle/make_golden.py``).
"""
