"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference (ICEORY/PMF) algorithms on the hot path, used as the parity checker.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``pmf_b200``, ``pc_processor`` shim) never does.

Pinning status: the reference ships no tests, golden vectors or fixtures, so the
restatements are pinned against OUTPUTS OF THE REFERENCE ITSELF, imported read-only from /root/reference
in the build container by ``tests/golden/make_golden.py`` (committed together with the fixtures it wrote
under ``tests/golden/``).  ``tests/test_oracle_pinning.py`` re-checks oracle == reference live whenever
/root/reference is present, and oracle == committed fixtures everywhere (including the GPU box).
"""
