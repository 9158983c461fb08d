"""CPU oracle for the audio-features -> fusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the
timed CPU baseline.  The product path (``msa_b200``) never imports it and
fails loudly when the CUDA library is missing.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for
this path, so the oracle is pinned against outputs of the
reference itself, executed in the build container by ``oracle/make_golden.py``
(imports ``/root/reference`` by file path) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those vectors.
"""
