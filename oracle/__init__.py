"""CPU oracle for the MatchNeRF per-ray hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``matchnerf_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker (or
as the timed CPU baseline), never as the product path.

Parity pin: ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (dev container only) and checks this restatement against
it; the resulting vectors live in ``tests/golden/``.
"""
