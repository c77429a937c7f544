"""CPU oracle for the dsp_core hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement of the reference's
``modules/dsp_core.py`` numeric path (numpy / scipy, float64).  It exists only
as the checker for the CUDA path:

* ``tests/``                      -- parity tests compare the CUDA path with it
* ``__graft_entry__.smoke()``     -- one tiny check on ``cuda:0``
* ``bench.py``                    -- the ``cpu_baseline`` leg and ``--impl reference``

Nothing under ``dsp_audio_project_b200/`` (the product) imports it, and the
product has no CPU fallback: it raises when the CUDA library is missing.

Parity pinning: the reference ships NO tests, golden vectors or fixtures
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports
``/root/reference/modules/dsp_core.py`` (with a stub ``soundfile`` module),
runs it on seeded inputs and commits the input/output vectors under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those vectors.
"""
