"""CPU oracle for the SMART hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``smartpy_b200`` never does: the
product path has no CPU fallback.

* :mod:`oracle.model`  -- ctypes binding of ``smart_oracle.c`` (restates
  ``smartpy/structure.py:30-503`` of the reference operation for operation).
* :mod:`oracle.scores` -- numpy restatement of the objective functions called at
  ``smartpy/montecarlo/montecarlo.py:193-209`` (third-party ``spotpy``, not vendored in the
  reference) and of ``smartpy/objfunctions.py:20-24``.

Parity pin: reference-generated fixtures in ``tests/golden`` (see
``tests/golden/make_golden.py``) and the reference's printed known answers.  The scoring
half is pinned only to float32 precision by ``examples/out/ExampleDaily/ExampleDaily.SMART.lhs``
(spotpy itself is not installable here), see DESIGN.md.
"""
from .model import build, onestep, allsteps, run, run_members  # noqa: F401
from . import scores  # noqa: F401
