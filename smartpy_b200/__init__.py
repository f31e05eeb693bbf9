"""smartpy_b200 -- the SMART rainfall-runoff hot path on NVIDIA B200 (sm_100a).

Drop-in for the public API of ThibHlln/smartpy v0.2.2 on the simulation path
(``smartpy/__init__.py:19-21`` exports ``SMART``, ``objfunctions``, ``__version__``):

    import smartpy_b200 as smartpy
    sm = smartpy.SMART(...); sm.simulate(sm.parameters.values)
    smartpy.montecarlo.LHS(...).run()

Every simulation goes through hand-written CUDA kernels behind the C ABI declared in
``include/smart_b200.h``.  There is no CPU fallback: importing works anywhere (so the host
logic can be tested), running needs the built library and a GPU.
"""
from .smart import SMART
from . import objfunctions
from .version import __version__

__all__ = ['SMART', 'objfunctions', '__version__']
