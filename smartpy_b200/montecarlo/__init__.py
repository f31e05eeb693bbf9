"""``smartpy.montecarlo`` surface (smartpy/montecarlo/__init__.py:19-22)."""
from .lhs import LHS
from .glue import GLUE
from .best import Best
from .total import Total

__all__ = ['LHS', 'GLUE', 'Best', 'Total']
