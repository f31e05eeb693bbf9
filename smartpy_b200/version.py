__version__ = '0.1.0'
# version of the reference API this package mirrors (smartpy/version.py)
__reference_version__ = '0.2.2'
