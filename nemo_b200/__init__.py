"""
nemo_b200 -- B200-native (sm_100a CUDA behind a C ABI) implementation of Nemo's per-tile
filter -> noise / S-N -> detect -> photometry hot path, keeping the ``nemo.filters`` / ``nemo.photometry``
call surface (reference: borisbolliet/nemo-1, nemo/filters.py, nemo/photometry.py, nemo/pipelines.py:112-354).

There is no CPU fallback: compute entry points raise if ``libnemo_b200.so`` is missing or no CUDA device is
present.  Importing the package itself needs neither.
"""

__version__ = "0.1.0"
