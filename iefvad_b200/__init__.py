"""Importable alias of the product package.

The product lives in the directory the project layout names, `ief-vad_b200/`, which is not a valid Python
identifier; this one-file package extends its `__path__` to that directory so that
`import iefvad_b200` / `from iefvad_b200.imf_vad import MMFMIL` resolve to `ief-vad_b200/*.py`."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ief-vad_b200"))

from .api import *  # noqa: E402,F401,F403  (ief-vad_b200/api.py)
