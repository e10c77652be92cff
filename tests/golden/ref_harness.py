"""Thin alias kept for the golden generators: the loader of the unmodified reference lives in ``oracle/ref_loader.py``."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_loader import *            # noqa: F401,F403,E402
from oracle.ref_loader import REF_PKG, REF_ROOT, available, install, modules, supplied_categorical, supplied_rand, supplied_uniforms   # noqa: F401,E402
