"""Importable alias of the `lit-llama_b200/` package directory (a hyphen cannot be
imported).  All code lives in `lit-llama_b200/`; this file only points the package
search path there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "lit-llama_b200"))

from ._init import *  # noqa: F401,F403,E402
from ._init import __all__  # noqa: F401,E402
