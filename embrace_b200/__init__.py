"""Importable alias of the product package
`prediction-of-active-and-inactive-regulatory-regions-with-embracenet-multimodal-neural-network-_b200/`
(whose name is not a valid Python identifier).  Sub-modules resolve inside that directory."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         'prediction-of-active-and-inactive-regulatory-regions-with-embracenet-multimodal-neural-network-_b200')
__path__ = [_PKG_DIR]

from ._native import EmbError, build, lib          # noqa: E402,F401
from .archspec import ArchSpec                      # noqa: E402,F401
from .engine import Engine                          # noqa: E402,F401
