"""Importable alias for the package directory ``behavior-driven-video-synthesis_b200/``.

The task fixes the package directory name, and a hyphenated name is not a Python identifier, so this
shim points its ``__path__`` at that directory: ``import bdvs_b200`` and ``from bdvs_b200.vunets import
VunetAlter`` resolve to the files under ``behavior-driven-video-synthesis_b200/``.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "behavior-driven-video-synthesis_b200",
)
__path__.insert(0, _PKG_DIR)

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
