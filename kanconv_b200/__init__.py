"""``import kanconv_b200`` - import name of the package whose sources live in
``convolutional-kan-for-image-classification_b200/`` (that directory name is not a valid Python identifier, so this
shim points the package search path at it; every submodule - ``layers``, ``models``, ``functional``, ``build`` - is
found there)."""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "convolutional-kan-for-image-classification_b200")
if not _os.path.isdir(_SRC):
    raise ImportError("kanconv_b200: source directory %s is missing" % _SRC)
__path__ = [_SRC]

from ._exports import *  # noqa: E402,F401,F403
from ._exports import build  # noqa: E402,F401
