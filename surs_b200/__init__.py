"""Import alias for the product package.

The product lives in ``super-resolution-3d-human-shape-from-a-single-low-resolution-image_b200/``
(the directory name the project layout prescribes); hyphens make that name
unimportable, so this shim exposes it as ``surs_b200`` by extending ``__path__``.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "super-resolution-3d-human-shape-from-a-single-low-resolution-image_b200",
)
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR

from ._version import __version__  # noqa: E402,F401
