"""Importable alias of the product package.

The product lives in ``embodied-one-shot-video-recognition_b200/`` (the directory name the
project layout prescribes; hyphens make it un-importable by name).  ``import eosvr_b200``
resolves its submodules from that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "embodied-one-shot-video-recognition_b200")
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
