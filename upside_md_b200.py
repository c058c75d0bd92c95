"""Import shim: the package directory is ``upside-md_b200/`` (hyphenated, per the project layout),
which Python cannot import by name.  Giving this module a ``__path__`` makes it behave as that package:
``import upside_md_b200.h5lite`` resolves to ``upside-md_b200/h5lite.py``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "upside-md_b200")]
