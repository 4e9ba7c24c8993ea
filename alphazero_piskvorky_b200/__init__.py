"""Import alias: the real package lives in ``alphazero-piskvorky_b200/`` (directory name fixed
by the build contract, not a valid identifier).  This module points its ``__path__`` there and
executes that package's ``__init__`` so ``import alphazero_piskvorky_b200`` works."""

import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "alphazero-piskvorky_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
