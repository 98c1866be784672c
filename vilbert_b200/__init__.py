"""Importable alias for the product package, whose directory name (``vilbert-multi-task_b200``,
fixed by the repo layout contract) is not a valid Python identifier. This shim points ``__path__``
at that directory and executes its ``__init__``; all code lives there."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(__file__), "..", "vilbert-multi-task_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
