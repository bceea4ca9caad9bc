"""Importable alias of the package directory ``vision-language-model-framework_b200/`` (a hyphen is not a valid
Python identifier, so the real sources live there and this stub points ``__path__`` at them)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vision-language-model-framework_b200")
__path__.insert(0, _real)  # type: ignore[name-defined]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
