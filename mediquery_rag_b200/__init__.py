"""Importable alias of the hyphenated package directory `mediquery-rag_b200/` (same modules, one copy)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mediquery-rag_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
