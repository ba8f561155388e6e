"""Importable alias of the hyphenated package directory `vectordb-from-scratch_b200/`.

`import vectordb_from_scratch_b200` resolves every submodule inside that directory (Python
cannot import a name containing '-'; the directory name is fixed by the project layout)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vectordb-from-scratch_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
