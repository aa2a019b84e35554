"""Import alias: `import gym_multigrid_b200` -> the package that lives in `gym-multigrid_b200/`.

The package directory is named after the reference repo (`gym-multigrid` + `_b200`), and a
hyphen is not importable; this stub points the import system at that directory and runs
its `__init__.py` in this module's namespace.  All code lives in `gym-multigrid_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gym-multigrid_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
