"""Importable name of the package that lives in ``sequential-inverse-kinematics_b200/``."""
from pathlib import Path as _Path

_impl = _Path(__file__).resolve().parent.parent / "sequential-inverse-kinematics_b200"
__path__.insert(0, str(_impl))
exec(compile((_impl / "__init__.py").read_text(), str(_impl / "__init__.py"), "exec"))
