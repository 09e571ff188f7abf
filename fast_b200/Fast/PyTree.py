"""``Fast.PyTree`` re-exports used by Fast/valid scripts (reference: Fast/Fast/Fast/PyTree.py:16-19)."""
from ..FastC.PyTree import _setNum2Zones, _setNum2Base, setNum2Zones, setNum2Base  # noqa: F401
