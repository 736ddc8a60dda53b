"""Windowed-attention HTR-VT variant (reference: model_window/model/HTR_VT.py)."""
from . import HTR_VT  # noqa: F401
