"""Drop-in for the reference's ``tools/closing.py`` (class_closing :9, smooth_out :26), computed in CUDA."""
from ....ops import class_closing, smooth_out

__all__ = ["class_closing", "smooth_out"]
