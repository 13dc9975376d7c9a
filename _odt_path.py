"""Puts object-detection-tensorflow_b200/ (the product package directory; its
name is not a valid Python identifier) on sys.path."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "object-detection-tensorflow_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
