"""Drop-in for the reference's YOLOv3.py (ref testYOLOv3.py:8,71)."""
import _odt_path  # noqa: F401
from odt_b200.api import YOLOv3  # noqa: F401
