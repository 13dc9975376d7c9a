"""Drop-in for the reference's SSD512.py (ref testSSD512.py:7,60)."""
import _odt_path  # noqa: F401
from odt_b200.api import SSD512  # noqa: F401
