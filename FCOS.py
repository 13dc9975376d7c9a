"""Drop-in for the reference's FCOS.py (ref testfcos.py:7,60)."""
import _odt_path  # noqa: F401
from odt_b200.api import FCOS  # noqa: F401
