"""Drop-in for the reference's RetinaNet.py (ref testretinanet.py:7,73)."""
import _odt_path  # noqa: F401
from odt_b200.api import RetinaNet  # noqa: F401
