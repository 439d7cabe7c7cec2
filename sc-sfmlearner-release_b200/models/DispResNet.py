"""`from models.DispResNet import DispResNet` keeps working (reference models/DispResNet.py:104)."""
from scsfm.nets import DepthDecoder, DispResNet  # noqa: F401
