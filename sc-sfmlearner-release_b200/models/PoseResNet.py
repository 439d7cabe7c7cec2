"""`from models.PoseResNet import PoseResNet` keeps working (reference models/PoseResNet.py:54)."""
from scsfm.nets import PoseDecoder, PoseResNet  # noqa: F401
