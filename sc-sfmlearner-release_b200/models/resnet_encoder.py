"""`from models.resnet_encoder import ResnetEncoder` keeps working (reference models/resnet_encoder.py:62)."""
from scsfm.nets import ResnetEncoder  # noqa: F401
