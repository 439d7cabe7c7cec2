"""Stand-in for imageio.imread (not installed) on top of PIL."""
import numpy as np
from PIL import Image


def imread(path):
    return np.asarray(Image.open(path))
