import numpy as np

from .colors import ListedColormap


def get_cmap(name, lutsize=256):
    g = np.linspace(0, 1, lutsize)
    return ListedColormap(np.stack([g, g, g, np.ones_like(g)], 1))
