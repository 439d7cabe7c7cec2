import numpy as np


class ListedColormap:
    def __init__(self, colors, *a, **k):
        self.colors = np.asarray(colors)
        self.N = len(self.colors)

    def __call__(self, x):
        x = np.asarray(x)
        idx = np.clip((x * (self.N - 1)).astype(int), 0, self.N - 1)
        return self.colors[idx]


class LinearSegmentedColormap(ListedColormap):
    @staticmethod
    def from_list(name, colors, N=256):
        pts = [(c[0], c[1]) if isinstance(c[0], float) else (i / (len(colors) - 1), c) for i, c in enumerate(colors)]
        xs = np.array([p[0] for p in pts])
        cs = np.array([p[1] for p in pts], dtype=float)
        x = np.linspace(0, 1, N)
        out = np.stack([np.interp(x, xs, cs[:, i]) for i in range(cs.shape[1])] + [np.ones(N)], 1)
        return ListedColormap(out)
