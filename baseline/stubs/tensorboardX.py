"""No-op stand-in for tensorboardX (not installed): the reference only calls add_scalar / add_image."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass

    def close(self):
        pass
