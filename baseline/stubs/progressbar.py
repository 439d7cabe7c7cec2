"""Stand-in for progressbar2 (not installed): silent bars."""


class ProgressBar:
    def __init__(self, max_value=None, fd=None, **k):
        self.max_value = max_value

    def start(self):
        return self

    def update(self, *a, **k):
        pass

    def finish(self):
        pass
