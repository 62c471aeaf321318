"""No-op stand-in for tensorboardX (the reference's Recorder only writes scalars/images to it)."""


class SummaryWriter(object):
    def __init__(self, *args, **kwargs):
        self.calls = 0

    def __getattr__(self, name):
        def _noop(*args, **kwargs):
            self.calls += 1
        return _noop
