"""nuScenes devkit stub: any name imports; using one raises (the nuScenes loaders are outside the tested path)."""


class _Missing(object):
    def __init__(self, *args, **kwargs):
        raise RuntimeError("nuscenes-devkit stub")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Missing
