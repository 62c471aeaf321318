class Quaternion(object):
    def __init__(self, *args, **kwargs):
        raise RuntimeError("pyquaternion stub: the nuScenes loaders are outside the tested path")
