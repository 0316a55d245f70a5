"""TEST INFRASTRUCTURE ONLY -- import-time placeholder for torchio (see README.md)."""


def __getattr__(name):
    def _missing(*a, **k):
        raise RuntimeError("torchio.%s: torchio is not installed in this image" % name)
    return _missing
