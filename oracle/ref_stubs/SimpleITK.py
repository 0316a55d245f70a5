"""TEST INFRASTRUCTURE ONLY -- import-time placeholder for SimpleITK (NIfTI IO, out of scope; see README.md)."""


class Image:       # referenced in type positions only (ctunet/utilities.py:196-212)
    pass


def __getattr__(name):
    def _missing(*a, **k):
        raise RuntimeError("SimpleITK.%s: SimpleITK is not installed in this image (file IO is out of scope)" % name)
    return _missing
