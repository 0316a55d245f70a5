"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *real* reference (vfmatzkin/ct-unet, mounted read-only at
/root/reference) inside the build container so that `oracle/make_golden.py`
can generate golden vectors and so that `tests/test_oracle_vs_reference.py`
can pin the restatement in `oracle/unet_oracle.py` against it.

/root/reference does not exist on the GPU box; everything that runs there
uses the committed fixtures under tests/golden/ instead.

The reference imports four packages that are not installable here
(SimpleITK, raster_geometry, monai, torchio -- see SURVEY.md section 8c); empty
stub modules are inserted so that `ctunet.pytorch.models`,
`ctunet.pytorch.ProblemHandler` and `ctunet.utilities` import.  None of the
stubs is ever called on the paths the oracle exercises.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CTUNET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ctunet", "pytorch", "models.py"))


def load_reference():
    """Returns (models_module, ProblemHandler_module, utilities_module, transforms_module)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    for name in ["SimpleITK", "raster_geometry", "monai", "monai.metrics", "torchio"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["raster_geometry"].cylinder = None
    sys.modules["raster_geometry"].cube = None
    sys.modules["SimpleITK"].Image = type("Image", (), {})
    sys.modules["monai"].metrics = sys.modules["monai.metrics"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    anomaly = torch.is_anomaly_enabled()
    ctunet = importlib.import_module("ctunet")  # noqa: F841  (Model.py:20 switches anomaly mode on)
    torch.autograd.set_detect_anomaly(anomaly)
    models = importlib.import_module("ctunet.pytorch.models")
    handler = importlib.import_module("ctunet.pytorch.ProblemHandler")
    utilities = importlib.import_module("ctunet.utilities")
    transforms = importlib.import_module("ctunet.pytorch.transforms")
    return models, handler, utilities, transforms
