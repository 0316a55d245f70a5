"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *real*, UNMODIFIED reference (vfmatzkin/ct-unet): from /root/reference inside the build container, or
from ``oracle/_ref`` (the pip-installed copy written by ``oracle/build_ref.py``; git-ignored, it travels to the GPU
box with the gpurun snapshot) anywhere else.  Used by ``oracle/make_golden*.py`` to generate the golden vectors,
by the tests that run the reference's own ``Model.forward_pass`` with the drop-in installed, and by the
``--impl reference`` / ``cpu_baseline`` legs of bench.py.

The reference imports four packages that are not installable here (SimpleITK, raster_geometry, monai, torchio --
SURVEY.md section 8c); ``oracle/ref_stubs`` is appended to ``sys.path`` so they resolve to the stand-ins described
in oracle/ref_stubs/README.md unless a real installation exists.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("CTUNET_REFERENCE_ROOT", "/root/reference")
INSTALLED_ROOT = os.environ.get("CTUNET_REFERENCE_INSTALL", os.path.join(HERE, "_ref"))
STUBS = os.path.join(HERE, "ref_stubs")


def reference_root():
    """Directory to put on sys.path, or None when the reference is nowhere to be found."""
    for root in (REFERENCE_ROOT, INSTALLED_ROOT):
        if os.path.isfile(os.path.join(root, "ctunet", "pytorch", "models.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def reference_kind() -> str:
    root = reference_root()
    return "none" if root is None else ("source tree" if root == REFERENCE_ROOT else "oracle/_ref install")


def load_reference(with_trainer: bool = False):
    """Returns (models, ProblemHandler, utilities, transforms) modules of the reference; with ``with_trainer`` also
    ``ctunet.pytorch.Model`` as a fifth element.  Autograd anomaly mode, which ``import ctunet`` switches on globally
    (Model.py:20), is restored to what it was."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present (neither %s nor %s)" % (REFERENCE_ROOT, INSTALLED_ROOT))
    import torch

    if STUBS not in sys.path:
        sys.path.append(STUBS)               # at the END: a real SimpleITK / monai / ... always wins
    if root not in sys.path:
        sys.path.insert(0, root)
    anomaly = torch.is_anomaly_enabled()
    importlib.import_module("ctunet")        # Model.py:20 switches anomaly mode on
    torch.autograd.set_detect_anomaly(anomaly)
    mods = [importlib.import_module(n) for n in ("ctunet.pytorch.models", "ctunet.pytorch.ProblemHandler",
                                                 "ctunet.utilities", "ctunet.pytorch.transforms")]
    if with_trainer:
        mods.append(importlib.import_module("ctunet.pytorch.Model"))
    return tuple(mods)
