"""TEST INFRASTRUCTURE ONLY -- placeholder package; only ``monai.metrics`` is provided (see ../README.md)."""
from . import metrics  # noqa: F401
