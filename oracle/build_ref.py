"""TEST INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference package into ``oracle/_ref/`` (git-ignored, travels to
the GPU box with gpurun) so the reference's own ``ctunet.pytorch.Model.forward_pass`` / model classes / loss handlers
can run there: as the checker of the drop-in tests and as the CPU ``--impl reference`` arm of bench.py.

    python oracle/build_ref.py            # needs /root/reference (build container only)

Recipe: copy the read-only tree to a scratch directory (setuptools writes build/ and egg-info into the source tree),
then ``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy>``.  ``--no-deps`` because
SimpleITK / raster_geometry / torchio / medpy / pynrrd / scikit-image are not in the offline wheelhouse; the four that
are imported at module level are stood in for by ``oracle/ref_stubs`` at import time (oracle/reference_loader.py).
No reference source is committed: ``oracle/_ref/`` is listed in .gitignore."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("CTUNET_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(HERE, "_ref")


def build_ref(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the installed reference afterwards."""
    marker = os.path.join(TARGET, "ctunet", "pytorch", "Model.py")
    if os.path.isfile(marker) and not force:
        return True
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "setup.py")):
        return os.path.isfile(marker)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "ctunet_src")
        shutil.copytree(REFERENCE_ROOT, src)
        shutil.rmtree(TARGET, ignore_errors=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
                        "--no-compile", "--find-links", "/opt/wheelhouse", "--target", TARGET, src], check=True)
    shutil.copytree(os.path.join(REFERENCE_ROOT, "examples"), os.path.join(TARGET, "examples"), dirs_exist_ok=True)
    return os.path.isfile(marker)


if __name__ == "__main__":
    ok = build_ref(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "reference tree not available")
