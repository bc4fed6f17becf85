"""Import the UNMODIFIED reference from /root/reference (read-only) in the build
container.  Only golden-vector generation scripts use this; nothing under
tests/ that runs on the GPU box may import it (the path does not exist there)."""
import contextlib
import io
import os
import sys

REF = os.environ.get("CAL_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "models", "hrnet"))


def setup():
    if not available():
        raise RuntimeError(f"reference not found at {REF}")
    for p in (os.path.join(HERE, "_shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)


@contextlib.contextmanager
def quiet():
    """The reference prints heavily inside the camera solve (prediction.py:260...)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
