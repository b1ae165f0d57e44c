"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (``/root/reference``) in this container so that
golden vectors can be generated from the reference's own code.
``/root/reference`` does not exist on the GPU box, therefore nothing under
``tests/`` (gpu marker), ``bench.py`` or ``__graft_entry__.smoke()`` may import
this module at run time; only ``tests/golden/make_golden.py`` (run here, output
committed) and CPU-side cross-check tests that skip when the mount is absent.

The reference cannot be imported raw: ``lib/model/SuRSNet.py:35`` ->
``lib/net_util.py:8`` -> ``lib/mesh_util.py:1`` imports scikit-image, which is
not installed.  We pre-seed ``sys.modules`` with empty ``skimage`` modules.
"""
import contextlib
import io
import os
import sys
import types
import argparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_MOUNT = os.environ.get("SURS_REFERENCE_ROOT", "/root/reference")
# the read-only mount when it exists (build container), else the verbatim copy oracle/make_ref.py made of it
# (oracle/_ref: git-ignored, travels to the GPU box with gpurun; only bench.py's CPU legs use it there)
REF_ROOT = _MOUNT if os.path.isdir(os.path.join(_MOUNT, "lib")) else os.path.join(_HERE, "_ref" if not os.environ.get("SURS_NO_REF_COPY") else "_none")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib"))


def is_mount() -> bool:
    return REF_ROOT == _MOUNT


def use_marching_cubes(fn):
    """Routes the reference's ``measure.marching_cubes_lewiner`` (lib/mesh_util.py:40,45) to ``fn`` -- the oracle's C
    twin when scikit-image is absent (labelled as such wherever a number is reported)."""
    _stub_skimage()
    sys.modules["skimage.measure"].marching_cubes_lewiner = fn
    if "lib.mesh_util" in sys.modules:
        sys.modules["lib.mesh_util"].measure.marching_cubes_lewiner = fn


def have_real_skimage() -> bool:
    m = sys.modules.get("skimage")
    return m is not None and getattr(m, "__file__", None) is not None


def _stub_skimage():
    if "skimage" in sys.modules:
        return
    try:  # pragma: no cover - not installed in this image
        import skimage  # noqa: F401
        return
    except Exception:
        pass
    sk = types.ModuleType("skimage")
    measure = types.ModuleType("skimage.measure")

    def _no_mc(*a, **k):
        raise RuntimeError("scikit-image is not installed: marching cubes parity is unpinned")

    measure.marching_cubes_lewiner = _no_mc
    sk.measure = measure
    sys.modules["skimage"] = sk
    sys.modules["skimage.measure"] = measure


def import_reference():
    """Returns the reference's ``lib`` package (imported in place, unmodified)."""
    if not available():
        raise RuntimeError("reference mount %s not present" % REF_ROOT)
    _stub_skimage()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import lib  # noqa: F401
        import lib.sdf  # noqa: F401
        import lib.geometry  # noqa: F401
        import lib.options  # noqa: F401
        import lib.model.SuRSNet  # noqa: F401
        import lib.mesh_util  # noqa: F401
    return sys.modules["lib"]


def make_opt(extra=()):
    """``opt`` exactly as the reference CLI would build it (lib/options.py:9-185)."""
    lib = import_reference()
    parser = argparse.ArgumentParser()
    parser = lib.options.BaseOptions().initialize(parser)
    return parser.parse_args(list(extra))


@contextlib.contextmanager
def quiet():
    """Silences DepthNormalizer.forward's stray print (lib/model/DepthNormalizer.py:17)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
