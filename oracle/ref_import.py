"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (liuyuan-pal/NeuRay) from
/root/reference so that golden vectors can be generated from it (oracle/gen_golden.py).

The reference is pure Python/PyTorch but `network/renderer.py` transitively imports packages that
are absent from this image (skimage, easydict, h5py, plyfile, transforms3d, imageio, matplotlib,
tensorboardX, inplace_abn, kornia).  None of them is touched by the per-ray rendering path, so we
register empty module stubs (with a __spec__, otherwise torch._dynamo's find_spec probes break) and
let the reference import.  `IBRNetWithNeuRay.posenc` pins its table to "cuda:0"
(network/ibrnet.py:312); on a GPU-less box we make that particular `.to("cuda:0")` a no-op.

Nothing in neuray_b200/ may import this file.  It only works where /root/reference exists (the build
container) -- never on the GPU box.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NEURAY_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "easydict", "skimage", "skimage.io", "skimage.metrics", "kornia", "kornia.utils", "h5py", "plyfile",
    "transforms3d", "transforms3d.axangles", "transforms3d.euler", "imageio", "matplotlib",
    "matplotlib.pyplot", "matplotlib.lines", "matplotlib.cm", "tensorboardX", "lpips", "inplace_abn",
]


class _Anything:
    """Placeholder for names imported from stubbed packages; never called on the rendering path."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party symbol was called")


def _make_stub(name):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything
    mod.__getattr__ = _getattr
    return mod


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "network"))


def load_reference():
    """Returns the reference's `network.renderer` module (imports it under stubs)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torch
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _make_stub(name)
    if not torch.cuda.is_available():
        _orig_to = torch.Tensor.to

        def _to(self, *args, **kwargs):
            if args and isinstance(args[0], str) and args[0].startswith("cuda"):
                args = args[1:]
                if not args and not kwargs:
                    return self
            return _orig_to(self, *args, **kwargs)
        torch.Tensor.to = _to
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)  # the reference opens relative paths (configs/, network/mvsnet/*.ckpt)
    try:
        import network.renderer as ref_renderer
    finally:
        os.chdir(cwd)
    return ref_renderer
