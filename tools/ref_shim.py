"""Import shim for the read-only reference checkout (golden generation only).

This module is *tooling*: it is used by ``tools/make_goldens.py`` in the build
container, where ``/root/reference`` exists, to run the reference's own PyTorch
code on seeded inputs.  Nothing under ``tscd_b200/``, ``tests/`` (at run time),
``bench.py`` or ``__graft_entry__.py`` imports it, and it never copies reference
source: it only arranges for ``import yolox`` to succeed.

Recipe:
  1. permissive stub packages for the optional deps the reference imports at
     module scope but never needs on the aggregation path
     (thop, matplotlib, timm, pycocotools, seaborn);
  2. a Haar-only ``pywt`` stub (only needed to *construct* WaveletsHFBlock);
  3. ``sys.path.insert(0, reference_root)``;
  4. on CPU-only hosts redirect ``Tensor.to('cuda')`` to a no-op, because the
     reference hard-codes ``.to('cuda')`` (post_trans.py:694-695 and siblings);
  5. pin ``torchvision.ops.batched_nms`` to the coordinate-trick path, the one
     the reference takes on CUDA for every size this stage produces
     (torchvision/ops/boxes.py: numel > 100000 switches on CUDA, > 4000 on CPU).
"""
import importlib.abc
import importlib.machinery
import math
import sys
import types

REFERENCE_ROOT = "/root/reference"
_STUB_ROOTS = {"thop", "matplotlib", "pycocotools", "timm", "seaborn"}


class _Anything:
    """Object that absorbs attribute access / calls (for never-executed imports)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def _install_pywt_stub():
    if "pywt" in sys.modules:
        return
    s = 1.0 / math.sqrt(2.0)
    mod = types.ModuleType("pywt")

    class Wavelet:
        def __init__(self, name):
            assert name == "haar", "pywt stub only knows the Haar wavelet"
            self.dec_lo, self.dec_hi = [s, s], [-s, s]
            self.rec_lo, self.rec_hi = [s, s], [s, -s]

    mod.Wavelet = Wavelet
    sys.modules["pywt"] = mod


_installed = False


def install(reference_root=REFERENCE_ROOT):
    """Make ``import yolox`` work against the read-only reference checkout."""
    global _installed
    if _installed:
        return
    import torch
    import torchvision

    sys.meta_path.insert(0, _StubFinder())
    _install_pywt_stub()
    sys.path.insert(0, reference_root)

    if not torch.cuda.is_available():
        _orig_to = torch.Tensor.to

        def _to(self, *args, **kwargs):
            if args and isinstance(args[0], str) and args[0].startswith("cuda"):
                args = ("cpu",) + tuple(args[1:])
            if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
                kwargs["device"] = "cpu"
            return _orig_to(self, *args, **kwargs)

        torch.Tensor.to = _to

    from torchvision.ops import boxes as _tvb

    def _pinned_batched_nms(boxes, scores, idxs, iou_threshold):
        return _tvb._batched_nms_coordinate_trick(boxes.float(), scores.float(), idxs, iou_threshold)

    torchvision.ops.batched_nms = _pinned_batched_nms
    _tvb.batched_nms = _pinned_batched_nms
    _installed = True
