"""Import shim for the reference package (golden generation, reference arm of bench.py, reference-vs-drop-in tests).

This module is *tooling*: it arranges for ``import yolox`` to succeed against either the read-only checkout
(``/root/reference``, build container) or the pip-installed copy under ``baseline/_ref`` (git-ignored; made by
``baseline/install_reference.sh``; travels to the GPU box).  It contains no reference source and nothing under
``tscd_b200/`` imports it.

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

import os

REFERENCE_ROOT = "/root/reference"
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
# The reference hard-codes .to('cuda') in its attention modules (post_trans.py:694-695 and siblings).  True: redirect
# those moves to the CPU (CPU runs, also on a box that HAS a GPU); False: leave them alone (CUDA eager runs).
CPU_REDIRECT = True


def reference_root():
    """The checkout if present, else the installed copy; None if neither exists."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "yolox")):
        return REFERENCE_ROOT
    if os.path.isdir(os.path.join(INSTALLED_ROOT, "yolox")):
        return INSTALLED_ROOT
    return None
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


def install(root=None, cpu_redirect=None):
    """Make ``import yolox`` work.  cpu_redirect: initial value of CPU_REDIRECT (default: True iff no GPU is visible)."""
    global _installed, CPU_REDIRECT
    import torch
    if cpu_redirect is None:          # keep the current setting once installed; first install: redirect iff no GPU is visible
        cpu_redirect = CPU_REDIRECT if _installed else not torch.cuda.is_available()
    CPU_REDIRECT = bool(cpu_redirect)
    if _installed:
        return
    import torchvision

    root = root or reference_root()
    if root is None:
        raise RuntimeError("reference package not found: neither /root/reference nor baseline/_ref exists "
                           "(run baseline/install_reference.sh where /root/reference is available)")
    sys.meta_path.insert(0, _StubFinder())
    _install_pywt_stub()
    sys.path.insert(0, root)

    _orig_to = torch.Tensor.to

    def _to(self, *args, **kwargs):
        if CPU_REDIRECT:
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
