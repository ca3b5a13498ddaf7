"""Exp file for the reference's tools (tools/tscd_eval.py -f exps_b200/ovis_tscd_large_b200.py ...): identical to
exps/TSCD_OVIS/ovis_tscd_large.py except that get_model() swaps the head class for TSCDHeadB200.
Run from a reference checkout with this repository on PYTHONPATH."""
import importlib.util
import os


def _reference_exp():
    here = os.environ.get("TSCD_REFERENCE_ROOT", os.getcwd())
    path = os.path.join(here, "exps", "TSCD_OVIS", "ovis_tscd_large.py")
    spec = importlib.util.spec_from_file_location("ovis_tscd_large_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Exp


class Exp(_reference_exp()):
    def get_model(self):
        import yolox.models.tscd_head as ref_head
        from tscd_b200.head import make_head_class
        original = ref_head.TSCDHead
        ref_head.TSCDHead = make_head_class()       # get_model() imports the class by name (ovis_tscd_large.py:106)
        try:
            return super().get_model()
        finally:
            ref_head.TSCDHead = original
