"""CPU test: the C-ABI library builds, loads and exports every symbol include/tscd_b200.h declares
(no compute calls without a GPU)."""
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from tscd_b200 import build, _lib
    build.build()
    handle = _lib.lib()
    header = open(os.path.join(ROOT, "include", "tscd_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(tscd_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(handle, name) is not None
    assert b"sm_100a" in handle.tscd_version()


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (compile a tiny probe with gcc)."""
    import ctypes
    import subprocess
    import tempfile
    from tscd_b200 import _lib
    pairs = [("tscd_view", _lib.View), ("tscd_anchors", _lib.Anchors), ("tscd_select_args", _lib.SelectArgs), ("tscd_pack_head_args", _lib.PackHeadArgs),
             ("tscd_nms_args", _lib.NmsArgs), ("tscd_gather_args", _lib.GatherArgs), ("tscd_linear_args", _lib.LinearArgs),
             ("tscd_attn_layout", _lib.AttnLayout), ("tscd_attn_prep_args", _lib.AttnPrepArgs),
             ("tscd_attn_pv_args", _lib.AttnPvArgs), ("tscd_attn_round2_args", _lib.AttnRound2Args),
             ("tscd_transpose_args", _lib.TransposeArgs), ("tscd_cafm_prep_args", _lib.CafmPrepArgs), ("tscd_cafm_chain_args", _lib.CafmChainArgs), ("tscd_cafm_cost_args", _lib.CafmCostArgs), ("tscd_cafm_lap_args", _lib.CafmLapArgs),
             ("tscd_frame_attention_args", _lib.FrameAttentionArgs), ("tscd_residual_ln2_args", _lib.ResidualLn2Args),
             ("tscd_final_expand_args", _lib.FinalExpandArgs), ("tscd_final_rows_args", _lib.FinalRowsArgs),
             ("tscd_bank_pack_args", _lib.BankPackArgs), ("tscd_bank_unpack_args", _lib.BankUnpackArgs),
             ("tscd_qkv_project_args", _lib.QkvProjectArgs), ("tscd_attn_rowmeta_args", _lib.AttnRowmetaArgs),
             ("tscd_local_offsets_args", _lib.LocalOffsetsArgs), ("tscd_pack_rows_args", _lib.PackRowsArgs),
             ("tscd_pack_detections_args", _lib.PackDetectionsArgs), ("tscd_repp_link_args", _lib.ReppLinkArgs),
             ("tscd_cafm_wide_args", _lib.CafmWideArgs), ("tscd_frame_flash_args", _lib.FrameFlashArgs),
             ("tscd_edge_patches_args", _lib.EdgePatchesArgs), ("tscd_edge_combine_args", _lib.EdgeCombineArgs)]
    body = "".join(f'printf("%zu\\n", sizeof({c}));' for c, _ in pairs)
    src = '#include <stdio.h>\n#include "tscd_b200.h"\nint main(){' + body + 'return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(t) for _, t in pairs]
