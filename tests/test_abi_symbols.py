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
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(tscd_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(handle, name) is not None
    assert b"sm_100a" in handle.tscd_version()


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (compile a tiny probe with gcc)."""
    import subprocess
    import tempfile
    from tscd_b200 import _lib
    src = '#include <stdio.h>\n#include "tscd_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(tscd_view),sizeof(tscd_anchors),sizeof(tscd_select_args),sizeof(tscd_nms_args),sizeof(tscd_gather_args),sizeof(tscd_linear_args));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    import ctypes
    mine = [ctypes.sizeof(x) for x in (_lib.View, _lib.Anchors, _lib.SelectArgs, _lib.NmsArgs, _lib.GatherArgs, _lib.LinearArgs)]
    assert sizes == mine
