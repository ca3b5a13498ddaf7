"""CPU oracle for the TSCD aggregation stage -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``tscd_b200`` never does; it fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests, golden vectors or weights
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference's
own PyTorch code run in the build container (``tools/make_goldens.py`` ->
``tests/golden/*.npz``) and, for the two un-vendored third-party algorithms
(torchvision ``nms``, SciPy ``linear_sum_assignment``), black-box against the
installed binaries.
"""
from .tscd_oracle import *  # noqa: F401,F403
