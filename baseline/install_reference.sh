#!/usr/bin/env bash
# Installs the UNMODIFIED reference (Video-Object-Detection/TSCD, a YOLOX fork packaged as `yolox`) into baseline/_ref
# with pip, offline.  baseline/_ref is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so the
# installed copy travels to the GPU box, where /root/reference does not exist.
#   --no-deps: the reference pins numpy==1.23 / protobuf<=3.20 and lists packages this image lacks (thop, lap, motmetrics,
#   timm, pycocotools ...); none of them is needed on the aggregation path and tools/ref_shim.py stubs the imports.
#   The build (FastCOCOEvalOp pre-compile hook) writes into the source tree, so it runs from a copy under /tmp.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/yolox" ] || { echo "no reference checkout at $SRC" >&2; exit 1; }
TMP="$(mktemp -d /tmp/tscd_ref.XXXXXX)"
cp -r "$SRC" "$TMP/src"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/src"
rm -rf "$TMP"
echo "installed: $HERE/_ref"
