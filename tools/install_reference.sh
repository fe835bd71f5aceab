#!/bin/bash
# Offline install of the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with gpurun).
# The reference's setup.py uses find_packages(), which skips the four directories that have no __init__.py
# (models/odom, models/pc_transform, experiments, data/datasets) although the code imports them -- so the install is
# made from a copy under /tmp in which those directories get an EMPTY __init__.py (no source line is changed).
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/panoptic_forecasting" ] || { echo "no reference tree at $SRC"; exit 0; }
TMP="$(mktemp -d /tmp/pf_ref.XXXXXX)"
cp -r "$SRC/." "$TMP/"
find "$TMP/panoptic_forecasting" -type d | while read d; do [ -f "$d/__init__.py" ] || : > "$d/__init__.py"; done
rm -rf "$REPO/baseline/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
       --target "$REPO/baseline/_ref" "$TMP"
rm -rf "$TMP"
ls "$REPO/baseline/_ref/panoptic_forecasting/data/datasets" > /dev/null && echo "reference installed into baseline/_ref"
