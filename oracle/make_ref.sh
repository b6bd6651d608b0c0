#!/bin/bash
# TEST / BASELINE INFRASTRUCTURE ONLY.  Copies the reference's own Python package for the hot path (lib/: utils.py, layers.py,
# filtering/oanet.py, pairwise/, config.py and what they import) from the read-only reference tree into the git-ignored
# oracle/_ref/ so that it travels to the GPU box with gpurun: `bench.py --impl reference` then times the UNMODIFIED reference code
# (imported through oracle/refimport.py with LMPCR_REFERENCE_ROOT=oracle/_ref) on the box's host cores instead of the numpy/C port.
# The reference is pure Python: there is nothing to compile.  Nothing under oracle/_ref is committed (.gitignore) and the product
# package never imports it.
#   oracle/make_ref.sh [reference root, default /root/reference]
set -e
SRC="${1:-/root/reference}"
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/lib" ]; then
  echo "make_ref: no reference tree at $SRC (nothing to do)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/configs/pairwise_registration/eval"
cp -r "$SRC/lib" "$DST/lib"
find "$DST/lib" -name "__pycache__" -type d -prune -exec rm -rf {} +
cp "$SRC/configs/pairwise_registration/eval/RegBlock.yaml" "$DST/configs/pairwise_registration/eval/"
cp "$SRC/LICENSE" "$DST/LICENSE"
echo "make_ref: reference package copied to $DST ($(du -sh "$DST" | cut -f1))"
