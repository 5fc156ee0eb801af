#!/bin/sh
# Copy the UNMODIFIED reference tree to baseline/_ref so that bench.py's reference arms can run it on the GPU box, where
# /root/reference does not exist.  baseline/_ref is git-ignored (like the built .so files: no reference source enters
# the history) but not gpurun-ignored, so it travels with the snapshot.  The reference ships no setup.py/pyproject, so
# `pip install --target baseline/_ref /root/reference` has nothing to install: a verbatim copy is the install.
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -f "$SRC/diffusion/gaussian_diffusion.py" ]; then
  echo "install_ref: no reference tree at $SRC (nothing to do)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
# code only: images/ (README figures) are not needed
(cd "$SRC" && tar cf - --exclude=images --exclude=.git --exclude=__pycache__ .) | (cd "$DST" && tar xf -)
(cd "$SRC" && find . -name '*.py' -not -path './.git/*' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
echo "install_ref: copied $(find "$DST" -name '*.py' | wc -l) python files to $DST"
