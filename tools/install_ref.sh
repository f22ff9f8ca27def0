#!/bin/bash
# Copies the UNMODIFIED reference sources the CPU arm and the drop-in test need from /root/reference (read-only, present in the build
# container only) into baseline/_ref/ -- git-ignored, NOT gpurun-ignored, so it travels to the GPU box with the snapshot -- plus a
# 4-line matplotlib stub (the reference imports matplotlib.pyplot at module top, multiTransformer.py:7 / train.py:15; matplotlib is not
# installed and only the plotting helpers use it).  Nothing under baseline/_ref is ever tracked or edited.
#   usage: tools/install_ref.sh [reference root, default /root/reference]
set -e
SRC="${1:-/root/reference}/transformer"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$SRC" ]; then
  echo "install_ref: $SRC not found (GPU box?) -- keeping whatever is in $DST" >&2
  exit 0
fi
mkdir -p "$DST"
for d in MFT SFT B2-Trans B3-MFN; do
  mkdir -p "$DST/$d"
  for f in multiTransformer.py models.py train.py datasets.py; do
    [ -f "$SRC/$d/$f" ] && cp "$SRC/$d/$f" "$DST/$d/$f"
  done
done
mkdir -p "$DST/_stubs/matplotlib"
cat > "$DST/_stubs/matplotlib/__init__.py" <<'PY'
"""Stub: the reference imports matplotlib at module top; only its plotting helpers (out of scope) would use it."""
PY
cat > "$DST/_stubs/matplotlib/pyplot.py" <<'PY'
def __getattr__(name):
    raise RuntimeError('matplotlib is stubbed out (baseline/_ref/_stubs): plotting is outside the benchmarked path')
PY
( cd "$DST" && find . -name '*.py' | sort | xargs sha256sum > MANIFEST.sha256 )
echo "install_ref: copied $(find "$DST" -name '*.py' | wc -l) files into $DST"
