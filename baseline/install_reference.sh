#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the gpurun snapshot) so that
# `bench.py --impl reference` and bench.py's cpu_baseline leg drive the reference's own code on the host cores.
# The reference is pure Python without a setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference`
# has nothing to build ("does not appear to be a Python project"); this script copies the package directories the hot path
# imports, byte for byte.  Nothing under baseline/_ref is ever committed or edited.
set -eu
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/modules/core/ddpm.py" ]; then
    echo "install_reference: $SRC is not the reference checkout" >&2
    exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
for d in modules utils inference basics configs; do
    cp -r "$SRC/$d" "$DST/$d"
done
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find modules utils inference basics configs -type f -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "install_reference: $(find "$DST" -name '*.py' | wc -l) python files -> $DST"
