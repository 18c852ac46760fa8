#!/usr/bin/env bash
# Install the UNMODIFIED reference (blobctrl/ + its vendored diffusers 0.30.0 fork) into baseline/_ref with pip, so the
# cfg4 harness (scripts/cfg4_harness.py) and bench.py's reference arm can import it on the GPU box, where /root/reference
# does not exist.  baseline/_ref is git-ignored (never part of the history) but travels with the gpurun snapshot.
#
#   scripts/install_reference.sh [/root/reference]
#
# /root/reference is read-only and the diffusers fork ships without a setup.py / [project] table, so both trees are
# copied to a scratch directory first; the fork gets a four-line pyproject there.  --no-deps: the image already has
# torch / transformers / numpy / cv2, and the index is unreachable.
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DEST="$ROOT/baseline/_ref"
TMP="$(mktemp -d /tmp/refinstall.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT

[ -d "$REF/blobctrl" ] || { echo "no reference tree at $REF" >&2; exit 1; }
rm -rf "$DEST"; mkdir -p "$DEST"

mkdir -p "$TMP/blobctrl_pkg" "$TMP/diffusers_pkg"
cp -r "$REF/blobctrl" "$REF/pyproject.toml" "$REF/README.md" "$TMP/blobctrl_pkg/"
cp -r "$REF/diffusers/src" "$TMP/diffusers_pkg/src"
cat > "$TMP/diffusers_pkg/pyproject.toml" <<'EOF'
[build-system]
requires = ["setuptools>=61.0"]
build-backend = "setuptools.build_meta"
[project]
name = "diffusers"
version = "0.30.0"
[tool.setuptools.packages.find]
where = ["src"]
EOF

PIP="python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target $DEST -q"
$PIP "$TMP/blobctrl_pkg"
$PIP "$TMP/diffusers_pkg"
# the demo ellipses of the ten example states are the only reference data the harness reads (fixtures already hold them)
python - <<EOF
import hashlib, json, os
dest = "$DEST"
files = sorted(os.path.join(d, f) for d, _, fs in os.walk(dest) for f in fs if f.endswith(".py"))
h = hashlib.sha256()
for f in files:
    h.update(open(f, "rb").read())
json.dump({"source": "$REF", "py_files": len(files), "sha256": h.hexdigest()}, open(os.path.join(dest, "INSTALL.json"), "w"))
print("installed", len(files), "python files into", dest)
EOF
