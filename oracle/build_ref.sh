#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the unmodified reference (mars-project/mars) in a scratch directory so
# that oracle/ref_runner.py and oracle/make_golden.py can import it (build container only; the GPU box has
# neither /root/reference nor this build -- tests there use the committed tests/golden/*.npz fixtures).
# No reference source is copied into the repository: the build lives in $MARS_REF_DIR (default /tmp/mars_ref).
set -euo pipefail
SRC=${MARS_REF_SRC:-/root/reference}
DEST=${MARS_REF_DIR:-/tmp/mars_ref}
if [ -d "$DEST/mars" ] && ls "$DEST"/mars/_utils*.so >/dev/null 2>&1; then
  echo "reference build already present in $DEST"; exit 0
fi
mkdir -p "$DEST"
for f in mars setup.py setup.cfg versioneer.py pyproject.toml README.rst MANIFEST.in; do
  cp -r "$SRC/$f" "$DEST/"
done
chmod -R u+w "$DEST"
cd "$DEST"
NO_WEB_UI=1 python setup.py build_ext --inplace -j8 > build.log 2>&1
echo "built reference extensions in $DEST (log: $DEST/build.log)"
