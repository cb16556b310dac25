#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the unmodified reference (mars-project/mars) into baseline/_ref/ so that
#   * oracle/ref_runner.py and oracle/make_golden.py can import it (golden fixtures, oracle pinning),
#   * bench.py --impl reference can time the reference's OWN operands on the GPU box's host cores, and
#   * tests/test_reference_plugin.py can drive the real reference chunk graph through our executors.
# baseline/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored: the built
# copy travels to the GPU box with the snapshot, where /root/reference does not exist.
# Recipe: SURVEY.md Appendix B (copy tree -> build_ext --inplace; runtime shim in oracle/ref_shim.py).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC=${MARS_REF_SRC:-/root/reference}
DEST=${MARS_REF_DIR:-$HERE/../baseline/_ref}
if [ -d "$DEST/mars" ] && ls "$DEST"/mars/_utils*.so >/dev/null 2>&1; then
  echo "reference build already present in $DEST"; exit 0
fi
if [ ! -d "$SRC/mars" ]; then
  echo "no reference checkout at $SRC and no build in $DEST" >&2; exit 1
fi
mkdir -p "$DEST"
for f in mars setup.py setup.cfg versioneer.py pyproject.toml README.rst MANIFEST.in; do
  cp -r "$SRC/$f" "$DEST/"
done
chmod -R u+w "$DEST"
cd "$DEST"
NO_WEB_UI=1 python setup.py build_ext --inplace -j8 > build.log 2>&1
# keep the snapshot small: compiler intermediates and the generated C/C++ of the Cython modules are not needed
rm -rf "$DEST/build"
find "$DEST/mars" \( -name '*.c' -o -name '*.cpp' \) -newer "$DEST/setup.py" -delete 2>/dev/null || true
find "$DEST/mars" -type d \( -name tests -o -name __pycache__ \) -prune -exec rm -rf {} + 2>/dev/null || true
echo "built reference extensions in $DEST (log: $DEST/build.log)"
