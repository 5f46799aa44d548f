#!/usr/bin/env bash
# Runs the reference's OWN pytest files against the drop-in modules of this package (not against the reference's copies).
#
# The reference keeps a private copy of its modules next to every test file, so pointing pytest at its test directories
# would test the reference itself.  This script stages ONLY the test_*.py files into a scratch directory
# (_reftests/, git-ignored; nothing of the reference is committed), with this package's Vch_control_{1D,2D} directories
# and the matplotlib stand-in on PYTHONPATH, so that every bare `from Forward2_solver import ...` resolves to the B200
# drop-in.
#
#   stage (needs the reference checkout, i.e. this container):  scripts/run_reference_tests.sh stage [/root/reference]
#   run   (needs a GPU; the staged files travel with gpurun):   scripts/run_reference_tests.sh run [pytest args]
#   clean:                                                      scripts/run_reference_tests.sh clean
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"
STAGE="$ROOT/_reftests"
cmd="${1:-run}"; shift || true
case "$cmd" in
  stage)
    REF="${1:-/root/reference}"
    rm -rf "$STAGE"; mkdir -p "$STAGE/1D" "$STAGE/2D"
    for f in "$REF"/src/1D/tests_1D/*/test_*.py; do cp "$f" "$STAGE/1D/"; done
    for f in "$REF"/src/2D/tests_2D/*/test_*.py; do cp "$f" "$STAGE/2D/"; done
    ls "$STAGE/1D" "$STAGE/2D" ;;
  run)
    rc=0
    for dim in 1D 2D; do
      ( cd "$STAGE/$dim" && MPLBACKEND=Agg PYTHONPATH="$PKG/Vch_control_$dim:$ROOT/oracle/_mpl_shim" \
          python -m pytest -q -p no:cacheprovider --rootdir "$STAGE/$dim" -o python_files='test_*.py' "$@" . ) || rc=1
    done
    exit $rc ;;
  clean) rm -rf "$STAGE" ;;
  *) echo "usage: $0 stage|run|clean"; exit 2 ;;
esac
