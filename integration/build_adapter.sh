#!/bin/bash
# Compile the NGSolve-side adapter (integration/ngsb200_ngla.cpp) against an installed NGSolve -- here the reference
# build of oracle/build_reference.sh -- and link it with libngsb200.so.  Output: integration/_build/_ngsb200*.so
# (git-ignored; it travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
PFX=${PFX:-$ROOT/oracle/_ref/ngs}
mkdir -p "$HERE/_build"
SUFFIX=$(python3-config --extension-suffix 2>/dev/null || python -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
"$PFX/bin/ngscxx" -shared -I"$ROOT/include" "$HERE/ngsb200_ngla.cpp" \
    -L"$PFX/lib" -lngla -lngstd -lngbla -L"$PFX/lib/python3.12/site-packages/netgen" -lngcore \
    -L"$ROOT/ngsolve_b200/lib" -lngsb200 '-Wl,-rpath,$ORIGIN/../../ngsolve_b200/lib' \
    -o "$HERE/_build/_ngsb200$SUFFIX"
echo "built $HERE/_build/_ngsb200$SUFFIX"
