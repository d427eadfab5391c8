#!/bin/bash
# Build the REFERENCE (netgen + NGSolve, CPU only, no MPI, no CUDA) from /root/reference into oracle/_ref/ngs.
#
# Test infrastructure only.  This is the build that (a) generates the golden fixtures (tests/golden/make_golden*.py),
# (b) compiles the NGSolve-side adapter integration/ngsb200_ngla.cpp against the real headers and (c) is the
# `cpu_baseline.kind == "reference"` leg of bench.py when oracle/_ref/ngs is present on the box.
# NOT part of __graft_entry__.build(): it needs cmake and ~25 min on 8 cores (SURVEY.md 8c recipe).  Sources are
# compiled where they lie (a symlink farm gives the tree the directory name the version script wants); only build
# products are written, under $BUILD (scratch) and oracle/_ref/ngs (git-ignored install prefix).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference}
PFX=${PFX:-$HERE/_ref/ngs}
BUILD=${BUILD:-/tmp/ngsbuild}
JOBS=${JOBS:-6}
# portable instruction set (the GPU box's host CPU is not this container's): x86-64-v3 = AVX2+FMA, no -march=native
ARCH=${ARCH:--march=x86-64-v3}
LIBS=$(python -c "import os,cv2;print(os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)),'opencv_python_headless.libs'))" 2>/dev/null \
       || echo /opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs)
OPENBLAS=$(ls $LIBS/libopenblasp-*.so | head -1)
GFORTRAN=$(ls $LIBS/libgfortran-*.so.* | head -1)
QUADMATH=$(ls $LIBS/libquadmath-*.so.* | head -1)

farm() {  # farm <src> <dst>: dst is a real directory whose entries are symlinks into src; cmake/ is a real directory too
    mkdir -p "$2"
    for e in "$1"/* ; do
        b=$(basename "$e")
        if [ "$b" = cmake ]; then mkdir -p "$2/cmake"; for c in "$e"/*; do ln -sfn "$c" "$2/cmake/$(basename "$c")"; done
        else ln -sfn "$e" "$2/$b"; fi
    done
}
mkdir -p "$BUILD" "$PFX"
NGSRC=$BUILD/src/ngsolve_v6.2.2506-0-g0000000
farm "$REF" "$NGSRC"
# cmake/generate_version_file.cmake reads <source>/version.txt when the tree is not a git checkout
echo "v6.2.2506-0-g0000000" > "$NGSRC/version.txt"
# netgen lives below external_dependencies/ (a symlink into the read-only tree) -- build it from there
NETGEN=$REF/external_dependencies/netgen

if [ ! -f "$PFX/lib/cmake/netgen/NetgenConfig.cmake" ]; then
    cmake -S "$NETGEN" -B "$BUILD/netgen" -G Ninja -DUSE_SUPERBUILD=OFF -DUSE_GUI=OFF -DUSE_OCC=OFF -DUSE_PYTHON=ON \
          -DUSE_MPI=OFF -DBUILD_STUB_FILES=OFF -DCMAKE_BUILD_TYPE=Release -DCMAKE_INSTALL_PREFIX="$PFX" \
          -DUSE_NATIVE_ARCH=OFF "-DCMAKE_CXX_FLAGS=$ARCH" -DNETGEN_VERSION_GIT=v6.2.2506-0-g0000000
    ninja -C "$BUILD/netgen" -j"$JOBS" install
fi
PYBIND_INC=$NETGEN/external_dependencies/pybind11/include
cmake -S "$NGSRC" -B "$BUILD/ngsolve" -G Ninja -DUSE_SUPERBUILD=OFF -DNetgen_DIR="$PFX/lib/cmake/netgen" -DUSE_UMFPACK=OFF \
      -DUSE_MKL=OFF -DUSE_CUDA=OFF -DBUILD_STUB_FILES=OFF -DCMAKE_BUILD_TYPE=Release -DCMAKE_INSTALL_PREFIX="$PFX" \
      "-DCMAKE_CXX_FLAGS=-I$PYBIND_INC $ARCH" -DUSE_NATIVE_ARCH=OFF -DUSE_LAPACK=ON "-DLAPACK_LIBRARIES=$OPENBLAS;$GFORTRAN;$QUADMATH"
ninja -C "$BUILD/ngsolve" -j"$JOBS" install
cat > "$PFX/env.sh" <<'ENVEOF'
# source this: the reference build of oracle/build_reference.sh (paths relative to this file, so that it works on the GPU box)
_NGS_PFX="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
_NGS_LIBS=$(python -c "import os,importlib.util as u;s=u.find_spec('cv2');print(os.path.join(os.path.dirname(os.path.dirname(s.origin)),'opencv_python_headless.libs'))")
export PYTHONPATH=$_NGS_PFX/lib/python3.12/site-packages${PYTHONPATH:+:$PYTHONPATH}
export LD_LIBRARY_PATH=$_NGS_PFX/lib:$_NGS_PFX/lib/python3.12/site-packages/netgen:$_NGS_LIBS${LD_LIBRARY_PATH:+:$LD_LIBRARY_PATH}
ENVEOF
echo "reference installed under $PFX"
