#!/bin/bash
# Build the REFERENCE's own device layer -- ngscuda (cuSPARSE DevSparseMatrix, cuBLAS UnifiedVector, DevCGSolver) -- from
# /root/reference/ngscuda against the reference install of build_reference.sh, for sm_100.  Its CMakeLists.txt supports
# exactly this stand-alone mode (project(ngscuda), find_package(NGSolve)).  Test / measurement infrastructure only:
# bench.py's `gpu_reference` block times "that kernel"  on the same B200.  nvcc cross-compiles here.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference}
PFX=${PFX:-$HERE/_ref/ngs}
BUILD=${BUILD:-/tmp/ngscuda_build}
JOBS=${JOBS:-6}
PYBIND_INC=$REF/external_dependencies/netgen/external_dependencies/pybind11/include
PY_INC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
mkdir -p "$BUILD/dummy"
# stand-alone mode still names the in-tree interface target `netgen_python` on the link line: satisfy it with an empty archive
ar rc "$BUILD/dummy/libnetgen_python.a"
cmake -S "$REF/ngscuda" -B "$BUILD" -G Ninja -DNGSolve_DIR="$PFX/lib/cmake/ngsolve" -DNetgen_DIR="$PFX/lib/cmake/netgen" \
      -DCMAKE_BUILD_TYPE=Release -DCMAKE_CUDA_ARCHITECTURES=100 -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc \
      -DCMAKE_CUDA_STANDARD=20 -DCMAKE_CXX_STANDARD=20 \
      "-DCMAKE_CXX_FLAGS=-I$PYBIND_INC -I$PY_INC -march=x86-64-v3" "-DCMAKE_CUDA_FLAGS=-I$PYBIND_INC -I$PY_INC -Xcompiler -march=x86-64-v3" \
      "-DCMAKE_SHARED_LINKER_FLAGS=-L$BUILD/dummy" "-DCMAKE_MODULE_LINKER_FLAGS=-L$BUILD/dummy" -DCMAKE_INSTALL_PREFIX="$PFX"
ninja -C "$BUILD" -j"$JOBS"
# the stand-alone project has no install rule for the library: place the two products by hand
cp "$BUILD"/libngscudalib_local.so "$PFX/lib/"
cp "$BUILD"/_ngscuda*.so "$PFX/lib/python3.12/site-packages/ngsolve/"
echo "ngscuda installed under $PFX"
