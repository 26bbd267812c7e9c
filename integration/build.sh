#!/bin/sh
# integration/build.sh -- proves that the CUDA plugins drop into the reference: copies the reference tree to a scratch
# directory, adds src/kernels/cuda-spmv.{hpp,cpp}, applies reference.patch (kernels.hpp, main.cpp, Makefile,
# util/perf-events.cpp), builds the stock binary with `make USE_CUDA=1` (recipe of SURVEY 8c: NO_LIBPFM, NO_LIBNUMA,
# three forced includes for GCC 13) and links the harness against the same archives.
#
# Only BINARIES land in the repository tree (integration/_build/, git-ignored; they travel to the GPU box like
# oracle/_ref): no reference source is copied into the repository.
#
#   integration/build.sh [REFERENCE_DIR]        default /root/reference
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(dirname "$HERE")
REF=${1:-/root/reference}
OUT="$HERE/_build"
LIBDIR="$REPO/spmv_cache_trace_b200/lib"
if [ ! -d "$REF/src/kernels" ]; then
    echo "integration/build.sh: $REF not present (GPU box?) -- keeping any prebuilt binaries in $OUT"
    exit 0
fi
if [ ! -f "$LIBDIR/libspmvb200.so" ]; then
    echo "integration/build.sh: build libspmvb200.so first (python -m spmv_cache_trace_b200.build)" >&2
    exit 1
fi
WORK=$(mktemp -d /tmp/spmv-integration.XXXXXX)
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF" "$WORK/tree"
chmod -R u+w "$WORK/tree"
cp "$HERE/src/kernels/cuda-spmv.hpp" "$HERE/src/kernels/cuda-spmv.cpp" "$WORK/tree/src/kernels/"
(cd "$WORK/tree" && patch -s -p1 < "$HERE/reference.patch")
CXX_CMD="/usr/bin/g++ -include cstdint -include cstring -include mutex"
# the binaries find the library relative to themselves, wherever the repository is unpacked
RPATH="-Wl,-rpath,'\$\$ORIGIN/../../spmv_cache_trace_b200/lib'"
make -s -C "$WORK/tree" -j8 USE_CUDA=1 NO_LIBPFM=1 NO_LIBNUMA=1 CC=/usr/bin/gcc CXX="$CXX_CMD" \
    SPMVB200="$REPO" SPMVB200_LIBDIR="$LIBDIR" LDFLAGS="-lz -L$LIBDIR -lspmvb200 $RPATH" spmv-cache-trace
mkdir -p "$OUT"
cp "$WORK/tree/spmv-cache-trace" "$OUT/spmv-cache-trace"
T="$WORK/tree"
$CXX_CMD -std=c++14 -O2 -fopenmp -DUSE_OPENMP -DUSE_POSIX_MEMALIGN -DUSE_CUDA -I"$T/src" -I"$REPO/include" \
    "$HERE/harness.cpp" "$T/src/trace-config.o" "$T/src/cache-simulation/kernels.a" "$T/src/cache-simulation/cache-simulation.a" \
    "$T/src/matrix/matrix.a" "$T/src/util/util.a" -lz -L"$LIBDIR" -lspmvb200 '-Wl,-rpath,$ORIGIN/../../spmv_cache_trace_b200/lib' \
    -o "$OUT/harness"
echo "built $OUT/spmv-cache-trace and $OUT/harness"
