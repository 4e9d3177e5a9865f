#!/usr/bin/env bash
# Builds the UNMODIFIED reference (cg-saarland/hagrid) for sm_100a from the
# sources where they lie under /root/reference; outputs go only to oracle/_ref/
# (git-ignored, but shipped to the GPU box by gpurun). Test oracle + reference
# bench arm only: nothing in the product links or loads these files.
#
#   oracle/_ref/libhagrid_ref.so  include/hagrid_b200.h ABI over the reference
#                                 (hagrid_b200/csrc/c_api.cpp, -DHGB_REFERENCE_BUILD)
#   oracle/_ref/hagrid_ref        the reference's own main.cpp executable
#
# Semantics-neutral compile-compat edits applied to a scratch copy:
#   parallel.cuh:13  private: -> public:   (nvcc host stubs name ResultType)
#   build.cu:508     device lambda gets `-> int`, build.cu:727 gets `-> BBox`
#   (CUB 2.8 cannot query return types of un-annotated extended lambdas)
# The vendored CUB 1.8.0 does not compile with CUDA 12.9; the toolkit's CUB is used.
# A second traverse object is built from a copy with `hit.id = steps;`
# (traverse.cu:93) deleted and its kernel + two entry points renamed *_pid
# (template kernels have vague linkage: without the rename the linker would
# fold both instantiations into one), so prim-id
# parity can be checked against the reference's own arithmetic.
set -euo pipefail
REF=${HAGRID_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
OUT=$HERE/_ref
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
[ -d "$REF/src" ] || { echo "reference sources not found at $REF (prebuilt oracle/_ref is used as is)"; exit 0; }
mkdir -p "$OUT"
W=$(mktemp -d)
trap 'rm -rf "$W"' EXIT
cp "$REF"/src/*.cu "$REF"/src/*.h "$REF"/src/*.cuh "$REF"/src/*.cpp "$W"/
cd "$W"
sed -i '13s/private:/public:/' parallel.cuh
sed -i '508s/\[\] __device__ (int a, int b) {/[] __device__ (int a, int b) -> int {/' build.cu
sed -i '727s/\[\] __device__ (BBox a, const BBox\& b) {/[] __device__ (BBox a, const BBox\& b) -> BBox {/' build.cu
grep -q -- '-> int' build.cu && grep -q -- '-> BBox' build.cu && grep -q 'public:' parallel.cuh
sed -e '/hit.id = steps;/d' -e 's/void setup_traversal(/void setup_traversal_pid(/' \
    -e 's/void traverse_grid(/void traverse_grid_pid(/' \
    -e 's/__global__ void traverse(/__global__ void traverse_pid(/' -e 's/traverse<<</traverse_pid<<</g' \
    traverse.cu > traverse_pid.cu
if grep -q 'hit.id = steps' traverse_pid.cu; then echo "pid patch failed"; exit 1; fi
[ "$(grep -c 'traverse_pid<<<' traverse_pid.cu)" = 2 ] || { echo "kernel rename failed"; exit 1; }

NVFLAGS="-std=c++17 --expt-extended-lambda -lineinfo --use_fast_math -O3 -DNDEBUG -DHOST=__host__ -DDEVICE=__device__ \
 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wno-deprecated-declarations,-fvisibility=hidden -w"
for f in build merge flatten expand compress mem_manager profile; do
  $NVCC $NVFLAGS -c $f.cu -o $f.o &
done
$NVCC $NVFLAGS --maxrregcount=40 -c traverse.cu -o traverse.o &
$NVCC $NVFLAGS --maxrregcount=40 -c traverse_pid.cu -o traverse_pid.o &
wait
CXXFLAGS="-O2 -DNDEBUG -DHOST= -DDEVICE= -fPIC -w -I/usr/local/cuda/include"
g++ -std=c++11 $CXXFLAGS -I"$W" -I"$ROOT/include" -DHGB_REFERENCE_BUILD -fvisibility=hidden \
    -c "$ROOT/hagrid_b200/csrc/c_api.cpp" -o c_api.o
# the reference's front-end functions (gen_camera, gen_rays, update_surface): src/main.cpp included unmodified
g++ -std=c++11 $CXXFLAGS -I"$HERE/sdl_stub" -I"$W" -fvisibility=hidden -c "$HERE/ref_frontend.cpp" -o ref_frontend.o
g++ -std=c++11 $CXXFLAGS -I"$W" -fvisibility=hidden -c load_obj.cpp -o load_obj_pic.o
# c_api.o needs default visibility for the hagrid:: symbols it imports from the objects above
g++ -shared -o "$OUT/libhagrid_ref.so" c_api.o ref_frontend.o load_obj_pic.o build.o merge.o flatten.o expand.o compress.o mem_manager.o profile.o \
    traverse.o traverse_pid.o -Wl,-Bsymbolic -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
g++ -std=c++11 $CXXFLAGS -I"$HERE/sdl_stub" -I"$W" -c main.cpp -o main.o
g++ -std=c++11 $CXXFLAGS -I"$W" -c load_obj.cpp -o load_obj.o
g++ -o "$OUT/hagrid_ref" main.o load_obj.o build.o merge.o flatten.o expand.o compress.o mem_manager.o profile.o traverse.o \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
# Drop-in check: the reference's UNMODIFIED front end (main.cpp, load_obj.cpp/.h) compiled against
# THIS repository's headers (hagrid_b200/include/hagrid) and linked with THIS repository's kernels.
# The three files are compiled from a scratch directory that holds nothing else, so every
# "build.h"/"grid.h"/... they include resolves to our headers, not the reference's.
OURS="$ROOT/hagrid_b200"
if ls "$OURS"/_build/ray_traverse.o >/dev/null 2>&1; then
  D=$(mktemp -d)
  cp "$REF"/src/main.cpp "$REF"/src/load_obj.cpp "$REF"/src/load_obj.h "$D"/
  (cd "$D" && g++ -std=c++11 $CXXFLAGS -I"$HERE/sdl_stub" -I"$OURS/include/hagrid" -c main.cpp -o main.o \
           && g++ -std=c++11 $CXXFLAGS -I"$OURS/include/hagrid" -c load_obj.cpp -o load_obj.o)
  g++ -o "$OUT/hagrid_dropin" "$D"/main.o "$D"/load_obj.o $(ls "$OURS"/_build/*.o | grep -v c_api.o) \
      -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
  rm -rf "$D"
fi
# SASS of the reference traversal, for expression-shape comparison (not shipped in git)
/usr/local/cuda/bin/cuobjdump -sass traverse_pid.o > "$OUT/traverse_pid.sass" 2>/dev/null || true
echo "built: $(ls "$OUT")"
