#!/usr/bin/env bash
# oracle/build_ref.sh — compile the REFERENCE's own CPU implementation of the two hot paths,
# straight from the sources where they lie under /root/reference, into oracle/_ref/libpyg_ref.so.
#
# TEST INFRASTRUCTURE ONLY.  Output goes to oracle/_ref/ (git-ignored, NOT gpurun-ignored so the
# prebuilt .so travels to the GPU box).  No reference source is copied into this repo; the only
# file we supply is a 2-line stand-in for the cmake-generated pyg_lib/csrc/config.h
# (config.h.in: WITH_MKL_BLAS()/NO_METIS() both 0 == the reference's default build).
#
# The reference's own build system (cmake + METIS + CUTLASS ...) is NOT run: g++ on the 14 files
# below is enough for pyg::neighbor_sample, pyg::hetero_neighbor_sample, pyg::subgraph, pyg::relabel_neighborhood, pyg::merge_sampler_outputs, pyg::segment_matmul,
# pyg::grouped_matmul (CPU + Autograd keys).
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
[ -d "$REF/pyg_lib/csrc" ] || { echo "no reference tree at $REF; keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT/stub/pyg_lib/csrc" "$OUT/obj"
cat > "$OUT/stub/pyg_lib/csrc/config.h" <<'EOF'
#pragma once
#define WITH_MKL_BLAS() 0
#define NO_METIS() 1
EOF
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
ABI=$($PY -c 'import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')
CXXFLAGS="-O3 -fPIC -std=c++17 -fopenmp -D_GLIBCXX_USE_CXX11_ABI=$ABI -w \
  -I$OUT/stub -I$REF -I$REF/third_party/parallel-hashmap \
  -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC"
SRCS="pyg_lib/csrc/library.cpp
pyg_lib/csrc/utils/check.cpp
pyg_lib/csrc/utils/convert.cpp
pyg_lib/csrc/sampler/neighbor.cpp
pyg_lib/csrc/sampler/cpu/neighbor_kernel.cpp
pyg_lib/csrc/sampler/dist_merge_outputs.cpp
pyg_lib/csrc/sampler/cpu/dist_merge_outputs_kernel.cpp
pyg_lib/csrc/sampler/dist_relabel.cpp
pyg_lib/csrc/sampler/cpu/dist_relabel_kernel.cpp
pyg_lib/csrc/sampler/subgraph.cpp
pyg_lib/csrc/sampler/cpu/subgraph_kernel.cpp
pyg_lib/csrc/ops/matmul.cpp
pyg_lib/csrc/ops/cpu/matmul_kernel.cpp
pyg_lib/csrc/ops/autograd/matmul_kernel.cpp"
OBJS=""
pids=""
for s in $SRCS; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_').o"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    ( g++ $CXXFLAGS -c "$REF/$s" -o "$o" ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait "$p"; done
g++ -shared -fopenmp -o "$OUT/libpyg_ref.so" $OBJS \
  -L"$TORCH_DIR/lib" -ltorch -ltorch_cpu -lc10 -Wl,-rpath,"$TORCH_DIR/lib"
echo "built $OUT/libpyg_ref.so"
