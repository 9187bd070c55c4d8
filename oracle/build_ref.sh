#!/usr/bin/env bash
# TEST INFRASTRUCTURE — not product code.
#
# Compiles the reference's own four native translation units, unmodified and
# from where they lie under /root/reference/cuda_kernel/, into
#   oracle/_ref/grouped_cumprod_ref.so
# (a torch/pybind extension module named `grouped_cumprod_ref`; the module name
# is injected with -DTORCH_EXTENSION_NAME, exactly what torch's BuildExtension
# does for setup.py:9 of the reference).  No reference source is copied into
# this repository; only the built .so lands in oracle/_ref/ (git-ignored, NOT
# gpurun-ignored, so it travels to the GPU box).
#
# The reference's own build (setup.py) is unusable here: Windows back-slash
# source paths (setup.py:11-14) and os.path.join(None, ...) when CUDA_PATH is
# unset (setup.py:4,17).  This recipe is the equivalent nvcc command line.
#
# Used by tests/ (-m gpu) as a second parity checker beside the fp64 oracle,
# (tests/test_gpu_parity.py, tests/test_reference_function.py) and by bench.py's `ref_cuda_ops` /
# `reference_function` legs, which time the reference ops beside ours on the same box.
set -euo pipefail
REF=${REF:-/root/reference/cuda_kernel}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present (GPU box?) - keeping prebuilt files" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_LIB PY_INC EXT_SUFFIX CXX11ABI < <($PY - <<'EOF'
import sysconfig, torch, os
from torch.utils import cpp_extension as ce
inc = ":".join(ce.include_paths())
lib = os.path.join(os.path.dirname(torch.__file__), "lib")
print(inc, lib, sysconfig.get_paths()["include"], sysconfig.get_config_var("EXT_SUFFIX"),
      int(torch._C._GLIBCXX_USE_CXX11_ABI))
EOF
)
INCS=""
IFS=':' read -ra PARTS <<< "$TORCH_INC"
for p in "${PARTS[@]}"; do INCS="$INCS -isystem $p"; done
INCS="$INCS -isystem $PY_INC"
DEFS="-DTORCH_EXTENSION_NAME=grouped_cumprod_ref -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$CXX11ABI"
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC"
pids=()
for f in grouped_cumprod_forward grouped_cumprod_backward grouped_cumsum_forward; do
  nvcc $NVFLAGS $DEFS $INCS -c "$REF/$f.cu" -o "$OUT/obj/$f.o" &
  pids+=($!)
done
g++ -O2 -std=c++17 -fPIC $DEFS $INCS -c "$REF/cuda_kernel.cpp" -o "$OUT/obj/cuda_kernel.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$OUT/grouped_cumprod_ref.so" "$OUT"/obj/*.o \
  -L"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python \
  -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TORCH_LIB"
rm -rf "$OUT/obj"
echo "[build_ref] built $OUT/grouped_cumprod_ref.so"
