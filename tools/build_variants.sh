#!/bin/bash
# Builds differently configured copies of the library into vkradixsort_b200/lib/variants/ (tuning runs; VKRS_LIB_PATH picks one).
#   tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."
set -e
cd "$(dirname "$0")/../vkradixsort_b200/csrc"
mkdir -p ../lib/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ \
      -shared -cudart static $flags vkrs_api.cu -o ../lib/variants/$name.so &
done
wait
ls -la ../lib/variants/
