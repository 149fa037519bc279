#!/bin/bash
# A/B of variant libraries in ONE process each, interleaved twice: per-kernel times of the bucket schedule at 1e8 uniform keys
for rep in 1 2; do for v in "$@"; do
  VKRS_LIB_PATH=$PWD/vkradixsort_b200/lib/variants/$v.so PROBE_QUICK=1 PROBE_SKIP_CORRECTNESS=1 PROBE_ONLY_UNIFORM=1 python tools/bucket_probe.py 1e8 10 2>/dev/null | grep '"kind": "timing"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$v', d['dist'], d['ok'], d['ms_median'], {k: v for k, v in d['kernels_us'].items() if v > 20})"
done; done
