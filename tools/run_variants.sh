#!/bin/bash
# GPU box: bucket_probe (quick mode) for each variant library named on the command line; one JSONL per variant in gpurun_out/.
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  VKRS_LIB_PATH=$PWD/vkradixsort_b200/lib/variants/$v.so PROBE_QUICK=1 timeout 300 python tools/bucket_probe.py 1e8 10 > gpurun_out/probe_$v.jsonl 2> gpurun_out/probe_$v.err
  echo "rc=$?"; grep -h '"kind": "timing"' gpurun_out/probe_$v.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['dist'], d['n'], d['ok'], d['ms_median'], d['gkeys_s'], {k: v for k, v in d['kernels_us'].items() if v > 20})"
  grep -h '"kind": "mismatch"\|"kind": "error"\|"kind": "correctness"' gpurun_out/probe_$v.jsonl | cut -c1-300
  tail -3 gpurun_out/probe_$v.err
done
