#!/bin/bash
# A/B the library variants under variants/ on one box: flow fwd/bwd times from the phase timeline + bench step time
for v in "$@"; do
  export DPI_B200_LIB=$PWD/variants/lib_$v.so
  echo "== variant $v"
  timeout 120 python bench.py --steps 40 --warmup 5 --cpu-steps 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('step ms', round(d['ms_per_step'],4), 'fwd', k['flow_reverse_fwd']['ms'], 'bwd', k['flow_reverse_bwd']['ms'])"
done
