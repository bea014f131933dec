#!/bin/bash
# cluster-size sweep of the small-batch flow kernel: step time of C1 / C2 for every cluster size
for wl in c1 c2; do
  for cl in auto 8 4 2; do
    if [ $cl = auto ]; then unset DPI_B200_SMALL_CL; else export DPI_B200_SMALL_CL=$cl; fi
    DPI_B200_SMALL_VERBOSE=1 timeout 150 python bench.py --workload $wl --steps 30 --warmup 5 --cpu-steps 1 2> /tmp/err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$wl cl=$cl', 'step', round(d['ms_per_step'],4), 'fwd', k['flow_reverse_fwd']['ms'], 'bwd', k['flow_reverse_bwd']['ms'])"
    grep -m1 "small plan" /tmp/err.log
  done
done
