#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "2 0" "2 1"; do
  set -- $cfg
  PMB_READOUT_VARIANT=$1 PMB_PAINT_PREFETCH=$2 timeout 300 python bench.py --nmesh 1024 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ab_$1_$2.log 2>&1
  echo "READOUT_VARIANT=$1 PAINT_PREFETCH=$2: $(python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('step', d['value'], 'paint', d['roofline']['paint_ms'], 'readout', d['roofline']['readout_ms'])
" gpurun_out/ab_$1_$2.log)"
done
