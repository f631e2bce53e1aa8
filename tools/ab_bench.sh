#!/bin/bash
# A/B/A/B of the default benchmark step under an environment switch: usage  bash tools/ab_bench.sh VAR [rounds]
mkdir -p gpurun_out
VAR=$1; R=${2:-3}
for i in $(seq 1 $R); do
  for ON in "" 1; do
    env ${ON:+$VAR=1} timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$VAR=$ON value %.0f ms/step %.2f e2e %.0f clocks %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks']['sm_mhz']))"
  done
done 2>&1 | tee gpurun_out/ab_$VAR.log
