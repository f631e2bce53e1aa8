#!/bin/bash
# Programmatic-dependent-launch check: GPU tests, the B=128 step with and without PDL, a short benchmark run with and without.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -8
for NP in "" 1; do
  echo "== DX_NO_PDL=$NP"
  ${NP:+env DX_NO_PDL=1} python tools/host_bound.py 128 3xtf32 2>&1 | tail -2
  ${NP:+env DX_NO_PDL=1} python tools/small_batch_probe.py 128 3xtf32 2>&1 | tail -2
  ${NP:+env DX_NO_PDL=1} timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('bench value %.0f ms/step %.2f e2e %.0f' % (j['value'], j['ms_per_step'], j['e2e']['value']))"
done 2>&1 | tee gpurun_out/pdl_iter.log
