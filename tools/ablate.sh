#!/bin/bash
for d in 0 1 2 4 3 5 6 7; do
  echo "== TVC_TC_DBG=$d"
  TVC_TC_DBG=$d python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['roofline']['per_kernel_ms_per_step'];print(d['ms_per_step'], {n:k[n] for n in ('tc_up4_c1','tc_up4_c2','tc_up4_c5','tc_up0_c1','tc_up2_c2','tc_down0','tc_idft')})"
done
