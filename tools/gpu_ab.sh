#!/bin/bash
for d in tmp_old tmp_9f2ad37 tmp_80d77f0 .; do
  (cd $d && timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --chunk 512 2>/dev/null | D="$d" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernels']
print('[%s] value %d  sum %.3f ms | '%(os.environ['D'],d['value'],sum(k['ms'] for k in ks)) + ' '.join('%.0f'%(k['ms']*1e3) for k in ks))
")
done
