# inference bench for every EMB_POOL_PARTS setting (epilogue warps sharing one tile in the pooled kernels; 0 = heuristic)
for pp in ${PARTS:-0 1 2 4}; do
  EMB_POOL_PARTS=$pp timeout 120 python bench.py --workload infer --cpu-baseline 0 2>/dev/null > /tmp/parts_$pp.json
  python -c "
import json,sys;d=json.loads(open('/tmp/parts_$pp.json').read().strip().splitlines()[-1]);print('parts',$pp,round(d['value']),d['ms_per_step'])"
done
