for f in 0 1 2 3 4 6 7; do TTG_DBG_FWD=$f timeout 120 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --no-graph > /tmp/o.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/o.json')); print('FWD dbg=$f fwd_us=%.1f' % (1000*d['kernels_ms']['sorted_fwd_kernel']))"; done
for f in 0 1 2 3 4 6 7 8 15; do TTG_DBG_BWD=$f timeout 120 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --no-graph > /tmp/o.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/o.json')); print('BWD dbg=$f bwd_rows_us=%.1f' % (1000*d['kernels_ms']['sorted_bwd_rows_kernel']))"; done
