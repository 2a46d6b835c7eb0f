for fl in 0 8; do for f in 0 1 2 4 6; do TTG_DBG_BWD=$f timeout 120 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --no-graph --flags $fl > /tmp/o.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/o.json')); print('flags=$fl BWD dbg=$f bwd_rows_us=%.1f' % (1000*d['kernels_ms']['sorted_bwd_rows_kernel']))"; done; done
