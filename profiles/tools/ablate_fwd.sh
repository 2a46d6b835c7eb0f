for fl in 0 8; do for f in 0 8; do TTG_DBG_FWD=$f timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --no-graph --flags $fl > /tmp/o.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/o.json')); print('flags=$fl FWD dbg=$f fwd_us=%.1f step_us=%.1f' % (1000*d['kernels_ms']['fwd_rows_kernel'], 1000*d['ms_per_step']))"; done; done
