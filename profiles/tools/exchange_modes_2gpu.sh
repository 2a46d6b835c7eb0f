python -m pytest tests/test_gpu_peer_exchange.py -x -q | tail -3
for sc in 0 1; do
  TTG_PEER_SCATTER=$sc python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$sc bench.py --gpus 2 --no-extra --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/x$sc.json
  python -c "import json; d=json.load(open('gpurun_out/x$sc.json')); print('scatter $sc', d['ms_per_step'], d['replicas_bit_identical'], d['exchange_failed'], d['kernels_ms'].get('optimizer_kernel'))"
done
