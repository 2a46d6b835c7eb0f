import os, sys, time, torch
ROOT='/root/repo'
sys.path[:0]=[ROOT, ROOT+'/falcon-ttdforgnns_b200']
from FBTT.tt_embeddings_ops import TTEmbeddingBag, OptimType
dev=torch.device('cuda',0)
m=TTEmbeddingBag(2449029,100,[16,16],[125,140,140],[4,5,5],optimizer=OptimType.SGD,sparse=True,use_cache=False,weight_dist='normal')
N=2449029
def t(fn,n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
idx=torch.arange(N,device=dev); off=torch.arange(N+1,device=dev)
with torch.no_grad():
    a=t(lambda: m.rows_range(0,N)); b=t(lambda: m(idx,off))
print('rows_range(0,N): %.3f ms (%.0f GB/s of output)   forward(arange): %.3f ms'%(a, N*400/a/1e6, b))
