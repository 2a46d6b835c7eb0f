import sys; sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/falcon-ttdforgnns_b200")
import torch, numpy as np, tt_embeddings as te
from FBTT.tt_embeddings_ops import TTEmbeddingBag
dev="cuda:0"; N=2449029; nnz=1310000
m=TTEmbeddingBag(N,100,[16,16],[125,140,140],[4,5,5],sparse=True,use_cache=True,cache_size=N//10,hashtbl_size=4*(N//10),weight_dist="normal").to(dev)
g=torch.Generator().manual_seed(0)
idx=torch.randint(0,N,(nnz,),generator=g).to(dev); off=torch.arange(nnz+1,device=dev)
for _ in range(3): m(idx,off)
m.cache_populate()
def t(f,n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
print("update_cache us", t(lambda: m.update_cache(idx)))
print("rowidx (preprocess warm) us", t(lambda: te.preprocess_indices_sync(idx,off,1,True,m.hashtbl,m.cache_state)))
print("cache_mark us", t(lambda: te.cache_mark(idx,m.hashtbl,m.cache_state)))
tt,loc=te.cache_mark(idx,m.hashtbl,m.cache_state); row=torch.arange(nnz,device=dev)
out=torch.zeros(nnz,100,device=dev)
print("cached frac", float((loc>=0).float().mean()))
print("cache_forward us", t(lambda: te.cache_forward(nnz,nnz,loc,row,m.cache_weight,out)))
print("cache_backward_sgd us", t(lambda: te.cache_backward_sgd(nnz,out,loc,row,0.0,m.cache_weight)))
print("preprocess with sync us", t(lambda: te.preprocess_indices_sync(idx,off,1,False,m.hashtbl,m.cache_state)))
