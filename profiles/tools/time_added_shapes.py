import sys; sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/falcon-ttdforgnns_b200")
import torch, _ttg, tt_embeddings as te
dev="cuda:0"
for name,p,q,N,nnz in [("q844",[55,55,56],[8,4,4],169343,60000),("q545",[125,140,140],[5,4,5],2449029,262144)]:
    rr=[1,16,16,1]; D=q[0]*q[1]*q[2]
    g=torch.Generator().manual_seed(1)
    cores=[(torch.randn(1,p[t],rr[t]*q[t]*rr[t+1],generator=g)/N**0.25).to(dev) for t in range(3)]
    idx=torch.randperm(N,generator=g)[:nnz].to(dev); row=torch.arange(nnz,device=dev); tb=torch.zeros_like(row); dO=(torch.rand(1,nnz,D,generator=g)*0.1).to(dev)
    for fl in (0,16,1):
        te.EXTRA_FLAGS=fl
        for i in range(3):
            te.tt_forward(1000,1,nnz,D,p,q,rr,None,nnz,idx,row,tb,cores); te.tt_dense_backward(1000,D,p,q,rr,None,nnz,idx,row,tb,dO,cores)
        torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
        for i in range(10):
            te.tt_forward(1000,1,nnz,D,p,q,rr,None,nnz,idx,row,tb,cores); te.tt_dense_backward(1000,D,p,q,rr,None,nnz,idx,row,tb,dO,cores)
        e1.record(); torch.cuda.synchronize(); print(name,"flags",fl,"us/step",round(e0.elapsed_time(e1)*100,1))
