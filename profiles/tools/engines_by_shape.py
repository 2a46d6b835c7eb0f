import sys; sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/falcon-ttdforgnns_b200")
import torch, tt_embeddings as te
dev="cuda:0"
for name,p,q,N,sizes in [("arxiv",[55,55,56],[4,4,8],169343,[40000,80000,169343]),("cora",[14,14,14],[4,4,8],2708,[2708])]:
    rr=[1,16,16,1]; D=q[0]*q[1]*q[2]
    g=torch.Generator().manual_seed(1)
    cores=[(torch.randn(1,p[t],rr[t]*q[t]*rr[t+1],generator=g)/N**0.25).to(dev) for t in range(3)]
    for nnz in sizes:
        idx=torch.randperm(N,generator=g)[:nnz].to(dev); row=torch.arange(nnz,device=dev); tb=torch.zeros_like(row); dO=(torch.rand(1,nnz,D,generator=g)*0.1).to(dev)
        res=[]
        for fl in (32,1024):
            te.EXTRA_FLAGS=fl
            for i in range(3):
                te.tt_forward(1000,1,nnz,D,p,q,rr,None,nnz,idx,row,tb,cores); te.tt_sgd_backward(1000,D,0.0,p,q,rr,None,nnz,idx,row,tb,dO,cores)
            torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
            for i in range(20):
                te.tt_forward(1000,1,nnz,D,p,q,rr,None,nnz,idx,row,tb,cores); te.tt_sgd_backward(1000,D,0.0,p,q,rr,None,nnz,idx,row,tb,dO,cores)
            e1.record(); torch.cuda.synchronize(); res.append(round(e0.elapsed_time(e1)*50,1))
        te.EXTRA_FLAGS=0
        print(name,"rows",nnz,"rows/group(R)",round(nnz/(p[1]*p[2]),1),"left us",res[0],"right us",res[1])
