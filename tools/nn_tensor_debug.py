import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from oracle import lmpcr_oracle as O, nn_c
import synthdata
from util import cabi, cu
for (n,m,seed) in [(128,256,1),(300,700,11),(1000,1000,12)]:
    feats,_,_=synthdata.synth_scene(2,max(n,m),seed=seed)
    fs,ft=feats[0,:n],feats[1,:m]
    jobs=torch.zeros((1,2),dtype=torch.int32,device='cuda')
    idx,dist,sc,amin=cabi.nn_tensor_debug(cu(fs[None]),cu(ft[None]),jobs)
    torch.cuda.synchronize()
    sc=sc[0].cpu().numpy()[:, :m]
    ah=fs.astype(np.float16).astype(np.float64); bh=ft.astype(np.float16).astype(np.float64)
    dn=nn_c.sqnorm(ft).astype(np.float64)
    exp=dn[None,:]-2*ah@bh.T
    err=np.abs(sc-exp)
    print(n,m,"score max err",err.max(),"mean",err.mean(), "sc sample", sc[0,:4], "exp", exp[0,:4])
    ri,rd=nn_c.nn_argmin(fs,ft)
    print("   idx equal", np.array_equal(idx[0].cpu().numpy(),ri), "mismatch", (idx[0].cpu().numpy()!=ri).sum(), "dist equal", np.array_equal(dist[0].cpu().numpy(),rd))
    print("   approx min vs exp min", np.abs(amin[0].cpu().numpy()-exp.min(1)).max())
