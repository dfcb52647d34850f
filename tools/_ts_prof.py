import os, sys, torch
sys.path.insert(0, os.getcwd())
import simspread_b200 as ss
from simspread_b200._lib import check, SS_PREDICT_CLEAN
ctx = ss.Context(0); L = ss.lib(); dev = torch.device("cuda:0")
def cm(rows, cols, fill):
    ld=(rows+15)//16*16
    b=torch.empty((cols,ld),dtype=torch.float64,device=dev)
    for c0 in range(0,cols,2000): fill(b[c0:c0+2000])
    torch.cuda.synchronize()
    return b, ss.DMat.wrap(ctx,b.data_ptr(),rows,cols,ld)
uni=lambda b: b.copy_(torch.round(torch.rand(b.shape,device=dev,dtype=torch.float64)*1e6)/1e6)
bern=lambda b: b.copy_((torch.rand(b.shape,device=dev)<0.05).to(torch.float64))
ns,nf,nt,nq=20000,20000,6400,128
bXs,mXs=cm(ns,nf,uni); bY,mY=cm(ns,nt,bern); bXq,mXq=cm(nq,nf,uni); bR,mR=cm(nq,nt,lambda b:b.zero_())
os.environ["SS_T_FORM"]="sparse"
for _ in range(2):
    check(L.ss_predict_query(ctx.h,mXq.h,mXs.h,mY.h,mR.h,SS_PREDICT_CLEAN,None))
ctx.sync()
