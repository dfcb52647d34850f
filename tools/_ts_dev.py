import json, os, sys, time, torch
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
dens=float(os.environ.get("YD","0.05"))
bern=lambda b: b.copy_((torch.rand(b.shape,device=dev)<dens).to(torch.float64))
ns,nf,nt,nq=20000,20000,50000,256
bXs,mXs=cm(ns,nf,uni); bY,mY=cm(ns,nt,bern); bXq,mXq=cm(nq,nf,uni)
out={}; Rs={}
for form in ("dense","sparse"):
    os.environ["SS_T_FORM"]=form
    bR,mR=cm(nq,nt,lambda b:b.zero_())
    for rep in range(2):
        ctx.profile(True)
        t0=time.perf_counter()
        check(L.ss_predict_query(ctx.h,mXq.h,mXs.h,mY.h,mR.h,SS_PREDICT_CLEAN,None))
        wall=time.perf_counter()-t0
        prof=ctx.profile_read(); ctx.profile(False)
    out[form]={"wall_ms":wall*1e3,"records":[(round(ms,2),fl) for ms,fl in prof]}
    Rs[form]=bR[:, :nq].clone()
os.environ.pop("SS_T_FORM")
a,b=Rs["dense"],Rs["sparse"]
m=(a!=-99)&(a!=0)
out["max_rel_diff"]=float(((a-b).abs()[m]/a.abs()[m]).max())
out["flags_equal"]=bool(torch.equal(a==-99,b==-99))
print(json.dumps(out))
