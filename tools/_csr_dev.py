import ctypes as C, json, os, statistics, sys, torch
sys.path.insert(0, os.getcwd())
import simspread_b200 as ss
from simspread_b200._lib import check
ctx = ss.Context(0); L = ss.lib(); dev = torch.device("cuda:0")
ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
rows, cols = 100_000, 20_000
ld = (rows + 15)//16*16
buf = torch.empty((cols, ld), dtype=torch.float64, device=dev)
for c0 in range(0, cols, 1000):
    buf[c0:c0+1000].copy_(torch.round(torch.rand((min(1000, cols-c0), ld), device=dev, dtype=torch.float64)*1e6)/1e6)
torch.cuda.synchronize()
mS = ss.DMat.wrap(ctx, buf.data_ptr(), rows, cols, ld)
def csr(alpha, w):
    h = C.c_void_p(); check(L.ss_featurize_csr(ctx.h, mS.h, alpha, w, C.byref(h))); L.ss_csr_destroy(h)
def timed(fn, reps=8):
    for _ in range(3): fn()
    ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(ext); fn(); e1.record(ext); ctx.sync(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
out={}
for name,(a,w) in {"weighted_4pct":(0.96,1),"binary_4pct":(0.96,0),"weighted_1pct":(0.99,1),"weighted_9pct":(0.91,1)}.items():
    out[name]=timed(lambda: csr(a,w))
print(json.dumps(out))
