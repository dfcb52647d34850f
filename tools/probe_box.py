"""One-off probe of the GPU box (run under gpurun): host cores / memory, GPU properties, measured
cuBLAS FP64 GEMM rate and pinned PCIe bandwidth.  Writes gpurun_out/probe.json."""
import json
import os
import subprocess
import time

import torch

out = {"cpu_count": os.cpu_count()}
for l in open("/proc/meminfo"):
    if l.split(":")[0] in ("MemTotal", "MemAvailable"):
        out[l.split(":")[0]] = int(l.split()[1]) * 1024
try:
    out["lscpu"] = subprocess.run("lscpu | grep -E 'Model name|Socket|Core|Thread'", shell=True,
                                  capture_output=True, text=True).stdout
    out["nvidia_smi"] = subprocess.run(["nvidia-smi", "--query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit",
                                        "--format=csv"], capture_output=True, text=True).stdout
except Exception as e:
    out["err"] = str(e)
p = torch.cuda.get_device_properties(0)
out["gpu"] = {"name": p.name, "sms": p.multi_processor_count, "mem": p.total_memory, "cc": [p.major, p.minor]}
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192, 16384):
    a = torch.rand((n, n), dtype=torch.float64, device=dev)
    b = torch.rand((n, n), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    res[str(n)] = best
    del a, b
out["cublas_fp64_tflops"] = res
# sustained: 3 seconds back to back at 8192
n = 8192
a = torch.rand((n, n), dtype=torch.float64, device=dev); b = torch.rand((n, n), dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t0 = time.time(); it = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5):
        torch.matmul(a, b)
    it += 5
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["cublas_fp64_tflops_sustained"] = 2 * n ** 3 * it / (e0.elapsed_time(e1) * 1e-3) / 1e12
del a, b
# pinned PCIe bandwidth
h = torch.empty(1 << 28, dtype=torch.float64).pin_memory()  # 2 GiB
d = torch.empty_like(h, device=dev)
for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
    out[f"pcie_{name}_gbs"] = h.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
