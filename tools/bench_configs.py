"""Wall-clock of the other BASELINE configs through the public host API (they are parity-test cases,
not the bench line): C2 = Enzyme-shaped 10-fold CV, C3 = 21-point alpha sweep on 5k queries x 2k
targets.  The reference's literal CPU path is timed beside them by `bench.py --impl reference
--reference-workload C2|C3` (the only place outside tests/ that runs the oracle).  Writes
gpurun_out/configs.json; run under gpurun."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess

import simspread_b200 as ss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_arm(workload):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reference-workload", workload],
                       capture_output=True, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])

out = {}
rng = np.random.default_rng(20241)
N, Nt = 445, 664
S = np.round(rng.beta(2, 5, size=(N, N)), 6)
np.fill_diagonal(S, 1.0)
Y = (rng.random((N, Nt)) < 0.0099).astype(float)
names = [f"D{i:04d}" for i in range(N)]
tn = [f"T{j:04d}" for j in range(Nt)]
DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Y, (names, tn))
ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=1)  # warm-up (context, workspaces)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    res = ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=1)
    ts.append(time.perf_counter() - t0)
t_gpu = float(np.median(ts))
ref2 = reference_arm("C2")
t_cpu = ref2["cpu_baseline"]["seconds_per_step"]
out["C2_enzyme_10fold_cv"] = {"shape": [N, Nt], "alpha": 0.35, "weighted": False, "gpu_wall_s_median_of_5": t_gpu,
                              "scores_per_s": N * Nt / t_gpu, "cpu_oracle_literal_wall_s": t_cpu,
                              "cpu_cores": os.cpu_count(), "AuROC": res["AuROC"], "AuPRC": res["AuPRC"],
                              "includes": "upload, featurize, 10 x (gather, predict+clean), AuROC/AuPRC, recall/precision@20, download"}
print(json.dumps(out["C2_enzyme_10fold_cv"]), flush=True)

nq, ns, nt = 5000, 5000, 2000
Nn = nq + ns
S3 = np.round(rng.random((Nn, Nn)), 6)
Y3 = (rng.random((Nn, nt)) < 0.01).astype(float)
n3 = [f"n{i}" for i in range(Nn)]
t3 = [f"t{j}" for j in range(nt)]
DD3, DT3 = ss.NamedArray(S3, (n3, n3)), ss.NamedArray(Y3, (n3, t3))
alphas = [round(0.05 * i, 2) for i in range(21)]
ss.alpha_sweep(DT3, DD3, n3[:nq], alphas[:2])
t0 = time.perf_counter()
sw = ss.alpha_sweep(DT3, DD3, n3[:nq], alphas)
t_sw = time.perf_counter() - t0
out["C3_alpha_sweep_21_points"] = {"shape": {"nq": nq, "ns": ns, "nt": nt}, "gpu_wall_s": t_sw,
                                   "scores_per_s": 21 * nq * nt / t_sw, "points": sw,
                                   "note": "dense DMMA chain for every alpha (alpha_sweep gathers dense blocks); "
                                           "the sparse chain is selected by predict(layout='auto')"}
print(json.dumps({k: v for k, v in out["C3_alpha_sweep_21_points"].items() if k != "points"}), flush=True)

# the reference's own formulation beside it -- (i) the literal CPU path (dense n x n construct ->
# spread -> A*(W*W), NumPy/OpenBLAS restatement, all host cores) for ONE alpha of C3 (the cost of the dense
# DGEMMs does not depend on alpha); (ii) torch.matmul in Float32 on the GPU for the padded n x n problem, the
# proxy for the reference's `GPU=true` cuBLAS SGEMM path (src/core.jl:404-419), at the C2 and C3 sizes.
if os.environ.get("SS_SKIP_REFERENCE_FORMS") != "1":
    import torch
    ref3 = reference_arm("C3")
    out["C3_cpu_literal_one_alpha"] = {"alpha": 0.5, "wall_s": ref3["cpu_baseline"]["seconds_per_step"],
                                       "scores_per_s": ref3["value"], "cores": ref3["cpu_baseline"]["cores"],
                                       "sample": ref3["cpu_baseline"]["sample"],
                                       "kind": "NumPy/OpenBLAS restatement of the reference CPU path (not Julia)"}
    print(json.dumps(out["C3_cpu_literal_one_alpha"]), flush=True)
    proxy = {}
    for label, n_ in (("C2_fold_n1510", 1510), ("C3_n17000", 17000)):
        A32 = torch.rand((n_, n_), dtype=torch.float32, device="cuda")
        W32 = torch.rand((n_, n_), dtype=torch.float32, device="cuda")
        torch.backends.cuda.matmul.allow_tf32 = False
        for _ in range(2):
            F32 = A32 @ (W32 @ W32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            F32 = A32 @ (W32 @ W32)
        e1.record()
        torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 3
        proxy[label] = {"n": n_, "ms": ms_, "tflops_fp32": 4 * n_ ** 3 / (ms_ * 1e-3) / 1e12}
        del A32, W32, F32
    out["reference_gpu_true_proxy_torch_fp32_dense_nxn"] = proxy
    print(json.dumps(proxy), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
