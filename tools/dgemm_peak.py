"""cuBLAS DGEMM peak (FP64 roofline denominator): torch.matmul float64 8192^3, best of 10 + sustained."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best / 1e9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = 60
e0.record()
for _ in range(reps):
    torch.matmul(a, b, out=c)
e1.record(); torch.cuda.synchronize()
sust = 2 * n**3 * reps / e0.elapsed_time(e1) / 1e9
print(json.dumps({"fp64_tflops": burst, "fp64_tflops_sustained": sust, "n": n,
                  "how": "torch.matmul float64 8192^3 (2*N^3): best of 10 (burst) and 60 back to back (sustained)"}))
