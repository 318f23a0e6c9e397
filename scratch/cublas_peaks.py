"""cuBLAS reference rates next to MEASURED_PEAKS.json: fp32 (no tf32), tf32, fp16 and bf16 at 8192^3, burst (best of 10)
and sustained (back to back for ~2 s).  Prints one JSON object."""
import json, time, torch
def rate(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(2000 / best))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): a @ b
    e1.record(); e1.synchronize()
    sus = e0.elapsed_time(e1) / reps
    f = 2.0 * n ** 3 / 1e12
    return {"burst_tflops": round(f / (best * 1e-3), 1), "sustained_tflops": round(f / (sus * 1e-3), 1)}
out = {"gpu": torch.cuda.get_device_name(0), "how": "torch.matmul 8192^3, CUDA events",
       "cublas_fp32": rate(torch.float32, False), "cublas_tf32": rate(torch.float32, True),
       "cublas_fp16": rate(torch.float16, False), "cublas_bf16": rate(torch.bfloat16, False)}
print(json.dumps(out))
