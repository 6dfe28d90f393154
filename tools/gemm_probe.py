"""Policy-head GEMM variants (4096 x 528 x 11584 bf16) through the library: which call shape is fastest."""
import torch
import torch.nn.functional as F

dev = "cuda:0"
M, K, N = 4096, 528, 11584
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
b = torch.randn(N, device=dev, dtype=torch.bfloat16)
wt = w.t().contiguous()          # [K, N]
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:40s} median {ts[len(ts) // 2]:7.1f} us  min {ts[0]:7.1f} us  {2 * M * K * N / ts[len(ts) // 2] / 1e6:7.1f} TFLOP/s (full-size flops)", flush=True)


bench("F.linear(x, w, b)", lambda: F.linear(x, w, b))
bench("F.linear(x, w)", lambda: F.linear(x, w))
bench("addmm(b, x, wt) [K,N] weights", lambda: torch.addmm(b, x, wt))
bench("mm(x, wt)", lambda: torch.mm(x, wt))
bench("mm(x, wt, out=)", lambda: torch.mm(x, wt, out=out))
bench("mm(w, x.t()) -> [N, M]", lambda: torch.mm(w, x.t()))
for mm in (2048, 1024):
    xs = x[:mm]
    bench(f"F.linear M={mm}", lambda: F.linear(xs, w, b))
for nn in (5760, 2896):   # fewer planes
    ws, bs = w[:nn], b[:nn]
    bench(f"F.linear N={nn}", lambda: F.linear(x, ws, bs))
try:
    torch.backends.cuda.preferred_blas_library("cublas")
    bench("cublas: F.linear(x, w, b)", lambda: F.linear(x, w, b))
    bench("cublas: mm(x, wt)", lambda: torch.mm(x, wt))
except Exception as e:  # noqa: BLE001
    print("preferred_blas_library failed", e)
try:
    x8, w8 = x.to(torch.float8_e4m3fn), w.to(torch.float8_e4m3fn)
    one = torch.ones((), device=dev)
    bench("fp8 _scaled_mm (reference point only)", lambda: torch._scaled_mm(x8, w8.t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
except Exception as e:  # noqa: BLE001
    print("fp8 failed", e)
