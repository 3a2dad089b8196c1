"""Separate main-loop / load / epilogue time of the tcgen05 GEMM at the bench shapes."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
lib = P._lib.load()
dev = torch.device("cuda")
M = 131328
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for name, N, K, epi, od in (("qkv", 1536, 512, 0, torch.bfloat16), ("proj", 512, 512, 2, torch.float32),
                            ("fc1", 2048, 512, 1, torch.bfloat16), ("fc2", 512, 2048, 2, torch.float32)):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    bias = torch.zeros(N, device=dev)
    res = torch.randn(M, N, device=dev) if epi == 2 else None
    out = torch.empty(M, N, device=dev, dtype=od)
    import ctypes as C
    g = P._lib.GemmArgs()
    g.A, g.lda, g.W, g.ldw, g.bias = P._lib.ptr(a), K, P._lib.ptr(w), K, P._lib.ptr(bias)
    g.residual, g.ldr = P._lib.ptr(res), N
    g.C, g.ldc, g.out_precision = P._lib.ptr(out), N, (P._lib.PCD_BF16 if od == torch.bfloat16 else P._lib.PCD_F32)
    g.M, g.N, g.K, g.epilogue = M, N, K, epi

    def fn():
        P._lib.check(lib.pcd_gemm_bf16_ex(C.byref(g), P._lib.stream_ptr()), "gemm")
    r = {}
    for flags in (0, 1, 2, 3, 10):  # per-call profiling switches (pcd_gemm_args.debug)
        g.debug = flags
        r[flags] = t(fn)
    g.debug = 0
    fl = 2.0 * M * N * K
    print(f"{name:5s} N={N} K={K}: full {r[0]*1e3:7.1f}us ({fl/r[0]/1e9:6.0f} TF) | no-epilogue {r[1]*1e3:7.1f}us ({fl/r[1]/1e9:6.0f} TF) | "
          f"no-TMA {r[2]*1e3:7.1f}us | MMA-only {r[3]*1e3:7.1f}us ({fl/r[3]/1e9:6.0f} TF) | epilogue-only {r[10]*1e3:7.1f}us")
