"""Run the node-level GEMM kernel (sndvae_debug_gemm) at the step's dominant shapes; for `ncu` captures and quick timing.
usage: python tools/gemm_probe.py [reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sndvae_b200 as sv

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = sv.Engine(sv.make_config(8, 2, "disentangled", sampling_num=2))
rows = 655360
SHAPES = [  # name, tA, tB, M, N, K
    ("wgrad coef3^T dy", 1, 0, 71, 50, rows), ("wgrad coef2^T dm2s", 1, 0, 92, 50, rows), ("wgrad conv", 1, 0, 250, 50, 65536),
    ("fwd coef2.W2", 0, 0, rows, 50, 92), ("dgrad dm2s.W2^T", 0, 1, rows, 92, 50), ("fwd coef3.W3", 0, 0, rows, 50, 71),
    ("head fwd", 0, 0, 2560, 100, 12800), ("head wgrad", 1, 0, 12800, 100, 2560), ("conv fwd", 0, 0, 65536, 50, 250),
]
g = torch.Generator().manual_seed(0)
for name, tA, tB, M, N, K in SHAPES:
    A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tB else (K, N), generator=g).cuda()
    C0 = torch.zeros((M, N)).cuda()
    eng.debug_gemm(A, B, tA=bool(tA), tB=bool(tB), beta=1.0 if tA else 0.0, C0=C0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Cm = C0.clone()
    import ctypes as C
    e0.record()
    for _ in range(reps):
        eng.lib.sndvae_debug_gemm(eng._h, tA, tB, M, N, K, 1.0, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1],
                                  1.0 if tA else 0.0, Cm.data_ptr(), N, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = 4.0 * (A.numel() + B.numel() + M * N)
    print(f"{name:22s} tA={tA} tB={tB} M={M} N={N} K={K}: {ms:.4f} ms  {byts / ms / 1e6:.0f} GB/s", flush=True)
