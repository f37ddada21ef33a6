"""Diagnostic: k-block rate of ONE tcgen05 GEMM CTA running alone on the GPU (unloaded L2), vs. ring depth.
Separates a per-CTA pipeline limit from a machine-wide L2 / HBM limit.  python tests/diag_gemm_single_cta.py"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def child():
    import torch
    import air_b200 as ab
    from air_b200 import ops
    ws = torch.zeros(16 << 20, device="cuda"); 
    for (M, N, K) in [(128, 128, 8192), (128, 64, 8192), (128, 128, 32768), (256, 128, 8192), (128 * 148, 128, 8192), (128 * 148, 128, 2048)]:
        A = torch.randn(M, K, device="cuda"); Bm = torch.randn(K, N, device="cuda"); out = torch.empty(M, N, device="cuda")
        run = lambda: ops.gemm(A, Bm, out, mode=1)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"stages={os.environ.get('AIR_TC_STAGES','-')} BN={os.environ.get('AIR_TC_BN','-')} M={M} N={N} K={K}: {us:8.1f} us  {K/32/us:6.2f} kb/us per CTA  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s")

if len(sys.argv) > 1:
    child()
else:
    for st in ("2", "4", "6"):
        env = dict(os.environ, AIR_TC_STAGES=st, AIR_TC_BN="128")
        subprocess.run([sys.executable, __file__, "child"], env=env)
