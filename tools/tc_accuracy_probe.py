"""Probe the accumulation behaviour of tcgen05 kind::tf32 through b2m_debug_tc_gemm."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlx_mcmc_b200 import _cabi
lib = _cabi.load(build_if_missing=False)
lib.b2m_debug_tc_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]

def run(A, B, chunk, mask):
    M, K = A.shape; N = B.shape[0]
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.empty(M, N, device="cuda")
    rc = lib.b2m_debug_tc_gemm(a.data_ptr(), b.data_ptr(), M, N, K, c.data_ptr(), chunk, mask, None)
    assert rc == 0, lib.b2m_last_error()
    torch.cuda.synchronize()
    return c.cpu().numpy()

rng = np.random.default_rng(0)
M, N = 256, 256
for K in (1024,):
    for kind in ("random", "positive"):
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = rng.standard_normal((N, K)).astype(np.float32)
        if kind == "positive":
            A, B = np.abs(A), np.abs(B)
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        f32 = (torch.from_numpy(A) @ torch.from_numpy(B).T).numpy()
        scale = np.max(np.abs(ref))
        row = [f"K={K} {kind:8s} scale={scale:9.1f} fp32-cpu err={np.max(np.abs(f32-ref))/scale:.2e}"]
        for mask in (1, 7):
            for chunk in (1, 4):
                out = run(A, B, chunk, mask)
                err = out - ref
                row.append(f"m{mask}c{chunk}: max {np.max(np.abs(err))/scale:.2e} bias {np.mean(err*np.sign(ref))/scale:+.2e}")
        print("\n   ".join(row))
