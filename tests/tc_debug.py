"""Debug harness (not a pytest): decode how the tcgen05 kernel maps MN-major operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keras_rs_b200 as K

def run(tA, tB, M=128, N=256, Kd=16, tag=""):
    K.set_gemm_engine("tcgen05")
    k = np.arange(Kd)
    # logical A (M,K): one-hot selecting k = m % Kd ; logical B (K,N): B[k][n] = k*N + n
    A = np.zeros((M, Kd), np.float32); A[np.arange(M), np.arange(M) % Kd] = 1.0
    B = (np.arange(Kd)[:, None] * N + np.arange(N)[None, :]).astype(np.float32)
    As = np.ascontiguousarray(A.T) if tA else A
    Bs = np.ascontiguousarray(B.T) if tB else B
    c0 = K._lib.lib.krs_gemm_tc_launch_count()
    got = K.ops.sgemm(torch.tensor(As).cuda(), torch.tensor(Bs).cuda(), tA, tB).cpu().numpy()
    torch.cuda.synchronize()
    ran = K._lib.lib.krs_gemm_tc_launch_count() - c0
    ref = A @ B
    ok = np.array_equal(got, ref)
    print(f"[{tag}] tA={tA} tB={tB} tc_ran={ran} exact={ok} maxerr={np.abs(got-ref).max():.1f}")
    if not ok:
        for m in (0, 1, 2, 9, 17, 33, 127):
            row = got[m]
            dec = [(int(v) // N, int(v) % N) if v == int(v) and v >= 0 else ("?", float(v)) for v in row[:40]]
            print(f"   m={m} expect k={m % Kd}: first (k',n') = {dec[:12]} ... n=32..: {dec[32:36]}")
    return ok

if __name__ == "__main__":
    tag = os.environ.get("TAG", "")
    run(False, True, tag=tag)
    run(False, False, tag=tag)
    run(True, True, tag=tag)
    run(True, False, tag=tag)
