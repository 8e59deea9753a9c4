"""Time sga_gemm_tf32x3 on a loss-sized Gram (tools only).  SGA_GEMM_DBG switches parts of the kernel off."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import ops
dev = torch.device('cuda:0')
M, N, K = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 12288, 100))]
Z = torch.randn(16384, K, device=dev)
ia = torch.randint(0, 16384, (M,), device=dev, dtype=torch.int32)
ib = torch.randint(0, 16384, (N,), device=dev, dtype=torch.int32)
out = torch.empty(M, N, device=dev)
dbg = int(os.environ.get('SGA_GEMM_DBG', '0'))      # read once by the library: one process per setting
if len(sys.argv) <= 4 and 'SGA_GEMM_DBG' not in os.environ:
    import subprocess
    for d in (0, 1, 2, 3, 4, 8, 12, 15):
        subprocess.run([sys.executable, __file__] + sys.argv[1:4], env=dict(os.environ, SGA_GEMM_DBG=str(d)), check=True)
    sys.exit(0)
for dbg in (dbg,):
    for _ in range(2):
        ops.gemm_tf32x3(Z, Z, M, N, K, a_idx=ia, b_idx=ib, out=out)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.gemm_tf32x3(Z, Z, M, N, K, a_idx=ia, b_idx=ib, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print('dbg %2d: %.3f ms  (%.1f algorithmic TFLOP/s)' % (dbg, ms, 2.0 * M * N * K / ms / 1e9))
