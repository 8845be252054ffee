import sys, os, numpy as np, torch, ctypes as C
sys.path.insert(0, os.getcwd())
from kektordb_b200 import GpuIndex, ffi
N, D = 1_000_000, 768
gi = GpuIndex(D, "cosine", 32, N)
X = torch.randn(N, D, device="cuda")
ffi.check(ffi.lib().kdbgpu_upload_vectors_device(gi._h, 1, N, C.c_void_p(X.data_ptr()), D))
q = np.random.default_rng(0).standard_normal(D).astype(np.float32)
for n in (262144, 2_000_000):
    ids = np.random.default_rng(1).integers(1, N + 1, n).astype(np.uint32)
    out = gi.distance_batch(q, ids)
    out = gi.distance_batch(q, ids)
print("done", out[:3])
