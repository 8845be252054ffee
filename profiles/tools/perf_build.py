import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from kektordb_b200 import GpuIndex
N = int(os.environ.get("N", 100000)); D = 768; M = 32; R = 32; EFC = int(os.environ.get("EFC", 200)); BMAX = int(os.environ.get("BMAX", 16384))
torch.manual_seed(0)
W = torch.randn(R, D, device="cuda") / R ** 0.5
X = torch.randn(N, R, device="cuda") @ W + 0.1 * torch.randn(N, D, device="cuda")
Q = (torch.randn(1024, R, device="cuda") @ W + 0.1 * torch.randn(1024, D, device="cuda")).cpu().numpy()
u = np.random.default_rng(1).random(N)
gi = GpuIndex(D, "cosine", M, N)
pos = 0; t0 = time.time()
sched = [256]
while sum(sched) < N: sched.append(min(BMAX, sum(sched), N - sum(sched)))
for b in sched:
    torch.cuda.synchronize(); t = time.time()
    gi.add_batch_device(X[pos:pos+b].data_ptr(), b, D, u[pos:pos+b], EFC)
    dt = time.time() - t; pos += b
    print(f"batch {b:6d} -> n={pos:7d}  {dt*1e3:9.1f} ms  {b/dt:9.0f} inserts/s", flush=True)
print("total build s", time.time() - t0)
gt, _, _, _ = gi.flat_search(Q[:256], 10, 1)
for ef in (64, 128, 256):
    ids, sc, cnt, st = gi.SearchWithScores(Q, 10, None, ef)
    ids, sc, cnt, st = gi.SearchWithScores(Q, 10, None, ef)
    rec = np.mean([len(set(ids[i]) & set(gt[i]))/10 for i in range(256)])
    print(f"ef {ef}: recall {rec:.4f} kernel {st.kernel_ms:.3f} ms QPS {1024/st.kernel_ms*1e3:.0f} E/q {st.dist_evals/1024:.0f} H/q {st.hops/1024:.0f}")
