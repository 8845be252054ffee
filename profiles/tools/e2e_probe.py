import sys, os, time, threading, numpy as np, torch
sys.path.insert(0, os.getcwd())
from kektordb_b200 import GpuIndex
N = 200000; D = 768; M = 32; R = 32; B = 1024
torch.manual_seed(0)
W = torch.randn(R, D, device="cuda") / R ** 0.5
X = torch.randn(N, R, device="cuda") @ W + 0.1 * torch.randn(N, D, device="cuda")
u = np.random.default_rng(1).random(N)
gi = GpuIndex(D, "cosine", M, N)
pos = 0
sched = [200]
while sum(sched) < N: sched.append(min(16384, sum(sched), N - sum(sched)))
for b in sched:
    gi.add_batch_device(X[pos:pos+b].data_ptr(), b, D, u[pos:pos+b], 200); pos += b
Qd = torch.randn(32 * B, R, device="cuda") @ W + 0.1 * torch.randn(32 * B, D, device="cuda")
Qh = torch.empty((32 * B, D), dtype=torch.float32, pin_memory=True); Qh.copy_(Qd); torch.cuda.synchronize()
Q = Qh.numpy()
gi.set_tuning(int(os.environ.get("SLOTS", 4)), 0, -1)
for _ in range(3): gi.SearchWithScores(Q[:B], 10, None, 128)
for nthreads in (1, 2, 3, 4):
    stats = []
    def worker(j):
        for i in range(j, 32, nthreads):
            t = time.perf_counter()
            ids, sc, cnt, st = gi.SearchWithScores(Q[i*B:(i+1)*B], 10, None, 128)
            stats.append((j, i, time.perf_counter() - t, st.kernel_ms, st.total_ms))
    ths = [threading.Thread(target=worker, args=(j,)) for j in range(nthreads)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    el = time.perf_counter() - t0
    wall = np.mean([s[2] for s in stats]) * 1e3; km = np.mean([s[3] for s in stats]); tm = np.mean([s[4] for s in stats])
    print(f"threads {nthreads}: QPS {32*B/el:.0f}  mean call wall {wall:.3f} ms  kernel_ms {km:.3f}  total_ms(dev) {tm:.3f}")
