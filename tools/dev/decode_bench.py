"""CPU micro-benchmark of the delta decoder (mg_host_apply_delta) on synthetic records: no GPU needed."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(ROOT, "gym-multigrid_b200", "libmultigrid_b200.so"))
lib.mg_host_apply_delta.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
lib.mg_delta_record_bytes.argtypes = [C.c_int, C.c_int]
N, cells, A = 65536, 100, 2
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2       # mirrors rotated (working set = B x 19.6 MB)
R = lib.mg_delta_record_bytes(cells, A)
rng = np.random.default_rng(0)
recs = []
for b in range(B):
    r = np.zeros((N, R), np.uint8)
    r[:, 0] = rng.integers(2, 5, N)                     # 2-4 patched cells
    r[:, 1:1 + A] = rng.integers(0, 4, (N, A))
    ent = r[:, 1 + A:1 + A + 12].reshape(N, 6, 2)
    ent[:, :, 0] = rng.integers(0, cells, (N, 6)); ent[:, :, 1] = rng.integers(0, 256, (N, 6))
    recs.append(r)
tab = np.arange(33, dtype=np.float64)
obs = [np.zeros((N, cells, 3), np.uint8) for _ in range(B)]
rew = np.zeros((N, A)); te = np.zeros(N, np.uint8); tr = np.zeros(N, np.uint8)
p = lambda a: a.ctypes.data_as(C.c_void_p)
for threads in (1, 2, 4, 8):
    for i in range(2 * B):
        lib.mg_host_apply_delta(p(recs[i % B]), N, cells, A, p(tab), p(obs[i % B]), p(rew), p(te), p(tr), None, threads)
    K = 40
    t0 = time.perf_counter()
    for i in range(K):
        lib.mg_host_apply_delta(p(recs[i % B]), N, cells, A, p(tab), p(obs[i % B]), p(rew), p(te), p(tr), None, threads)
    dt = (time.perf_counter() - t0) / K
    print(f"mirrors={B} threads={threads}: {dt*1e6:8.1f} us per 65536-env decode = {dt/N*threads*1e9:.1f} ns per env-thread", flush=True)
