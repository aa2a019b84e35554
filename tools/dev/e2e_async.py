"""e2e throughput of the host-buffer path (step_async / step_wait, delta transport) for a given number of env batches in flight and
host threads; MG_HOST_ASYNC / MG_HOST_SYNC_SPIN / taskset are set by the caller (one process per setting: they are read once)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gym_multigrid_b200 as mg
n, RING, steps = 65536, 32, int(os.environ.get("E2E_STEPS", "400"))
if os.environ.get("E2E_PIN"):   # rank r of G keeps to its own slice of the host's cores (threads created later inherit the mask)
    r, G = int(os.environ.get("CUDA_VISIBLE_DEVICES", "0")), int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    cpus = sorted(os.sched_getaffinity(0)); k = len(cpus) // G
    if os.environ["E2E_PIN"] == "ht":   # sibling hyperthreads are i and i + n/2 on most boxes: k/2 cores with both their threads
        half = len(cpus) // 2; h = k // 2
        mine = cpus[r * h:(r + 1) * h] + cpus[half + r * h:half + (r + 1) * h]
    else:
        mine = cpus[r * k:(r + 1) * k]
    os.sched_setaffinity(0, mine)
tag = f"pin={os.environ.get('E2E_PIN', '-')} gpu={os.environ.get('CUDA_VISIBLE_DEVICES', '-')} async={os.environ.get('MG_HOST_ASYNC', '1')} spin={os.environ.get('MG_HOST_SYNC_SPIN', '0')} cores={len(os.sched_getaffinity(0))}"
for spec in sys.argv[1:]:
    EB, threads = (int(x) for x in spec.split(","))
    if os.environ.get("E2E_ALIGN"):   # several processes (one per GPU) started together: begin each configuration at the same wall-clock tick
        q = float(os.environ["E2E_ALIGN"]); time.sleep(q - time.time() % q)
    envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=0, env_id_base=b * n, host_threads=(threads or None)) for b in range(EB)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 4, size=(RING, n, 2)).astype(np.int8) for _ in range(EB)]
    for i in range(3 * EB):
        envs[i % EB].step(acts[i % EB][i % RING])
    torch.cuda.synchronize()
    best = 0.0
    for rep in range(int(os.environ.get('E2E_REPS', '3'))):
        tw = ta = 0.0
        t0 = time.perf_counter()
        for b in range(EB):
            envs[b].step_async(acts[b][0])
        chk = 0.0
        for i in range(steps):
            b = i % EB
            t1 = time.perf_counter()
            obs, rew, term, trunc, _ = envs[b].step_wait()
            chk += float(rew[0, 0]) + float(obs[0, 1, 8, 0])
            t2 = time.perf_counter()
            envs[b].step_async(acts[b][(i // EB + 1) % RING])
            t3 = time.perf_counter()
            tw += t2 - t1; ta += t3 - t2
        for b in range(EB):
            envs[b].step_wait()
        dt = time.perf_counter() - t0
        best = max(best, n * steps / dt)
    print(f"{tag} EB={EB} threads={threads:2d}: best {best:.3e} env-steps/s; last rep per step {dt/steps*1e6:.1f} us = wait {tw/steps*1e6:.1f} + enqueue {ta/steps*1e6:.1f}", flush=True)
    for e in envs:
        e.close()

# ---- where the enqueue time goes (one batch, blocking waits in between so nothing else runs)
if os.environ.get("E2E_BREAKDOWN"):
    import ctypes as C
    e = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=0, host_threads=16)
    e.reset()
    a = np.zeros((n, 2), np.int8)
    for i in range(5):
        e.step_async(a); e.step_wait()
    t = [0.0] * 5
    K = 200
    for i in range(K):
        t0 = time.perf_counter()
        io = e._host_io(a)
        t1 = time.perf_counter()
        cur = torch.cuda.current_stream(e.device)
        idle = cur.query()
        t2 = time.perf_counter()
        rc = e._lib.mg_step_host_async(e._h, C.c_void_p(e.state.data_ptr()), C.byref(io), C.c_void_p(e._host_stream.cuda_stream))
        t3 = time.perf_counter()
        e._host_pending = True
        out = e.step_wait()
        t4 = time.perf_counter()
        t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2; t[3] += t4 - t3
    print(f"{tag} breakdown per step: _host_io {t[0]/K*1e6:.1f} us, stream query {t[1]/K*1e6:.1f} us, mg_step_host_async {t[2]/K*1e6:.1f} us, step_wait {t[3]/K*1e6:.1f} us", flush=True)
