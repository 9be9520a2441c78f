import os, sys, time
sys.path.insert(0, "multimodal-ghm_b200")
import numpy as np, torch, torch.distributed as dist
from ghm_b200.data_random_GHM import ClipSampler, GenTransition
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
u = np.ones(10) / 10
s = ClipSampler([4, 4], [3, 3], [u, u], [.2, .2], K=4, device=dev, rng="philox", seed=1234)
grid = []
for p in [0.02 * (i + 1) for i in range(20)]:
    np.random.seed(42); grid.append((p, GenTransition(4, 3, 10, p, 1.0), GenTransition(4, 3, 10, p, 1.0)))
n = 65536 * world
def call(k, lazy):
    p, tt, it = grid[k % 20]
    t0 = time.perf_counter(); s.reparameterize([p, p], transitions=(tt, it)); t1 = time.perf_counter()
    s.tree_offset = 0
    h = s.get_Bayes(n_eval=n, keep_batch=True, lazy=lazy, distributed=True); t2 = time.perf_counter()
    return h, t1 - t0, t2 - t1
for lazy in (False, True, False, True):
    for k in range(5): h, _, _ = call(k, False)
    dist.barrier(); torch.cuda.synchronize()
    T0 = time.perf_counter(); pend = None; a = b = c = 0.0
    for k in range(40):
        h, ta, tb = call(k, lazy); a += ta; b += tb
        t0 = time.perf_counter()
        if lazy:
            if pend is not None: pend.result()
            pend = h
        c += time.perf_counter() - t0
    if pend is not None: pend.result()
    torch.cuda.synchronize(); el = time.perf_counter() - T0
    if rank == 0: print("lazy=%s: %.3f ms/call; host: reparam %.3f, get_Bayes %.3f, result %.3f ms" % (lazy, el / 40 * 1e3, a / 40 * 1e3, b / 40 * 1e3, c / 40 * 1e3), flush=True)
dist.destroy_process_group()
