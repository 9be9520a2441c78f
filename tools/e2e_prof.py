"""Host-side profile of the bench's end-to-end step (development aid): times `reparameterize(p_k)` + `get_Bayes(n_eval)`
through the facade, each half separately, and prints the cProfile top entries of 200 steps.

    python tools/e2e_prof.py          # needs a GPU
"""
import sys, time, cProfile, pstats, io
sys.path.insert(0, "multimodal-ghm_b200")
import numpy as np, torch
from ghm_b200 import data_random_GHM as G
u = np.ones(10) / 10
s = G.ClipSampler([4, 4], [3, 3], [u, u], [.2, .2], rng="philox", seed=1)
grid = []
for p in [0.02 * (i + 1) for i in range(20)]:
    np.random.seed(42)
    grid.append((p, G.GenTransition(4, 3, 10, p, 1.0), G.GenTransition(4, 3, 10, p, 1.0)))
n = 65536
LAZY = "--lazy" in sys.argv
pend = [None]
def step(k):
    p, tt, it = grid[k % 20]
    s.reparameterize([p, p], transitions=(tt, it))
    if not LAZY:
        return s.get_Bayes(n_eval=n, keep_batch=True)
    h = s.get_Bayes(n_eval=n, keep_batch=True, lazy=True)      # read evaluation k-1 while k runs (bench.py e2e)
    if pend[0] is not None:
        pend[0].result()
    pend[0] = h
for k in range(5): step(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(100): step(k)
torch.cuda.synchronize(); el = time.perf_counter() - t0
print("step %.1f us" % (el / 100 * 1e6))
t0 = time.perf_counter()
for k in range(100):
    p, tt, it = grid[k % 20]; s.reparameterize([p, p], transitions=(tt, it))
torch.cuda.synchronize(); print("reparameterize %.1f us" % ((time.perf_counter() - t0) / 100 * 1e6))
t0 = time.perf_counter()
for k in range(100): s.get_Bayes(n_eval=n, keep_batch=True, lazy=LAZY)
torch.cuda.synchronize(); print("get_Bayes %.1f us" % ((time.perf_counter() - t0) / 100 * 1e6))
pr = cProfile.Profile(); pr.enable()
for k in range(200): step(k)
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22); print(st.getvalue()[:5000])
