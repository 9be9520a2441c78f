import sys, time
sys.path.insert(0, "multimodal-ghm_b200")
import numpy as np, torch
from ghm_b200 import ops
from ghm_b200.data_random_GHM import ConditionalDenoiseSampler
u = np.ones(10) / 10
s = ConditionalDenoiseSampler([4, 4], [3, 3], [u, u], [.2, .2], sigma=1.0, rng="philox", seed=4321)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B = 262144
print("get_Bayes 1M lazy", t(lambda: s.get_Bayes(n_eval=1048576, lazy=True)))
print("get_Bayes 1M", t(lambda: s.get_Bayes(n_eval=1048576)))
tm, im = s.t_model, s.i_model
out = tm.sample(B, seed=1, root_mode=ops.ROOT_UNIFORM, want_post=True, want_root_hd=True)
print("text sample+bp", t(lambda: tm.sample(B, seed=1, root_mode=ops.ROOT_UNIFORM, want_post=True, want_root_hd=True)))
print("text sample", t(lambda: tm.sample(B, seed=1, root_mode=ops.ROOT_UNIFORM)))
print("image sample given root", t(lambda: im.sample(B, root=out["root"], seed=2)))
io = im.sample(B, root=out["root"], seed=2)
print("noise", t(lambda: im.gauss_noise(io["leaves"], 1.0, seed=3)))
z = im.gauss_noise(io["leaves"], 1.0, seed=3)
print("bp_dns", t(lambda: im.bp_dns(z, 1.0, out["root_hd"])))
mean = im.bp_dns(z, 1.0, out["root_hd"])
print("risk", t(lambda: ops.risk_cdm(mean, io["leaves"])))
print("_run", t(lambda: s._run(B)))
for n in (262144, 524288, 1048576):
    print("get_Bayes", n, t(lambda: s.get_Bayes(n_eval=n, lazy=True)))
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
for _ in range(3): s.get_Bayes(n_eval=1048576, lazy=True)
torch.cuda.synchronize()
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(12); print(st.getvalue()[:3000])
