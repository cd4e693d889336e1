"""Developer tool: GPU-side view of a SHORT host-fed call (nncf_train_steps_host, 20 steps after a 200-step warm-up): per
step, when its gather passed griddepcontrol.wait (relative to the first step) and the period to the next step, from the
in-kernel global-timer stamps (NNCF_TIMELINE).  Shows where a short call loses time: host enqueue rate, chunk boundaries."""
import os, sys, time
os.environ["NNCF_TIMELINE"] = "/tmp/nncf_tl.txt"
os.environ["NNCF_HOST_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, '.')
from nncf_b200.ops import FusedStep, StepSpec
steps, warm = 20, 200
R, B, d, n = 37, 512, 128, 1_000_000
g = torch.Generator(device="cuda").manual_seed(0)
EU = (torch.rand((n, d), device="cuda", generator=g) - 0.5) * 0.1
EV = (torch.rand((n, d), device="cuda", generator=g) - 0.5) * 0.1
tot = (steps + warm) * R * B
h_uid = torch.randint(0, n, (tot,), generator=torch.Generator().manual_seed(1), dtype=torch.int32).pin_memory()
h_cid = torch.randint(0, n, (tot,), generator=torch.Generator().manual_seed(2), dtype=torch.int32).pin_memory()
h_loss = torch.empty((steps + warm) * R, dtype=torch.float32).pin_memory()
st = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="sgd", learn_rate=0.01,
                        replicas=R, neg_loss_weight=128.0, loss_gamma=10.0, u_reg=1e-6))
st.run_host(EU, EV, h_uid, h_cid, warm, h_loss)
torch.cuda.synchronize()
t0 = time.perf_counter()
st.run_host(EU, EV, h_uid[warm * R * B:], h_cid[warm * R * B:], steps, h_loss)
dt = time.perf_counter() - t0
print("20-step call: %.1f us wall = %.2f us per step" % (dt * 1e6, dt / steps * 1e6))
del st
import gc; gc.collect()
rows = [list(map(int, l.split())) for l in open("/tmp/nncf_tl.txt") if l.strip() and not l.startswith('#')]
a = np.array([r for r in rows if r[5] > 0 and r[0] < 2 ** 63], dtype=np.float64)[-steps:]
gw, se = a[:, 1], a[:, 5]
print("step: gather past wait (us since step 0), period to next (us), score end -> next gather past wait (us)")
for i in range(len(a)):
    nxt = (gw[i + 1] - gw[i]) / 1e3 if i + 1 < len(a) else float('nan')
    gap = (gw[i + 1] - se[i]) / 1e3 if i + 1 < len(a) else float('nan')
    print("%3d  %8.1f  %6.1f  %6.1f" % (i, (gw[i] - gw[0]) / 1e3, nxt, gap))
print("first gather wait -> last score end: %.1f us" % ((se[-1] - gw[0]) / 1e3))
