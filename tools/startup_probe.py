"""Developer probe: device time of K consecutive steps right after a host synchronisation, with and without a short busy
kernel in front of the first event (is there a fixed start-up cost inside a short timed window?)."""
import sys, torch
sys.path.insert(0, '.')
from nncf_b200.ops import FusedStep, StepSpec
R, B, d = 37, 512, 128
g = torch.Generator(device="cuda").manual_seed(0)
EU = (torch.rand((1_000_000, d), device="cuda", generator=g) - 0.5) * 0.1
EV = (torch.rand((1_000_000, d), device="cuda", generator=g) - 0.5) * 0.1
n = 6000 * R * B
uid = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
cid = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
st = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="sgd", learn_rate=0.01, replicas=R, neg_loss_weight=128.0, u_reg=1e-6))
loss = torch.empty(6000 * R, device="cuda")
st.run(EU, EV, uid, cid, 2000, loss_out=loss)
torch.cuda.synchronize()
for spin in (0, 200000, 2000000):
    for K in (20, 20, 40, 100, 300, 1000):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if spin:
            torch.cuda._sleep(spin)
        e0.record()
        st.run(EU, EV, uid, cid, K, loss_out=loss)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("spin %7d  K %4d: %.1f us total, %.2f us/step" % (spin, K, ms * 1e3, ms * 1e3 / K))
