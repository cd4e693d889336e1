"""Times one training configuration (device-resident ids): python tools/config_bench.py scheme loss B d R steps [norm] [adam] [ureg]"""
import sys, torch
sys.path.insert(0, '.')
from nncf_b200.ops import FusedStep, StepSpec
scheme, loss, B, d, R, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
norm = "norm" in sys.argv[7:]
adam = "adam" in sys.argv[7:]
u_reg = 1e-6 if "ureg" in sys.argv[7:] else 0.0          # the reference's default activity regulariser
import os
nu = int(os.environ.get("NU", 1_000_000)); ni = int(os.environ.get("NI", 1_000_000))      # table sizes (developer switches)
g = torch.Generator(device="cuda").manual_seed(0)
EU = (torch.rand((nu, d), device="cuda", generator=g) - 0.5) * 0.1
EV = (torch.rand((ni, d), device="cuda", generator=g) - 0.5) * 0.1
n = (steps + 5) * R * B
uid = torch.randint(0, nu, (n,), device="cuda", generator=g, dtype=torch.int32)
cid = torch.randint(0, ni if scheme != "group_neg_shared" else 20000, (n,), device="cuda", generator=g, dtype=torch.int32)
if os.environ.get("ZIPF"):          # power-law ids with offsets ZIPF = "user_offset,item_offset" (bench.py's block distributions)
    sys.path.insert(0, '.')
    from bench import synth_ids_device
    uo, io_ = [float(x) for x in os.environ["ZIPF"].split(",")]
    uid, cid = synth_ids_device(n, nu, ni, 5, torch, user_offset=uo, item_offset=io_)
lam, gamma = (8.0 if loss == "mse" else 128.0), (0.1 if loss == "max-margin" else 10.0)
st = FusedStep(StepSpec(scheme=scheme, loss=loss, precision="bf16", batch_size_p=B, dim=d, norm_u=norm, norm_v=norm, optimizer=("lazy_adam" if adam else "sgd"),
                        learn_rate=(0.001 if adam else 0.01), replicas=R, neg_loss_weight=lam, loss_gamma=gamma, u_reg=u_reg))
state = [torch.zeros_like(EU), torch.zeros_like(EU), torch.zeros_like(EV), torch.zeros_like(EV)] if adam else None
st.run(EU, EV, uid, cid, 5, adam_state=state)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = st.run(EU, EV, uid[5 * R * B:], cid[5 * R * B:], steps, adam_state=state); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
st.set_profile(True); st.run(EU, EV, uid, cid, min(steps, 20), adam_state=state); torch.cuda.synchronize(); ph, ns = st.get_profile()
fl = 6.0 * B * B * d * R
print("%s %s B=%d d=%d R=%d norm=%s: %.1f us/step  %.3e links/s  loss %.4f  phases(us) %s  score-kernel %.1f TFLOP/s algorithmic" % (
    scheme, loss, B, d, R, norm, ms * 1e3, R * B / ms * 1e3, float(out["loss"][-1]), [round(x / ns * 1e3, 1) for x in ph], fl / (ph[1] / ns * 1e-3) / 1e12))
