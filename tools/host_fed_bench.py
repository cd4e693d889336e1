"""Developer tool: the host-fed loop (nncf_train_steps_host) against the device-fed loop on the C3 step: wall clock per
step for a few chunk lengths (NNCF_HOST_CHUNK is read when a trainer's host state is created, so every variant gets its own
process: this script re-executes itself)."""
import os, subprocess, sys, time
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import torch
    sys.path.insert(0, '.')
    from nncf_b200.ops import FusedStep, StepSpec
    steps = int(sys.argv[2])
    R, B, d, n = 37, 512, 128, 1_000_000
    g = torch.Generator(device="cuda").manual_seed(0)
    EU = (torch.rand((n, d), device="cuda", generator=g) - 0.5) * 0.1
    EV = (torch.rand((n, d), device="cuda", generator=g) - 0.5) * 0.1
    tot = (max(steps, 200) + 5) * R * B
    uid = torch.randint(0, n, (tot,), device="cuda", generator=g, dtype=torch.int32)
    cid = torch.randint(0, n, (tot,), device="cuda", generator=g, dtype=torch.int32)
    st = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="sgd", learn_rate=0.01,
                            replicas=R, neg_loss_weight=128.0, loss_gamma=10.0, u_reg=1e-6))
    h_uid, h_cid = uid.cpu().pin_memory(), cid.cpu().pin_memory()
    h_loss = torch.empty((steps + 5) * R, dtype=torch.float32).pin_memory()
    st.run(EU, EV, uid, cid, 200); torch.cuda.synchronize()
    t0 = time.perf_counter(); st.run(EU, EV, uid, cid, steps); t_enq = time.perf_counter() - t0; torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
    st.run_host(EU, EV, h_uid, h_cid, 5, h_loss)
    res = []
    for rep in range(3):
        t0 = time.perf_counter(); st.run_host(EU, EV, h_uid, h_cid, steps, h_loss); res.append((time.perf_counter() - t0) / steps * 1e6)
    print("steps %5d chunk %-4s: device-fed %.2f us/step (host enqueue %.2f)   host-fed %s us/step" % (
        steps, os.environ.get("NNCF_HOST_CHUNK", "dflt"), t_dev / steps * 1e6, t_enq / steps * 1e6,
        " ".join("%.2f" % x for x in res)), flush=True)
else:
    for steps in (20, 1000):
        for chunk in ("1", "4", ""):
            env = dict(os.environ)
            if chunk: env["NNCF_HOST_CHUNK"] = chunk
            subprocess.run([sys.executable, __file__, "one", str(steps)], env=env, timeout=300)
