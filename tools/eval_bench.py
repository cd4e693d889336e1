import sys, torch
sys.path.insert(0, '.')
from nncf_b200 import ops
nu, ni, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(1)
U = torch.randn((nu, 128), device="cuda", generator=g) / 128 ** 0.5
V = torch.randn((ni, 128), device="cuda", generator=g) / 128 ** 0.5
ops.eval_topk(U[:1024], V, k, "bf16"); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.eval_topk(U, V, k, "bf16"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("users %d items %d k %d: %.2f ms  %.0f users/s  %.1f TFLOP/s  (%.0f cycles/tile/CTA-slot)" % (nu, ni, k, ms, nu / ms * 1e3, 2.0 * nu * ni * 128 / ms / 1e9, ms * 1e-3 * 1.9e9 / ((ni / 128) * ((nu / 128 + 147) // 148))))
