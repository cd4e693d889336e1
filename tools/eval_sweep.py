"""In-process sweep of the whole@k kernels' developer switches: python tools/eval_sweep.py users items [k]
Prints ms, TFLOP/s and cycles per (128 x 128) tile per SM for each (generation, stages, cap, dbg, pipe) setting."""
import os, sys, torch
sys.path.insert(0, '.')
from nncf_b200 import ops
nu, ni = int(sys.argv[1]), int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 50
g = torch.Generator(device="cuda").manual_seed(1)
U = torch.randn((nu, 128), device="cuda", generator=g) / 128 ** 0.5
V = torch.randn((ni, 128), device="cuda", generator=g) / 128 ** 0.5
KEYS = ("NNCF_EVAL_GEN", "NNCF_EVAL_NST", "NNCF_EVAL_NST3", "NNCF_EVAL_CAP", "NNCF_EVAL_DBG", "NNCF_EVAL_PIPE", "NNCF_EVAL_CL")
def run(tag, kk=k, **env):
    for key in KEYS: os.environ.pop(key, None)
    for key, val in env.items(): os.environ["NNCF_EVAL_" + key] = str(val)
    try:
        ops.eval_topk(U[:2048], V[:65536], kk, "bf16"); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(2):
            e0.record(); ops.eval_topk(U, V, kk, "bf16"); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        waves = (nu / 128 + 147) // 148
        print("%-44s k=%3d %8.2f ms %7.1f TFLOP/s %6.0f cyc/tile" % (tag, kk, best, 2.0 * nu * ni * 128 / best / 1e9, best * 1e-3 * 1.9e9 / ((ni / 128) * waves)), flush=True)
    except Exception as ex:
        print("%-44s FAILED %s" % (tag, str(ex)[:200]), flush=True)
        raise
which = sys.argv[4] if len(sys.argv) > 4 else "all"
if which in ("all", "g2"):
    run("g2 default")
    for nst, cap in ((2, 96), (3, 96), (4, 96)):
        run("g2 nst=%d cap=%d pipeline only (dbg2)" % (nst, cap), NST=nst, CAP=cap, DBG=2)
    run("g2 nst=4 cap=96 dbg2 no copies", NST=4, CAP=96, DBG=2, PIPE=1)
    run("g2 nst=4 cap=96 dbg2 no MMAs", NST=4, CAP=96, DBG=2, PIPE=2)
    run("g2 nst=4 cap=96 dbg2 cluster 1", NST=4, CAP=96, DBG=2, CL=1)
    run("g2 nst=4 cap=96 + TMEM loads (dbg1)", NST=4, CAP=96, DBG=1)
    run("g2 nst=4 cap=96 + maxima/ballot (dbg3)", NST=4, CAP=96, DBG=3)
    run("g2 nst=4 cap=96 full", NST=4, CAP=96)
    run("g2 default, old filter (dbg4)", DBG=4)
if which in ("all", "g3"):
    for nst in (2, 4, 6, 8):
        run("g3 nst=%d cap=96 pipeline only (dbg2)" % nst, GEN=3, NST3=nst, CAP=96, DBG=2)
    run("g3 cap=96 dbg2 no copies", GEN=3, CAP=96, DBG=2, PIPE=1)
    run("g3 cap=96 dbg2 no MMAs", GEN=3, CAP=96, DBG=2, PIPE=2)
    run("g3 cap=96 + TMEM loads (dbg1)", GEN=3, CAP=96, DBG=1)
    run("g3 cap=96 + maxima/ballot (dbg3)", GEN=3, CAP=96, DBG=3)
    run("g3 cap=96 full", GEN=3, CAP=96)
    run("g3 default full", GEN=3)
    run("g3 default, old filter (dbg4)", GEN=3, DBG=4)
    for cap in (96, 128, 160, 192):
        run("g3 cap=%d full" % cap, GEN=3, CAP=cap)
    for kk in (10, 100):
        run("g3 default full", kk=kk, GEN=3)
        run("g2 default full", kk=kk)
if which == "quick":
    run("g3 default, exact compaction (pipe4)", GEN=3, PIPE=4)
    run("g3 old filter, exact compaction", GEN=3, DBG=4, PIPE=4)
    run("g3 cap=128 nst=4", GEN=3, CAP=128, NST3=4)
    run("g2 default full"); run("g2 default, old filter (dbg4)", DBG=4)
    run("g3 default full", GEN=3); run("g3 default, old filter (dbg4)", GEN=3, DBG=4)
    run("g3 + maxima/ballot only (dbg3)", GEN=3, DBG=3)
    for cap in (96, 160, 192):
        run("g3 cap=%d full" % cap, GEN=3, CAP=cap)
    for kk in (10, 100):
        run("g3 default full", kk=kk, GEN=3)
        run("g3 default, old filter (dbg4)", kk=kk, GEN=3, DBG=4)
if which == "stats":
    os.environ["NNCF_EVAL_STATS"] = "1"
    run("g3 old filter (dbg4)", GEN=3, DBG=4)
    run("g3 new filter", GEN=3)
    run("g3 old filter (dbg4) exact compaction", GEN=3, DBG=4, PIPE=4)
    run("g3 old filter (dbg4) k=10", kk=10, GEN=3, DBG=4)
if which == "quick2":
    for kk in (10, 50, 100):
        run("g2 default", kk=kk)
        run("g3 default", kk=kk, GEN=3)
        run("g3 default (again)", kk=kk, GEN=3)
    run("g3 exact compaction", GEN=3, PIPE=4)
    run("g3 maxima only (dbg3)", GEN=3, DBG=3)
    run("g3 pipeline only (dbg2)", GEN=3, DBG=2)
