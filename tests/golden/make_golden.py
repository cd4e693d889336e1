"""make_golden.py — generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE in this container.

The reference is Python 2 on Keras 1.2.2 / TensorFlow 1.0 and cannot be imported as a package, but the functions on
the hot path are plain Python whose only dependencies are a handful of backend calls.  This script slices those
functions out of the reference files where they lie (nothing is copied into the repo), executes them with
NumPy-backed shims for the `K.*` / `tf.*` calls they make, and stores their inputs and outputs:

  utils/objectives.py      get_original_loss, get_neg_shared_loss, get_group_neg_shared_loss   (loss VALUES)
  configs/data_utils.py    group_shuffle_train                                                 (batch order)
  utils/metrics_ranking.py eval_multiple, eval_multiple_original  (with bottleneck -> numpy argpartition)
  sampler/nodesampler.cpp  via oracle/_ref (table for a small distribution is covered by tests/test_oracle.py)

Run:  python tests/golden/make_golden.py      (needs /root/reference; the GPU box only uses the committed .npz)
"""
import json
import os
import re
import sys
import types

import numpy as np

REF = os.environ.get("NNCF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def slice_function(path, name):
    """Returns the source text of top-level function `name` in a (Python-2) file without parsing the whole file."""
    lines = open(os.path.join(REF, path)).read().split("\n")
    start = next(i for i, l in enumerate(lines) if re.match(r"def %s\(" % re.escape(name), l))
    end = start + 1
    while end < len(lines) and (lines[end].strip() == "" or lines[end][0] in " \t#"):
        end += 1
    return "\n".join(lines[start:end]) + "\n"


# ---------------------------------------------------------------------------------------------- shims
class _K:
    """NumPy realisation of the Keras-backend calls made by utils/objectives.py:35-220."""
    @staticmethod
    def diag(x, size=None):            # utils/utilities.py:74-83 tensorflow_diag -> gather_nd on (i, i)
        return np.diagonal(x).copy()
    mean = staticmethod(lambda x: np.mean(x))
    sum = staticmethod(lambda x: np.sum(x))
    relu = staticmethod(lambda x: np.maximum(x, 0.0))
    log = staticmethod(np.log)
    sigmoid = staticmethod(lambda x: 1.0 / (1.0 + np.exp(-x)))
    reshape = staticmethod(lambda x, s: np.reshape(x, s))
    @staticmethod
    def repeat(x, n):                  # Keras K.repeat: (samples, dim) -> (samples, n, dim)
        return np.repeat(x[:, None, :], n, axis=1)


class _TF:
    """NumPy realisation of the TensorFlow calls made by get_group_neg_shared_loss (utils/objectives.py:163-220)."""
    float32 = np.float64
    constant = staticmethod(lambda x: np.array(x, dtype=np.float64))
    Variable = staticmethod(lambda x, name=None: np.array(x, dtype=np.float64))
    reshape = staticmethod(lambda x, s: np.reshape(x, s))
    shape = staticmethod(lambda x: np.array(x.shape))
    cast = staticmethod(lambda x, t: np.float64(x))
    assign = staticmethod(lambda var, val: np.array(val, dtype=np.float64))
    @staticmethod
    def gather_nd(x, idx):
        idx = np.asarray(idx)
        return x[idx[:, 0], idx[:, 1]]
    @staticmethod
    def scatter_nd_add(var, idx, upd):
        out = np.array(var, dtype=np.float64)
        idx = np.asarray(idx)
        np.add.at(out, (idx[:, 0], idx[:, 1]), upd)
        return out
    @staticmethod
    def slice(x, begin, size):
        return x[begin[0]:begin[0] + int(size[0]), begin[1]:begin[1] + int(size[1])]


class Conf:
    def __init__(self, lam, gamma):
        self.neg_loss_weight = lam
        self.loss_gamma = gamma


def reference_loss_namespace():
    ns = {"np": np, "K": _K, "tf": _TF,
          # utils/objectives.py:15-32 read these two attributes (their fallback branches print in py2 syntax)
          "_get_neg_loss_weight": lambda conf: np.array(conf.neg_loss_weight, dtype=np.float64),
          "_get_gamma": lambda conf, verbose=True: np.array(conf.loss_gamma, dtype=np.float64)}
    for fn in ("get_original_loss", "get_neg_shared_loss", "get_group_neg_shared_loss"):
        exec(compile(slice_function("utils/objectives.py", fn), "ref:utils/objectives.py:" + fn, "exec"), ns)
    return ns


def main():
    from oracle import nncf_oracle as O
    out = {}
    meta = {"loss_cases": [], "original_cases": []}
    ns = reference_loss_namespace()
    rng = np.random.RandomState(2017)
    # ---- neg_shared / group_neg_shared loss values from the reference code
    i = 0
    for loss in ("skip-gram", "mse", "log-loss", "max-margin"):
        for B in (2, 5, 33):
            lam, gamma = (8.0 if loss == "mse" else 128.0), (0.1 if loss == "max-margin" else 10.0)
            S = rng.normal(size=(B, B)) * 0.5
            f = ns["get_neg_shared_loss"](loss, B, Conf(lam, gamma))
            Lvec = f(None, S)
            assert Lvec.shape == (B, 1)
            L = float(np.mean(Lvec))                     # Keras averages the returned (B,1) tensor
            _, G = O.neg_shared_loss_grad(S, loss, lam, gamma)
            out["S_%d" % i], out["L_%d" % i], out["G_%d" % i] = S, L, G
            meta["loss_cases"].append({"scheme": "neg_shared", "loss": loss, "lam": lam, "gamma": gamma})
            i += 1
            nu = max(2, B // 2 + 1)
            P = rng.normal(size=(B, nu)) * 0.5
            pos = rng.randint(0, nu, size=B)
            pos_idxs = np.stack([np.arange(B), pos], 1)
            # the reference pads its masks to (B, B) and slices to pred's shape (objectives.py:169-192)
            Lvec = ns["get_group_neg_shared_loss"](P, pos_idxs, loss, B, Conf(lam, gamma))
            L = float(np.mean(Lvec))
            _, G = O.group_neg_shared_loss_grad(P, pos, loss, lam, gamma)
            out["S_%d" % i], out["L_%d" % i], out["G_%d" % i], out["pos_%d" % i] = P, L, G, pos
            meta["loss_cases"].append({"scheme": "group_neg_shared", "loss": loss, "lam": lam, "gamma": gamma})
            i += 1
    # ---- original loss values
    j = 0
    for loss in ("skip-gram", "mse", "log-loss", "max-margin"):
        B, k = 7, 4
        lam, gamma = (8.0 if loss == "mse" else 128.0), (0.1 if loss == "max-margin" else 10.0)
        s = rng.normal(size=((1 + k) * B, 1)) * 0.5
        y = np.ones(((1 + k) * B, 1)); y[B:] = -1.0 if loss == "skip-gram" else 0.0
        Lvec = ns["get_original_loss"](loss, B, k, Conf(lam, gamma))(y, s)
        out["os_%d" % j], out["oL_%d" % j] = s[:, 0], float(np.mean(Lvec))
        meta["original_cases"].append({"loss": loss, "B": B, "k": k, "lam": lam, "gamma": gamma})
        j += 1
    # ---- group_shuffle_train from the reference (its argsort made stable: the declared tie rule)
    src = slice_function("configs/data_utils.py", "group_shuffle_train")
    assert "train[train[:, -1].argsort()]" in src
    src = src.replace("train[train[:, -1].argsort()]", "train[train[:, -1].argsort(kind='stable')]")
    gns = {"np": np}
    exec(compile(src, "ref:configs/data_utils.py:group_shuffle_train", "exec"), gns)
    train = np.stack([rng.randint(0, 40, 1003), rng.randint(0, 57, 1003), np.ones(1003, dtype=np.int64)], 1)
    nkeys = int(train[:, 1].max()) + 1
    np.random.seed(2017)                                  # the reference draws from the global legacy stream
    gs_out = gns["group_shuffle_train"](train.copy(), by="item", chop=4, iidx=np.arange(nkeys))
    out["gs_train"], out["gs_out"], out["gs_nkeys"] = train, gs_out, nkeys
    # ---- metrics from the reference (continuous scores: its random tie-break never fires)
    bn = types.ModuleType("bottleneck")
    bn.argpartition = lambda a, kth: np.argpartition(a, kth)   # bottleneck >= 1.0 semantics
    sys.modules["bottleneck"] = bn
    mns = {}
    exec(compile(open(os.path.join(REF, "utils/metrics_ranking.py")).read(), "ref:utils/metrics_ranking.py", "exec"), mns)
    nuse, nit = 40, 300
    pred = rng.normal(size=(nuse, nit))
    truth = (rng.uniform(size=(nuse, nit)) < 0.03).astype(np.int32)
    truth[3] = 0
    per = [mns["eval_multiple"](truth[u], pred[u], 10) for u in range(nuse) if truth[u].sum() > 0]
    out["ev_truth"], out["ev_pred"] = truth, pred
    out["ev_map"], out["ev_recall"] = float(np.mean([p[0] for p in per])), float(np.mean([p[1] for p in per]))
    out["ev_per_user"] = np.array(per)
    pero = [mns["eval_multiple_original"](truth[u], pred[u], -1) for u in range(nuse) if truth[u].sum() > 0]
    out["evo_per_user"] = np.array(pero)
    # ---- sampler table restatement at a small size (the full-size table is pinned against oracle/_ref in tests)
    deg = rng.randint(0, 30, size=200).astype(np.float64)
    out["sm_deg"], out["sm_table"] = deg, O.sampler_build_table(deg, 0.75, 10 ** 5)
    out["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_golden.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
