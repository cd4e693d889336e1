"""Developer diagnostic: graph-captured mean-pool step vs the eager autograd step on the same batches - how far apart do the
two end up, and where (user table, tower parameters, BatchNorm statistics, tower output in train / eval mode)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from nncf_b200.conf import Conf
from nncf_b200.data_utils import get_data
from nncf_b200.model_framework import get_model

def run(mode, scheme, loss, nb=7):
    os.environ["NNCF_TOWER_GRAPH"] = mode
    conf = Conf('synthetic_small', {'loss': loss, 'batch_size_p': 128, 'user_dim': 32, 'item_dim': 32, 'word_dim': 32, 'learn_rate': 0.01, 'seed': 3})
    np.random.seed(0); torch.manual_seed(0)
    dh = get_data('synthetic_small', conf, reverse_samping=True)
    md = get_model(conf, dh, 'basic_embedding')
    view = md['model_neg_shared' if scheme == 'neg_shared' else 'model_group_neg_shared']
    train = torch.from_numpy(np.ascontiguousarray(dh.data['train'][:128 * nb], dtype=np.int32)).cuda()
    cost, n = view.train_tower_batches(train[:, 0].contiguous(), train[:, 1].contiguous(), 128)
    torch.cuda.synchronize()
    st = md['_state']
    ids = torch.arange(64, device="cuda", dtype=torch.int32)
    with torch.no_grad():
        st.tower.eval(); e_eval = st.tower(ids).clone()
        rm, rv = st.tower.bn.running_mean.clone(), st.tower.bn.running_var.clone()
        st.tower.train(); e_train = st.tower(ids).clone()
    out = {'cost': torch.tensor(cost), 'user_table': st.user_table.clone(), 'running_mean': rm, 'running_var': rv, 'emb_eval': e_eval, 'emb_train': e_train}
    for n_, p in st.tower.named_parameters(): out['p.' + n_] = p.detach().clone()
    return out

for scheme, loss in (("neg_shared", "skip-gram"), ("group_neg_shared", "log-loss")):
    for rep in range(2):
        for nb in (1, 3, 4, 7):
            a, b = run("1", scheme, loss, nb), run("0", scheme, loss, nb)
            line = []
            for k in a:
                d = (a[k].double() - b[k].double()).abs()
                tol = 2e-5 + 1e-4 * b[k].double().abs()
                line.append("%s max %.2e bad %.1e" % (k, float(d.max()), float((d > tol).double().mean())))
            print(scheme, loss, "rep", rep, "batches", nb, "|", " | ".join(line), flush=True)
        g0, g1 = run("1", scheme, loss), run("1", scheme, loss)
        e0, e1 = run("0", scheme, loss), run("0", scheme, loss)
        print("   graph vs graph: user max %.2e   eager vs eager: user max %.2e" % (float((g0['user_table'] - g1['user_table']).abs().max()), float((e0['user_table'] - e1['user_table']).abs().max())), flush=True)
