"""Times an epoch of the content-tower step (basic_embedding, C1 shape: 5,551 users x 16,980 items, ~205k links, B = 512,
d = dw = 50, L = 300, vocabulary 8,000): python tools/tower_bench.py [neg_shared|group_neg_shared] [graph 1|0]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
scheme = sys.argv[1] if len(sys.argv) > 1 else "neg_shared"
os.environ["NNCF_TOWER_GRAPH"] = sys.argv[2] if len(sys.argv) > 2 else "1"
from nncf_b200.conf import Conf
from nncf_b200.data_utils import get_data
from nncf_b200.model_framework import get_model
conf = Conf('synthetic_citeulike', {'loss': 'skip-gram' if scheme == 'neg_shared' else 'log-loss'})
np.random.seed(0)
dh = get_data('synthetic_citeulike', conf, reverse_samping=True)
md = get_model(conf, dh, 'basic_embedding')
view = md['model_neg_shared' if scheme == 'neg_shared' else 'model_group_neg_shared']
train = torch.from_numpy(np.ascontiguousarray(dh.data['train'], dtype=np.int32)).cuda()
B = conf.batch_size_p
nb = train.shape[0] // B
u, c = train[:nb * B, 0].contiguous(), train[:nb * B, 1].contiguous()
view.train_tower_batches(u[:B * 8], c[:B * 8], B)          # warm-up + capture
torch.cuda.synchronize()
t0 = time.perf_counter()
cost, it = view.train_tower_batches(u, c, B)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
L, dw = dh.data['C'].shape[1], conf.word_dim
print("%s graph=%s: %d it in %.3f s = %.1f us per step = %.3e links/s, mean loss %.4f; mean-pool algorithmic bytes per step ~%.1f MB each way (n_u <= %d, L = %d, dw = %d)" % (
    scheme, os.environ["NNCF_TOWER_GRAPH"], it, dt, dt / it * 1e6, it * B / dt, cost / it, B * L * (4 + 4 * dw) / 1e6, B, L, dw))
