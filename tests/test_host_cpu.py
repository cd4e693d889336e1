"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute calls): config surface, data
contract, CSR truth construction, CLI parser, and that the C-ABI library loads and exports every declared symbol."""
import os
import re

import numpy as np
import pytest

from oracle import nncf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    import ctypes
    from nncf_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "nncf_b200.h")).read()
    declared = set(re.findall(r"\b(nncf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(L.EXPORTS), (declared - set(L.EXPORTS), set(L.EXPORTS) - declared)
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert L.lib.nncf_version() >= 100
    assert L.lib.nncf_launch_count() == 0          # nothing launched: no compute without a GPU


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle)|oracle/", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nncf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), (dirpath, f)


def test_no_cpu_fallback_without_cuda():
    import torch
    from nncf_b200._lib import NNCFError
    from nncf_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("this is the CPU-box check")
    with pytest.raises(NNCFError):
        ops.FusedStep(ops.StepSpec())
    with pytest.raises(NNCFError):
        ops.eval_topk(torch.zeros(4, 8), torch.zeros(16, 8), 2)


def test_conf_defaults_and_reference_quirks():
    from nncf_b200.conf import Conf, get_conf
    c = Conf("citeulike_title_only_fold1")
    assert (c.max_epoch, c.batch_size_p, c.num_negatives, c.loss) == (40, 512, 10, "skip-gram")
    assert (c.learn_rate, c.loss_gamma, c.neg_loss_weight) == (0.01, 10, 128)
    assert (c.neg_dist, c.neg_sampling_power, c.shuffle_st, c.chop_size, c.group_shuffling_trick) == ("unigram", 1, "by_item_chop", 2, True)
    assert (c.user_dim, c.item_dim, c.word_dim, c.u_reg, c.eval_topk) == (50, 50, 50, 1e-6, 50)
    assert c.emb_normalization is False
    # quirk kept: overriding `loss` alone does not switch lr / gamma / neg weight (they are evaluated before param_dict)
    c = Conf("x", {"loss": "max-margin"})
    assert (c.learn_rate, c.loss_gamma, c.neg_loss_weight) == (0.01, 10, 128)
    assert c.emb_normalization is True                      # but normalisation follows the loss (basic_embedding_conf.py:53-56)
    # get_conf_best applies its tuned values, then re-applies param_dict only when it carries reset_after_getconf
    pd = {"max_epoch": 3, "word_emb_dropout_rate": 0.1}
    assert get_conf("citeulike_title_only_fold1", "best", pd).max_epoch == 30
    pd["reset_after_getconf"] = True
    best = get_conf("citeulike_title_only_fold1", "best", pd)
    assert best.max_epoch == 3 and best.word_emb_dropout_rate == 0.1
    assert get_conf("news_title_only_fold1", "best", None).max_epoch == 20
    with pytest.raises(AssertionError):
        get_conf("x", "nonsense")


def test_cli_surface_matches_reference():
    from nncf_b200.main import build_parser
    p = build_parser()
    a = p.parse_args(["--data_name", "d", "--model_choice", "mf", "--conf_choice", "best"])
    assert (a.train_scheme, a.eval_scheme, a.param_dict, a.pred_name, a.gpu) == ("original", "given", None, None, None)
    with pytest.raises(SystemExit):
        p.parse_args(["--data_name", "d"])


def test_synthetic_data_contract():
    from nncf_b200.data_utils import make_synthetic
    d = make_synthetic(300, 700, 5000, content_len=20, vocab=100, seed=1)
    assert set(d) == {"C", "train", "test", "test_seen", "train_items", "test_items"}          # data/readme.txt:3
    assert d["train"].shape[1] == 3 and np.all(d["train"][:, 2] == 1)
    assert d["C"].shape == (700, 20)
    assert not set(d["train_items"]) & set(d["test_items"])                                    # cold-start split
    for row in d["C"][:50]:                                                                     # zero padding in the beginning
        nz = np.nonzero(row)[0]
        assert nz.size > 0 and np.all(row[nz[0]:] != 0)


def test_csr_truth_equals_reference_dense_truth():
    from nncf_b200.data_utils import make_synthetic
    from nncf_b200.objectives import _csr_truth
    d = make_synthetic(120, 260, 3000, content_len=8, vocab=50, seed=2)
    user_count = int(max(d["train"][:, 0].max(), d["test"][:, 0].max()) + 1)
    tr_items, te_items, tr_true, te_true = O.prepare_whole_eval(d["train"], d["test"], user_count)
    n_items = 260
    for links, items, dense in ((d["train"], tr_items, tr_true), (d["test"][d["test"][:, 2] == 1], te_items, te_true)):
        pos = np.full(n_items, -1, dtype=np.int64)
        pos[items] = np.arange(items.size)
        indptr, cols = _csr_truth(links, user_count, pos)
        rebuilt = np.zeros_like(dense)
        for u in range(user_count):
            rebuilt[u, cols[indptr[u]:indptr[u + 1]]] = 1
            assert np.all(np.diff(cols[indptr[u]:indptr[u + 1]]) > 0)       # sorted, unique (binary search in the kernel)
        np.testing.assert_array_equal(rebuilt, dense)


def test_golden_original_and_metric_cases():
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
    import json
    meta = json.loads(str(g["meta"]))
    for j, case in enumerate(meta["original_cases"]):
        L, _ = O.original_loss_grad(g["os_%d" % j], case["B"], case["k"], case["loss"], case["lam"], case["gamma"])
        assert L == pytest.approx(float(g["oL_%d" % j]), rel=1e-12)
    truth, pred = g["ev_truth"], g["ev_pred"]
    kept = [u for u in range(truth.shape[0]) if truth[u].sum() > 0]
    got = np.array([O.eval_multiple(truth[u], pred[u], 10) for u in kept])
    np.testing.assert_allclose(got, g["ev_per_user"], rtol=1e-12)
    goto = np.array([O.eval_multiple_original(truth[u], pred[u], -1) for u in kept])
    np.testing.assert_allclose(goto, g["evo_per_user"], rtol=1e-12)


def test_create_val_follows_reference_split():
    """data/create_val.py: 10% of the train items become cold test items; nothing lost, nothing shared"""
    from nncf_b200.data_utils import create_val, make_synthetic
    data = make_synthetic(200, 300, 5000, content_len=10, vocab=50, seed=1)
    n_train, items = data['train'].shape[0], sorted(data['train_items'])
    rs = np.random.RandomState(5)
    expect = list(items); rs.shuffle(expect)                    # the reference shuffles train_items with np.random
    create_val(data, rng=np.random.RandomState(5))
    assert data['test_items'] == expect[:int(len(items) * 0.1)] and data['train_items'] == expect[int(len(items) * 0.1):]
    assert data['train'].shape[0] + data['test'].shape[0] == n_train
    assert set(data['test'][:, 1]) <= set(data['test_items']) and not (set(data['train'][:, 1]) & set(data['test_items']))
    assert np.array_equal(data['test_seen'], data['train'][:data['test'].shape[0]])


def test_cnn_and_rnn_towers_shapes_and_gradients():
    """the CNN / RNN item towers (plain torch modules) honour the tower contract: ids -> [n, d], differentiable"""
    import torch
    from nncf_b200.conf import get_conf
    from nncf_b200.towers import CNNTower, RNNTower

    class DS:
        word_count = 120
        W_pretrain = None
    content = torch.from_numpy(np.random.RandomState(0).randint(0, 120, (40, 12)).astype(np.int32))
    g = torch.Generator().manual_seed(0)
    cases = [('cnn_embedding', CNNTower, {}),
             ('cnn_embedding', CNNTower, {'filter_lengths': [3, 5], 'num_filters': [6, 5], 'poolings': ['max', 'average'], 'pool_lengths': [2, -1]}),
             ('rnn_embedding', RNNTower, {'lstm_dims': [8]}),
             ('rnn_embedding', RNNTower, {'rnn': 'gru', 'bidirection': False, 'use_seq_for_dnn': False, 'lstm_dims': [8, 6]})]
    for mc, T, extra in cases:
        pd = {'user_dim': 16, 'item_dim': 16, 'word_dim': 8}
        pd.update(extra)
        conf = get_conf('synthetic_small', 'default', pd, mc)
        assert conf.u_reg == 1e-5                              # cnn / rnn conf default (configs/cnn_embedding_conf.py:34)
        tower = T(DS, conf, content, g)
        tower.train()
        out = tower(torch.arange(9))
        assert out.shape == (9, 16)
        out.square().sum().backward()
        assert tower.word_embedding.grad is not None and torch.isfinite(tower.word_embedding.grad).all()
        tower.eval()
        with torch.no_grad():
            a, b = tower(torch.arange(5)), tower(torch.arange(5))
        assert torch.equal(a, b)                               # test phase: dropout off, BN running statistics


def test_mf_conf_follows_pretrained_conf():
    """main.py:47-48 gives --model_choice mf the Conf of configs/pretrained_conf.py"""
    from nncf_b200.conf import get_conf
    c = get_conf('citeulike_title_only_fold1', 'default', None, 'mf')
    assert (c.max_epoch, c.num_negatives, c.batch_size_p, c.neg_loss_weight, c.interaction_bias) == (20, 5, 64, 1, 'item')
    assert (c.neg_dist, c.chop_size, c.shuffle_st, c.u_reg) == ('uniform', 1, 'by_item', 1e-5)
    b = get_conf('citeulike_title_only_fold1', 'best', None, 'mf')             # pretrained_conf.py:121-142
    assert (b.max_epoch, b.num_negatives, b.interaction_bias, b.u_reg) == (30, 10, None, 1e-6)
    assert get_conf('news_title_only_fold1', 'best', None, 'mf').u_reg == 1e-5
    assert get_conf('x', 'default', {'interaction_bias': None, 'batch_size_p': 512}, 'mf').batch_size_p == 512


def test_content_id_and_pretrain_combination_towers():
    """use_content_id (ref: modules/content/mean_pool.py:102-108) and ItemCombination (ref: modules/shared/vec2vec.py:17-64)
    as wrappers round any item tower: shapes, the v_reg activity regulariser, merge modes, frozen pretrained vectors."""
    import torch
    from nncf_b200.conf import get_conf
    from nncf_b200.towers import ContentIdTower, PretrainCombinedTower, FrozenItemTable

    class Inner(torch.nn.Module):
        def __init__(self, n, d):
            super().__init__()
            self.w = torch.nn.Parameter(torch.arange(n * d, dtype=torch.float32).reshape(n, d) / (n * d))

        def forward(self, ids):
            return self.w[ids.long()]

    g = torch.Generator().manual_seed(0)
    n_items, d = 30, 8
    ids = torch.tensor([3, 7, 7, 11], dtype=torch.int32)
    t = ContentIdTower(Inner(n_items, d), n_items, d, v_reg=0.5, generator=g, device='cpu')
    t.train()
    out = t(ids)
    e = t.emb_cid[ids.long()]
    assert torch.allclose(out, t.tower(ids) + e)
    # utils/utilities.py:129-135: v_reg * sum_d mean_b E[b, d]^2
    assert torch.allclose(t.reg_loss, 0.5 * (e * e).mean(0).sum())
    assert (t.emb_cid.abs() <= 0.05).all()                                   # Keras-1 'uniform'
    t.eval()
    t(ids)
    assert t.reg_loss is None

    C = np.random.RandomState(1).randn(n_items, 5).astype(np.float32)
    for mode, pdim in (('concat', 5), ('sum', d), ('mul', d), ('ave', d), ('max', d)):
        Cm = C if pdim == 5 else np.random.RandomState(2).randn(n_items, d).astype(np.float32)
        conf = get_conf('synthetic_small', 'default', {'user_dim': 6, 'item_dim': 6, 'pretrain': {
            'C_pretrain': Cm, 'pretrain_combine_mode': mode, 'pretrain_combine_dropout': 0.0, 'pretrain_combine_actv': 'relu'}})
        assert conf.pretrain['pretrain_combine_actv'] == 'relu'
        pt = PretrainCombinedTower(Inner(n_items, d), Cm, conf, tower_dim=d)
        pt.train()
        o = pt(ids)
        assert o.shape == (4, 6) and (o >= 0).all()
        assert not pt.c_pretrain.requires_grad and 'c_pretrain' not in dict(pt.named_parameters())   # trainable=False
        o.sum().backward()
        assert pt.dense.weight.grad is not None and pt.tower.w.grad is not None
        h, p = pt.tower(ids), torch.from_numpy(Cm)[ids.long()]
        merged = {'concat': lambda: torch.cat([h, p], 1), 'sum': lambda: h + p, 'mul': lambda: h * p, 'ave': lambda: 0.5 * (h + p),
                  'max': lambda: torch.maximum(h, p)}[mode]()
        assert torch.allclose(o, torch.relu(pt.dense(merged)), atol=1e-6)
    # dropout >= 1: the pretrained vectors are not used at all (vec2vec.py:29,45-47)
    conf = get_conf('synthetic_small', 'default', {'user_dim': 6, 'item_dim': 6, 'pretrain': {'C_pretrain': C, 'pretrain_combine_dropout': 1.0}})
    pt = PretrainCombinedTower(Inner(n_items, d), C, conf, tower_dim=d)
    assert pt.dense.in_features == d
    # row dropout: whole table rows, rescaled by 1 / (1 - p); the two occurrences of item 7 share the mask
    conf = get_conf('synthetic_small', 'default', {'user_dim': 6, 'item_dim': 6, 'pretrain': {'C_pretrain': C, 'pretrain_combine_dropout': 0.5}})
    pt = PretrainCombinedTower(None, C, conf)
    pt.train()
    torch.manual_seed(0)
    pt.dense.weight.data = torch.eye(6, 5)
    pt.dense.bias.data.zero_()
    pt.actv = lambda x: x
    o = pt(ids)
    ref = torch.from_numpy(C)[ids.long()]
    for r in range(4):
        assert torch.allclose(o[r, :5], torch.zeros(5)) or torch.allclose(o[r, :5], 2.0 * ref[r], atol=1e-6)
    assert torch.allclose(o[1], o[2])
    ft = FrozenItemTable(C)
    assert torch.equal(ft(ids), torch.from_numpy(C)[ids.long()]) and len(list(ft.parameters())) == 0


def test_pretrained_vectors_loader_and_conf(tmp_path):
    """configs/data_utils.py:129-185 (pkl and text formats) and the pretrain keys of the Conf classes"""
    import pickle
    from nncf_b200 import data_utils as DU
    from nncf_b200.conf import get_conf

    class Spec:
        word_count, item_count = 6, 5

    class Helper:
        word2id = {'alpha': 1, 'beta': 4, 'zzz': 99}

    W = np.arange(12, dtype=np.float64).reshape(6, 2)
    wp = tmp_path / 'word_vectors_50d.pkl'
    with open(wp, 'wb') as fp:
        pickle.dump(W, fp, protocol=2)
    sp = tmp_path / 'sentence_vectors_50d.txt'
    sp.write_text('_*0 0.5 1.5 \n_*3 -1.0 2.0 \n_*9 7.0 7.0 \n')
    conf = get_conf('synthetic_small', 'default', {'pretrain': {'wordvec_filepath': str(wp), 'sentvec_filepath': str(sp)}})
    Wl, Cl = DU.get_pretrained_vectors(conf, Spec, Helper)
    assert np.array_equal(Wl, W)
    assert Cl.shape == (5, 2) and np.array_equal(Cl[3], [-1.0, 2.0]) and np.array_equal(Cl[1], [0.0, 0.0])   # id 9 >= item_count dropped
    wt = tmp_path / 'words.txt'
    wt.write_text('2 2\nalpha 1.0 2.0 \nbeta 3.0 4.0 \nzzz 5.0 6.0 \n')
    conf = get_conf('synthetic_small', 'default', {'pretrain': {'wordvec_filepath': str(wt), 'sentvec_filepath': None}})
    Wl, Cl = DU.get_pretrained_vectors(conf, Spec, Helper)
    assert Cl is None and np.array_equal(Wl[1], [1.0, 2.0]) and np.array_equal(Wl[4], [3.0, 4.0]) and not Wl[0].any()
    assert DU.get_pretrained_vectors(get_conf('synthetic_small', 'default', None), Spec, Helper) == (None, None)
    # get_conf_best: per-dataset combine dropout (basic_embedding_conf.py:109-128) and the conf_var switches (:131-136)
    pre = {'sentvec_filepath': str(sp)}
    assert get_conf('citeulike_title_only', 'best', {'pretrain': dict(pre)}).pretrain['pretrain_combine_dropout'] == 0.3
    assert get_conf('news_title_only', 'best', {'pretrain': dict(pre)}).pretrain['pretrain_combine_dropout'] == 0.1
    assert get_conf('citeulike_title_only', 'best', {'pretrain': dict(pre), 'conf_var': 'sup'}).pretrain is None
    assert get_conf('citeulike_title_only', 'best', {'pretrain': dict(pre), 'conf_var': 'unsup_dropout=0.7'}).pretrain['pretrain_combine_dropout'] == 0.7
    # model_choice 'pretrained' (configs/pretrained_conf.py:57-65): transform off by default
    c = get_conf('citeulike_title_only', 'default', None, 'pretrained')
    assert c.pretrain['transform'] is False and c.pretrain['sentvec_filepath'].endswith('pretrain//sentence_vectors_50d.txt')
    assert DU.get_pretrain_folder('news_title_and_abstract_fold2', aug=True).endswith('news/title_and_abstract/pretrain/aug/')


def test_contextual_gating_layers():
    """modules/shared/gatings.py:63-118 as torch modules, wired into the CNN / RNN towers in the reference's order"""
    import torch
    from nncf_b200.conf import get_conf
    from nncf_b200.towers import SpatialGate, TemporalGate, CNNTower
    torch.manual_seed(0)
    x = torch.randn(5, 7, 6)
    sg = SpatialGate(6, {'gating_hidden_dim': 4, 'gating_hidden_actv': 'tanh'})
    g = torch.sigmoid(sg.out(torch.tanh(sg.hidden(x)).mean(1)))
    assert torch.allclose(sg(x), x * g[:, None, :])
    for nl, scale in (('nl', False), ('bn+nl', True), ('bn+l', True)):
        tg = TemporalGate(6, {'gating_hidden_dim': 4, 'gating_hidden_actv': 'relu', 'scale': scale, 'nl_choice': nl})
        tg.eval()
        y = tg(x)
        assert y.shape == x.shape
        ratio = (y / x)                                              # one gate per (sample, step), shared by the dims
        assert torch.allclose(ratio, ratio[:, :, :1].expand_as(ratio), atol=1e-5)
        assert (ratio > 0).all() and (ratio < 1.0 + 1e-6).all()
        assert (tg.c is not None) == scale

    class DS:
        word_count = 50
        W_pretrain = None
    content = torch.from_numpy(np.random.RandomState(0).randint(0, 50, (20, 9)).astype(np.int32))
    gd = {'gating_hidden_dim': 5, 'gating_hidden_actv': 'relu', 'scale': True, 'nl_choice': 'nl'}
    conf = get_conf('synthetic_small', 'default', {'user_dim': 8, 'item_dim': 8, 'word_dim': 6}, 'cnn_embedding')
    conf.contextual_temporal_gated_input = gd
    conf.contextual_spatial_gated_input = gd
    t = CNNTower(DS, conf, content, torch.Generator().manual_seed(0))
    out = t(torch.arange(4))
    assert out.shape == (4, 8)
    out.sum().backward()
    assert t.temporal_gate.hidden.weight.grad is not None and t.spatial_gate.out.weight.grad is not None


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE JSON line (library banners go to stderr); the reference arm runs without a GPU"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # (also: the reference arm must not map the product library; bench.py prints the modules it loaded on request)
    probe = ("import runpy, sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '3', '--warmup', '1', '--replicas', '2'];\n"
             "runpy.run_path('bench.py', run_name='__main__')\n"
             "bad = [m for m in sys.modules if m == 'nncf_b200' or m.startswith('nncf_b200.')]\n"
             "sys.stderr.write('PRODUCT_MODULES=%r\\n' % bad)\n")
    q = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, timeout=600, cwd=root)
    assert q.returncode == 0 and "PRODUCT_MODULES=[]" in q.stderr, q.stderr[-800:]
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:500]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "links/s" and j["higher_is_better"] is True
    assert j["steps"] == 3 and j["warmup"] == 1 and j["n_gpus"] == 1 and j["value"] > 0
    assert j["config"]["replicas_per_step"] == 37 and j["config"]["u_reg"] == 1e-6      # the same step as the GPU arm
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert j["config"]["workload"].startswith("C3")


def test_gpu_runner_scripts_parse():
    """the documented GPU runners (tools/gpu_prof_r02.sh, tools/gpu_experiments_r02.sh) are shell scripts nobody can run here:
    at least their syntax is checked, and every experiment named in the header of the experiment runner has a case branch"""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("gpu_prof_r02.sh", "gpu_experiments_r02.sh", "ablate.sh"):
        subprocess.run(["bash", "-n", os.path.join(root, "tools", name)], check=True)
    src = open(os.path.join(root, "tools", "gpu_experiments_r02.sh")).read()
    named = re.findall(r"^#   (\w+)\s{2,}", src, flags=re.M)
    assert len(named) >= 8
    for n in named:
        assert re.search(r"^\s+%s\)" % re.escape(n), src, flags=re.M), "no case branch for experiment %r" % n
