#!/bin/bash
# C1 / C2 of BASELINE.json at full CiteULike shape (synthetic stand-in): one training epoch + whole@50 eval
mkdir -p gpurun_out
PD1="{'reset_after_getconf': True, 'max_epoch': 1, 'loss': 'skip-gram', 'num_negatives': 10, 'neg_loss_weight': 128, 'loss_gamma': 10, 'learn_rate': 0.01, 'neg_dist': 'unigram', 'neg_sampling_power': 1, 'batch_size_p': 512}"
PD2="{'reset_after_getconf': True, 'max_epoch': 1, 'loss': 'log-loss', 'num_negatives': 10, 'neg_loss_weight': 128, 'loss_gamma': 10, 'learn_rate': 0.01, 'neg_dist': 'unigram', 'neg_sampling_power': 1, 'batch_size_p': 512, 'chop_size': 4}"
( time timeout 900 python -m nncf_b200.main --data_name citeulike_title_only_fold1 --model_choice basic_embedding --conf_choice best --train_scheme neg_shared --eval_scheme whole@50 --param_dict "$PD1" ) > gpurun_out/c1_basic_neg_shared.log 2>&1
( time timeout 900 python -m nncf_b200.main --data_name citeulike_title_only_fold1 --model_choice basic_embedding --conf_choice best --train_scheme group_neg_shared --eval_scheme whole@50 --param_dict "$PD2" ) > gpurun_out/c2_basic_group_neg_shared.log 2>&1
( time timeout 900 python -m nncf_b200.main --data_name citeulike_title_only_fold1 --model_choice mf --conf_choice best --train_scheme group_neg_shared --eval_scheme whole@50 --param_dict "$PD2" ) > gpurun_out/c2_mf_group_neg_shared.log 2>&1
( time timeout 900 python -m nncf_b200.main --data_name citeulike_title_only_fold1 --model_choice mf --conf_choice best --train_scheme original --eval_scheme given@-1 --param_dict "$PD1" ) > gpurun_out/c1_mf_original_given.log 2>&1
for f in c1_basic_neg_shared c2_basic_group_neg_shared c2_mf_group_neg_shared c1_mf_original_given; do echo "== $f"; grep -E "epoch|Training time|real|Error|error|Traceback" gpurun_out/$f.log | tail -6; done
