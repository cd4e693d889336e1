"""Entry point with the reference's flag surface.   ref: main.py:13-105

  python -m nncf_b200.main --data_name citeulike_title_only_fold1 --model_choice mf --conf_choice best \
      --train_scheme neg_shared --eval_scheme whole@50 --param_dict "{'reset_after_getconf': True, 'max_epoch': 5}"

Same flags, same `param_dict` python-literal, same `name@k` eval_scheme split, same trainer protocol.  --gpu sets
CUDA_VISIBLE_DEVICES exactly like the reference (main.py:34-35).
"""
import argparse
import ast
import os


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--data_name', required=True)
    parser.add_argument('--model_choice', required=True)
    parser.add_argument('--conf_choice', required=True)
    parser.add_argument('--train_scheme', default='original')
    parser.add_argument('--eval_scheme', default='given')
    parser.add_argument('--param_dict', default=None)
    parser.add_argument('--pred_name', default=None)
    parser.add_argument('--gpu', default=None, type=str)
    return parser


def run(argv=None):
    args_config = build_parser().parse_args(argv)
    data_name = args_config.data_name
    model_choice = args_config.model_choice
    conf_choice = args_config.conf_choice
    train_scheme = args_config.train_scheme
    eval_scheme = args_config.eval_scheme
    pred_filename = args_config.pred_name
    param_dict = None if args_config.param_dict is None else ast.literal_eval(args_config.param_dict)
    if args_config.gpu is not None:
        os.environ['CUDA_VISIBLE_DEVICES'] = args_config.gpu
    print('model_choice: %s \nconf_choice: %s' % (model_choice, conf_choice))

    # load confs and related (main.py:46-58)
    if model_choice in ('mf', 'pretrained', 'basic_embedding', 'cnn_embedding', 'rnn_embedding'):
        from .conf import get_conf                         # ('pretrained' needs conf.pretrain to name / carry the vectors)
    else:
        assert False, 'model choice %s not defined' % model_choice
    conf = get_conf(data_name, conf_choice, param_dict, model_choice)
    # basic postprocessing (main.py:60-63)
    if eval_scheme.find('@') > 0:
        p = eval_scheme.find('@')
        conf.eval_topk = int(eval_scheme[p + 1:])
        eval_scheme = eval_scheme[:p]

    from .data_utils import get_data
    data_helper = get_data(data_name, conf, reverse_samping=True)
    print(conf.__dict__)

    from .model_framework import get_model
    model_dict = get_model(conf, data_helper, model_choice)

    from .trainers import get_trainer
    Trainer = get_trainer(train_scheme)
    trainer = Trainer(model_dict, conf, data_helper)
    trainer.train(eval_scheme)
    if pred_filename is not None:
        _ = trainer.predict(eval_scheme, pred_filename)
    return trainer


if __name__ == '__main__':
    run()
