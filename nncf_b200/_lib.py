"""ctypes binding of libnncf_b200.so (the C-ABI declared in include/nncf_b200.h).

There is NO fallback: if the shared library is missing the import fails loudly, and every compute entry point
needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NNCF_LIB_PATH: developer switch, loads an alternative build of the same library (tools/ablate.sh builds ablated variants)
LIB_PATH = os.environ.get("NNCF_LIB_PATH") or os.path.join(_HERE, "libnncf_b200.so")

# every symbol include/nncf_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "nncf_last_error", "nncf_version", "nncf_launch_count",
    "nncf_sampler_create", "nncf_sampler_destroy", "nncf_sampler_sample_batch_dev", "nncf_sampler_sample_batch_host",
    "nncf_sampler_seek", "nncf_sampler_export_table",
    "nncf_group_sampler_create", "nncf_group_sampler_destroy", "nncf_group_sampler_sample",
    "nncf_group_sampler_sample_with_negs", "nncf_group_sampler_check",
    "nncf_permute_rows", "nncf_group_shuffle_workspace_bytes", "nncf_group_shuffle", "nncf_assemble_pairs_batch",
    "nncf_presample_assemble", "nncf_assemble_sns_batches",
    "nncf_trainer_create", "nncf_trainer_destroy", "nncf_train_steps", "nncf_train_steps_host", "nncf_trainer_set_profile", "nncf_trainer_set_device_clock",
    "nncf_trainer_get_profile", "nncf_unique_first_occurrence", "nncf_gather_rows", "nncf_updater_create",
    "nncf_updater_destroy", "nncf_updater_begin_step", "nncf_updater_apply",
    "nncf_meanpool_fwd", "nncf_meanpool_bwd", "nncf_meanpool_fwd_n", "nncf_meanpool_bwd_n", "nncf_tower_bn_act_fwd", "nncf_tower_bn_act_bwd", "nncf_dense_adam_step",
    "nncf_peer_alloc", "nncf_peer_open", "nncf_peer_close", "nncf_peer_free", "nncf_peer_barrier", "nncf_peer_copy", "nncf_peer_signal", "nncf_peer_wait", "nncf_trainer_set_shards",
    "nncf_eval_topk_workspace_bytes", "nncf_eval_topk", "nncf_eval_metrics", "nncf_score_pairs", "nncf_eval_given",
]

SCHEMES = {"neg_shared": 0, "group_neg_shared": 1, "pairs": 2, "sampled_neg_shared": 3}
LOSSES = {"skip-gram": 0, "mse": 1, "log-loss": 2, "max-margin": 3}
PRECISIONS = {"fp32": 0, "bf16": 1}
OPTIMIZERS = {"none": 0, "sgd": 1, "lazy_adam": 2}
BIASES = {None: 0, "user": 1, "item": 2, "both": 3}


class StepConfig(C.Structure):
    _fields_ = [
        ("scheme", C.c_int32), ("loss", C.c_int32), ("precision", C.c_int32), ("batch_size_p", C.c_int32),
        ("num_negatives", C.c_int32), ("dim", C.c_int32), ("norm_u", C.c_int32), ("norm_v", C.c_int32),
        ("optimizer", C.c_int32), ("replicas", C.c_int32),
        ("neg_loss_weight", C.c_float), ("loss_gamma", C.c_float), ("u_reg", C.c_float), ("learn_rate", C.c_float),
        ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float), ("interaction_bias", C.c_int32),
    ]


class Tables(C.Structure):
    _fields_ = [
        ("user_table", C.c_void_p), ("user_m", C.c_void_p), ("user_v", C.c_void_p), ("n_users", C.c_int64),
        ("item_table", C.c_void_p), ("item_m", C.c_void_p), ("item_v", C.c_void_p), ("n_items", C.c_int64),
    ]


class StepIO(C.Structure):
    _fields_ = [
        ("loss_out_dev", C.c_void_p), ("grad_user_rows_dev", C.c_void_p), ("grad_item_rows_dev", C.c_void_p),
        ("unique_ids_dev", C.c_void_p), ("inverse_dev", C.c_void_p), ("n_unique_dev", C.c_void_p),
        ("item_rows_dev", C.c_void_p), ("response_dev", C.c_void_p),
        ("next_user_ids_dev", C.c_void_p), ("next_item_ids_dev", C.c_void_p),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "nncf_b200: %s is missing. Build it with ./build.sh (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
    sig = {
        "nncf_last_error": (C.c_char_p, []),
        "nncf_version": (i32, []),
        "nncf_launch_count": (i64, []),
        "nncf_sampler_create": (i32, [vp, i32, C.c_double, C.c_uint64, C.POINTER(vp)]),
        "nncf_sampler_destroy": (i32, [vp]),
        "nncf_sampler_sample_batch_dev": (i32, [vp, i64, vp, vp]),
        "nncf_sampler_sample_batch_host": (i32, [vp, i64, vp]),
        "nncf_sampler_seek": (i32, [vp, C.c_uint64]),
        "nncf_sampler_export_table": (i32, [vp, vp, vp]),
        "nncf_group_sampler_create": (i32, [vp, i64, i32, i32, i32, i32, C.c_double, C.c_uint64, C.POINTER(vp)]),
        "nncf_group_sampler_destroy": (i32, [vp]),
        "nncf_group_sampler_sample": (i32, [vp, i32, i32, vp, vp]),
        "nncf_group_sampler_sample_with_negs": (i32, [vp, i32, i32, i32, vp, vp, vp]),
        "nncf_group_sampler_check": (i32, [vp, C.POINTER(C.c_int)]),
        "nncf_permute_rows": (i32, [vp, i64, vp, vp, vp]),
        "nncf_group_shuffle_workspace_bytes": (sz, [i64, i64]),
        "nncf_group_shuffle": (i32, [vp, i64, i32, vp, i64, vp, vp, i32, vp, vp, sz, vp]),
        "nncf_assemble_pairs_batch": (i32, [vp, i32, i32, vp, i32, i32, vp, vp]),
        "nncf_presample_assemble": (i32, [vp, i64, i32, vp, i32, i32, i32, vp, vp]),
        "nncf_assemble_sns_batches": (i32, [vp, i64, i32, i32, vp, vp, vp, vp]),
        "nncf_trainer_create": (i32, [C.POINTER(StepConfig), C.POINTER(vp)]),
        "nncf_trainer_destroy": (i32, [vp]),
        "nncf_train_steps": (i32, [vp, C.POINTER(Tables), vp, vp, i64, C.POINTER(StepIO), vp]),
        "nncf_train_steps_host": (i32, [vp, C.POINTER(Tables), vp, vp, i64, vp, vp]),
        "nncf_trainer_set_profile": (i32, [vp, i32]),
        "nncf_trainer_set_device_clock": (i32, [vp, i32]),
        "nncf_trainer_get_profile": (i32, [vp, vp, vp]),
        "nncf_unique_first_occurrence": (i32, [vp, i32, vp, vp, vp, vp]),
        "nncf_gather_rows": (i32, [vp, i32, vp, i64, vp, vp]),
        "nncf_updater_create": (i32, [i32, f32, f32, f32, f32, C.POINTER(vp)]),
        "nncf_updater_destroy": (i32, [vp]),
        "nncf_updater_begin_step": (i32, [vp]),
        "nncf_updater_apply": (i32, [vp, vp, vp, vp, i64, i32, vp, i64, vp, vp]),
        "nncf_peer_alloc": (i32, [sz, C.POINTER(vp), vp]),
        "nncf_peer_open": (i32, [vp, C.POINTER(vp)]),
        "nncf_peer_close": (i32, [vp]),
        "nncf_peer_free": (i32, [vp]),
        "nncf_peer_barrier": (i32, [vp, i32, i32, C.c_uint, vp]),
        "nncf_peer_copy": (i32, [vp, vp, sz, vp]),
        "nncf_peer_signal": (i32, [vp, C.c_uint, vp]),
        "nncf_peer_wait": (i32, [vp, C.c_uint, vp]),
        "nncf_trainer_set_shards": (i32, [vp, i32, i32, vp, vp, vp]),
        "nncf_meanpool_fwd": (i32, [vp, i32, vp, i32, vp, i32, vp, vp]),
        "nncf_meanpool_bwd": (i32, [vp, i32, vp, i32, vp, i32, vp, vp]),
        "nncf_meanpool_fwd_n": (i32, [vp, i32, vp, i32, vp, i32, vp, vp, vp]),
        "nncf_meanpool_bwd_n": (i32, [vp, i32, vp, i32, vp, i32, vp, vp, vp]),
        "nncf_tower_bn_act_fwd": (i32, [vp, i32, i32, vp, i32, i32, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp]),
        "nncf_tower_bn_act_bwd": (i32, [vp, vp, vp, vp, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp]),
        "nncf_dense_adam_step": (i32, [i32, vp, vp, vp, vp, vp, f32, f32, f32, f32, vp, vp, vp]),
        "nncf_eval_topk_workspace_bytes": (sz, [i64, i64, i32, i32, i32]),
        "nncf_eval_topk": (i32, [vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, sz, vp]),
        "nncf_eval_metrics": (i32, [vp, i64, i32, vp, vp, vp, vp, vp]),
        "nncf_score_pairs": (i32, [vp, vp, i32, vp, vp, i64, vp, vp]),
        "nncf_eval_given": (i32, [vp, vp, vp, i64, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class NNCFError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.nncf_last_error()
        raise NNCFError("libnncf_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
