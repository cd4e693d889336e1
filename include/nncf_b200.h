/*
 * nncf_b200.h — C-ABI of libnncf_b200.so: the B200-native (sm_100a) implementation of NNCF's
 * sampling-and-scoring training loop and whole@k / given@k evaluation.
 *
 * Conventions
 *   - every function returns 0 on success or a negative NNCF_E* code; nncf_last_error() gives the text
 *     (thread-local).  The reference aborts with assert() (sampler/nodesampler.cpp:48,65; main.py:57,100);
 *     the host wrappers turn a non-zero status into a Python exception.
 *   - pointers named *_dev are device pointers on the current CUDA device, *_host are host pointers.
 *     All buffers are caller-owned (torch allocates); kernels never allocate except inside the opaque
 *     handles created by *_create and released by *_destroy.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are asynchronous
 *     with respect to the host unless stated otherwise.
 *   - no torch types appear in any signature.
 *
 * Paths cited as "ref:" are relative to the reference tree (chentingpc/NNCF).
 */
#ifndef NNCF_B200_H
#define NNCF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNCF_OK            0
#define NNCF_EINVAL       -1   /* bad argument */
#define NNCF_ECUDA        -2   /* CUDA runtime error (text in nncf_last_error) */
#define NNCF_ENOMEM       -3
#define NNCF_EUNSUPPORTED -4   /* shape / mode outside what the kernels implement */

/* enums are plain ints in the ABI */
enum { NNCF_SCHEME_NEG_SHARED = 0, NNCF_SCHEME_GROUP_NEG_SHARED = 1, NNCF_SCHEME_PAIRS = 2,
       NNCF_SCHEME_SAMPLED_NEG_SHARED = 3 /* B positives + k shared sampled negatives: rows_per_batch = B + k,
                                             ref: models/model_framework.py:138-143, utils/objectives.py:120-161 */ };
enum { NNCF_LOSS_SKIP_GRAM = 0, NNCF_LOSS_MSE = 1, NNCF_LOSS_LOG_LOSS = 2, NNCF_LOSS_MAX_MARGIN = 3 };
enum { NNCF_PREC_FP32 = 0,   /* CUDA-core fp32 FMA, exact fp32 accumulate (1e-4 parity mode)            */
       NNCF_PREC_BF16 = 1 }; /* tcgen05 tensor cores: bf16 operands, fp32 accumulate in TMEM (1e-2 mode) */
enum { NNCF_OPT_NONE = 0, NNCF_OPT_SGD = 1, NNCF_OPT_LAZY_ADAM = 2 };

const char* nncf_last_error(void);
int nncf_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
int64_t nncf_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (1) Negative sampler.   ref: sampler/nodesampler.cpp:55-76 (NodeSampler ctor / sample / sample_batch),
 *     bound by sampler/sampler.pyx:19-38 (MultinomialSampler).  Same target distribution
 *     p_i ∝ dist[i]^power for dist[i] > 0 (nodesampler.cpp:29-49); the 1e8-entry lookup table + 64-bit
 *     LCG become a Walker/Vose alias table in HBM + counter-based Philox4x32-10.
 * ---------------------------------------------------------------------------------------------- */
typedef struct nncf_sampler nncf_sampler_t;
int nncf_sampler_create(const double* dist_host, int dist_size, double neg_sampling_power, uint64_t rand_seed,
                        nncf_sampler_t** out);
int nncf_sampler_destroy(nncf_sampler_t* s);
/* n draws into a device buffer; consumes Philox counters [c, c + n) of this sampler's stream */
int nncf_sampler_sample_batch_dev(nncf_sampler_t* s, int64_t n, int32_t* out_dev, void* stream);
/* n draws into a host buffer (the drop-in for NodeSampler::sample_batch(int n, int* result)); synchronous */
int nncf_sampler_sample_batch_host(nncf_sampler_t* s, int64_t n, int32_t* out_host);
/* position the stream explicitly (multi-GPU: same key, disjoint counter ranges per rank) */
int nncf_sampler_seek(nncf_sampler_t* s, uint64_t counter);
/* copies the alias table to host for inspection by tests: prob[dist_size] (float), alias[dist_size] (int32) */
int nncf_sampler_export_table(nncf_sampler_t* s, float* prob_host, int32_t* alias_host);

/* ------------------------------------------------------------------------------------------------
 * (2) Batch-index builders.   ref: configs/data_utils.py:218-241 (group_shuffle_train), the per-scheme
 *     assembly in models/train_original.py:49-56 and models/train_group_sample.py:75-85.
 *     The three permutations come from the host's shared random stream (same draw order as the
 *     reference); the device does the gathers, the STABLE sort by key and the chop-block permutation.
 * ---------------------------------------------------------------------------------------------- */
/* out[r, :] = train[row_perm[r], :]            (np.random.shuffle(train), train rows are int32[3]) */
int nncf_permute_rows(const int32_t* train_dev, int64_t n_rows, const int64_t* row_perm_dev, int32_t* out_dev,
                      void* stream);
size_t nncf_group_shuffle_workspace_bytes(int64_t n_rows, int64_t n_keys);
/* Full group_shuffle_train: key[r] = iidx[train[row_perm[r], col]]; stable sort by key; if chop > 0 the
 * first (n_rows / chop) blocks of `chop` rows are permuted by block_perm (out block b = sorted block
 * block_perm[b]) and the tail is appended unshuffled.  iidx_dev holds the ALREADY shuffled iidx array. */
int nncf_group_shuffle(const int32_t* train_dev, int64_t n_rows, int col, const int64_t* iidx_dev, int64_t n_keys,
                       const int64_t* row_perm_dev, const int64_t* block_perm_dev, int chop, int32_t* out_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream);
/* (1+k)B-row batch of the 'original' / 'group_sample' schemes: rows [0,B) = positives; row B + p*k + n is
 * positive p with column `neg_col` (1 = item, 0 = user) replaced by negs[p*k + n] and column 2 by neg_sign. */
int nncf_assemble_pairs_batch(const int32_t* pos_dev, int B, int k, const int32_t* negs_dev, int neg_col,
                              int neg_sign, int32_t* out_dev, void* stream);

/* presample: the whole epoch's N links with their k sampled negatives each, (1+k)N rows.  layout 0 = each positive
 * followed by its k negatives (shuffle_st 'original' / 'reverse', ref: models/train_presample.py:46-60); layout 1 =
 * the N positives, then positive p's negatives at N + p k + j (np.vstack((train_p, train_n)), :61-65).
 * negs_dev [N k]; column neg_col of a negative row <- its sample, column 2 <- neg_sign. */
int nncf_presample_assemble(const int32_t* pos_dev, int64_t n_links, int k, const int32_t* negs_dev, int neg_col,
                            int neg_sign, int layout, int32_t* out_dev, void* stream);
/* sampled_neg_shared: id arrays of n_batches batches of B + k rows: the B positives of train rows [b B, (b+1) B), then
 * k rows (user 0, item negs[b k + j]).   ref: models/train_sampled_neg_shared.py:28,46-49 */
int nncf_assemble_sns_batches(const int32_t* train_dev, int64_t n_batches, int B, int k, const int32_t* negs_dev,
                              int32_t* user_ids_dev, int32_t* item_ids_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2b) GroupSampler (the batch source when group_shuffling_trick is False).   ref: configs/data_utils.py:244-408
 *      (class GroupSampler: __init__ :248-304, sample :306-335, sample_with_negs :337-408), constructed at
 *      models/train_group_neg_shared.py:33-37 and models/train_group_sample.py:31-36.
 *      group_by: 0 = 'item' (groups are items, members users), 1 = 'user'.  neg_dist: 0 = 'unigram',
 *      1 = 'uniform' (p_n/p_d corrected), 2 = 'uniform_no_correction'.  train_host = int32 [n_links, 3] HOST rows
 *      (user, item, label), read once to build the CSR-by-group and the alias tables in device memory.
 *      Every call draws a fresh block of the Philox stream selected by rand_seed.
 * ---------------------------------------------------------------------------------------------- */
typedef struct nncf_group_sampler nncf_group_sampler_t;
int nncf_group_sampler_create(const int32_t* train_host, int64_t n_links, int group_by, int chop, int neg_dist,
                              int neg_sign, double neg_sampling_power, uint64_t rand_seed, nncf_group_sampler_t** out);
int nncf_group_sampler_destroy(nncf_group_sampler_t* g);
/* n_batches independent results of GroupSampler.sample(batch_size_p) (strict_return_shape=True):
 * out_dev int32 [n_batches][batch_size_p][3], rows (member, group, 1) (columns swapped for group_by user) */
int nncf_group_sampler_sample(nncf_group_sampler_t* g, int batch_size_p, int n_batches, int32_t* out_dev, void* stream);
/* n_batches independent results of GroupSampler.sample_with_negs(batch_size_p, k): out_dev int32
 * [n_batches][batch_size_p * (1 + k)][3]: positives first (label 1), then negatives (label neg_sign), truncated to
 * batch_size_p * (1 + k) rows; n_pos_dev[n_batches] (optional) = number of positive rows of each batch */
int nncf_group_sampler_sample_with_negs(nncf_group_sampler_t* g, int batch_size_p, int k, int n_batches,
                                        int32_t* out_dev, int32_t* n_pos_dev, void* stream);
/* *failed_out = 1 if some sample_with_negs batch could not be filled within the reference's 10 top-up rounds (the
 * reference asserts there, configs/data_utils.py:368-370).  Synchronises the device and clears the flag. */
int nncf_group_sampler_check(nncf_group_sampler_t* g, int* failed_out);

/* ------------------------------------------------------------------------------------------------
 * (3) Fused training step.   ref: models/model_framework.py:40-65,85-143 (graph), modules/interaction/
 *     interaction_dot.py:92-107 (scores), utils/objectives.py:35-220 (losses), utils/utilities.py:122-135
 *     (activity regulariser), utils/optimizer.py:108-147 (lazy Adam); driven like Keras'
 *     Model.train_on_batch([user_ids, item_ids], [response]) at models/train_neg_shared.py:50.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t scheme;          /* NNCF_SCHEME_*  (PAIRS = row-wise 'mul' view used by original / group_sample) */
  int32_t loss;            /* NNCF_LOSS_* */
  int32_t precision;       /* NNCF_PREC_* */
  int32_t batch_size_p;    /* B: positives per batch */
  int32_t num_negatives;   /* k (PAIRS: the batch has (1+k)B rows; SAMPLED_NEG_SHARED: B + k rows) */
  int32_t dim;             /* d = user_dim = item_dim, 1..256 */
  int32_t norm_u;          /* l2-normalise user rows (emb_normalization, model_framework.py:62-63) */
  int32_t norm_v;          /* l2-normalise item rows (model_framework.py:109-111; not for 'mf' / 'pretrained': :85-88) */
  int32_t optimizer;       /* NNCF_OPT_* */
  int32_t replicas;        /* R >= 1 independent batches per call (synchronous data-parallel virtual workers
                              on one GPU; R = 1 is the reference's sequential loop) */
  float neg_loss_weight;   /* lambda */
  float loss_gamma;        /* gamma */
  float u_reg;             /* activity L2 on the un-normalised user rows */
  float learn_rate;
  float beta1, beta2, epsilon;   /* lazy Adam */
  int32_t interaction_bias;      /* 0 none, 1 user, 2 item, 3 both (ref: modules/interaction/interaction_dot.py:26-34,
                                    96-107).  When non-zero `dim` = embedding dim + 2 and every table row carries two
                                    extra columns: users (ubias, 1), items (1, cbias), so the bias terms come out of the
                                    same contraction; those columns are excluded from l2-normalisation and the regulariser
                                    and the gradient of the constant column (and of an unused bias) is dropped. */
} nncf_step_config;

typedef struct nncf_trainer nncf_trainer_t;
int nncf_trainer_create(const nncf_step_config* cfg, nncf_trainer_t** out);
int nncf_trainer_destroy(nncf_trainer_t* t);

/* Embedding-table operands of one step.  Any of the *_m / *_v pointers may be NULL unless optimizer is
 * LAZY_ADAM.  If item_table is NULL the item side is "dense": item_rows_dev [n_cols, d] holds the item
 * embeddings produced by a framework tower (mean-pool/CNN/RNN) for the batch's columns and the item
 * gradient is only returned (grad_item_rows_dev), never applied. */
typedef struct {
  float* user_table;  float* user_m;  float* user_v;   int64_t n_users;
  float* item_table;  float* item_m;  float* item_v;   int64_t n_items;
} nncf_tables;

/* Runs `n_steps` consecutive steps; step s, replica r reads ids at  ids + (s * R + r) * rows_per_batch,
 * rows_per_batch = B (matmul schemes), (1+k)B (PAIRS) or B + k (SAMPLED_NEG_SHARED).  loss_out_dev[n_steps * R] receives each batch's
 * loss (task loss + regulariser = what Keras' train_on_batch returns).  Lazy Adam's step counter t is kept
 * in the handle (advanced once per step).
 * Optional outputs of the LAST step, replica 0 (NULL to skip), used by parity tests and by framework
 * towers: grad_user_rows_dev [rows, d] and grad_item_rows_dev [n_cols, d] are dLoss/d(raw gathered row),
 * per batch position (duplicates not merged); n_cols = rows (neg_shared, PAIRS) or n_unique
 * (group_neg_shared; unique_ids_dev[B], inverse_dev[B], n_unique_dev[1] receive tf.unique's outputs). */
typedef struct {
  float* loss_out_dev;
  float* grad_user_rows_dev;
  float* grad_item_rows_dev;
  int32_t* unique_ids_dev;
  int32_t* inverse_dev;
  int32_t* n_unique_dev;
  const float* item_rows_dev;     /* dense item side only: [n_cols, d] */
  const int32_t* response_dev;    /* PAIRS scheme, pointwise losses, optional: [n_steps * R * rows] response labels
                                     (1 = positive; -1 / 0 = negative), the y_true of ref utils/objectives.py:59-70.
                                     NULL = rows [0, B) of every batch are the positives (the reference's layout for
                                     train_original / train_group_sample); presample's shuffled batches need it. */
  const int32_t* next_user_ids_dev;  /* optional hint: ids of the step that FOLLOWS this call's last step (the caller's next */
  const int32_t* next_item_ids_dev;  /* call), [R * rows] each; their rows are pulled into L2 while the last step computes.
                                     A hint only: never dereferenced for results, stale or NULL values are harmless. */
} nncf_step_io;

int nncf_train_steps(nncf_trainer_t* t, const nncf_tables* tables, const int32_t* user_ids_dev,
                     const int32_t* item_ids_dev, int64_t n_steps, const nncf_step_io* io, void* stream);
/* Host-fed variant of nncf_train_steps: the ids of all n_steps x R batches live in HOST memory (pinned memory lets the
 * copies overlap the kernels), as the reference's `train` array does (ref: models/train_neg_shared.py:46-50 slices it per
 * batch and feeds it through feed_dict), and loss_out_host[n_steps * R] receives every batch's loss (what Keras'
 * train_on_batch returns, ref: models/train_neg_shared.py:50).  Every step's ids are copied H2D on an internal copy
 * stream, in chunks of 1, 4, 16, 16, ... steps that overlap the kernels of earlier steps; the steps run on `stream`; the
 * R losses of every step of a chunk are copied to loss_out_host on a second copy stream as soon as the chunk has finished.
 * Returns when all steps and copies have completed.  Embedding-table models only. */
int nncf_train_steps_host(nncf_trainer_t* t, const nncf_tables* tables, const int32_t* user_ids_host,
                          const int32_t* item_ids_host, int64_t n_steps, float* loss_out_host, void* stream);
/* Optional per-phase device timing with CUDA events on the launching stream (used by bench.py's roofline leg; it
 * synchronises after every step, so never enable it inside a throughput measurement).
 * phase_ms_out[3] = accumulated ms of { gather/prepare, score+gradient kernel, finalize/optimizer }. */
int nncf_trainer_set_profile(nncf_trainer_t* t, int enable);
/* Device step clock for CUDA-graph capture: with it enabled the lazy-Adam step count and the bias-corrected rate
 * lr_t (ref: utils/optimizer.py:109-111) are kept in device memory and advanced by a kernel of the step itself, so a
 * captured step stays correct when the graph is replayed (a host-side count would be frozen into the graph). */
int nncf_trainer_set_device_clock(nncf_trainer_t* t, int enable);
int nncf_trainer_get_profile(nncf_trainer_t* t, double* phase_ms_out, int64_t* steps_out);
/* tf.unique on device (first-occurrence order), exposed because group_neg_shared towers need it before
 * they can produce item_rows_dev.   ref: models/model_framework.py:45-48 */
int nncf_unique_first_occurrence(const int32_t* ids_dev, int n, int32_t* unique_ids_dev, int32_t* inverse_dev,
                                 int32_t* n_unique_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3b) Multi-GPU: row-sharded embedding tables addressed directly over NVLink / NVSwitch peer memory.
 *      The reference is single-device; this is the extension the north_star prescribes (tables sharded by
 *      row, rows and gradients exchanged over NVLink).  Row `id` of a table lives on rank (id mod N) at local
 *      row (id div N).  Each rank allocates its shards with nncf_peer_alloc, exchanges the 64-byte handles
 *      (e.g. torch.distributed.all_gather_object), opens the others with nncf_peer_open and registers the N
 *      pointers with nncf_trainer_set_shards.  The gather kernel then LOADS remote rows and the update kernel
 *      issues red.global.add.v4 to remote rows inside the same kernels that do the local work; two
 *      device-side barriers per step (nncf_peer_barrier, flag words in peer memory) keep the step synchronous:
 *      nobody updates before everybody has gathered, nobody gathers before every update has landed.
 *      Sparse SGD only (optimizer state is not sharded yet); matmul schemes; dim % 4 == 0.
 * ---------------------------------------------------------------------------------------------- */
int nncf_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out /* [64] */);
int nncf_peer_open(const unsigned char* handle /* [64] */, void** ptr_out);
int nncf_peer_close(void* ptr);
int nncf_peer_free(void* ptr);
int nncf_peer_barrier(void* const* flag_ptrs /* [n_ranks] 64-byte flag arrays */, int n_ranks, int rank, unsigned int epoch,
                      void* stream);
/* Stream-ordered transfer between two device pointers, either of which may be a peer mapping (nncf_peer_open): runs on
 * the copy engines over NVLink, no SM involved.  With the two flag calls below it is the transport of the pipelined
 * stratum rotation (nncf_b200/parallel.py): copy, then signal the receiver's flag; the receiver's stream waits on it. */
int nncf_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
/* *flag = value, ordered after the stream's earlier work (system-scope fences); flag may live on a peer. */
int nncf_peer_signal(void* flag_ptr, unsigned int value, void* stream);
/* blocks the STREAM (not the host) until *flag >= value (wrap-safe compare); bounded spin, traps after ~30 s. */
int nncf_peer_wait(const void* flag_ptr, unsigned int value, void* stream);
int nncf_trainer_set_shards(nncf_trainer_t* t, int n_shards, int rank, void* const* user_shards, void* const* item_shards,
                            void* const* barrier_flags);

/* Stand-alone row gather and sparse row update: the two halves of the step that framework towers and the
 * row-sharded multi-GPU path run on their own (owners gather rows for peers / apply the gradients they get back).
 *   nncf_gather_rows:    out[r, :] = table[ids[r], :]
 *   nncf_updater_apply:  per-position row gradients grads[n, dim] (clobbered) applied to table rows ids[n]:
 *                        SGD = atomic scatter-add, duplicates sum; lazy Adam = duplicates summed, then the
 *                        _apply_sparse rule of ref utils/optimizer.py:108-147 with step count t
 *                        (advance t once per training step with nncf_updater_begin_step). */
int nncf_gather_rows(const float* table_dev, int dim, const int32_t* ids_dev, int64_t n, float* out_dev, void* stream);
typedef struct nncf_updater nncf_updater_t;
int nncf_updater_create(int optimizer, float learn_rate, float beta1, float beta2, float epsilon, nncf_updater_t** out);
int nncf_updater_destroy(nncf_updater_t* u);
int nncf_updater_begin_step(nncf_updater_t* u);
int nncf_updater_apply(nncf_updater_t* u, float* table_dev, float* m_dev, float* v_dev, int64_t n_table_rows, int dim,
                       const int32_t* ids_dev, int64_t n, float* grads_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4) Mean-of-word-vectors item encoder.   ref: modules/content/mean_pool.py:27-33 (AverageEmbeddings:
 *     mean over ALL L positions, pad id 0 included) on word rows gathered per models/model_framework.py:51-56.
 * ---------------------------------------------------------------------------------------------- */
/* out[n, :] = mean_l word_table[content[item_ids[n], l], :]    (item_ids may be NULL: rows 0..n-1) */
int nncf_meanpool_fwd(const float* word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                      const int32_t* item_ids_dev, int n_items, float* out_dev, void* stream);
/* grad_word_table[content[item_ids[n], l], :] += grad_out[n, :] / L   (atomic scatter-add) */
int nncf_meanpool_bwd(float* grad_word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                      const int32_t* item_ids_dev, int n_items, const float* grad_out_dev, void* stream);
/* The same two with the number of valid slots in DEVICE memory (slots [*n_valid, n_slots) read no id: the forward writes
 * zero rows, the backward skips them): the unique-item count of tf.unique never has to reach the host, so a training
 * step of the content model can be captured and replayed as a CUDA graph. */
int nncf_meanpool_fwd_n(const float* word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                        const int32_t* item_ids_dev, int n_slots, const int32_t* n_valid_dev, float* out_dev, void* stream);
int nncf_meanpool_bwd_n(float* grad_word_table_dev, int word_dim, const int32_t* content_dev, int content_len,
                        const int32_t* item_ids_dev, int n_slots, const int32_t* n_valid_dev, const float* grad_out_dev, void* stream);
/* Tail of the content towers on the block of unique items: BatchNorm over the first *n_valid rows (batch statistics,
 * Keras defaults; running statistics updated in place; ref: modules/content/mean_pool.py:90-97) followed by the
 * activation (0 linear, 1 relu, 2 tanh), and its backward.  h / y / xhat / dy / dh are [rows, dim] row-major, rows beyond
 * *n_valid come out as zeros; rstd / dgamma / dbeta are [dim].  use_bn = 0: activation only. */
int nncf_tower_bn_act_fwd(const float* h_dev, int rows, int dim, const int32_t* n_valid_dev, int use_bn, int activation,
                          const float* gamma_dev, const float* beta_dev, float eps, float momentum, float* running_mean_dev,
                          float* running_var_dev, float* y_dev, float* xhat_dev, float* rstd_dev, void* stream);
int nncf_tower_bn_act_bwd(const float* dy_dev, const float* y_dev, const float* xhat_dev, const float* rstd_dev, int rows, int dim,
                          const int32_t* n_valid_dev, int use_bn, int activation, const float* gamma_dev, float* dh_dev,
                          float* dgamma_dev, float* dbeta_dev, void* stream);

/* Dense Adam step on n_tensors parameter tensors (host arrays of device pointers; sizes in elements), Keras-1 form:
 * lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t), p -= lr_t m / (sqrt(v) + eps) - what the reference's towers train with
 * (ref: configs/basic_embedding_conf.py `optimizer = Adam(lr)`; utils/optimizer.py:108-147 restates the rule).  The step
 * count t lives in *step_dev (int64, zero at the start) and is advanced by the call itself on the stream, lr_t_dev is one
 * float of scratch: a call captured into a CUDA graph stays correct when replayed. */
int nncf_dense_adam_step(int n_tensors, float* const* params_dev, const float* const* grads_dev, float* const* m_dev,
                         float* const* v_dev, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                         long long* step_dev, float* lr_t_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) Evaluation.   ref: utils/objectives.py:296-321 (test_eval_mat: all users x candidate items),
 *     utils/metrics_ranking.py:6-35 (eval_multiple), utils/objectives.py:333-370 (evaluate_mat),
 *     utils/objectives.py:231-294 + utils/metrics_ranking.py:38-61 (given@-1).
 * ---------------------------------------------------------------------------------------------- */
size_t nncf_eval_topk_workspace_bytes(int64_t n_users, int64_t n_items, int dim, int topk, int precision);
/* For each of n_users rows of user_rows_dev [n_users, d] score all n_items rows of item_rows_dev
 * [n_items, d] and return the k best columns: topk_ids_dev [n_users, k] (column index into item_rows),
 * topk_scores_dev [n_users, k]; order = score descending, ties by LOWEST column index.  The score matrix
 * is never written to memory. */
int nncf_eval_topk(const float* user_rows_dev, int64_t n_users, const float* item_rows_dev, int64_t n_items, int dim,
                   int topk, int precision, int32_t* topk_ids_dev, float* topk_scores_dev, void* workspace_dev,
                   size_t workspace_bytes, void* stream);
/* Ranking metrics from top-k ids and CSR truth (truth_indptr [n_users+1], truth_cols sorted ascending within a
 * row, in candidate-column space).  per_user_dev [n_users, 3] = (AP@k, recall@k, precision@k), zeros for users
 * with no relevant candidate; sums_dev[4] (double) = (sum AP, sum recall, sum precision, #users kept). */
int nncf_eval_metrics(const int32_t* topk_ids_dev, int64_t n_users, int topk, const int64_t* truth_indptr_dev,
                      const int32_t* truth_cols_dev, float* per_user_dev, double* sums_dev, void* stream);
/* given@k: scores of listed (user, item) pairs = row-wise dot (interaction_dot.py:92-99). */
int nncf_score_pairs(const float* user_table_dev, const float* item_table_dev, int dim, const int32_t* user_ids_dev,
                     const int32_t* item_ids_dev, int64_t n_pairs, float* scores_dev, void* stream);
/* given@k metrics (utils/objectives.py:259-294 with metrics_ranking.py:6-61): pairs grouped by user (seg_indptr
 * [n_groups+1] over the pair list), full descending sort per group (ties: lowest position first).  topk = -1: AP over
 * the whole list (eval_multiple_original); topk >= 1: AP / recall / precision over the first k ranked entries with
 * nhits counted over the whole list and denominator min(nhits, k) (eval_multiple; a list shorter than k is ranked
 * whole).  AUC (Mann-Whitney, average ranks) is over the whole list either way.
 * per_group_dev [n_groups, 4] = (AP@k, AUC, recall@k, precision@k). */
int nncf_eval_given(const float* scores_dev, const int32_t* truth_dev, const int64_t* seg_indptr_dev, int64_t n_groups,
                    int topk, float* per_group_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NNCF_B200_H */
