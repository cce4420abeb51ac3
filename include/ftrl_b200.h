/*
 * ftrl_b200.h -- C ABI of the B200-native FTRL trainer for LR / FM / FFM.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (massquantity/Ftrl-FFM) has no FFI layer; the boundary its callers use is the
 * C++ virtual interface ftrl::FtrlModel (src/include/model/ftrl_model.h:14-51)
 * plus the public weight members.  Each entry point below names the reference
 * interface it replaces.  Plain pointers and sizes only; no exceptions cross
 * this boundary; every function returns 0 on success or a negative ftrl_status.
 *
 * Threading: one handle <-> one host control thread.  Concurrency comes from the
 * CUDA streams inside the handle (they replace the reference's ThreadPool /
 * Hogwild workers, src/include/concurrent/thread_pool.h, src/task/ftrl_offline.cpp:85-91).
 *
 * There is no CPU fallback: if no CUDA device is usable every call fails with
 * FTRL_ERR_CUDA and ftrl_last_error() says why.
 */
#ifndef FTRL_B200_H
#define FTRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTRL_B200_ABI_VERSION 1

typedef enum ftrl_status {
  FTRL_OK = 0,
  FTRL_ERR_ARG = -1,   /* bad argument (reference: std::invalid_argument, ftrl_offline.cpp:29-32) */
  FTRL_ERR_CUDA = -2,  /* CUDA runtime / driver error, or no device                                 */
  FTRL_ERR_IO = -3,    /* file error (reference: exit(EXIT_FAILURE), compression/file_ops.c)        */
  FTRL_ERR_NCCL = -4,  /* multi-GPU exchange error                                                  */
  FTRL_ERR_STATE = -5  /* call order error                                                          */
} ftrl_status;

/* model_type switch of the reference: "LR" | "FM" | "FFM" (cmd_option.cpp:69-70,
 * enum ModelType in src/include/utils/types.h:21-25). */
typedef enum ftrl_model_type { FTRL_LR = 0, FTRL_FM = 1, FTRL_FFM = 2 } ftrl_model_type;

/* How ftrl_train_batch treats the samples of one call.
 *  FTRL_MODE_BATCH      : minibatch semantics.  All samples of the call
 *                         read the w materialised from (n,z) at call start -- what
 *                         concurrent Hogwild workers of the reference observe -- and
 *                         every coordinate receives the telescoped reference recurrence
 *                         n' = n + sum g^2, z' = z + sum g - w (sqrt(n') - sqrt(n)) / alpha.
 *                         Duplicates are resolved by sort-by-key segmented reduction.
 *  FTRL_MODE_SEQUENTIAL : reference-exact.  Samples are applied one after another in
 *                         row order with the reference's fp32 operation order
 *                         (incl. ffm.cpp:118); equals n_threads=1 of the reference. */
typedef enum ftrl_mode { FTRL_MODE_BATCH = 0, FTRL_MODE_SEQUENTIAL = 1 } ftrl_mode;

/* POD mirror of config_options (src/include/utils/cmd_option.h:29-63) restricted to
 * what the models read, plus device-side knobs.  Zero-initialise, then call
 * ftrl_config_default(). */
typedef struct ftrl_config {
  int32_t model_type;   /* ftrl_model_type                       (--model_type)  */
  int32_t n_feats;      /* table height                           (--n_feats)     */
  int32_t n_fields;     /* FFM only                               (--n_fields)    */
  int32_t n_factors;    /* FM / FFM                               (--n_factors)   */
  float init_mean;      /* Gaussian init of w                     (--init_mean)   */
  float init_stddev;    /*                                        (--init_stddev) */
  float w_alpha;        /*                                        (--w_alpha)     */
  float w_beta;         /*                                        (--w_beta)      */
  float w_l1;           /*                                        (--w_l1)        */
  float w_l2;           /*                                        (--w_l2)        */
  int32_t mode;         /* ftrl_mode                                              */
  int32_t device;       /* CUDA device ordinal                                    */
  uint64_t seed;        /* Philox seed for the Gaussian init (the reference seeds
                           from std::random_device per weight, utils.h:31-36)     */
  int64_t max_batch_rows; /* capacity hints for the CSR staging slots; 0 = grow   */
  int64_t max_batch_nnz;  /* on demand                                            */
  /* feature-sharded multi-GPU run: this process owns rows with
   * feat % world_size == rank.  world_size <= 1: single GPU.                      */
  int32_t rank;
  int32_t world_size;
  /* reserved[0] bit 0: the caller of ftrl_train_batch_device promises that the CSR arrays of a call are complete
   * before the PREVIOUS train call on this handle was made (e.g. a data set resident in HBM).  The library may
   * then read them -- not the weights -- while the previous batch is still being trained: the sort / segment work
   * of batch i+1 overlaps the forward / update kernels of batch i.  ftrl_train_batch (host pointers) always
   * does this; without the bit the device-pointer variant reads its inputs in stream order only. */
  int32_t reserved[8];
} ftrl_config;

typedef struct ftrl_handle ftrl_handle;

/* fills the reference defaults (cmd_option.h:49-63): FFM, n_fields 8, n_feats 10000,
 * n_factors 16, init N(0, 0.02), alpha 1e-4, beta 1, l1 0.1, l2 5; mode BATCH, device 0 */
void ftrl_config_default(ftrl_config *cfg);

int ftrl_abi_version(void);

/* Replaces LR/FM/FFM(const config_options&) (ftrl_model.cpp:12-34, fm.cpp:9-19,
 * ffm.cpp:17-28): allocates bias/lin/vec w,n,z in HBM, w ~ N(init_mean, init_stddev)
 * generated on device, n = z = 0. */
int ftrl_create(const ftrl_config *cfg, ftrl_handle **out);
/* Replaces the C++ destructor. */
void ftrl_destroy(ftrl_handle *h);
/* Message of the last failing call on this handle (h == NULL: last ftrl_create failure). */
const char *ftrl_last_error(const ftrl_handle *h);

/* Replaces `float FtrlModel::train(feat_vec&, int)` (lr.cpp:9-18, fm.cpp:21-32,
 * ffm.cpp:38-49) called per sample by ftrl_offline.cpp:74-83 / ftrl_online.cpp:70-80:
 * one call trains on n_rows samples given as CSR over HOST memory (pinned or pageable):
 *   row_ptr[n_rows+1] (int64), field/feat/val[row_ptr[n_rows]], label[n_rows] (0/1).
 * Out-of-range features are masked on device by the rule of remove_out_range
 * (ftrl_model.cpp:36-42, ffm.cpp:30-36).  The call is asynchronous: the CSR is copied
 * into an internal device slot and kernels are enqueued; logits_out (pre-update logits,
 * nullable, host) and *loss_sum_out (fp64 sum of eval/loss.h:8-12, nullable, host) are
 * valid after ftrl_sync().  Input buffers may be reused as soon as the call returns when
 * pageable, after ftrl_sync() or THREE further train / predict calls when pinned (three
 * CSR slots rotate; the copy of call i is only known to be complete when call i + 3 starts). */
int ftrl_train_batch(ftrl_handle *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field,
                     const int32_t *feat, const float *val, const int32_t *label,
                     float *logits_out, double *loss_sum_out);

/* Same, with every pointer in DEVICE memory of cfg.device (inputs already resident in
 * HBM).  logits_out / loss_sum_out are device pointers (nullable). */
int ftrl_train_batch_device(ftrl_handle *h, int64_t n_rows, int64_t nnz, const int64_t *row_ptr,
                            const int32_t *field, const int32_t *feat, const float *val,
                            const int32_t *label, float *logits_out, double *loss_sum_out);

/* Replaces `float FtrlModel::predict(feat_vec&, bool)` (lr.cpp:20-24, fm.cpp:34-38,
 * ffm.cpp:51-55) as called by evaluate.cpp:23-33 / ftrl_offline.cpp:56-61.  Reads the
 * stored w (stale by one update, like the reference).  out[n_rows] receives logits or
 * probabilities; label / loss_sum_out are optional (NULL). Host pointers. */
int ftrl_predict_batch(ftrl_handle *h, int64_t n_rows, const int64_t *row_ptr, const int32_t *field,
                       const int32_t *feat, const float *val, const int32_t *label, int output_prob,
                       float *out, double *loss_sum_out);
int ftrl_predict_batch_device(ftrl_handle *h, int64_t n_rows, int64_t nnz, const int64_t *row_ptr,
                              const int32_t *field, const int32_t *feat, const float *val,
                              const int32_t *label, int output_prob, float *out,
                              double *loss_sum_out);

/* ROC AUC of n scores (logits or probabilities: any monotone score) against 0/1 labels, computed on the
 * device: radix sort + Mann-Whitney rank sum with average ranks over ties (exact integer arithmetic).
 * The reference reports only the mean log-loss (src/include/eval/loss.h:8-12, evaluate.cpp:39-49); this is
 * the second quality metric the evaluation pass reports (`main --auc true`).  Synchronous; *auc_out is a
 * host double, NaN when only one class is present.  Host pointers / device pointers. */
int ftrl_eval_auc(ftrl_handle *h, int64_t n, const float *scores, const int32_t *label, double *auc_out);
int ftrl_eval_auc_device(ftrl_handle *h, int64_t n, const float *scores, const int32_t *label,
                         double *auc_out);

/* Blocks until every enqueued batch has finished and host outputs are written.
 * Replaces ThreadPool::synchronize (thread_pool.h:82-88) / the epoch join. */
int ftrl_sync(ftrl_handle *h);

/* Replaces reads/writes of the public members bias, lin_w, vec_w (ftrl_model.h:36-37,
 * fm.h:22, ffm.h:25) in the reference layout: lin_w[n_feats], vec_w[n_feats][row_len],
 * row_len = n_factors (FM) | n_fields*n_factors (FFM, index field*k+f), 0 (LR).
 * Any pointer may be NULL to skip that part.  Host pointers; synchronous. */
int ftrl_get_weights(ftrl_handle *h, float *bias, float *lin_w, float *vec_w);
int ftrl_set_weights(ftrl_handle *h, const float *bias, const float *lin_w, const float *vec_w);
/* The protected/private FTRL accumulators (ftrl_model.h:45-48, fm.h:25-26, ffm.h:29-30),
 * needed by parity tests and resumable checkpoints.  which: 1 = n, 2 = z. */
int ftrl_get_state(ftrl_handle *h, int which, float *bias_s, float *lin_s, float *vec_s);
int ftrl_set_state(ftrl_handle *h, int which, const float *bias_s, const float *lin_s,
                   const float *vec_s);
/* Row-range variants for tables that do not fit host memory at once:
 * rows [row0, row0 + n_rows) of lin / vec.  which: 0 = w, 1 = n, 2 = z. */
int ftrl_get_rows(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, float *lin, float *vec);
int ftrl_set_rows(ftrl_handle *h, int which, int64_t row0, int64_t n_rows, const float *lin,
                  const float *vec);
int64_t ftrl_row_len(const ftrl_handle *h);

/* Replaces utils::has_zero_weights over lin_w / vec_w (utils.h:63-76,
 * ftrl_offline.cpp:105-119): *out = 1 if any stored weight == 0. */
int ftrl_has_zero_weights(ftrl_handle *h, int *out);

/* Replaces LR/FFM::save_compressed_model / load_compressed_model (lr.cpp:26-39,
 * ffm.cpp:138-159, compression/compress.cpp:15-51): one zstd frame (content size in
 * the frame header) of little-endian fp32 [bias][lin_w][vec_w rows].  FM uses the same
 * layout with row_len = n_factors (the reference has no FM save).
 * Multi-GPU handles write / read the SAME single-model file: ftrl_save_model is called on ONE attached
 * rank (it reads every shard through peer memory; the other ranks must be idle, i.e. after a host barrier
 * behind ftrl_sync), ftrl_load_model on EVERY rank (each keeps the rows it owns and the replicated bias). */
int ftrl_save_model(ftrl_handle *h, const char *path, int compress_level);
int ftrl_load_model(ftrl_handle *h, const char *path);
/* Replaces FFM::save_model / load_model (ffm.cpp:161-200): text, line 1 bias, n_feats
 * lines lin_w, n_feats lines of row_len space-separated values. */
int ftrl_save_model_text(ftrl_handle *h, const char *path);
int ftrl_load_model_text(ftrl_handle *h, const char *path);

/* Page-locked host memory for the CSR staging buffers of the host program (the parser packs
 * samples straight into these so ftrl_train_batch's copies are truly asynchronous). */
void *ftrl_alloc_pinned(size_t bytes);
void ftrl_free_pinned(void *p);

/* ---- measurement hooks (no reference counterpart) ------------------------------ */
/* Run all work of this handle on an existing CUDA stream (cudaStream_t passed as
 * void*), e.g. the caller's current stream, so caller-side CUDA events bracket it. */
int ftrl_set_stream(ftrl_handle *h, void *cuda_stream);
/* Per-phase device timing with CUDA events on the launching stream.
 * ftrl_profile_enable(h, 1) starts collecting; ftrl_profile_read returns, for phase i,
 * its name, accumulated milliseconds and launch count; returns FTRL_ERR_ARG past the end. */
int ftrl_profile_enable(ftrl_handle *h, int on);
int ftrl_profile_reset(ftrl_handle *h);
int ftrl_profile_read(ftrl_handle *h, int i, const char **name, double *ms, int64_t *launches);
/* Statistics of the last trained batch: distinct feature rows U, valid occurrences nnz,
 * rows finalised in the fused per-sample kernel, rows sent through the segmented path. */
typedef struct ftrl_batch_stats {
  int64_t n_rows, nnz_valid, n_unique, n_fused_rows, n_segmented_rows, n_chunks;
  int64_t kernel_launches; /* kernels launched by the last train call */
  int64_t reserved[5];
} ftrl_batch_stats;
int ftrl_last_batch_stats(ftrl_handle *h, ftrl_batch_stats *out);

/* Overwrites the FTRL accumulators with a synthetic warm state: z ~ N(0, z_scale), n ~ U(n_lo, n_hi)
 * for bias, linear and latent coordinates (generated on device).  From a cold start the reference's
 * latent vectors never leave 0; benchmarks and large-scale parity tests use this to
 * exercise live latent arithmetic without shipping hundreds of GB through the host. */
int ftrl_randomize_state(ftrl_handle *h, uint64_t seed, float z_scale, float n_lo, float n_hi);

/* ---- multi-GPU (feature-sharded tables, peer memory over NVLink) ----------------- */
/* Replaces the shared-memory Hogwild workers of src/task/ftrl_offline.cpp:63-103 across GPUs: the samples of a
 * global minibatch are split over the ranks, rows live on rank feat % world_size (config.rank / world_size,
 * LR, FM or FFM in minibatch mode, max_batch_rows / max_batch_nnz fixed at ftrl_create; any samples: an FFM
 * batch in which a sample repeats a field takes the generic kernels on every rank).  After ftrl_attach_peers every
 * ftrl_train_batch(_device) call is COLLECTIVE: all ranks call it once per step with their share of the
 * minibatch; the result equals one GPU training on the concatenated minibatch up to fp32 re-association of
 * the per-rank partial sums (deterministic for a given world_size).  A rank that stops calling makes the
 * others fail with FTRL_ERR_STATE after FTRL_B200_BARRIER_TIMEOUT_S (default 20 s) instead of hanging.
 *
 * ftrl_export_peer_blob: opaque, fixed-size descriptor of this rank's device buffers that peers can map
 * (cudaIpcMemHandle-based).  Exchange the blobs out of band (e.g. an all-gather over torch.distributed /
 * MPI), then hand all world_size blobs, in rank order, to ftrl_attach_peers.  No host-side barrier is needed
 * between ftrl_attach_peers and the first collective call: a rank that is still attaching is waited for by
 * the device-side barrier of the step. */
#define FTRL_PEER_BLOB_BYTES 1024
int ftrl_export_peer_blob(ftrl_handle *h, void *blob /* FTRL_PEER_BLOB_BYTES */);
int ftrl_attach_peers(ftrl_handle *h, const void *blobs /* world_size * FTRL_PEER_BLOB_BYTES */);

#ifdef __cplusplus
}
#endif
#endif /* FTRL_B200_H */
