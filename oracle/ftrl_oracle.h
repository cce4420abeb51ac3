/*
 * ftrl_oracle.h -- CPU oracle for the FTRL LR / FM / FFM training path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * algorithm (massquantity/Ftrl-FFM, src/model/{ftrl_model,lr,fm,ffm}.cpp) used
 * as the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * ftrl-ffm_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *  (a) bit-for-bit against the reference itself compiled from
 *      /root/reference into oracle/_ref/libftrl_ref.so (when present), and
 *  (b) bit-for-bit against tests/golden/ fixtures generated from that build by
 *      tests/golden/make_golden.py, which travel to the GPU box.
 *
 * The very same C interface (prefix ftrl_ref_ instead of ftrl_oracle_) is
 * exported by oracle/ref_shim.cpp around the reference's own C++ classes.
 */
#ifndef FTRL_ORACLE_H
#define FTRL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FTRL_ORACLE_LR = 0, FTRL_ORACLE_FM = 1, FTRL_ORACLE_FFM = 2 };

typedef struct ftrl_oracle ftrl_oracle;

/* mirrors FtrlModel/FM/FFM constructors (ftrl_model.cpp:12-34, fm.cpp:9-19,
 * ffm.cpp:17-28) except that w is zero-filled instead of Gaussian; callers set
 * w through the state pointers. */
ftrl_oracle *ftrl_oracle_create(int model_type, int n_feats, int n_fields, int n_factors,
                                float w_alpha, float w_beta, float w_l1, float w_l2);
void ftrl_oracle_destroy(ftrl_oracle *o);

/* direct views of the state.  which: 0 = w, 1 = n, 2 = z.
 *  bias : 3 floats {bias, bias_n, bias_z}
 *  lin  : n_feats floats
 *  vec  : n_feats * row_len floats, row_len = n_factors (FM) or
 *         n_fields*n_factors (FFM, index field*k+f), NULL for LR        */
float *ftrl_oracle_bias(ftrl_oracle *o);
float *ftrl_oracle_lin(ftrl_oracle *o, int which);
float *ftrl_oracle_vec(ftrl_oracle *o, int which);
int64_t ftrl_oracle_row_len(const ftrl_oracle *o);

/* One reference `train(feat_vec&, int)` call (lr.cpp:9-18, fm.cpp:21-32,
 * ffm.cpp:38-49): returns the pre-update logit. */
float ftrl_oracle_train(ftrl_oracle *o, int nnz, const int32_t *field, const int32_t *feat,
                        const float *val, int label);
/* One reference `predict(feat_vec&, bool)` call. */
float ftrl_oracle_predict(ftrl_oracle *o, int nnz, const int32_t *field, const int32_t *feat,
                          const float *val, int output_prob);

/* Sequential loop over a CSR block in row order: what one reference worker
 * does (ftrl_offline.cpp:74-83).  logits_out may be NULL.  Returns the fp64
 * sum of loss(y, logit) (eval/loss.h:8-12). */
double ftrl_oracle_train_csr(ftrl_oracle *o, int64_t n_rows, const int64_t *row_ptr,
                             const int32_t *field, const int32_t *feat, const float *val,
                             const int32_t *label, float *logits_out);
double ftrl_oracle_predict_csr(ftrl_oracle *o, int64_t n_rows, const int64_t *row_ptr,
                               const int32_t *field, const int32_t *feat, const float *val,
                               const int32_t *label, int output_prob, float *out);

/* DERIVED (not in the reference): minibatch semantics of SURVEY.md 8(a).
 * Every sample of the block sees w materialised from the (n,z) at block start
 * (what concurrent Hogwild workers of the reference observe); each coordinate
 * then receives the telescoped reference recurrence
 *   n' = n + sum g^2 ,  z' = (z + sum g) - w * (sqrt(n') - sqrt(n)) / alpha.
 * Sums are accumulated in fp64 in sample order.  At n_rows == 1 this equals
 * ftrl_oracle_train up to the ffm.cpp:118 quirk. */
double ftrl_oracle_train_batch_csr(ftrl_oracle *o, int64_t n_rows, const int64_t *row_ptr,
                                   const int32_t *field, const int32_t *feat, const float *val,
                                   const int32_t *label, float *logits_out);

/* scalar helpers pinned by the reference's tests/test_utils.cpp */
double ftrl_oracle_loss(int y, double logit);     /* eval/loss.h:8-12   */
float ftrl_oracle_sigmoid(float x);               /* utils.h:20-23      */
float ftrl_oracle_sgn(float x);                   /* utils.h:16-18      */
float ftrl_oracle_weight(const ftrl_oracle *o, float n, float z); /* ftrl_model.h:29-33 */

#ifdef __cplusplus
}
#endif
#endif
