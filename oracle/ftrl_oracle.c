/*
 * ftrl_oracle.c -- plain-C restatement of the reference's per-sample FTRL
 * training path for LR / FM / FFM.  TEST INFRASTRUCTURE ONLY (see ftrl_oracle.h).
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  The arithmetic is fp32 with each operation individually
 * rounded and in the reference's association order; build with
 * -ffp-contract=off and without -ffast-math (oracle/Makefile does).
 * The reference binary is built -O3 for baseline x86-64 (no FMA), so a
 * bit-for-bit match with oracle/_ref is expected and tested.
 *
 * Layout: flat arrays instead of the reference's vector<vector<float>>;
 * vec[which][feat * row_len + field * k + f] for FFM, [feat * k + f] for FM.
 */
#include "ftrl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct ftrl_oracle {
  int model_type, n_feats, n_fields, k;
  int64_t row_len;
  float alpha, beta, l1, l2;
  float bias[3];  /* bias, bias_n, bias_z  (ftrl_model.h:36,45,46) */
  float *lin[3];  /* w, n, z               (ftrl_model.h:37,47,48) */
  float *vec[3];  /* w, n, z               (fm.h:22-27, ffm.h:25-31) */
  float *sum_vx;  /* FM scratch member     (fm.h:24) */
  /* per-call scratch: filtered sample */
  int cap;
  int32_t *s_field, *s_feat;
  float *s_val;
  /* batch-mode scratch (derived semantics) */
  double *acc_lin;     /* 2 per feature: sum g, sum g^2 */
  double *acc_vec;     /* 2 per latent coordinate        */
  double acc_bias[2];
  unsigned char *touched_lin, *touched_vec;
  int64_t *list_lin, *list_vec;
  int64_t n_list_lin, n_list_vec, cap_list_lin, cap_list_vec;
};

/* utils.h:16-18 -- sgn(0) is -1 */
float ftrl_oracle_sgn(float x) { return x > 0 ? 1.0f : -1.0f; }

/* utils.h:20-23 with T = float: 1 / (1 + std::exp(-x)) -> expf */
float ftrl_oracle_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

/* eval/loss.h:8-12: fp64, no clipping */
double ftrl_oracle_loss(int y, double logit) {
  const double s = 1.0 / (1.0 + exp(-logit));
  return -y * log(s) - (1 - y) * log(1 - s);
}

/* ftrl_model.h:29-33, T = float.  The `-1.0 *` literal promotes numerator and
 * quotient to double; the result is narrowed to float on return. */
static inline float weight_of(const ftrl_oracle *o, float n, float z) {
  if (fabsf(z) <= o->l1) return (float)0.0;
  const float num = z - ftrl_oracle_sgn(z) * o->l1;
  const float den = o->l2 + (o->beta + sqrtf(n)) / o->alpha;
  return (float)((-1.0 * (double)num) / (double)den);
}
float ftrl_oracle_weight(const ftrl_oracle *o, float n, float z) { return weight_of(o, n, z); }

ftrl_oracle *ftrl_oracle_create(int model_type, int n_feats, int n_fields, int n_factors,
                                float w_alpha, float w_beta, float w_l1, float w_l2) {
  if (model_type < 0 || model_type > 2 || n_feats <= 0) return NULL;
  ftrl_oracle *o = (ftrl_oracle *)calloc(1, sizeof(*o));
  if (!o) return NULL;
  o->model_type = model_type;
  o->n_feats = n_feats;
  o->n_fields = n_fields;
  o->k = n_factors;
  o->alpha = w_alpha;
  o->beta = w_beta;
  o->l1 = w_l1;
  o->l2 = w_l2;
  o->row_len = model_type == FTRL_ORACLE_FFM  ? (int64_t)n_fields * n_factors
               : model_type == FTRL_ORACLE_FM ? (int64_t)n_factors
                                              : 0;
  for (int t = 0; t < 3; t++) {
    o->lin[t] = (float *)calloc((size_t)n_feats, sizeof(float));
    if (o->row_len) o->vec[t] = (float *)calloc((size_t)n_feats * (size_t)o->row_len, sizeof(float));
  }
  if (model_type == FTRL_ORACLE_FM) o->sum_vx = (float *)calloc((size_t)n_factors, sizeof(float));
  return o;
}

void ftrl_oracle_destroy(ftrl_oracle *o) {
  if (!o) return;
  for (int t = 0; t < 3; t++) {
    free(o->lin[t]);
    free(o->vec[t]);
  }
  free(o->sum_vx);
  free(o->s_field);
  free(o->s_feat);
  free(o->s_val);
  free(o->acc_lin);
  free(o->acc_vec);
  free(o->touched_lin);
  free(o->touched_vec);
  free(o->list_lin);
  free(o->list_vec);
  free(o);
}

float *ftrl_oracle_bias(ftrl_oracle *o) { return o->bias; }
float *ftrl_oracle_lin(ftrl_oracle *o, int which) { return o->lin[which]; }
float *ftrl_oracle_vec(ftrl_oracle *o, int which) { return o->vec[which]; }
int64_t ftrl_oracle_row_len(const ftrl_oracle *o) { return o->row_len; }

/* ftrl_model.cpp:36-42 (LR, FM) and ffm.cpp:30-36 (FFM): drop out-of-range
 * features, keep order.  Returns the filtered length; output in o->s_*. */
static int filter_sample(ftrl_oracle *o, int nnz, const int32_t *field, const int32_t *feat,
                         const float *val) {
  if (nnz > o->cap) {
    o->cap = nnz + 16;
    o->s_field = (int32_t *)realloc(o->s_field, sizeof(int32_t) * (size_t)o->cap);
    o->s_feat = (int32_t *)realloc(o->s_feat, sizeof(int32_t) * (size_t)o->cap);
    o->s_val = (float *)realloc(o->s_val, sizeof(float) * (size_t)o->cap);
  }
  int m = 0;
  for (int t = 0; t < nnz; t++) {
    const int32_t i = feat[t], fl = field[t];
    int drop = i < 0 || i >= o->n_feats;
    if (o->model_type == FTRL_ORACLE_FFM) drop = drop || fl < 0 || fl >= o->n_fields;
    if (drop) continue;
    o->s_field[m] = fl;
    o->s_feat[m] = i;
    o->s_val[m] = val[t];
    m++;
  }
  return m;
}

/* ftrl_model.cpp:52-64 */
static void materialise_linear(ftrl_oracle *o, int F) {
  for (int t = 0; t < F; t++) {
    const int i = o->s_feat[t];
    o->lin[0][i] = weight_of(o, o->lin[1][i], o->lin[2][i]);
  }
  o->bias[0] = weight_of(o, o->bias[1], o->bias[2]);
}

/* ftrl_model.cpp:44-50 */
static float linear_logit(const ftrl_oracle *o, int F) {
  float acc = o->bias[0];
  for (int t = 0; t < F; t++) acc = acc + o->lin[0][o->s_feat[t]] * o->s_val[t];
  return acc;
}

/* ftrl_model.cpp:66-85 */
static void update_linear_nz(ftrl_oracle *o, int F, float g) {
  for (int t = 0; t < F; t++) {
    const int i = o->s_feat[t];
    const float wi = o->lin[0][i];
    const float ni = o->lin[1][i];
    const float gi = g * o->s_val[t];
    const float si = (sqrtf(ni + gi * gi) - sqrtf(ni)) / o->alpha;
    o->lin[2][i] += gi - si * wi;
    o->lin[1][i] += gi * gi;
  }
  {
    const float gi = g;
    const float si = (sqrtf(o->bias[1] + gi * gi) - sqrtf(o->bias[1])) / o->alpha;
    o->bias[2] += gi - si * o->bias[0];
    o->bias[1] += gi * gi;
  }
}

/* ffm.cpp:72-88 */
static void ffm_materialise(ftrl_oracle *o, int F) {
  const int k = o->k;
  float *w = o->vec[0], *nn = o->vec[1], *zz = o->vec[2];
  for (int m = 0; m < F; m++)
    for (int n = m + 1; n < F; n++) {
      const int64_t a = (int64_t)o->s_feat[m] * o->row_len + (int64_t)o->s_field[n] * k;
      const int64_t b = (int64_t)o->s_feat[n] * o->row_len + (int64_t)o->s_field[m] * k;
      for (int f = 0; f < k; f++) {
        w[a + f] = weight_of(o, nn[a + f], zz[a + f]);
        w[b + f] = weight_of(o, nn[b + f], zz[b + f]);
      }
    }
}

/* ffm.cpp:57-70 */
static float ffm_logit(const ftrl_oracle *o, int F) {
  const int k = o->k;
  const float *w = o->vec[0];
  float result = linear_logit(o, F);
  for (int m = 0; m < F; m++)
    for (int n = m + 1; n < F; n++) {
      const float *a = w + (int64_t)o->s_feat[m] * o->row_len + (int64_t)o->s_field[n] * k;
      const float *b = w + (int64_t)o->s_feat[n] * o->row_len + (int64_t)o->s_field[m] * k;
      float dot = 0.0f;
      for (int f = 0; f < k; f++) dot = dot + a[f] * b[f];
      result += dot * o->s_val[m] * o->s_val[n];
    }
  return result;
}

/* ffm.cpp:90-136, including the line-118 term sqrtf(n2 + g2*g1). */
static void ffm_update_nz(ftrl_oracle *o, int F, float g) {
  const int k = o->k;
  float *w = o->vec[0], *nn = o->vec[1], *zz = o->vec[2];
  float *tmp = (float *)malloc(sizeof(float) * 4 * (size_t)(k > 0 ? k : 1));
  float *zi1 = tmp, *ni1 = tmp + k, *ni2 = tmp + 2 * k, *zi2 = tmp + 3 * k;
  for (int m = 0; m < F; m++)
    for (int n = m + 1; n < F; n++) {
      const float x = o->s_val[m] * o->s_val[n];
      const int64_t a = (int64_t)o->s_feat[m] * o->row_len + (int64_t)o->s_field[n] * k;
      const int64_t b = (int64_t)o->s_feat[n] * o->row_len + (int64_t)o->s_field[m] * k;
      for (int f = 0; f < k; f++) {
        const float vif1 = w[a + f], v_nif1 = nn[a + f], v_zif1 = zz[a + f];
        const float vif2 = w[b + f], v_nif2 = nn[b + f], v_zif2 = zz[b + f];
        const float v_gif1 = g * vif2 * x;
        const float v_sif1 = (sqrtf(v_nif1 + v_gif1 * v_gif1) - sqrtf(v_nif1)) / o->alpha;
        zi1[f] = v_zif1 + v_gif1 - v_sif1 * vif1;
        ni1[f] = v_nif1 + v_gif1 * v_gif1;
        const float v_gif2 = g * vif1 * x;
        const float v_sif2 = (sqrtf(v_nif2 + v_gif2 * v_gif1) - sqrtf(v_nif2)) / o->alpha;
        zi2[f] = v_zif2 + v_gif2 - v_sif2 * vif2;
        ni2[f] = v_nif2 + v_gif2 * v_gif2;
      }
      /* write-back order of ffm.cpp:129-132: slice 1 then slice 2 */
      memcpy(zz + a, zi1, sizeof(float) * (size_t)k);
      memcpy(nn + a, ni1, sizeof(float) * (size_t)k);
      memcpy(zz + b, zi2, sizeof(float) * (size_t)k);
      memcpy(nn + b, ni2, sizeof(float) * (size_t)k);
    }
  free(tmp);
}

/* fm.cpp:69-78 */
static void fm_materialise(ftrl_oracle *o, int F) {
  const int k = o->k;
  for (int t = 0; t < F; t++) {
    const int64_t a = (int64_t)o->s_feat[t] * k;
    for (int f = 0; f < k; f++) o->vec[0][a + f] = weight_of(o, o->vec[1][a + f], o->vec[2][a + f]);
  }
}

/* fm.cpp:40-67; sum_vx (when given) receives the per-factor sums (train path) */
static float fm_logit(const ftrl_oracle *o, int F, float *sum_vx) {
  const int k = o->k;
  float result = linear_logit(o, F);
  for (int f = 0; f < k; f++) {
    float s_vx = 0.0f, sum_sqr = 0.0f;
    for (int t = 0; t < F; t++) {
      const float vx = o->vec[0][(int64_t)o->s_feat[t] * k + f] * o->s_val[t];
      s_vx += vx;
      sum_sqr += vx * vx;
    }
    if (sum_vx) sum_vx[f] = s_vx;
    result += 0.5f * (s_vx * s_vx - sum_sqr);
  }
  return result;
}

/* fm.cpp:80-101 */
static void fm_update_nz(ftrl_oracle *o, int F, float g) {
  const int k = o->k;
  for (int t = 0; t < F; t++) {
    const int64_t a = (int64_t)o->s_feat[t] * k;
    const float x = o->s_val[t];
    for (int f = 0; f < k; f++) {
      const float vif = o->vec[0][a + f];
      const float v_nif = o->vec[1][a + f];
      const float v_zif = o->vec[2][a + f];
      const float s_vx = o->sum_vx[f];
      const float v_gif = g * (x * s_vx - vif * x * x);
      const float v_sif = (sqrtf(v_nif + v_gif * v_gif) - sqrtf(v_nif)) / o->alpha;
      o->vec[2][a + f] = v_zif + v_gif - v_sif * vif;
      o->vec[1][a + f] = v_nif + v_gif * v_gif;
    }
  }
}

/* lr.cpp:9-18, fm.cpp:21-32, ffm.cpp:38-49 */
float ftrl_oracle_train(ftrl_oracle *o, int nnz, const int32_t *field, const int32_t *feat,
                        const float *val, int label) {
  const int F = filter_sample(o, nnz, field, feat, val);
  materialise_linear(o, F);
  float logit;
  if (o->model_type == FTRL_ORACLE_FFM) {
    ffm_materialise(o, F);
    logit = ffm_logit(o, F);
  } else if (o->model_type == FTRL_ORACLE_FM) {
    fm_materialise(o, F);
    logit = fm_logit(o, F, o->sum_vx);
  } else {
    logit = linear_logit(o, F);
  }
  const float g = ftrl_oracle_sigmoid(logit) - (float)label;
  update_linear_nz(o, F, g);
  if (o->model_type == FTRL_ORACLE_FFM) ffm_update_nz(o, F, g);
  if (o->model_type == FTRL_ORACLE_FM) fm_update_nz(o, F, g);
  return logit;
}

/* lr.cpp:20-24, fm.cpp:34-38, ffm.cpp:51-55 */
float ftrl_oracle_predict(ftrl_oracle *o, int nnz, const int32_t *field, const int32_t *feat,
                          const float *val, int output_prob) {
  const int F = filter_sample(o, nnz, field, feat, val);
  float logit;
  if (o->model_type == FTRL_ORACLE_FFM)
    logit = ffm_logit(o, F);
  else if (o->model_type == FTRL_ORACLE_FM)
    logit = fm_logit(o, F, NULL);
  else
    logit = linear_logit(o, F);
  return output_prob ? ftrl_oracle_sigmoid(logit) : logit;
}

/* ftrl_offline.cpp:74-83 with one worker in row order */
double ftrl_oracle_train_csr(ftrl_oracle *o, int64_t n_rows, const int64_t *row_ptr,
                             const int32_t *field, const int32_t *feat, const float *val,
                             const int32_t *label, float *logits_out) {
  double tmp_loss = 0.0;
  for (int64_t r = 0; r < n_rows; r++) {
    const int64_t b = row_ptr[r];
    const int nnz = (int)(row_ptr[r + 1] - b);
    const float logit = ftrl_oracle_train(o, nnz, field + b, feat + b, val + b, label[r]);
    if (logits_out) logits_out[r] = logit;
    tmp_loss += ftrl_oracle_loss(label[r], logit);
  }
  return tmp_loss;
}

double ftrl_oracle_predict_csr(ftrl_oracle *o, int64_t n_rows, const int64_t *row_ptr,
                               const int32_t *field, const int32_t *feat, const float *val,
                               const int32_t *label, int output_prob, float *out) {
  double tmp_loss = 0.0;
  for (int64_t r = 0; r < n_rows; r++) {
    const int64_t b = row_ptr[r];
    const int nnz = (int)(row_ptr[r + 1] - b);
    const float logit = ftrl_oracle_predict(o, nnz, field + b, feat + b, val + b, 0);
    if (out) out[r] = output_prob ? ftrl_oracle_sigmoid(logit) : logit;
    if (label) tmp_loss += ftrl_oracle_loss(label[r], logit);
  }
  return tmp_loss;
}

/* ------------------------------------------------------------------------
 * DERIVED minibatch semantics -- see header.
 * ---------------------------------------------------------------------- */
static void push_idx(int64_t **list, int64_t *n, int64_t *cap, int64_t v) {
  if (*n == *cap) {
    *cap = *cap ? *cap * 2 : 1024;
    *list = (int64_t *)realloc(*list, sizeof(int64_t) * (size_t)*cap);
  }
  (*list)[(*n)++] = v;
}

static void touch_lin(ftrl_oracle *o, int64_t i) {
  if (!o->touched_lin[i]) {
    o->touched_lin[i] = 1;
    push_idx(&o->list_lin, &o->n_list_lin, &o->cap_list_lin, i);
  }
}
static void touch_vec(ftrl_oracle *o, int64_t c) {
  if (!o->touched_vec[c]) {
    o->touched_vec[c] = 1;
    push_idx(&o->list_vec, &o->n_list_vec, &o->cap_list_vec, c);
  }
}

static float telescoped(const ftrl_oracle *o, float *n, float *z, float w, double sg, double sg2) {
  const float fsg = (float)sg, fsg2 = (float)sg2;
  const float n_new = *n + fsg2;
  const float sigma = (sqrtf(n_new) - sqrtf(*n)) / o->alpha;
  *z = (*z + fsg) - sigma * w;
  *n = n_new;
  return n_new;
}

double ftrl_oracle_train_batch_csr(ftrl_oracle *o, int64_t n_rows, const int64_t *row_ptr,
                                   const int32_t *field, const int32_t *feat, const float *val,
                                   const int32_t *label, float *logits_out) {
  const int k = o->k;
  const int64_t n_coord = (int64_t)o->n_feats * o->row_len;
  if (!o->acc_lin) {
    o->acc_lin = (double *)calloc((size_t)o->n_feats * 2, sizeof(double));
    o->touched_lin = (unsigned char *)calloc((size_t)o->n_feats, 1);
    if (n_coord) {
      o->acc_vec = (double *)calloc((size_t)n_coord * 2, sizeof(double));
      o->touched_vec = (unsigned char *)calloc((size_t)n_coord, 1);
    }
  }
  o->n_list_lin = o->n_list_vec = 0;
  o->acc_bias[0] = o->acc_bias[1] = 0.0;
  double loss_sum = 0.0;
  float *sum_vx = k > 0 ? (float *)malloc(sizeof(float) * (size_t)k) : NULL;

  /* pass 1: materialise w for every coordinate any sample touches, from the
   * (n,z) at block start (n,z are not modified until pass 3). */
  for (int64_t r = 0; r < n_rows; r++) {
    const int64_t b = row_ptr[r];
    const int F = filter_sample(o, (int)(row_ptr[r + 1] - b), field + b, feat + b, val + b);
    materialise_linear(o, F);
    if (o->model_type == FTRL_ORACLE_FFM) ffm_materialise(o, F);
    if (o->model_type == FTRL_ORACLE_FM) fm_materialise(o, F);
  }
  /* pass 2: logits, g, gradient sums */
  for (int64_t r = 0; r < n_rows; r++) {
    const int64_t b = row_ptr[r];
    const int F = filter_sample(o, (int)(row_ptr[r + 1] - b), field + b, feat + b, val + b);
    float logit;
    if (o->model_type == FTRL_ORACLE_FFM)
      logit = ffm_logit(o, F);
    else if (o->model_type == FTRL_ORACLE_FM)
      logit = fm_logit(o, F, sum_vx);
    else
      logit = linear_logit(o, F);
    if (logits_out) logits_out[r] = logit;
    loss_sum += ftrl_oracle_loss(label[r], logit);
    const float g = ftrl_oracle_sigmoid(logit) - (float)label[r];
    o->acc_bias[0] += (double)g;
    o->acc_bias[1] += (double)(g * g);
    for (int t = 0; t < F; t++) {
      const int64_t i = o->s_feat[t];
      const float gi = g * o->s_val[t];
      touch_lin(o, i);
      o->acc_lin[2 * i] += (double)gi;
      o->acc_lin[2 * i + 1] += (double)(gi * gi);
    }
    if (o->model_type == FTRL_ORACLE_FM) {
      for (int t = 0; t < F; t++) {
        const int64_t a = (int64_t)o->s_feat[t] * k;
        const float x = o->s_val[t];
        for (int f = 0; f < k; f++) {
          const float vif = o->vec[0][a + f];
          const float gv = g * (x * sum_vx[f] - vif * x * x);
          touch_vec(o, a + f);
          o->acc_vec[2 * (a + f)] += (double)gv;
          o->acc_vec[2 * (a + f) + 1] += (double)(gv * gv);
        }
      }
    } else if (o->model_type == FTRL_ORACLE_FFM) {
      for (int m = 0; m < F; m++)
        for (int n = m + 1; n < F; n++) {
          const float x = o->s_val[m] * o->s_val[n];
          const int64_t a = (int64_t)o->s_feat[m] * o->row_len + (int64_t)o->s_field[n] * k;
          const int64_t c = (int64_t)o->s_feat[n] * o->row_len + (int64_t)o->s_field[m] * k;
          for (int f = 0; f < k; f++) {
            const float g1 = g * o->vec[0][c + f] * x;
            const float g2 = g * o->vec[0][a + f] * x;
            touch_vec(o, a + f);
            o->acc_vec[2 * (a + f)] += (double)g1;
            o->acc_vec[2 * (a + f) + 1] += (double)(g1 * g1);
            touch_vec(o, c + f);
            o->acc_vec[2 * (c + f)] += (double)g2;
            o->acc_vec[2 * (c + f) + 1] += (double)(g2 * g2);
          }
        }
    }
  }
  /* pass 3: one closed-form update per touched coordinate */
  if (n_rows > 0) telescoped(o, &o->bias[1], &o->bias[2], o->bias[0], o->acc_bias[0], o->acc_bias[1]);
  for (int64_t t = 0; t < o->n_list_lin; t++) {
    const int64_t i = o->list_lin[t];
    telescoped(o, &o->lin[1][i], &o->lin[2][i], o->lin[0][i], o->acc_lin[2 * i], o->acc_lin[2 * i + 1]);
    o->acc_lin[2 * i] = o->acc_lin[2 * i + 1] = 0.0;
    o->touched_lin[i] = 0;
  }
  for (int64_t t = 0; t < o->n_list_vec; t++) {
    const int64_t c = o->list_vec[t];
    telescoped(o, &o->vec[1][c], &o->vec[2][c], o->vec[0][c], o->acc_vec[2 * c], o->acc_vec[2 * c + 1]);
    o->acc_vec[2 * c] = o->acc_vec[2 * c + 1] = 0.0;
    o->touched_vec[c] = 0;
  }
  free(sum_vx);
  return loss_sum;
}
