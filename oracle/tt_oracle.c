/*
 * tt_oracle.c -- CPU restatement of the reference's TT-embedding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call it, and only as the checker (or as the timed CPU baseline), never as a
 * fallback of the CUDA path.
 *
 * Each function cites the reference lines it follows (paths relative to the reference tree).
 * Parity pin: oracle/pin_against_reference.py checks this file against the reference's own
 * pure-PyTorch contraction tt_matrix_to_full (FBTT/tt_embeddings_ops.py:80-127), autograd
 * through it (the commented assertions of sage_profiler.py:303-305,362-367), and the hash
 * known-answer vectors obtained from FBTT/hashtbl_cuda_utils.cuh; the vectors
 * are committed under tests/golden/.
 *
 * Arithmetic: values are accumulated in double and rounded to float once, so this is the
 * mathematically exact reference the fp32 GPU kernels are compared to within 1e-5 relative.
 * Index / hash / partition logic is integer and must match bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXT 4

typedef struct {
  int T, num_tables;
  int p[MAXT], q[MAXT], r[MAXT + 1];
  int cols[MAXT];
  int64_t L[MAXT];
  int64_t num_rows;
  int D;
} shape_t;

static void make_shape(shape_t* s, int T, int num_tables, const int* p, const int* q, const int* r) {
  memset(s, 0, sizeof(*s));
  s->T = T;
  s->num_tables = num_tables;
  s->D = 1;
  s->num_rows = 1;
  for (int t = 0; t < T; ++t) {
    s->p[t] = p[t];
    s->q[t] = q[t];
    s->r[t] = r[t];
    s->cols[t] = r[t] * q[t] * r[t + 1];
    s->D *= q[t];
    s->num_rows *= p[t];
  }
  s->r[T] = r[T];
  /* L = [p1*p2, p2, 1]  FBTT/tt_embeddings_ops.py:519-527 */
  int64_t L = 1;
  for (int t = T - 1; t >= 0; --t) {
    s->L[t] = L;
    L *= p[t];
  }
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline asks for every core */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* index split  FBTT/tt_embeddings_cuda.cu:798-802 */
static void split_index(const shape_t* s, int64_t idx, int* it) {
  int64_t rem = idx;
  for (int t = 0; t < s->T; ++t) {
    it[t] = (int)(rem / s->L[t]);
    rem = rem % s->L[t];
  }
}

/* One row of the chain: X_0 = core0[i0] ([q0, r1]); X_t = X_{t-1}.view[m, r_t] * core_t[i_t]
 * ([r_t, q_t r_{t+1}]).  This is what the T-1 cublasGemmBatchedEx calls compute per index
 * (FBTT/tt_embeddings_cuda.cu:1045-1061) and what tt_matrix_to_full does for the whole table
 * (FBTT/tt_embeddings_ops.py:106-126).  X[t] must hold prod(q[0..t]) * r[t+1] doubles. */
static void chain_forward(const shape_t* s, const float* const* cores, int64_t tidx, const int* it,
                          double** X) {
  const float* c0 = cores[0] + ((int64_t)tidx * s->p[0] + it[0]) * s->cols[0];
  for (int o = 0; o < s->cols[0]; ++o) X[0][o] = c0[o];
  int m = s->q[0];
  for (int t = 1; t < s->T; ++t) {
    const int k = s->r[t], n = s->q[t] * s->r[t + 1];
    const float* c = cores[t] + ((int64_t)tidx * s->p[t] + it[t]) * s->cols[t];
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < n; ++b) {
        double acc = 0.0;
        for (int kk = 0; kk < k; ++kk) acc += X[t - 1][a * k + kk] * (double)c[kk * n + b];
        X[t][a * n + b] = acc;
      }
    m *= s->q[t];
  }
}

static size_t x_len(const shape_t* s, int t) {
  size_t m = 1;
  for (int u = 0; u <= t; ++u) m *= s->q[u];
  return m * s->r[t + 1];
}

/* tt_embeddings_forward_cuda  FBTT/tt_embeddings_cuda.cu:967-1081:
 * output = zeros[num_tables, B, D]; output[tableidx[n], rowidx[n], :] += TT_row(indices[n]) */
int orc_tt_forward(int T, int num_tables, const int* p, const int* q, const int* r, int64_t B,
                   int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                   const int64_t* tableidx, const float* const* cores, float* output) {
  shape_t s;
  make_shape(&s, T, num_tables, p, q, r);
  const size_t out_elems = (size_t)num_tables * B * s.D;
  double* acc = (double*)calloc(out_elems, sizeof(double));
  if (!acc) return -1;
#pragma omp parallel
  {
    double* X[MAXT];
    for (int t = 0; t < T; ++t) X[t] = (double*)malloc(sizeof(double) * x_len(&s, t));
#pragma omp for schedule(static)
    for (int64_t n = 0; n < nnz; ++n) {
      int it[MAXT];
      split_index(&s, indices[n], it);
      chain_forward(&s, cores, tableidx[n], it, X);
      double* o = acc + ((size_t)tableidx[n] * B + rowidx[n]) * s.D;
      for (int d = 0; d < s.D; ++d) {
#pragma omp atomic
        o[d] += X[T - 1][d];
      }
    }
    for (int t = 0; t < T; ++t) free(X[t]);
  }
  for (size_t i = 0; i < out_elems; ++i) output[i] = (float)acc[i];
  free(acc);
  return 0;
}

/* fp32 variant with one index per output row and no accumulation buffer: the CPU baseline
 * that bench.py times (same arithmetic order as the per-row cuBLAS GEMMs, float FMAs). */
int orc_tt_forward_f32_rows(int T, const int* p, const int* q, const int* r, int64_t nnz,
                            const int64_t* indices, const float* const* cores, float* output) {
  shape_t s;
  make_shape(&s, T, 1, p, q, r);
#pragma omp parallel
  {
    float* X[MAXT];
    for (int t = 0; t < T; ++t) X[t] = (float*)malloc(sizeof(float) * x_len(&s, t));
#pragma omp for schedule(static)
    for (int64_t n = 0; n < nnz; ++n) {
      int it[MAXT];
      split_index(&s, indices[n], it);
      const float* c0 = cores[0] + (int64_t)it[0] * s.cols[0];
      memcpy(X[0], c0, sizeof(float) * s.cols[0]);
      int m = s.q[0];
      for (int t = 1; t < T; ++t) {
        const int k = s.r[t], nn = s.q[t] * s.r[t + 1];
        const float* c = cores[t] + (int64_t)it[t] * s.cols[t];
        float* dst = (t == T - 1) ? output + n * s.D : X[t];
        for (int a = 0; a < m; ++a) {
          for (int b = 0; b < nn; ++b) dst[a * nn + b] = 0.f;
          for (int kk = 0; kk < k; ++kk) {
            const float xv = X[t - 1][a * k + kk];
            for (int b = 0; b < nn; ++b) dst[a * nn + b] += xv * c[kk * nn + b];
          }
        }
        m *= s.q[t];
      }
    }
    for (int t = 0; t < T; ++t) free(X[t]);
  }
  return 0;
}

/* tt_embeddings_backward_cuda  FBTT/tt_embeddings_cuda.cu:421-654 (dense part):
 * recompute the chain (:531-547); for t = T-1..1: d_core_t[i_t] += X_{t-1}^T dX_t (:550-578),
 * dX_{t-1} = dX_t core_t[i_t]^T (:579-593); d_core_0[i_0] += dX_0 (:594-609).
 * d_cores[t] is [num_tables][p_t][cols_t], fully overwritten. */
int orc_tt_backward_dense(int T, int num_tables, const int* p, const int* q, const int* r, int64_t B,
                          int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                          const int64_t* tableidx, const float* d_output,
                          const float* const* cores, float* const* d_cores) {
  shape_t s;
  make_shape(&s, T, num_tables, p, q, r);
  size_t csz[MAXT], ctot = 0, coff[MAXT];
  for (int t = 0; t < T; ++t) {
    csz[t] = (size_t)num_tables * s.p[t] * s.cols[t];
    coff[t] = ctot;
    ctot += csz[t];
  }
  double* total = (double*)calloc(ctot, sizeof(double));
  if (!total) return -1;
#pragma omp parallel
  {
    double* X[MAXT];
    size_t maxlen = 0;
    for (int t = 0; t < T; ++t) {
      X[t] = (double*)malloc(sizeof(double) * x_len(&s, t));
      if (x_len(&s, t) > maxlen) maxlen = x_len(&s, t);
    }
    double* dA = (double*)malloc(sizeof(double) * maxlen);
    double* dB = (double*)malloc(sizeof(double) * maxlen);
    double* priv = (double*)calloc(ctot, sizeof(double));
#pragma omp for schedule(static)
    for (int64_t n = 0; n < nnz; ++n) {
      int it[MAXT];
      split_index(&s, indices[n], it);
      const int64_t tidx = tableidx[n];
      chain_forward(&s, cores, tidx, it, X);
      const float* g = d_output + ((size_t)tidx * B + rowidx[n]) * s.D;
      for (int d = 0; d < s.D; ++d) dA[d] = g[d];
      double *dcur = dA, *dnxt = dB;
      int m = 1;
      for (int t = 0; t < T - 1; ++t) m *= s.q[t];
      for (int t = T - 1; t >= 1; --t) {
        const int k = s.r[t], nn = s.q[t] * s.r[t + 1];
        const float* c = cores[t] + ((int64_t)tidx * s.p[t] + it[t]) * s.cols[t];
        double* gd = priv + coff[t] + ((size_t)tidx * s.p[t] + it[t]) * s.cols[t];
        for (int kk = 0; kk < k; ++kk)
          for (int b = 0; b < nn; ++b) {
            double acc = 0.0;
            for (int a = 0; a < m; ++a) acc += X[t - 1][a * k + kk] * dcur[a * nn + b];
            gd[kk * nn + b] += acc;
          }
        for (int a = 0; a < m; ++a)
          for (int kk = 0; kk < k; ++kk) {
            double acc = 0.0;
            for (int b = 0; b < nn; ++b) acc += dcur[a * nn + b] * (double)c[kk * nn + b];
            dnxt[a * k + kk] = acc;
          }
        double* tmp = dcur;
        dcur = dnxt;
        dnxt = tmp;
        if (t > 1) m /= s.q[t - 1];
      }
      double* g0 = priv + coff[0] + ((size_t)tidx * s.p[0] + it[0]) * s.cols[0];
      for (int o = 0; o < s.cols[0]; ++o) g0[o] += dcur[o];
    }
#pragma omp critical
    for (size_t i = 0; i < ctot; ++i) total[i] += priv[i];
    free(priv);
    free(dA);
    free(dB);
    for (int t = 0; t < T; ++t) free(X[t]);
  }
  for (int t = 0; t < T; ++t)
    for (size_t i = 0; i < csz[t]; ++i) d_cores[t][i] = (float)total[coff[t] + i];
  free(total);
  return 0;
}

/* fp32 backward for the timed CPU baseline: one index per row (rowidx[n] = n), one table, float
 * arithmetic like the reference's fp32 GEMMs, thread-private gradient copies instead of
 * atomics.  Same maths as orc_tt_backward_dense. */
int orc_tt_backward_f32_rows(int T, const int* p, const int* q, const int* r, int64_t nnz,
                             const int64_t* indices, const float* d_output,
                             const float* const* cores, float* const* d_cores) {
  shape_t s;
  make_shape(&s, T, 1, p, q, r);
  size_t csz[MAXT], ctot = 0, coff[MAXT];
  for (int t = 0; t < T; ++t) {
    csz[t] = (size_t)s.p[t] * s.cols[t];
    coff[t] = ctot;
    ctot += csz[t];
  }
  for (int t = 0; t < T; ++t) memset(d_cores[t], 0, sizeof(float) * csz[t]);
#pragma omp parallel
  {
    float* X[MAXT];
    size_t maxlen = 0;
    for (int t = 0; t < T; ++t) {
      X[t] = (float*)malloc(sizeof(float) * x_len(&s, t));
      if (x_len(&s, t) > maxlen) maxlen = x_len(&s, t);
    }
    float* dA = (float*)malloc(sizeof(float) * maxlen);
    float* dB = (float*)malloc(sizeof(float) * maxlen);
    float* priv = (float*)calloc(ctot, sizeof(float));
#pragma omp for schedule(static)
    for (int64_t n = 0; n < nnz; ++n) {
      int it[MAXT];
      split_index(&s, indices[n], it);
      memcpy(X[0], cores[0] + (int64_t)it[0] * s.cols[0], sizeof(float) * s.cols[0]);
      int m = s.q[0];
      for (int t = 1; t < T - 1; ++t) {
        const int k = s.r[t], nn = s.q[t] * s.r[t + 1];
        const float* c = cores[t] + (int64_t)it[t] * s.cols[t];
        for (int a = 0; a < m; ++a) {
          for (int b = 0; b < nn; ++b) X[t][a * nn + b] = 0.f;
          for (int kk = 0; kk < k; ++kk) {
            const float xv = X[t - 1][a * k + kk];
            for (int b = 0; b < nn; ++b) X[t][a * nn + b] += xv * c[kk * nn + b];
          }
        }
        m *= s.q[t];
      }
      memcpy(dA, d_output + n * s.D, sizeof(float) * s.D);
      float *dcur = dA, *dnxt = dB;
      for (int t = T - 1; t >= 1; --t) {
        const int k = s.r[t], nn = s.q[t] * s.r[t + 1];
        const float* c = cores[t] + (int64_t)it[t] * s.cols[t];
        float* gd = priv + coff[t] + (size_t)it[t] * s.cols[t];
        for (int a = 0; a < m; ++a)
          for (int kk = 0; kk < k; ++kk) {
            const float xv = X[t - 1][a * k + kk];
            float acc = 0.f;
            for (int b = 0; b < nn; ++b) {
              gd[kk * nn + b] += xv * dcur[a * nn + b];
              acc += dcur[a * nn + b] * c[kk * nn + b];
            }
            dnxt[a * k + kk] = acc;
          }
        float* tmp = dcur;
        dcur = dnxt;
        dnxt = tmp;
        if (t > 1) m /= s.q[t - 1];
      }
      float* g0 = priv + coff[0] + (size_t)it[0] * s.cols[0];
      for (int o = 0; o < s.cols[0]; ++o) g0[o] += dcur[o];
    }
#pragma omp critical
    for (int t = 0; t < T; ++t)
      for (size_t i = 0; i < csz[t]; ++i) d_cores[t][i] += priv[coff[t] + i];
    free(priv);
    free(dA);
    free(dB);
    for (int t = 0; t < T; ++t) free(X[t]);
  }
  return 0;
}

/* update_tt_cores_sgd_kernel / update_tt_cores_adagrad_kernel formulas
 * FBTT/tt_embeddings_cuda.cu:381-419, on every element (SURVEY.md 8a-6 documents that the
 * reference's launch skips tail rows; rows_limit reproduces that: only rows < rows_limit[t] of
 * core t are updated, pass a huge value for the intended full update). optim: 0 sgd, 1 adagrad */
void orc_apply_optimizer(int T, int num_tables, const int* p, const int* cols, int optim, float lr,
                         float eps, const int64_t* rows_limit, float* const* cores,
                         float* const* state, const float* const* d_cores) {
  for (int t = 0; t < T; ++t)
    for (int tb = 0; tb < num_tables; ++tb)
      for (int row = 0; row < p[t]; ++row) {
        if (rows_limit && row >= rows_limit[t]) continue;
        for (int d = 0; d < cols[t]; ++d) {
          const size_t o = ((size_t)tb * p[t] + row) * cols[t] + d;
          const float g = d_cores[t][o];
          if (optim == 0) {
            cores[t][o] -= lr * g;
          } else {
            state[t][o] += g * g;
            cores[t][o] -= lr * g / (sqrtf(state[t][o]) + eps);
          }
        }
      }
}

/* ---------------------------------------------------------------------------------------
 * hash table  FBTT/hashtbl_cuda_utils.cuh
 * ------------------------------------------------------------------------------------- */
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* murmor_hash_3_32(int64_t, int32_t)  FBTT/hashtbl_cuda_utils.cuh:48-76 */
uint32_t orc_hash(int64_t key, int32_t C) {
  uint32_t h = 0;
  uint32_t w[2];
  memcpy(w, &key, 8);
  for (int i = 0; i < 2; ++i) {
    uint32_t k = w[i];
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5 + 0xe6546b64u;
  }
  h ^= 2;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return (uint32_t)(((uint64_t)h * (uint64_t)(uint32_t)C) >> 32);
}

#define MAX_PROBES 3 /* FBTT/tt_embeddings_cuda.cu:31 */

/* hashtbl_find  FBTT/hashtbl_cuda_utils.cuh:135-154 */
static int32_t tbl_find(int64_t key, int32_t size, const int64_t* keys) {
  int32_t s = (int32_t)orc_hash(key, size);
  for (int c = 0; c < MAX_PROBES; ++c) {
    if (keys[s] == key) return s;
    if (key == -1) return -1;
    s = (s + 1) % size;
  }
  return -1;
}

/* update_cache_state_kernel + hashtbl_insert<accumulate=true>
 * FBTT/tt_embeddings_cuda.cu:1083-1095, FBTT/hashtbl_cuda_utils.cuh:102-133, in sequential
 * index order (the reference's winner under collisions is thread-order dependent). */
void orc_update_cache_state(int64_t nnz, const int64_t* indices, int32_t size, int64_t* keys,
                            int64_t* freq) {
  for (int64_t n = 0; n < nnz; ++n) {
    const int64_t key = indices[n];
    int32_t s = (int32_t)orc_hash(key, size);
    for (int c = 0; c < MAX_PROBES; ++c) {
      if (keys[s] == -1) keys[s] = key; /* atomicCAS(-1 -> key) */
      if (keys[s] == key) {
        freq[s] += 1;
        break;
      }
      s = (s + 1) % size;
    }
  }
}

typedef struct {
  int64_t freq, key;
  int32_t slot;
} fk_t;

static int cmp_freq_desc(const void* a, const void* b) {
  const fk_t* x = (const fk_t*)a;
  const fk_t* y = (const fk_t*)b;
  if (x->freq != y->freq) return x->freq > y->freq ? -1 : 1;
  return x->slot < y->slot ? -1 : (x->slot > y->slot ? 1 : 0); /* stable: slot order */
}

/* cache_populate_cuda  FBTT/tt_embeddings_cuda.cu:1270-1347: stable radix sort by freq
 * descending (:1286-1319), mark_popular_colidx_kernel (:1122-1149); sorted_keys_out[size]
 * receives the sorted key list after the "filler 0" fix-up (the rows prefetched into
 * cache_weight are TT_row(sorted_keys_out[n]) for n < cache_size, :1166-1268). */
void orc_cache_populate_index(int32_t size, int32_t cache_size, int64_t* keys, int64_t* freq,
                              int32_t* cache_state, int64_t* sorted_keys_out) {
  fk_t* a = (fk_t*)malloc(sizeof(fk_t) * (size_t)size);
  for (int32_t i = 0; i < size; ++i) {
    a[i].freq = freq[i];
    a[i].key = keys[i];
    a[i].slot = i;
  }
  qsort(a, (size_t)size, sizeof(fk_t), cmp_freq_desc);
  for (int32_t n = 0; n < size; ++n) sorted_keys_out[n] = a[n].key;
  free(a);
  for (int32_t n = 0; n < size; ++n) {
    if (sorted_keys_out[n] != -1) {
      const int32_t s = tbl_find(sorted_keys_out[n], size, keys);
      if (s < 0) continue;
      if (n < cache_size) {
        cache_state[s] = n;
      } else {
        keys[s] = -1;
        freq[s] = 0;
      }
    } else if (n < cache_size) {
      sorted_keys_out[n] = 0;
    }
  }
}

/* preprocess_indices_sync_cuda  FBTT/tt_embeddings_cuda.cu:1388-1507.
 * rowidx/tableidx from offsets (:1349-1365); if !warmup && num_tables == 1: lookup (:1367-1386)
 * and cub::DevicePartition::Flagged x3 (:1448-1490): selected (TT) items first in order, the
 * rejected (cached) items at the back in REVERSE order.  Returns nnz_tt. */
int64_t orc_preprocess_indices(int64_t nnz, int64_t num_offsets, const int64_t* colidx,
                               const int64_t* offsets, int num_tables, int warmup, int32_t size,
                               const int64_t* keys, const int32_t* cache_state, int64_t* rowidx,
                               int64_t* tableidx, int64_t* part_col, int64_t* part_row,
                               int32_t* part_loc) {
  const int64_t num_bags = num_offsets - 1;
  const int64_t B = num_bags / num_tables;
  for (int64_t b = 0; b < B * num_tables; ++b)
    for (int64_t l = offsets[b]; l < offsets[b + 1]; ++l) {
      rowidx[l] = b % B;
      tableidx[l] = b / B;
    }
  if (warmup || num_tables != 1) return nnz;
  int64_t nsel = 0, nrej = 0;
  for (int64_t n = 0; n < nnz; ++n) {
    const int32_t s = tbl_find(colidx[n], size, keys);
    int32_t loc = -1;
    if (s != -1 && cache_state[s] != -1) loc = cache_state[s];
    if (loc == -1) {
      part_col[nsel] = colidx[n];
      part_row[nsel] = rowidx[n];
      part_loc[nsel] = -1;
      ++nsel;
    } else {
      const int64_t pos = nnz - 1 - nrej;
      part_col[pos] = colidx[n];
      part_row[pos] = rowidx[n];
      part_loc[pos] = loc;
      ++nrej;
    }
  }
  return nsel;
}

/* cache_forward_kernel  FBTT/tt_embeddings_cuda.cu:1509-1549 (float, same order) */
void orc_cache_forward(int64_t nnz, int D, const int32_t* loc, const int64_t* rowidx,
                       const float* weight, float* output) {
  for (int64_t n = 0; n < nnz; ++n)
    for (int d = 0; d < D; ++d) output[rowidx[n] * D + d] += weight[(int64_t)loc[n] * D + d];
}

/* cache_backward_sgd_kernel :1585-1632 (mode 0) / cache_backward_dense_kernel :1670-1708 (mode 1) */
void orc_cache_backward(int64_t nnz, int D, const float* grad_output, const int32_t* loc,
                        const int64_t* rowidx, float lr, int mode, float* dst) {
  for (int64_t n = 0; n < nnz; ++n)
    for (int d = 0; d < D; ++d) {
      const float g = grad_output[rowidx[n] * D + d];
      dst[(int64_t)loc[n] * D + d] += (mode == 0) ? (-g * lr) : g;
    }
}

/* cache_backward_rowwise_adagrad_approx_kernel :1746-1806, sequential order */
void orc_cache_backward_rowwise_adagrad(int64_t nnz, int D, const float* grad_output,
                                        const int32_t* loc, const int64_t* rowidx, float lr,
                                        float eps, float* state, float* weight) {
  for (int64_t n = 0; n < nnz; ++n) {
    float sq = 0.f;
    for (int d = 0; d < D; ++d) {
      const float g = grad_output[rowidx[n] * D + d];
      sq += g * g;
    }
    const float g_avg = sq / D;
    const float old = state[loc[n]];
    state[loc[n]] = old + g_avg;
    const float mult = lr * (1.0f / (sqrtf(old + g_avg) + eps));
    for (int d = 0; d < D; ++d)
      weight[(int64_t)loc[n] * D + d] -= grad_output[rowidx[n] * D + d] * mult;
  }
}

/* Efficient_TT index decomposition  Efficient_TT/efficient_tt_cuda.cu:189-213 with the
 * reference's int/float arithmetic (use_float != 0) or exact integer arithmetic. out[n*4..] =
 * {group, I1, I2, I3}. */
void orc_eff_split(int64_t nnz, const int64_t* indices, const int64_t* p, int use_float,
                   int32_t* out) {
  for (int64_t n = 0; n < nnz; ++n) {
    int idx = (int)indices[n];
    int group, I1, I2, I3;
    if (use_float) {
      float tmp = (float)idx / (float)p[2];
      group = (int)floorf(tmp);
      I3 = (int)(idx % p[2]);
      I1 = (int)floorf((float)group / (float)p[1]);
      I2 = (int)(group % p[1]);
    } else {
      group = (int)(indices[n] / p[2]);
      I3 = (int)(indices[n] % p[2]);
      I1 = (int)(group / p[1]);
      I2 = (int)(group % p[1]);
    }
    out[n * 4 + 0] = group;
    out[n * 4 + 1] = I1;
    out[n * 4 + 2] = I2;
    out[n * 4 + 3] = I3;
  }
}

/* neighbour aggregation: out[v] = scale_v * sum_e w_e x[src(e)]   (SAGEConv 'mean' /
 * GraphConv sum; gnn_model.py:78-81,211-214,287 -- DGL 2.1.0 semantics restated, un-vendored) */
void orc_spmm_csr_fwd(int64_t num_dst, int F, const int64_t* indptr, const int32_t* indices,
                      const float* ew, int mean, const float* x, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < num_dst; ++v) {
    const int64_t e0 = indptr[v], e1 = indptr[v + 1];
    for (int d = 0; d < F; ++d) {
      double acc = 0.0;
      for (int64_t e = e0; e < e1; ++e)
        acc += (double)(ew ? ew[e] : 1.0f) * (double)x[(int64_t)indices[e] * F + d];
      if (mean && e1 > e0) acc /= (double)(e1 - e0);
      out[v * F + d] = (float)acc;
    }
  }
}

void orc_spmm_csr_bwd(int64_t num_dst, int64_t num_src, int F, const int64_t* indptr,
                      const int32_t* indices, const float* ew, int mean, const float* dout,
                      float* dx) {
  double* acc = (double*)calloc((size_t)num_src * F, sizeof(double));
  for (int64_t v = 0; v < num_dst; ++v) {
    const int64_t e0 = indptr[v], e1 = indptr[v + 1];
    if (e1 <= e0) continue;
    const double scale = mean ? 1.0 / (double)(e1 - e0) : 1.0;
    for (int64_t e = e0; e < e1; ++e)
      for (int d = 0; d < F; ++d)
        acc[(int64_t)indices[e] * F + d] +=
            scale * (double)(ew ? ew[e] : 1.0f) * (double)dout[v * F + d];
  }
  for (int64_t i = 0; i < num_src * F; ++i) dx[i] = (float)acc[i];
  free(acc);
}
