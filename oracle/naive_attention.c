/*
 * oracle/naive_attention.c — TEST INFRASTRUCTURE ONLY (the checker, never the product path).
 *
 * CPU restatement of the reference's naive attention, fp32 throughout:
 *   code/cuda_fa1/main.cu:165-202   the five steps of attention_baseline_kernel
 *       1. s_k = (q . k_k) * softmax_scale          (:166-174, scale 1/sqrt(d) from :216)
 *       2. max over k                                (:177-180)
 *       3. e_k = expf(s_k - max), sum                (:183-187)
 *       4. normalise                                 (:190-192)
 *       5. O = sum_k p_k v_k                         (:195-201)
 *   code/cuda_fa1/flashAttention.cu:115-120,137-138  meaning of the per-row outputs
 *       m_i = max_k s_ik,  l_i = sum_k exp(s_ik - m_i)      =>  lse_i = m_i + ln l_i
 *   code/triton_fa2/FA2-triton.py:70-73              causal rule: key col > query row -> -inf
 *       (generalised to N_kv != N with bottom-right alignment: col > row + (N_kv - N))
 *
 * The reference itself has no CPU implementation; its own check is the GPU
 * kernel above.  Parity pin: tests/test_oracle.py checks this file against golden vectors produced
 * by the reference's Python oracle sdpa_reference (FA2-triton.py:311-323) run in the build
 * container (tests/golden/make_golden.py), and against the reference's seeded input stream.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this code.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const float *Q, *K, *V;
  float *O, *lse, *l, *m;
  int BH, Nq, Nkv, d, causal;
  float scale;
  /* rows [row_begin, row_end) of every (b,h) slice are computed: lets bench.py time a bounded sample */
  int row_begin, row_end;
  int tid, nthreads;
} job_t;

static void attention_rows(const job_t* j) {
  const int Nq = j->Nq, Nkv = j->Nkv, d = j->d;
  const int off = Nkv - Nq;
  float* s = (float*)malloc(sizeof(float) * (size_t)(Nkv > 0 ? Nkv : 1));
  const int rows_per_bh = j->row_end - j->row_begin;
  const long total = (long)j->BH * rows_per_bh;
  for (long w = j->tid; w < total; w += j->nthreads) {
    const int bh = (int)(w / rows_per_bh);
    const int r = j->row_begin + (int)(w % rows_per_bh);
    const float* q = j->Q + ((size_t)bh * Nq + r) * d;
    const float* Kb = j->K + (size_t)bh * Nkv * d;
    const float* Vb = j->V + (size_t)bh * Nkv * d;
    float* o = j->O + ((size_t)bh * Nq + r) * d;
    int nk = Nkv;
    if (j->causal) {
      nk = r + off + 1;
      if (nk > Nkv) nk = Nkv;
      if (nk < 0) nk = 0;
    }
    /* step 1: scores */
    float mx = -INFINITY;
    for (int k = 0; k < nk; ++k) {
      const float* kr = Kb + (size_t)k * d;
      float acc = 0.f;
      for (int t = 0; t < d; ++t) acc += q[t] * kr[t];
      acc *= j->scale;
      s[k] = acc;
      if (acc > mx) mx = acc; /* step 2 */
    }
    /* step 3 */
    float sum = 0.f;
    for (int k = 0; k < nk; ++k) {
      s[k] = expf(s[k] - mx);
      sum += s[k];
    }
    /* steps 4 + 5 (normalisation folded into the output row) */
    for (int t = 0; t < d; ++t) o[t] = 0.f;
    for (int k = 0; k < nk; ++k) {
      const float p = s[k];
      const float* vr = Vb + (size_t)k * d;
      for (int t = 0; t < d; ++t) o[t] += p * vr[t];
    }
    const float inv = (sum > 0.f) ? 1.f / sum : 0.f; /* l == 0 -> 0, flashAttention.cu:146 */
    for (int t = 0; t < d; ++t) o[t] *= inv;
    const size_t so = (size_t)bh * Nq + r;
    if (j->m) j->m[so] = (nk > 0) ? mx : -INFINITY;
    if (j->l) j->l[so] = sum;
    if (j->lse) j->lse[so] = (nk > 0) ? mx + logf(sum) : -INFINITY;
  }
  free(s);
}

static void* worker(void* p) {
  attention_rows((const job_t*)p);
  return NULL;
}

/* O [BH,Nq,d]; lse/l/m [BH,Nq] may be NULL.  scale <= 0 -> 1/sqrt(d).  nthreads <= 0 -> 1.
 * Computes rows [row_begin,row_end) of every slice (row_end <= 0 -> Nq).  Returns 0. */
int oracle_attention_f32(const float* Q, const float* K, const float* V, float* O, float* lse, float* l,
                         float* m, int BH, int Nq, int Nkv, int d, int causal, float scale, int nthreads,
                         int row_begin, int row_end) {
  if (!Q || !K || !V || !O || BH <= 0 || Nq <= 0 || Nkv < 0 || d <= 0) return 1;
  if (row_end <= 0 || row_end > Nq) row_end = Nq;
  if (row_begin < 0 || row_begin >= row_end) row_begin = 0;
  if (nthreads <= 0) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (scale <= 0.f) scale = 1.0f / sqrtf((float)d);
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    job_t jb = {Q, K, V, O, lse, l, m, BH, Nq, Nkv, d, causal, scale, row_begin, row_end, t, nthreads};
    jobs[t] = jb;
  }
  if (nthreads == 1) {
    attention_rows(&jobs[0]);
    return 0;
  }
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* Symmetric relative error of the reference's verify step: max |a-b| / (|a|+|b|+1e-5)
 * (code/cuda_fa1/main.cu:319-347, gate < 0.02). */
float oracle_max_symmetric_rel_err(const float* a, const float* b, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) {
    const float e = fabsf(a[i] - b[i]) / (fabsf(a[i]) + fabsf(b[i]) + 1e-5f);
    if (e > mx) mx = e;
  }
  return mx;
}

/* Merge of two partials over disjoint key sets (ring attention): the fp64 restatement of
 * lse' = log(e^a + e^b), O' = O_a e^(a-lse') + O_b e^(b-lse').  In place on (O_a, lse_a). */
void oracle_merge_partial(float* O_a, float* lse_a, const float* O_b, const float* lse_b, size_t rows, int d) {
  for (size_t r = 0; r < rows; ++r) {
    const double a = lse_a[r], b = lse_b[r];
    if (b == -INFINITY) continue;
    const double mx = a > b ? a : b;
    const double ea = (a == -INFINITY) ? 0.0 : exp(a - mx), eb = exp(b - mx);
    const double wa = ea / (ea + eb), wb = eb / (ea + eb);
    for (int t = 0; t < d; ++t)
      O_a[r * d + t] = (float)(O_a[r * d + t] * wa + O_b[r * d + t] * wb);
    lse_a[r] = (float)(mx + log(ea + eb));
  }
}
