/* ORACLE (C restatement) — test infrastructure only; never linked into the product.
 *
 * Plain-C float32 restatement of the gradient path the reference delegates to torch, used (a) as a
 * second checker next to oracle/ddp_oracle.py and (b) inside the timed CPU baseline of bench.py
 * (`cpu_baseline.kind = "port"`), spread over the host cores with pthreads (this image's gcc has no
 * libgomp).  Citations as in ddp_oracle.py:
 *   ddp_mean_reduce : accelerator.py:1892 -> torch reducer.hpp:276 (pre-divide) + :524 (allreduce SUM)
 *   sumsq / clip    : accelerator.py:3007 -> torch/nn/utils/clip_grad.py
 *   unscale         : accelerator.py:2944 -> torch/amp/grad_scaler.py _unscale_grads_
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define MAX_THREADS 256

typedef struct {
  void (*fn)(size_t lo, size_t hi, void* arg, int tid);
  void* arg;
  size_t lo, hi;
  int tid;
} job_t;

static void* trampoline(void* p) {
  job_t* j = (job_t*)p;
  j->fn(j->lo, j->hi, j->arg, j->tid);
  return NULL;
}

static int thread_count(size_t n) {
  const char* e = getenv("ORACLE_THREADS");
  long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  if (t < 1) t = 1;
  if (t > MAX_THREADS) t = MAX_THREADS;
  if ((size_t)t > n / 65536 + 1) t = (long)(n / 65536 + 1); /* not worth a thread below 64K elements */
  return (int)t;
}

/* contiguous static partition: deterministic for a given thread count */
static int parallel_for(size_t n, void (*fn)(size_t, size_t, void*, int), void* arg) {
  int t = thread_count(n);
  pthread_t th[MAX_THREADS];
  job_t jobs[MAX_THREADS];
  size_t per = (n + t - 1) / t;
  for (int i = 0; i < t; ++i) {
    size_t lo = per * i, hi = lo + per < n ? lo + per : n;
    if (lo > n) lo = n;
    jobs[i] = (job_t){fn, arg, lo, hi, i};
    if (i > 0) pthread_create(&th[i], NULL, trampoline, &jobs[i]);
  }
  trampoline(&jobs[0]);
  for (int i = 1; i < t; ++i) pthread_join(th[i], NULL);
  return t;
}

int oracle_threads(size_t n) { return thread_count(n); }

/* ---- out[i] = sum_r fl(g_r[i] * fl(1/N)), r ascending, float32 arithmetic, no FMA contraction ---- */
typedef struct { const float* const* grads; int world; float* out; } reduce_arg;
static void reduce_body(size_t lo, size_t hi, void* p, int tid) {
  (void)tid;
  reduce_arg* a = (reduce_arg*)p;
  const float inv = 1.0f / (float)a->world;
  for (size_t i = lo; i < hi; ++i) {
    float acc = a->grads[0][i] * inv;
    for (int r = 1; r < a->world; ++r) {
      float term = a->grads[r][i] * inv; /* -ffp-contract=off: mul and add stay separate roundings */
      acc = acc + term;
    }
    a->out[i] = acc;
  }
}
void oracle_ddp_mean_reduce_f32(const float* const* grads, int world, size_t n, float* out) {
  reduce_arg a = {grads, world, out};
  parallel_for(n, reduce_body, &a);
}

typedef struct { const float* x; double part[MAX_THREADS]; float pmax[MAX_THREADS]; int bad[MAX_THREADS]; } stat_arg;
static void stat_body(size_t lo, size_t hi, void* p, int tid) {
  stat_arg* a = (stat_arg*)p;
  double s = 0.0;
  float m = 0.f;
  int bad = 0;
  for (size_t i = lo; i < hi; ++i) {
    float v = a->x[i];
    s += (double)v * (double)v;
    float av = fabsf(v);
    if (av > m) m = av;
    bad |= !isfinite(v);
  }
  a->part[tid] = s;
  a->pmax[tid] = m;
  a->bad[tid] = bad;
}
double oracle_sumsq_f32(const float* x, size_t n) {
  static __thread stat_arg a;
  a.x = x;
  int t = parallel_for(n, stat_body, &a);
  double s = 0.0;
  for (int i = 0; i < t; ++i) s += a.part[i];
  return s;
}
float oracle_maxabs_f32(const float* x, size_t n) {
  static __thread stat_arg a;
  a.x = x;
  int t = parallel_for(n, stat_body, &a);
  float m = 0.f;
  for (int i = 0; i < t; ++i) if (a.pmax[i] > m) m = a.pmax[i];
  return m;
}
int oracle_has_nonfinite_f32(const float* x, size_t n) {
  static __thread stat_arg a;
  a.x = x;
  int t = parallel_for(n, stat_body, &a);
  int bad = 0;
  for (int i = 0; i < t; ++i) bad |= a.bad[i];
  return bad;
}

/* torch: clamp(max_norm / (total_norm + 1e-6), max=1.0) in float32 */
float oracle_clip_coef(float total_norm, float max_norm) {
  float coef = max_norm / (total_norm + 1e-6f);
  return coef > 1.0f ? 1.0f : coef;
}

/* torch: scale.double().reciprocal().float() */
float oracle_inv_scale(float scale) { return (float)(1.0 / (double)scale); }

typedef struct { float* x; float m; } scale_arg;
static void scale_body(size_t lo, size_t hi, void* p, int tid) {
  (void)tid;
  scale_arg* a = (scale_arg*)p;
  for (size_t i = lo; i < hi; ++i) a->x[i] = a->x[i] * a->m;
}
void oracle_scale_f32(float* x, size_t n, float m) {
  scale_arg a = {x, m};
  parallel_for(n, scale_body, &a);
}
static void clamp_body(size_t lo, size_t hi, void* p, int tid) {
  (void)tid;
  scale_arg* a = (scale_arg*)p;
  const float v = a->m;
  for (size_t i = lo; i < hi; ++i) a->x[i] = a->x[i] < -v ? -v : (a->x[i] > v ? v : a->x[i]);
}
void oracle_clamp_f32(float* x, size_t n, float v) {
  scale_arg a = {x, v};
  parallel_for(n, clamp_body, &a);
}

/* One Adam / AdamW update in float32 with the roundings torch's CUDA foreach functors have (a + alpha*b inside one
 * functor is a single fma; sqrt and division correctly rounded) — reference: optimizer.py:145-181 -> torch/optim/adam.py
 * `_multi_tensor_adam`.  This file is built with -ffp-contract=off, so every fma below is explicit.
 *   decay_mul = 1 - lr*wd (AdamW) or 1; l2 = wd (Adam) or 0; lerp_w = float(1-beta1); omb2 = float(1-beta2);
 *   neg_step = -lr/(1-beta1^t); bc2_sqrt = sqrt(1-beta2^t); mult = clip coefficient x 1/loss-scale. */
typedef struct {
  float *p, *m, *v; const float* g;
  float decay_mul, l2, lerp_w, beta2, omb2, neg_step, bc2_sqrt, eps, mult;
} adam_arg;
static void adam_body(size_t lo, size_t hi, void* ptr, int tid) {
  (void)tid;
  adam_arg* a = (adam_arg*)ptr;
  for (size_t i = lo; i < hi; ++i) {
    float p = a->p[i], m = a->m[i], v = a->v[i];
    float g = a->g[i] * a->mult;
    if (a->decay_mul != 1.0f) p = p * a->decay_mul;
    if (a->l2 != 0.0f) g = fmaf(a->l2, p, g);
    const float diff = g - m;
    m = fabsf(a->lerp_w) < 0.5f ? fmaf(a->lerp_w, diff, m) : g - diff * (1.0f - a->lerp_w);
    v = v * a->beta2;
    v = fmaf(a->omb2, g * g, v);
    const float denom = sqrtf(v) / a->bc2_sqrt + a->eps;
    p = fmaf(a->neg_step, m / denom, p);
    a->p[i] = p; a->m[i] = m; a->v[i] = v;
  }
}
void oracle_adam_step_f32(float* p, const float* g, float* m, float* v, size_t n, double decay_mul, double l2,
                          double beta1, double beta2, double neg_step, double bc2_sqrt, double eps, float mult) {
  adam_arg a = {p, m, v, g, (float)decay_mul, (float)l2, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                (float)neg_step, (float)bc2_sqrt, (float)eps, mult};
  parallel_for(n, adam_body, &a);
}
