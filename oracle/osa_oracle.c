/*
 * oracle/osa_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see osa_oracle.h).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile)
 * -ffp-contract=off is REQUIRED: the replay must round every operation exactly
 * like the CUDA kernels, which use explicit __fmaf_rn/__fmul_rn/__fadd_rn.
 */
#include "osa_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

/* ------------------------------------------------------------------------- */
/* Philox4x32-10                                                             */
/* ------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_engine_draw(uint64_t seed, uint64_t traj, uint32_t stream, uint32_t c0, uint32_t c1,
                     uint32_t out[4]) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {c0, c1, (uint32_t)traj,
                     ((uint32_t)(traj >> 32) & 0x3fffffffu) | (stream << 30)};
  orc_philox4x32_10(ctr, key, out);
}

int orc_init_bit(uint64_t seed, uint64_t traj, uint32_t j) {
  uint32_t r[4];
  orc_engine_draw(seed, traj, ORC_STREAM_INIT, j >> 7, 0u, r);
  return (int)((r[(j >> 5) & 3u] >> (j & 31u)) & 1u);
}

/* -ln(u) for u = (2w+1)/2^33.  Mantissa truncated to 24 bits, then the classic
 * single-precision minimax polynomial for ln(1+f) on [sqrt(1/2)-1, sqrt(2)-1]
 * (coefficients as published in Cephes logf), evaluated with explicit fmaf so
 * host and device round identically.                                         */
float orc_neglogf(uint32_t w) {
  uint64_t v = ((uint64_t)w << 1) | 1u;
  int p = 63 - __builtin_clzll(v);
  uint32_t m24 = (uint32_t)((v << (63 - p)) >> 40);
  int e = p - 33;
  float mf = (float)m24 * 0x1p-23f;
  if (m24 > 0x00B504F3u) {
    mf = mf * 0.5f;
    e += 1;
  }
  float f = mf - 1.0f;
  float z = f * f;
  float y = 7.0376836292E-2f;
  y = fmaf(y, f, -1.1514610310E-1f);
  y = fmaf(y, f, 1.1676998740E-1f);
  y = fmaf(y, f, -1.2420140846E-1f);
  y = fmaf(y, f, 1.4249322787E-1f);
  y = fmaf(y, f, -1.6668057665E-1f);
  y = fmaf(y, f, 2.0000714765E-1f);
  y = fmaf(y, f, -2.4999993993E-1f);
  y = fmaf(y, f, 3.3333331174E-1f);
  y = y * f;
  y = y * z;
  float fe = (float)e;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(-0.5f, z, y);
  float r = f + y;
  r = fmaf(fe, 0.693359375f, r);
  return -r;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* (A) reference-faithful restatements                                       */
/* ------------------------------------------------------------------------- */

/* annealing.hpp:31-40 */
double orc_ref_energy(const double *flat_qubo, const char *state, int n) {
  double result = 0.0;
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      result += flat_qubo[(size_t)i * n + j] * state[i] * state[j];
    }
  }
  return result;
}

/* qubo_helpers.hpp:26-44.  The reference loops over all ordered (i, j), i != j,
 * and adds get_connection({i,j}) to BOTH [i+jN] and [j+iN]; a stored pair
 * (a, b) is therefore added once to each side, and a model holding both (a,b)
 * and (b,a) gets their sum on both sides.  Iterating over the stored couplings
 * gives the same result (additions of the same values; order only matters when
 * both orientations are stored, where (i,j) with i<j is visited first).      */
void orc_ref_flatten(int n, const int *lin_idx, const double *lin_val, int n_lin,
                     const int *quad_i, const int *quad_j, const double *quad_val, int n_quad,
                     double *out) {
  memset(out, 0, sizeof(double) * (size_t)n * n);
  for (int k = 0; k < n_lin; ++k) {
    int i = lin_idx[k];
    if (i >= 0 && i < n) out[(size_t)i + (size_t)i * n] = lin_val[k];
  }
  /* two passes keep the reference's visiting order: row-major over (i, j) means a
   * pair with i<j is added before its mirror (j, i) */
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = 0; k < n_quad; ++k) {
      int i = quad_i[k], j = quad_j[k];
      if (i == j || i < 0 || j < 0 || i >= n || j >= n) continue;
      if ((pass == 0) != (i < j)) continue;
      out[(size_t)i + (size_t)j * n] += quad_val[k];
      out[(size_t)j + (size_t)i * n] += quad_val[k];
    }
  }
}

/* one-solver-anneal.cpp:23-29 */
void orc_ref_schedule_linear(double *schedule, double beta_min, double beta_max,
                             unsigned num_iter) {
  for (unsigned i = 0; i < num_iter; i++) {
    schedule[i] = beta_min + beta_max * i / (double)(num_iter - 1);
  }
}

/* one-solver-anneal.cpp:31-39 */
void orc_ref_schedule_geometric(double *schedule, double beta_min, double beta_max,
                                unsigned num_iter) {
  schedule[0] = beta_min;
  double alpha = pow(beta_max / beta_min, 1.0 / (num_iter - 1));
  for (unsigned i = 1; i < num_iter; i++) {
    schedule[i] = (schedule[i - 1]) * alpha;
  }
}

/* annealing.hpp:85-126 for one trajectory */
static double ref_trajectory(const double *flat_qubo, int n, const double *beta_schedule,
                             int num_iter, int sweeps_per_beta, uint64_t seed, uint64_t traj,
                             char *current, char *best) {
  for (int j = 0; j < n; j++) { /* :90-92 */
    best[j] = current[j] = (char)orc_init_bit(seed, traj, (uint32_t)j);
  }
  double best_energy = orc_ref_energy(flat_qubo, current, n); /* :94 */
  double current_energy = best_energy;                        /* :95 */
  uint32_t step = 0;
  for (int iter = 0; iter < num_iter; iter++) { /* :97 */
    double beta = beta_schedule[iter];          /* :98 */
    for (int sweep = 0; sweep < sweeps_per_beta; sweep++, step++) { /* :100 */
      uint32_t r[4];
      orc_engine_draw(seed, traj, ORC_STREAM_RND, 0u, step, r);
      int spin_to_flip = (int)(((uint64_t)r[0] * (uint64_t)n) >> 32); /* :101 bit_index() */
      current[spin_to_flip] = (char)(1 - current[spin_to_flip]);      /* :102-103 */
      double new_energy = orc_ref_energy(flat_qubo, current, n);      /* :104 */
      /* :106-108; u is drawn unconditionally here (counter-based stream) */
      double u = ((double)r[1] + 0.5) * 0x1p-32;
      if ((new_energy < current_energy) || (exp((current_energy - new_energy) / beta) > u)) {
        current_energy = new_energy; /* :109 */
      } else {
        current[spin_to_flip] = (char)(1 - current[spin_to_flip]); /* :111-112 */
      }
      if (current_energy < best_energy) { /* :115-121 */
        best_energy = current_energy;
        for (int j = 0; j < n; j++) best[j] = current[j];
      }
    }
  }
  return best_energy; /* :125 */
}

int64_t orc_ref_anneal(const double *flat_qubo, int n, const double *beta_schedule, int num_iter,
                       uint64_t num_tries, int sweeps_per_beta, uint64_t seed, uint64_t first_try,
                       char *best_states, double *best_energies, int num_threads) {
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
  num_threads = 1;
#endif
  /* annealing.hpp:85-86: one work-item per trajectory */
#pragma omp parallel num_threads(num_threads)
  {
    char *current = (char *)malloc((size_t)n);
#pragma omp for schedule(dynamic, 1)
    for (int64_t t = 0; t < (int64_t)num_tries; ++t) {
      best_energies[t] =
          ref_trajectory(flat_qubo, n, beta_schedule, num_iter, sweeps_per_beta, seed,
                         first_try + (uint64_t)t, current, best_states + (size_t)t * n);
    }
    free(current);
  }
  /* annealing.hpp:134-135: std::min_element -> first minimum */
  int64_t best_idx = 0;
  for (int64_t t = 1; t < (int64_t)num_tries; ++t) {
    if (best_energies[t] < best_energies[best_idx]) best_idx = t;
  }
  return best_idx;
}

/* exhaustive.hpp:44-61 (upper-triangular matrix), :63-90 (ranges), :104-137
 * (kernel), :158-166 (host argmin), ulong_to_vec.hpp:23-32                   */
int orc_ref_exhaustive(const double *flat_qubo, int n, int num_ranges, char *best_state,
                       double *best_energy) {
  if (n <= 0 || n > 30 || num_ranges <= 0) return -1; /* 1 << n_bits is an int shift, :67 */
  uint64_t n_states = (uint64_t)1 << n;
  uint64_t per = n_states / (uint64_t)num_ranges, rem = n_states % (uint64_t)num_ranges;
  double *energies = (double *)malloc(sizeof(double) * (size_t)num_ranges);
  uint64_t *states = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)num_ranges);
  uint64_t *starts = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)num_ranges + 1));
  uint64_t count = 0, rem_count = 0;
  for (int i = 0; i < num_ranges; ++i) {
    starts[i] = count;
    count += per;
    if (rem_count < rem) { count += 1; rem_count++; }
  }
  starts[num_ranges] = count;
#pragma omp parallel for schedule(dynamic, 1)
  for (int item = 0; item < num_ranges; ++item) {
    double e_best = 1.7976931348623157e308; /* numeric_limits<double>::max(), :106 */
    uint64_t state_best = 0;
    for (uint64_t state = starts[item]; state < starts[item + 1]; ++state) {
      double e = 0.0;
      for (int i = 0; i < n; ++i) {
        if (!((state >> i) & 1)) continue; /* both branches of :115-128 need bit i */
        for (int j = i; j < n; ++j) {
          if ((state >> j) & 1) e += flat_qubo[(size_t)i * n + j];
        }
      }
      if (e < e_best) { e_best = e; state_best = state; }
    }
    energies[item] = e_best;
    states[item] = state_best;
  }
  int min_idx = 0;
  for (int i = 1; i < num_ranges; ++i)
    if (energies[i] < energies[min_idx]) min_idx = i;
  for (int i = 0; i < n; ++i) best_state[i] = (char)((states[min_idx] >> i) & 1);
  *best_energy = energies[min_idx];
  free(energies); free(states); free(starts);
  return 0;
}

/* solution.hpp:58-67.  ostream<<double with default flags == printf("%g").   */
size_t orc_ref_solution_csv(const char *state, int n, double energy, char *buf, size_t buflen) {
  size_t off = 0;
  for (int i = 0; i < n; ++i) off += (size_t)snprintf(buf + off, off < buflen ? buflen - off : 0, "%d,", i);
  off += (size_t)snprintf(buf + off, off < buflen ? buflen - off : 0, "energy\n");
  for (int i = 0; i < n; ++i)
    off += (size_t)snprintf(buf + off, off < buflen ? buflen - off : 0, "%d,", (int)state[i]);
  off += (size_t)snprintf(buf + off, off < buflen ? buflen - off : 0, "%g\n", energy);
  return off;
}

void orc_energy_packed(const double *qsym, int n, const uint32_t *states_packed, uint64_t count,
                       double *out) {
  int nw = (n + 31) / 32;
#pragma omp parallel
  {
    char *state = (char *)malloc((size_t)n);
#pragma omp for schedule(static)
    for (int64_t t = 0; t < (int64_t)count; ++t) {
      const uint32_t *w = states_packed + (size_t)t * nw;
      for (int i = 0; i < n; ++i) state[i] = (char)((w[i >> 5] >> (i & 31)) & 1u);
      out[t] = orc_ref_energy(qsym, state, n);
    }
    free(state);
  }
}

/* ------------------------------------------------------------------------- */
/* (B) bit-exact host replay                                                 */
/* ------------------------------------------------------------------------- */

/* Flip trace of a trajectory, the value osa_anneal_traced returns (include/onesolver_b200.h):
 * FNV-1a over (step, block of 32 sites, mask of accepted sites) of every block with an accepted
 * flip, in the order the flips happen.  orc_set_trace_output(buf) makes the replays below store
 * the hash of trajectory tl in buf[tl] (NULL switches it off again).                           */
static uint64_t *g_trace_out = NULL;
void orc_set_trace_output(uint64_t *buf) { g_trace_out = buf; }

typedef struct {
  uint64_t h;
  uint32_t step, blk, mask;
} orc_trace;

static inline void trace_init(orc_trace *t) {
  t->h = 0xcbf29ce484222325ull;
  t->step = t->blk = t->mask = 0;
}
static inline void trace_flush(orc_trace *t) {
  if (t->mask) {
    t->h = (t->h ^ (uint64_t)t->step) * 0x100000001b3ull;
    t->h = (t->h ^ (((uint64_t)t->blk << 32) | t->mask)) * 0x100000001b3ull;
  }
  t->mask = 0;
}
static inline void trace_flip(orc_trace *t, uint32_t step, int site) {
  const uint32_t blk = (uint32_t)site >> 5;
  if (t->mask && (t->step != step || t->blk != blk)) trace_flush(t);
  t->step = step;
  t->blk = blk;
  t->mask |= 1u << (site & 31);
}

static inline int getbit(const uint32_t *x, int i) { return (int)((x[i >> 5] >> (i & 31)) & 1u); }
static inline void flipbit(uint32_t *x, int i) { x[i >> 5] ^= (1u << (i & 31)); }

static void init_state_packed(uint64_t seed, uint64_t traj, int n, uint32_t *x) {
  int nw = (n + 31) / 32;
  for (int w = 0; w < nw; ++w) {
    uint32_t r[4];
    orc_engine_draw(seed, traj, ORC_STREAM_INIT, (uint32_t)w >> 2, 0u, r);
    uint32_t word = r[w & 3];
    int valid = n - w * 32;
    if (valid < 32) word &= (valid <= 0) ? 0u : ((1u << valid) - 1u);
    x[w] = word;
  }
}

/* The dense and CSR replays are generated for float and double from one macro
 * body so both precisions follow literally the same statement order.         */
/* The _round variant adds what a resumable launch of the CUDA kernel takes (osa_pt_anneal):    \
 * init_states (NULL: STREAM_INIT), tscale_traj (NULL: tscale[iter]), step_base.             */ \
#define DEFINE_REPLAY_DENSE(NAME, T, FMA)                                                       \
  int NAME##_round(const T *qoff, const T *diag, int n, size_t ld, const T *tscale, int num_iter,\
           int sweeps_per_beta, int mode, uint64_t seed, uint64_t first_try, uint64_t num_tries, \
           int batch_r, double *best_rel, uint32_t *best_states_packed,                          \
           uint32_t *final_states_packed, orc_counters *counters,                                \
           const uint32_t *init_states, const T *tscale_traj, uint32_t step_base);               \
  int NAME(const T *qoff, const T *diag, int n, size_t ld, const T *tscale, int num_iter,       \
           int sweeps_per_beta, int mode, uint64_t seed, uint64_t first_try, uint64_t num_tries, \
           int batch_r, double *best_rel, uint32_t *best_states_packed,                          \
           uint32_t *final_states_packed, orc_counters *counters) {                              \
    return NAME##_round(qoff, diag, n, ld, tscale, num_iter, sweeps_per_beta, mode, seed,        \
                        first_try, num_tries, batch_r, best_rel, best_states_packed,             \
                        final_states_packed, counters, NULL, NULL, 0u);                          \
  }                                                                                              \
  int NAME##_round(const T *qoff, const T *diag, int n, size_t ld, const T *tscale, int num_iter,\
           int sweeps_per_beta, int mode, uint64_t seed, uint64_t first_try, uint64_t num_tries, \
           int batch_r, double *best_rel, uint32_t *best_states_packed,                          \
           uint32_t *final_states_packed, orc_counters *counters,                                \
           const uint32_t *init_states, const T *tscale_traj, uint32_t step_base) {              \
    if (n <= 0 || num_iter <= 0 || sweeps_per_beta <= 0) return -1;                              \
    const int nw = (n + 31) / 32;                                                                \
    if (mode == ORC_MODE_RANDOM_SITE || batch_r < 1) batch_r = 1;                                \
    const uint64_t steps_per_traj =                                                              \
        (uint64_t)num_iter * (uint64_t)sweeps_per_beta *                                         \
        (mode == ORC_MODE_SEQUENTIAL_SWEEP ? (uint64_t)n : 1u);                                  \
    const int64_t n_batches = (int64_t)((num_tries + (uint64_t)batch_r - 1) / (uint64_t)batch_r);\
    uint64_t tot_acc = 0, tot_rows = 0, tot_init_rows = 0;                                       \
    const int want_rows = (counters != NULL) && batch_r > 1;                                     \
    _Pragma("omp parallel for schedule(dynamic,1) reduction(+:tot_acc,tot_rows,tot_init_rows)")  \
    for (int64_t b = 0; b < n_batches; ++b) {                                                    \
      T *h = (T *)malloc(sizeof(T) * (size_t)n);                                                 \
      uint32_t *x = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nw);                           \
      uint32_t *xb = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nw);                          \
      uint8_t *any = want_rows ? (uint8_t *)calloc(steps_per_traj + (uint64_t)n, 1) : NULL;      \
      for (int rr = 0; rr < batch_r; ++rr) {                                                     \
        uint64_t tl = (uint64_t)b * (uint64_t)batch_r + (uint64_t)rr;                            \
        if (tl >= num_tries) break;                                                              \
        uint64_t traj = first_try + tl;                                                          \
        if (init_states) memcpy(x, init_states + (size_t)tl * nw, sizeof(uint32_t) * (size_t)nw);\
        else init_state_packed(seed, traj, n, x);                                                \
        memcpy(xb, x, sizeof(uint32_t) * (size_t)nw);                                            \
        /* initial local field: h = diag, then add rows of set spins in index order */           \
        for (int j = 0; j < n; ++j) h[j] = diag[j];                                              \
        for (int i = 0; i < n; ++i) {                                                            \
          if (!getbit(x, i)) continue;                                                           \
          const T *row = qoff + (size_t)i * ld;                                                  \
          for (int j = 0; j < n; ++j) h[j] = FMA((T)1, row[j], h[j]);                            \
          if (any) any[steps_per_traj + (uint64_t)i] = 1;                                        \
        }                                                                                        \
        double erel = 0.0, best = 0.0;                                                           \
        uint64_t sidx = 0;                                                                       \
        uint32_t step = step_base;                                                               \
        orc_trace tr;                                                                            \
        trace_init(&tr);                                                                         \
        for (int iter = 0; iter < num_iter; ++iter) {                                            \
          const T ts = tscale_traj ? tscale_traj[tl] : tscale[iter];                             \
          for (int sw = 0; sw < sweeps_per_beta; ++sw, ++step) {                                 \
            const int n_sites = (mode == ORC_MODE_SEQUENTIAL_SWEEP) ? n : 1;                     \
            for (int s = 0; s < n_sites; ++s, ++sidx) {                                          \
              uint32_t r[4];                                                                     \
              int k;                                                                             \
              uint32_t wu;                                                                       \
              if (mode == ORC_MODE_SEQUENTIAL_SWEEP) {                                           \
                k = s;                                                                           \
                orc_engine_draw(seed, traj, ORC_STREAM_SEQ, (uint32_t)s >> 2, step, r);          \
                wu = r[s & 3];                                                                   \
              } else {                                                                           \
                orc_engine_draw(seed, traj, ORC_STREAM_RND, 0u, step, r);                        \
                k = (int)(((uint64_t)r[0] * (uint64_t)n) >> 32);                                 \
                wu = r[1];                                                                       \
              }                                                                                  \
              const T theta = ts * (T)orc_neglogf(wu);                                           \
              const int xk = getbit(x, k);                                                       \
              const T dE = xk ? -h[k] : h[k];                                                    \
              if (dE < theta) {                                                                  \
                const T sgn = xk ? (T)-1 : (T)1;                                                 \
                const T *row = qoff + (size_t)k * ld;                                            \
                for (int j = 0; j < n; ++j) h[j] = FMA(sgn, row[j], h[j]);                       \
                flipbit(x, k);                                                                   \
                trace_flip(&tr, step, k);                                                        \
                erel += (double)dE;                                                              \
                tot_acc++;                                                                       \
                if (any) any[sidx] = 1;                                                          \
                if (erel < best) {                                                               \
                  best = erel;                                                                   \
                  memcpy(xb, x, sizeof(uint32_t) * (size_t)nw);                                  \
                }                                                                                \
              }                                                                                  \
            }                                                                                    \
          }                                                                                      \
        }                                                                                        \
        best_rel[tl] = best;                                                                     \
        trace_flush(&tr);                                                                        \
        if (g_trace_out) g_trace_out[tl] = tr.h;                                                 \
        if (best_states_packed) memcpy(best_states_packed + (size_t)tl * nw, xb, sizeof(uint32_t) * (size_t)nw); \
        if (final_states_packed) memcpy(final_states_packed + (size_t)tl * nw, x, sizeof(uint32_t) * (size_t)nw); \
      }                                                                                          \
      if (any) {                                                                                 \
        for (uint64_t q = 0; q < steps_per_traj; ++q) tot_rows += any[q];                        \
        for (int q = 0; q < n; ++q) tot_init_rows += any[steps_per_traj + (uint64_t)q];          \
        free(any);                                                                               \
      }                                                                                          \
      free(h); free(x); free(xb);                                                                \
    }                                                                                            \
    if (counters) {                                                                              \
      counters->attempts = steps_per_traj * num_tries;                                           \
      counters->accepts = tot_acc;                                                               \
      counters->row_fetches = want_rows ? tot_rows : tot_acc;                                    \
      counters->init_row_fetches = tot_init_rows;                                                \
    }                                                                                            \
    return 0;                                                                                    \
  }

DEFINE_REPLAY_DENSE(orc_replay_dense_f32, float, fmaf)
DEFINE_REPLAY_DENSE(orc_replay_dense_f64, double, fma)

#define DEFINE_REPLAY_CSR(NAME, T)                                                               \
  int NAME(const int32_t *rowptr, const int32_t *col, const T *val, const T *diag, int n,        \
           const T *tscale, int num_iter, int sweeps_per_beta, int mode, uint64_t seed,          \
           uint64_t first_try, uint64_t num_tries, double *best_rel,                             \
           uint32_t *best_states_packed, uint32_t *final_states_packed,                          \
           orc_counters *counters) {                                                             \
    if (n <= 0 || num_iter <= 0 || sweeps_per_beta <= 0) return -1;                              \
    const int nw = (n + 31) / 32;                                                                \
    const uint64_t steps_per_traj =                                                              \
        (uint64_t)num_iter * (uint64_t)sweeps_per_beta *                                         \
        (mode == ORC_MODE_SEQUENTIAL_SWEEP ? (uint64_t)n : 1u);                                  \
    uint64_t tot_acc = 0;                                                                        \
    _Pragma("omp parallel for schedule(dynamic,4) reduction(+:tot_acc)")                         \
    for (int64_t tl = 0; tl < (int64_t)num_tries; ++tl) {                                        \
      uint32_t *x = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nw);                           \
      uint32_t *xb = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nw);                          \
      uint64_t traj = first_try + (uint64_t)tl;                                                  \
      init_state_packed(seed, traj, n, x);                                                       \
      memcpy(xb, x, sizeof(uint32_t) * (size_t)nw);                                              \
      double erel = 0.0, best = 0.0;                                                             \
      uint32_t step = 0;                                                                         \
      orc_trace tr;                                                                              \
      trace_init(&tr);                                                                           \
      for (int iter = 0; iter < num_iter; ++iter) {                                              \
        const T ts = tscale[iter];                                                               \
        for (int sw = 0; sw < sweeps_per_beta; ++sw, ++step) {                                   \
          const int n_sites = (mode == ORC_MODE_SEQUENTIAL_SWEEP) ? n : 1;                       \
          for (int s = 0; s < n_sites; ++s) {                                                    \
            uint32_t r[4];                                                                       \
            int k;                                                                               \
            uint32_t wu;                                                                         \
            if (mode == ORC_MODE_SEQUENTIAL_SWEEP) {                                             \
              k = s;                                                                             \
              orc_engine_draw(seed, traj, ORC_STREAM_SEQ, (uint32_t)s >> 2, step, r);            \
              wu = r[s & 3];                                                                     \
            } else {                                                                             \
              orc_engine_draw(seed, traj, ORC_STREAM_RND, 0u, step, r);                          \
              k = (int)(((uint64_t)r[0] * (uint64_t)n) >> 32);                                   \
              wu = r[1];                                                                         \
            }                                                                                    \
            const T theta = ts * (T)orc_neglogf(wu);                                             \
            /* local field recomputed in CSR order: h = diag + sum_{p} val[p]*x[col[p]] */       \
            T hk = diag[k];                                                                      \
            for (int32_t p = rowptr[k]; p < rowptr[k + 1]; ++p)                                  \
              if (getbit(x, col[p])) hk = hk + val[p];                                           \
            const int xk = getbit(x, k);                                                         \
            const T dE = xk ? -hk : hk;                                                          \
            if (dE < theta) {                                                                    \
              flipbit(x, k);                                                                     \
              trace_flip(&tr, step, k);                                                          \
              erel += (double)dE;                                                                \
              tot_acc++;                                                                         \
              if (erel < best) {                                                                 \
                best = erel;                                                                     \
                memcpy(xb, x, sizeof(uint32_t) * (size_t)nw);                                    \
              }                                                                                  \
            }                                                                                    \
          }                                                                                      \
        }                                                                                        \
      }                                                                                          \
      trace_flush(&tr);                                                                          \
      if (g_trace_out) g_trace_out[tl] = tr.h;                                                   \
      best_rel[tl] = best;                                                                       \
      if (best_states_packed) memcpy(best_states_packed + (size_t)tl * nw, xb, sizeof(uint32_t) * (size_t)nw); \
      if (final_states_packed) memcpy(final_states_packed + (size_t)tl * nw, x, sizeof(uint32_t) * (size_t)nw); \
      free(x); free(xb);                                                                         \
    }                                                                                            \
    if (counters) {                                                                              \
      counters->attempts = steps_per_traj * num_tries;                                           \
      counters->accepts = tot_acc;                                                               \
      counters->row_fetches = tot_acc;                                                           \
      counters->init_row_fetches = 0;                                                            \
    }                                                                                            \
    return 0;                                                                                    \
  }

DEFINE_REPLAY_CSR(orc_replay_csr_f32, float)
DEFINE_REPLAY_CSR(orc_replay_csr_f64, double)

/* ------------------------------------------------------------------------- */
/* (C) population annealing: the resampling step of osa_pa_anneal            */
/* ------------------------------------------------------------------------- */

/* exp(x) for x <= 0, the fixed operation sequence of the engine (onesolver_b200/csrc/osa_pa.cu,
 * det_exp): k = rint(x log2 e), two-part ln 2 reduction, degree-13 Taylor polynomial in Horner
 * form with fma, exact scaling by 2^k.  Written from the definition in the C ABI header; every
 * operation is correctly rounded, so host and device agree bit for bit.                        */
double orc_det_exp(double x) {
  if (!(x >= -60.0)) return 0.0;
  const double kf = rint(x * 1.4426950408889634074);
  double r = fma(-kf, 6.93147180369123816490e-01, x);
  r = fma(-kf, 1.90821492927058770002e-10, r);
  static const double c[12] = {2.08767569878681e-09,   2.505210838544172e-08, 2.755731922398589e-07,
                               2.7557319223985893e-06, 2.48015873015873e-05,  1.984126984126984e-04,
                               1.388888888888889e-03,  8.333333333333333e-03, 4.1666666666666664e-02,
                               1.6666666666666666e-01, 0.5,                   1.0};
  double p = 1.6059043836821613e-10;
  for (int i = 0; i < 12; ++i) p = fma(p, r, c[i]);
  p = fma(p, r, 1.0);
  const int64_t k = (int64_t)kf;
  union { uint64_t u; double d; } scale;
  scale.u = (uint64_t)(1023 + k) << 52;
  return p * scale.d;
}

/* One population, one resampling step (include/onesolver_b200.h, osa_pa_anneal):
 *   x_i = neg_db * E_i, q_i = floor(2^40 exp(x_i - max x)), C = prefix sums of q, W = C[M-1],
 *   r = floor(W u32 / 2^32), slot j continues from src[j] = min{ i : C_i > floor((j W + r) / M) }.
 * u32 = word 0 of the Philox draw (stream 3, c0 = 0x50410000, c1 = step, id = population).     */
void orc_pa_resample(const double *e, int m, double neg_db, uint64_t seed, uint64_t population,
                     uint32_t step, int32_t *src) {
  double mx = -INFINITY;
  for (int i = 0; i < m; ++i) {
    const double x = neg_db * e[i];
    if (x > mx) mx = x;
  }
  uint64_t *cum = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m);
  uint64_t run = 0;
  for (int i = 0; i < m; ++i) {
    const double w = orc_det_exp(neg_db * e[i] + (-mx));
    run += (uint64_t)(w * 1099511627776.0);
    cum[i] = run;
  }
  uint32_t d[4];
  orc_engine_draw(seed, population, 3u, 0x50410000u, step, d);
  const unsigned __int128 W = run;
  const uint64_t r = (uint64_t)((W * d[0]) >> 32);
  int i = 0;
  for (int j = 0; j < m; ++j) {
    const uint64_t pos = (uint64_t)(((unsigned __int128)j * W + r) / (unsigned __int128)m);
    while (cum[i] <= pos) ++i; /* pos < W = cum[m-1], so i stays in range; pos grows with j */
    src[j] = i;
  }
  free(cum);
}
