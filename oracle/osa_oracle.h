/*
 * oracle/osa_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C11) of the oneSolver annealing hot path, used only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm as the checker / baseline.  Nothing under onesolver_b200/, include/ or
 * app/ may include, link or call anything in this directory.
 *
 * Two families of functions live here:
 *
 *  (A) reference-faithful restatements ("orc_ref_*"): line-by-line CPU
 *      versions of the reference's own code, each citing the file:line it
 *      follows under /root/reference.  They keep the reference algorithm:
 *      one single-spin flip attempt per "iteration" at a random site, a FULL
 *      O(N^2) energy recompute per attempt, acceptance
 *      exp((E_cur - E_new)/beta) > u, strict-< best tracking, first-minimum
 *      argmin.
 *
 *  (B) the bit-exact host replay ("orc_replay_*"): a scalar twin of the new
 *      CUDA engine (bit-packed spins, local field h, dE = (1-2x_i) h_i,
 *      counter-based Philox keyed by (seed, trajectory, sweep, spin)), written
 *      independently of the CUDA sources.  GPU trajectories must match it
 *      bit for bit.
 *
 * PARITY PINNING: the reference cannot be compiled in this container (needs
 * DPC++ -fsycl, oneMKL device RNG, Boost).  Everything
 * deterministic-by-formula (energy, flatten, schedules, exhaustive ground
 * states, CSV writer, parser accept/reject set) is pinned against the
 * reference's own golden vectors (tests/golden/, generated from the files
 * under /root/reference by tests/golden/make_golden.py).  The RNG stream of
 * the reference (oneMKL philox4x32x10, headers not vendored, no KAT in the
 * reference) is "parity unpinned": trajectory-level parity is statistical.
 */
#ifndef OSA_ORACLE_H_
#define OSA_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG: Philox4x32-10 (Salmon et al. 2011), counter based ------------- */
enum { ORC_STREAM_INIT = 0, ORC_STREAM_SEQ = 1, ORC_STREAM_RND = 2 };

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* engine keying: key=(seed lo, seed hi); ctr=(c0, c1, traj lo, (traj hi & 0x3fffffff)|stream<<30) */
void orc_engine_draw(uint64_t seed, uint64_t traj, uint32_t stream, uint32_t c0, uint32_t c1,
                     uint32_t out[4]);

/* deterministic -ln(u), u = (2w+1)/2^33 in (0,1); only IEEE +,*,fma on floats */
float orc_neglogf(uint32_t w);

/* initial spin j of trajectory traj (replaces random.bit(), annealing.hpp:90-92) */
int orc_init_bit(uint64_t seed, uint64_t traj, uint32_t j);

/* ---- (A) reference-faithful restatements -------------------------------- */

/* annealing.hpp:31-40 -- sum_{i<=j} Q[i*N+j]*x_i*x_j, fixed i-then-j order */
double orc_ref_energy(const double *flat_qubo, const char *state, int n);

/* qubo_helpers.hpp:26-44 -- dense symmetric flatten from coordinate lists.
 * lin_idx/lin_val: n_lin linear terms; quad_i/quad_j/quad_val: n_quad couplings
 * exactly as stored in the model's maps (either orientation allowed).       */
void orc_ref_flatten(int n, const int *lin_idx, const double *lin_val, int n_lin,
                     const int *quad_i, const int *quad_j, const double *quad_val, int n_quad,
                     double *out /* n*n */);

/* one-solver-anneal.cpp:23-29 and :31-39 */
void orc_ref_schedule_linear(double *schedule, double beta_min, double beta_max, unsigned num_iter);
void orc_ref_schedule_geometric(double *schedule, double beta_min, double beta_max,
                                unsigned num_iter);

/* annealing.hpp:85-139 with the engine's STREAM_RND draws substituted for the
 * unpinned oneMKL stream.  Trajectory ids are first_try .. first_try+num_tries-1.
 * best_states: num_tries*n chars; best_energies: num_tries doubles.
 * Returns index (relative to first_try) of the first minimum (std::min_element).
 * num_threads<=0: use all OpenMP threads.                                     */
int64_t orc_ref_anneal(const double *flat_qubo, int n, const double *beta_schedule, int num_iter,
                       uint64_t num_tries, int sweeps_per_beta, uint64_t seed, uint64_t first_try,
                       char *best_states, double *best_energies, int num_threads);

/* exhaustive.hpp:29-167 -- brute force over 2^n states split in num_ranges
 * contiguous ranges (the reference uses max_compute_units); bit i <-> variable i
 * (ulong_to_vec.hpp:23-32).  n <= 30.                                          */
int orc_ref_exhaustive(const double *flat_qubo, int n, int num_ranges, char *best_state,
                       double *best_energy);

/* solution.hpp:58-67 -- two-line CSV, "%g"-style 6 significant digits          */
size_t orc_ref_solution_csv(const char *state, int n, double energy, char *buf, size_t buflen);

/* ---- (B) bit-exact host replay of the CUDA engine ------------------------ */

enum { ORC_MODE_RANDOM_SITE = 0, ORC_MODE_SEQUENTIAL_SWEEP = 1 };

typedef struct {
  uint64_t attempts;
  uint64_t accepts;
  uint64_t row_fetches;      /* (batch, step) pairs with >= 1 accept (batch = R consecutive ids) */
  uint64_t init_row_fetches; /* (batch, site) pairs with >= 1 set initial spin */
} orc_counters;

/* The replays below store the flip trace of trajectory tl (the value of osa_anneal_traced, see
 * include/onesolver_b200.h) in buf[tl] while buf is set; NULL switches it off.  Not thread-safe
 * across concurrent replay CALLS (the calls themselves are OpenMP-parallel inside).           */
void orc_set_trace_output(uint64_t *buf);

/* Dense replay.  qoff: n*ld row-major symmetric with ZERO diagonal, diag: n.
 * tscale[num_iter]: threshold scale per iteration (beta for the reference rule,
 * 1/beta for the Boltzmann rule), already rounded to the sweep precision.
 * Outputs per trajectory: best_rel (tracked best energy relative to the initial
 * state, double), best_states_packed [num_tries][ceil(n/32)], final states
 * (optional).  batch_r: trajectories per row-fetch batch for the counters.     */
int orc_replay_dense_f32(const float *qoff, const float *diag, int n, size_t ld,
                         const float *tscale, int num_iter, int sweeps_per_beta, int mode,
                         uint64_t seed, uint64_t first_try, uint64_t num_tries, int batch_r,
                         double *best_rel, uint32_t *best_states_packed,
                         uint32_t *final_states_packed, orc_counters *counters);
int orc_replay_dense_f64(const double *qoff, const double *diag, int n, size_t ld,
                         const double *tscale, int num_iter, int sweeps_per_beta, int mode,
                         uint64_t seed, uint64_t first_try, uint64_t num_tries, int batch_r,
                         double *best_rel, uint32_t *best_states_packed,
                         uint32_t *final_states_packed, orc_counters *counters);

/* resumable variants (osa_pt_anneal): init_states NULL = STREAM_INIT, tscale_traj NULL =
 * tscale[iter], step_base = first sweep number in the STREAM_SEQ counter            */
int orc_replay_dense_f32_round(const float *qoff, const float *diag, int n, size_t ld,
                               const float *tscale, int num_iter, int sweeps_per_beta, int mode,
                               uint64_t seed, uint64_t first_try, uint64_t num_tries, int batch_r,
                               double *best_rel, uint32_t *best_states_packed,
                               uint32_t *final_states_packed, orc_counters *counters,
                               const uint32_t *init_states, const float *tscale_traj,
                               uint32_t step_base);
int orc_replay_dense_f64_round(const double *qoff, const double *diag, int n, size_t ld,
                               const double *tscale, int num_iter, int sweeps_per_beta, int mode,
                               uint64_t seed, uint64_t first_try, uint64_t num_tries, int batch_r,
                               double *best_rel, uint32_t *best_states_packed,
                               uint32_t *final_states_packed, orc_counters *counters,
                               const uint32_t *init_states, const double *tscale_traj,
                               uint32_t step_base);

/* CSR replay (local field recomputed on demand in CSR order, like the CUDA
 * sparse kernel).  Symmetric adjacency, no diagonal entries.                   */
int orc_replay_csr_f32(const int32_t *rowptr, const int32_t *col, const float *val,
                       const float *diag, int n, const float *tscale, int num_iter,
                       int sweeps_per_beta, int mode, uint64_t seed, uint64_t first_try,
                       uint64_t num_tries, double *best_rel, uint32_t *best_states_packed,
                       uint32_t *final_states_packed, orc_counters *counters);
int orc_replay_csr_f64(const int32_t *rowptr, const int32_t *col, const double *val,
                       const double *diag, int n, const double *tscale, int num_iter,
                       int sweeps_per_beta, int mode, uint64_t seed, uint64_t first_try,
                       uint64_t num_tries, double *best_rel, uint32_t *best_states_packed,
                       uint32_t *final_states_packed, orc_counters *counters);

/* energy of packed states with the reference formula (upper triangle of the
 * symmetric dense matrix qsym, diag on the diagonal), fp64                     */
void orc_energy_packed(const double *qsym, int n, const uint32_t *states_packed, uint64_t count,
                       double *out);

int orc_num_threads(void);

/* (C) population annealing: deterministic exp and one resampling step of osa_pa_anneal
 * (include/onesolver_b200.h); src[m] receives the source replica of every slot.            */
double orc_det_exp(double x);
void orc_pa_resample(const double *e, int m, double neg_db, uint64_t seed, uint64_t population,
                     uint32_t step, int32_t *src);

#ifdef __cplusplus
}
#endif
#endif
