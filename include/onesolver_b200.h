/*
 * onesolver_b200.h -- C ABI of the B200-native simulated-annealing engine.
 *
 * This is the drop-in boundary for oneSolver's annealing hot path.  The
 * reference has no FFI: its boundary is the C++ template call
 *     sa::anneal(instance, queue, beta_schedule, num_iter, num_tries[, sweeps_per_beta])
 *         -> qubo::Solution                     (include/simulated_annealing/annealing.hpp:55-58)
 * invoked from app/one-solver-anneal.cpp:163-164.  The header-only C++ shim in
 * include/simulated_annealing/annealing.hpp keeps that signature and calls the
 * entry points below; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host
 * buffer; the library owns device memory behind the opaque osa_problem handle;
 * all calls block until the result is in the caller's buffers (the reference
 * waits on its single kernel, annealing.hpp:127); every function returns an
 * osa_status and osa_last_error() gives the message for the calling thread.
 * There is no CPU fallback: without a CUDA device every compute entry point
 * returns OSA_ERR_NO_DEVICE / OSA_ERR_CUDA.
 */
#ifndef ONESOLVER_B200_H_
#define ONESOLVER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSA_ABI_VERSION 3

typedef enum {
  OSA_OK = 0,
  OSA_ERR_INVALID = 1,     /* bad argument (message says which) */
  OSA_ERR_CUDA = 2,        /* a CUDA runtime call or kernel failed */
  OSA_ERR_NO_DEVICE = 3,   /* no usable CUDA device */
  OSA_ERR_UNSUPPORTED = 4, /* shape/mode combination not implemented */
  OSA_ERR_NOMEM = 5
} osa_status;

/* How sites are visited.
 * OSA_MODE_RANDOM_SITE: one attempt per (iter, sweep) at a uniformly random
 *   site drawn from the trajectory's stream -- the reference's loop
 *   (annealing.hpp:97-101).  Default of sa::anneal and the CLI.
 * OSA_MODE_SEQUENTIAL_SWEEP: each (iter, sweep) visits sites 0..N-1 in order
 *   (N attempts) -- the throughput mode named by BASELINE.json's north_star. */
typedef enum { OSA_MODE_RANDOM_SITE = 0, OSA_MODE_SEQUENTIAL_SWEEP = 1 } osa_mode;

/* Acceptance rule for an attempted flip with energy change dE at schedule value b.
 * OSA_ACCEPT_REFERENCE: dE < 0 or exp(-dE / b) > u  (annealing.hpp:106-108: the
 *   reference DIVIDES by beta).
 * OSA_ACCEPT_BOLTZMANN: dE < 0 or exp(-b * dE) > u  (conventional inverse temperature). */
typedef enum { OSA_ACCEPT_REFERENCE = 0, OSA_ACCEPT_BOLTZMANN = 1 } osa_accept;

/* Arithmetic of the sweep (local fields, dE, thresholds). Returned energies are
 * always recomputed in fp64 with the reference formula (annealing.hpp:31-40). */
typedef enum { OSA_SWEEP_F64 = 0, OSA_SWEEP_F32 = 1 } osa_precision;

typedef struct osa_problem osa_problem; /* opaque: device-resident Q (dense or CSR) */

typedef struct {
  uint64_t seed;           /* 1234 in the reference (annealing.hpp:87) */
  uint64_t first_try;      /* global id of this call's trajectory 0 (multi-GPU shards) */
  uint64_t num_tries;      /* number of independent trajectories (annealing.hpp:57) */
  int32_t num_iter;        /* length of the schedule (annealing.hpp:56) */
  int32_t sweeps_per_beta; /* annealing.hpp:58, default 1 */
  int32_t mode;            /* osa_mode */
  int32_t accept_rule;     /* osa_accept */
  int32_t kernel_variant;  /* 0 = auto; >0 forces a kernel (see osa_kernel_name) */
  int32_t flags;           /* must be 0 */
} osa_anneal_params;

typedef struct {
  uint64_t attempts;         /* spin-flip attempts executed */
  uint64_t accepts;          /* accepted flips */
  uint64_t row_fetches;      /* Q rows streamed by the sweep: (batch, step) pairs with >=1 accept */
  uint64_t init_row_fetches; /* Q rows streamed to build the initial local fields */
  float ms_total;            /* device time of the whole call (CUDA events, library stream) */
  float ms_sweep;            /* init + sweep kernel */
  float ms_energy;           /* exact fp64 energy recompute */
  float ms_reduce;           /* argmin + gather */
  int32_t kernel_id;         /* which sweep kernel ran */
  int32_t traj_per_batch;    /* R: trajectories sharing one row fetch */
  int32_t q_elem_bytes;      /* sizeof(element) of the Q copy the sweep streamed */
  int32_t grid;              /* CTAs launched by the sweep kernel */
  int32_t launches;          /* kernels launched by this call */
  int32_t reserved;
  /* diagnostics of the dense sequential kernel: SM cycles summed over CTAs (0 elsewhere) */
  uint64_t cyc_decide, cyc_apply, cyc_stage, cyc_init;
  uint64_t pt_swaps;         /* osa_pt_anneal: accepted replica exchanges; osa_pa_anneal: resampled slots */
} osa_stats;

/* ---- library / device ---------------------------------------------------- */
int osa_abi_version(void);
const char *osa_last_error(void);
/* devices::construct_device_selector("gpu") (include/helpers/devices.hpp:27-42) */
int osa_device_count(int *count);
int osa_device_name(int device, char *buf, size_t buflen);
const char *osa_kernel_name(int kernel_id);

/* ---- problem upload: replaces the sycl::buffer creation, annealing.hpp:65-72 ---- */
/* qsym: N x N row-major symmetric, diagonal = linear terms, off-diagonal = full
 * coupling on both sides -- exactly helpers::flatten_qubo's layout
 * (include/helpers/qubo_helpers.hpp:26-44).                                     */
int osa_problem_create_dense_f64(const double *qsym, int n, int device, int sweep_precision,
                                 osa_problem **out);
int osa_problem_create_dense_f32(const float *qsym, int n, int device, osa_problem **out);
/* CSR of the symmetric coupling graph: both directions stored, no diagonal
 * entries, columns ascending within a row; diag[n] holds the linear terms.      */
int osa_problem_create_csr_f64(const int32_t *rowptr, const int32_t *col, const double *val,
                               const double *diag, int n, int device, int sweep_precision,
                               osa_problem **out);
int osa_problem_destroy(osa_problem *p);
int osa_problem_size(const osa_problem *p, int *n, int *is_sparse, int *sweep_precision);

/* ---- the hot path: replaces q.submit(parallel_for<annealing>).wait() and the host
 *      argmin, annealing.hpp:74-139 -------------------------------------------- */
/* beta_schedule[num_iter] as built by one-solver-anneal.cpp:23-39.
 * Outputs (any may be NULL):
 *   best_energies[num_tries]                      per-trajectory best energy (annealing.hpp:125)
 *   best_states_packed[num_tries][ceil(N/32)]     bit i%32 of word i/32 = variable i
 *   best_state[N]                                 0/1 chars of the winning trajectory
 *   best_energy, best_index                       its energy and GLOBAL trajectory id; ties go
 *                                                 to the lowest id (std::min_element, :134)   */
int osa_anneal(osa_problem *p, const double *beta_schedule, const osa_anneal_params *params,
               double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
               double *best_energy, uint64_t *best_index, osa_stats *stats);

/* osa_anneal plus the FLIP TRACE of every trajectory: trace_hash[num_tries] (may be NULL) receives a
 * 64-bit FNV-1a hash over the accepted flips in the order they happen, one update per (step, block
 * of 32 sites) with at least one accepted flip:
 *     h = 0xcbf29ce484222325;  h = (h ^ step) * 0x100000001b3;  h = (h ^ (block << 32 | mask)) * 0x100000001b3
 * step = counter of the random stream (sequential mode: sweep number, random-site mode: attempt
 * number), mask = the accepted sites of the block.  It pins the whole spin sequence of a
 * trajectory, not only its best state; the host replay computes the same value
 * (oracle/osa_oracle.c), which is how BASELINE.json's "replayed trajectories" criterion is tested. */
int osa_anneal_traced(osa_problem *p, const double *beta_schedule, const osa_anneal_params *params,
                      double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                      double *best_energy, uint64_t *best_index, uint64_t *trace_hash,
                      osa_stats *stats);

/* ---- parallel tempering on top of the same sweep kernel.  The reference has no such sampler; its
 *      benchmark report recommends one (benchmarks/annealing/performance.md:54-59).
 * num_groups independent runs of num_replicas replicas each; betas[num_replicas] is the ladder
 * (strictly increasing, positive).  Per round every replica does sweeps_per_round sequential
 * sweeps at its current beta, then neighbouring rungs (even pairs in even rounds, odd pairs in odd
 * rounds) exchange configurations with the Metropolis probability computed from exact fp64
 * energies.  Trajectory id = first_group * num_replicas + group * num_replicas + slot; slot k
 * starts on rung k.  Dense problems with n <= 8192 (fp32 sweeps) / 4096 (fp64 sweeps) only.
 * Outputs as in osa_anneal, per trajectory (= per replica slot): the best state each one visited.
 * The local fields of every replica are carried from round to round in the sweep precision (a
 * round does not rebuild them from the spins), exactly as one long annealing run carries them;
 * every exchange decision and every returned energy uses energies recomputed exactly in fp64. */
typedef struct osa_pt_params {
  uint64_t seed;             /* 1234 like annealing.hpp:87 */
  uint64_t first_group;      /* id offset when a run is sharded over GPUs */
  uint64_t num_groups;
  int32_t num_replicas;
  int32_t num_rounds;
  int32_t sweeps_per_round;
  int32_t accept_rule;       /* OSA_ACCEPT_*: also selects the Boltzmann weight of the exchange */
  uint32_t flags;            /* must be 0 */
  int32_t reserved;
} osa_pt_params;

int osa_pt_anneal(osa_problem *p, const double *betas, const osa_pt_params *params,
                  double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                  double *best_energy, uint64_t *best_index, osa_stats *stats);

/* ---- population annealing on top of the same sweep kernel.  The reference has no such sampler;
 *      its benchmark report names it first among the ones it recommends
 *      (benchmarks/annealing/performance.md:54-59).
 * num_populations independent populations of population_size replicas; betas[num_steps] is the
 * annealing schedule of a population.  Step t: every replica does sweeps_per_step sequential
 * sweeps at betas[t]; then (t + 1 < num_steps) the population is resampled for the next
 * temperature with weights exp(-(b' - b) E) from exact fp64 energies, b = the inverse temperature
 * of the acceptance rule (1 / beta for the reference's rule, beta for the Boltzmann rule), keeping
 * its size: systematic resampling on 40-bit integer weights with ONE uniform per (population, step)
 * from the Philox stream, so the result does not depend on any reduction order (osa_pa.cu).
 * Trajectory id = (first_population + population) * population_size + slot.  Dense problems with
 * n <= 8192 (fp32 sweeps) / 4096 (fp64 sweeps) only.  Outputs as in osa_anneal, per replica slot:
 * the best state seen in that slot.  stats->pt_swaps = replicas overwritten by a copy of another.
 * The local fields travel with a replica through the resampling and are carried from step to
 * step in the sweep precision; weights and returned energies use exact fp64 energies. */
typedef struct osa_pa_params {
  uint64_t seed;             /* 1234 like annealing.hpp:87 */
  uint64_t first_population; /* id offset when a run is sharded over GPUs */
  uint64_t num_populations;
  int32_t population_size;   /* 1 .. 2^20 */
  int32_t num_steps;         /* temperatures */
  int32_t sweeps_per_step;
  int32_t accept_rule;       /* OSA_ACCEPT_* */
  uint32_t flags;            /* must be 0 */
  int32_t reserved;
} osa_pa_params;

int osa_pa_anneal(osa_problem *p, const double *betas, const osa_pa_params *params,
                  double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                  double *best_energy, uint64_t *best_index, osa_stats *stats);

/* ---- the same call sharded over the GPUs of one box (BASELINE config 5: 1M tries over 8 B200).
 * One process, one host thread and stream per GPU, Q replicated, device k of G runs the global
 * trajectory ids of shard k (contiguous ranges, remainder to the low devices); the random streams
 * are keyed by global ids, so every output is identical to the one-device call.  The only
 * exchange is ONE ncclAllGather of {best energy, global id, packed best state} per device at the
 * end; the winner is min energy, then min id (std::min_element, annealing.hpp:134).
 * devices: CUDA ordinals, or NULL for devices 0..num_devices-1 (num_devices <= 0: all visible).
 * stats: sums over the devices, times = the slowest device, reserved = number of devices;
 * device_stats[number of devices] (may be NULL): the per-device records.                        */
typedef struct osa_multi osa_multi;
int osa_multi_create_dense_f64(const double *qsym, int n, const int *devices, int num_devices,
                               int sweep_precision, osa_multi **out);
int osa_multi_create_dense_f32(const float *qsym, int n, const int *devices, int num_devices,
                               osa_multi **out);
int osa_multi_create_csr_f64(const int32_t *rowptr, const int32_t *col, const double *val,
                             const double *diag, int n, const int *devices, int num_devices,
                             int sweep_precision, osa_multi **out);
int osa_multi_destroy(osa_multi *m);
int osa_multi_devices(const osa_multi *m, int *num_devices, int *devices, int capacity);
int osa_multi_problem(osa_multi *m, int slot, osa_problem **out); /* the replica on device slot */
int osa_multi_anneal(osa_multi *m, const double *beta_schedule, const osa_anneal_params *params,
                     double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                     double *best_energy, uint64_t *best_index, osa_stats *stats,
                     osa_stats *device_stats);

/* sa::energy (annealing.hpp:31-40) for a batch of packed states, fp64 on the device */
int osa_energy_batch(osa_problem *p, const uint32_t *states_packed, uint64_t count, double *out);

/* ---- optional pinned host staging buffers (cudaMallocHost / cudaFreeHost) for callers that
 *      upload large Q matrices repeatedly; any host pointer is accepted by the calls above. */
int osa_host_alloc_pinned(size_t bytes, void **out);
int osa_host_free_pinned(void *ptr);

/* ---- brute-force ground state for n <= 40: CUDA counterpart of exhaustive::solve
 *      (include/exhaustive/exhaustive.hpp:29-167).  qsym in flatten_qubo layout; the winner is the
 *      LOWEST state integer among the minima (bit i = variable i), energy by the reference formula. */
int osa_exhaustive_dense_f64(const double *qsym, int n, int device, uint8_t *best_state,
                             double *best_energy);

/* ---- measurement helper: achieved read bandwidth of a `bytes`-sized buffer swept
 *      `iters` times by all SMs (L2-resident if it fits, HBM otherwise); GB/s.      */
int osa_measure_read_bandwidth(int device, size_t bytes, int iters, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* ONESOLVER_B200_H_ */
