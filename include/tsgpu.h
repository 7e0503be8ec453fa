/* tsgpu.h -- C ABI of the B200-native TeraStructure hot path (libtsgpu.so).
 *
 * The reference (StoreyLab/terastructure) has no plugin/FFI interface; its external contract
 * is the CLI + output files, and the internal seam this library replaces is the thread
 * hand-off between SNPSamplingE::optimize_lambda (src/snpsamplinge.cc:320-366) and the
 * PhiRunnerE workers (src/snpsamplinge.cc:649-759, src/snpsamplinge.hh:266-300,416-431).
 * Each entry point below names the reference code it stands in for.  INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative ts_status on failure, with a message available from ts_last_error() (the
 * reference's convention is lerr()+exit(-1)/assert; the CLI front end keeps that).  All
 * device work of one engine is ordered on that engine's own CUDA stream; calls that return
 * data to the host synchronise that stream.  An engine is not thread-safe.
 *
 * There is no CPU fallback: every compute entry point fails with TS_ERR_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef TSGPU_H
#define TSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSGPU_ABI_VERSION 1
#define TS_MAX_K 32 /* populations supported by the compiled kernels (reference: any K) */

typedef enum ts_status {
  TS_OK = 0,
  TS_ERR_ARG = -1,   /* bad argument / unsupported shape */
  TS_ERR_CUDA = -2,  /* CUDA runtime error or no usable device */
  TS_ERR_STATE = -3, /* call made in the wrong order (e.g. step before data upload) */
  TS_ERR_IO = -4
} ts_status;

/* Hyper-parameters and shard geometry.  Defaults are the reference's (src/env.hh:200-249,
 * src/snpsamplinge.cc:15-16): alpha = 1/K, eta0 = eta1 = 1, nodetau0 = env.nodetau0 + 1 = 2,
 * nodekappa = 0.5, online_iterations = 10 (100 under -compute-beta, snpsamplinge.cc:75),
 * meanchangethresh = 1e-3. */
typedef struct ts_config {
  uint64_t n_total;   /* N: individuals in the whole data set (env.n) */
  uint64_t n_begin;   /* first individual owned by this engine; must be a multiple of 4 */
  uint64_t n_local;   /* individuals owned by this engine (== n_total on one GPU) */
  uint64_t l;         /* L: loci (env.l) */
  uint32_t k;         /* K: populations (env.k), 1..TS_MAX_K */
  uint32_t online_iterations;
  double alpha;
  double eta0, eta1;
  double nodetau0;
  double nodekappa;
  double meanchangethresh;
  int32_t device;     /* CUDA device ordinal */
  int32_t rank;       /* position of this shard among nranks engines (0 on one GPU) */
  int32_t nranks;
  int32_t reserved;
} ts_config;

typedef struct ts_engine ts_engine;

/* Fill cfg with the reference defaults for a single-GPU run of shape (n, l, k). */
void ts_config_defaults(ts_config *cfg, uint64_t n, uint64_t l, uint32_t k);

const char *ts_last_error(void);
int ts_abi_version(void);
/* Number of usable CUDA devices (0 when there is none; never negative). */
int ts_device_count(void);

/* ---- engine life cycle --------------------------------------------------------------- */

/* Replaces the allocation part of the SNPSamplingE ctor (snpsamplinge.cc:6-37) and
 * start_threads (snpsamplinge.cc:252-265): allocates the device-resident state
 * (2-bit packed genotype shard, gamma, exp(psi(gamma)), per-individual step counts,
 * lambda) and the stream on which the persistent cooperative kernel that replaces the PhiRunnerE
 * pool is launched.  lambda is set to eta
 * (init_lambda, snpsamplinge.cc:239-250). */
int ts_create(const ts_config *cfg, ts_engine **out);
/* The reference never frees (infer() exits the process); we do. */
int ts_destroy(ts_engine *e);

/* Replaces SNP::read_bed's payload loop (src/snp.cc:186-229) for loci [loc_begin,
 * loc_begin+nloc).  `rows` points at PLINK .bed SNP-major rows of the WHOLE data set
 * (ceil(n_total/4) significant bytes each, `row_pitch` bytes apart, 3-byte header already
 * skipped); the engine copies only its shard's byte range to the device.  Genotypes stay
 * 2-bit packed in HBM (codes 00->0, 10->1, 11->2, 01->missing, snp.cc:203-216). */
int ts_load_bed(ts_engine *e, uint64_t loc_begin, uint64_t nloc, const uint8_t *rows,
                uint64_t row_pitch);
/* The same for all engines of one process in ONE pass over the rows (the CLI's -gpus N; `rows` points at
 * the row of locus loc_begin, typically inside a memory-mapped .bed of up to 250 GB): chunks are staged through pinned host buffers and
 * every engine copies its own byte range asynchronously while the next chunk is read.  Returns when
 * all engines hold their shards. */
int ts_load_bed_fanout(ts_engine **engines, int n, uint64_t loc_begin, uint64_t nloc, const uint8_t *rows,
                       uint64_t row_pitch);
/* Synthetic genotypes generated on the device (BASELINE.md section 4 / SURVEY.md 8d):
 * y[l][n] ~ Binomial(2, sum_k theta[n][k]*beta[l][k]) from a counter-based generator keyed
 * by (seed, l, global n); `missing_rate` of the entries become the missing code.
 * theta: n_local x K rows of this shard, beta: L x K, both row-major host arrays. */
int ts_synth_bed(ts_engine *e, uint64_t seed, const float *theta, const float *beta,
                 double missing_rate);
/* Copy this shard's packed bytes of one locus back (ceil(n_local/4) bytes). */
int ts_get_bed_row(ts_engine *e, uint64_t loc, uint8_t *out);

/* Replaces _validation_map (snpsamplinge.hh:190-191) and kv_ok's two std::map lookups
 * (snpsamplinge.hh:389-408).  CSR over validation loci in ASCENDING locus order (the order
 * compute_likelihood visits them, snpsamplinge.cc:478-499): val_loc[nval],
 * val_off[nval+1], val_indiv[val_off[nval]] = GLOBAL individual ids, ascending within a
 * locus.  Must be called after the bed rows of those loci are resident. */
int ts_set_validation(ts_engine *e, uint64_t nval, const uint32_t *val_loc,
                      const uint64_t *val_off, const uint32_t *val_indiv);

/* init_gamma's result (snpsamplinge.cc:226-237) or load_gamma's (snpsamplinge.cc:800-837):
 * n_local x K row-major.  Also performs estimate_all_theta/set_dir_exp
 * (snpsamplinge.cc:595-609, lib.hh:19-35) on the device. */
int ts_set_gamma(ts_engine *e, const double *gamma_rows);
/* Reset lambda to eta for all loci (init_lambda) and the per-individual step counts
 * _c_indiv (snpsamplinge.cc:688-693) to zero. */
int ts_reset_lambda(ts_engine *e);
int ts_reset_counts(ts_engine *e);

/* ---- the hot path -------------------------------------------------------------------- */

/* One SVI iteration on locus `loc`: optimize_lambda(loc) (snpsamplinge.cc:320-366: up to
 * online_iterations rounds of {E-step over all individuals, 2K-sum, lambda/Elogbeta update},
 * stopping early when mean|d lambda| < meanchangethresh) followed -- unless hol_mode -- by
 * the gamma natural-gradient step and Elogtheta refresh that the reference's workers apply
 * lazily at the start of the next SNP (snpsamplinge.cc:660-671, :695-740).  Eager and lazy
 * application are observationally identical (SURVEY.md App. A note 3).
 * rounds_out (nullable) receives the number of rounds run; passing it forces a stream sync. */
int ts_step(ts_engine *e, uint32_t loc, int hol_mode, int *rounds_out);

/* The infer() loop body (snpsamplinge.cc:422-434) for n pre-drawn loci: the host keeps the
 * GSL-exact RNG, draws the SNP indices and hands them over in one call; the engine runs
 * them back to back with no host round trip.  Asynchronous unless rounds_out != NULL. */
int ts_steps(ts_engine *e, const uint32_t *locs, uint64_t n, int hol_mode, uint32_t *rounds_out);

/* compute_likelihood's loop (snpsamplinge.cc:489-499) + snp_likelihood
 * (snpsamplinge.hh:322-361) over this shard: for every validation locus in ascending order,
 * first ? estimate_beta only : a full hol-mode optimize_lambda; then
 * sum_n log max(C(2,y) q^y (1-q)^(2-y), 1e-30), q = sum_k Ebeta[loc][k]*Etheta[n][k], over
 * the shard's held-out individuals.  per_locus_sum[nval] (nullable) receives the per-locus
 * sums so that a multi-GPU caller can add shards in rank order; *sum is their total in
 * locus order, *count the number of held-out genotypes of this shard.  The caller advances
 * its iteration counter by nval when !first (snpsamplinge.hh:333) and applies the stopping
 * rule (snpsamplinge.cc:510-541). */
int ts_heldout_ll(ts_engine *e, int first, double *sum, uint64_t *count, double *per_locus_sum);

/* ---- state read-back (save_gamma snpsamplinge.cc:546-576, save_beta :761-778) --------- */
int ts_get_gamma(ts_engine *e, double *out /* n_local x K */);
int ts_get_theta(ts_engine *e, double *out /* n_local x K: gamma / rowsum */);
int ts_get_elogtheta(ts_engine *e, double *out /* n_local x K: psi(gamma) - psi(rowsum) */);
int ts_get_counts(ts_engine *e, uint32_t *out /* n_local */);
int ts_get_lambda(ts_engine *e, uint64_t loc_begin, uint64_t nloc, double *out /* nloc x K x 2 */);
int ts_get_beta(ts_engine *e, uint64_t loc_begin, uint64_t nloc, double *out /* nloc x K */);
int ts_sync(ts_engine *e);

/* Waits inside the kernels (grid barrier, peer exchange) are bounded by a wall-clock limit
 * (TSGPU_TIMEOUT_S in the environment at ts_create, default 60 s).  All ranks of an exchange group
 * must therefore enter ts_steps / ts_heldout_ll with the same work within that time of each other
 * (a host barrier before the first exchange is enough).  After a timeout the call returns
 * TS_ERR_CUDA ("barrier timed out") from the next synchronising call and the engine's state is
 * undefined: destroy it. */

/* ---- multi-GPU exchange (replaces the main thread's sum over workers' lambdat,
 *      snpsamplinge.cc:337-352) ------------------------------------------------------------
 * Individuals are sharded; each round every engine adds its 2K partial sums, as fixed-point
 * integers, into an accumulator in every peer's exchange buffer (NVLink atomics, or one NVLS
 * multicast reduction), so all ranks hold bit-identical lambda whatever the order of arrival
 * (details with ts_comm_attach_symmetric below).  One process per GPU: export the
 * local buffer as an opaque 64-byte CUDA IPC handle, all-gather the handles by any means,
 * then connect.  One process driving several GPUs: ts_comm_connect_local. */
#define TS_COMM_HANDLE_BYTES 64
int ts_comm_export(ts_engine *e, void *handle_out);
int ts_comm_connect(ts_engine *e, const void *all_handles /* nranks x 64 bytes, rank order */);
int ts_comm_connect_local(ts_engine **engines, int n);
/* Symmetric-memory exchange, optionally through NVLS (NVSwitch multicast).  The caller owns one
 * buffer of ts_comm_state_bytes() bytes per rank, each mapped on this engine's device:
 * rank_ptrs[nranks] (rank order, own rank included) and, optionally, a multicast pointer that
 * aliases the same offsets of ALL ranks' buffers (e.g. torch.distributed._symmetric_memory:
 * empty() + rendezvous() give buffer_ptrs and multicast_ptr).  The engine's exchange state moves into
 * its own buffer.  Exchange schemes (ts_comm_mode; TSGPU_XCHG=gacc in the environment forces the second):
 * with a multicast pointer the default is "mcacc" -- on every GPU the CTA whose arrival completes a word (the
 * arrival is an atomic with return value) adds the GPU's total into an accumulator on every rank with ONE
 * multimem.red, every CTA polls one local word pair; without, "gacc" -- the same with one NVLink red.add per
 * peer (also what ts_comm_connect uses).  The schemes measured and rejected (slots written by every GPU, with
 * peer stores or one multimem.st; multimem.red from every CTA; a polling CTA 0 that forwards) are described in
 * profiles/r2_summary.md).  Call it on every rank, then
 * synchronise the ranks (a host barrier) before the first ts_steps.  total_ctas = sum of the ranks'
 * CTA counts (ts_get_plan), or 0 when all shards have this engine's geometry.  ts_comm_connect_local does
 * all of this by itself for the engines of one process when the devices support multicast. */
uint64_t ts_comm_state_bytes(void);
/* Exchange scheme in use: 3 = the last local arrival adds the GPU's totals into an accumulator on every rank
 * with one NVLink red.add per peer ("gacc"), 4 = the same with one multimem.red per word ("mcacc"); 0-2 were
 * the schemes of earlier builds and are no longer returned. */
int ts_comm_mode(const ts_engine *e);
int ts_comm_attach_symmetric(ts_engine *e, const void *const *rank_ptrs, void *multicast_ptr, uint64_t bytes,
                             uint32_t total_ctas);

/* Launch geometry the engine uses for a shard of n_local individuals on a device with num_sms SMs
 * (pure host arithmetic, no device needed): individuals per thread held in registers by the
 * persistent kernel, CTAs and threads per CTA.  The reference's counterpart is the static split of individuals over
 * `-nthreads` workers (split_all_indivs, snpsamplinge.cc:298-318). */
int ts_plan_shard(uint64_t n_local, int k, int num_sms, int *ind_per_thread, int *grid, int *block);
/* The geometry this engine runs with.  It differs from ts_plan_shard's only under the test knob
 * TSGPU_IPT=<i> (environment, read by ts_create), which pins the individuals per thread (0 = the
 * tiered kernel, see ts_get_tiers) so that parity tests reach every kernel instantiation at oracle-sized inputs. */
int ts_get_plan(const ts_engine *e, int *ind_per_thread, int *grid, int *block);
/* Shards beyond the register-resident capacity (148 x threads x individuals per thread) keep further
 * individuals in the CTAs' shared memory and stream the rest from L2/HBM every round.  ts_plan_tiers:
 * individuals per thread in shared memory and number of streamed individuals the engine would choose
 * (0, 0 for a register-resident shard).  ts_get_tiers: what this engine runs with; smem_per_thread = -1
 * means the register-only kernel.  Test knobs read by ts_create: TSGPU_IPT=0 forces the tiered kernel,
 * TSGPU_TIER_J / TSGPU_TIER_GRID / TSGPU_TIER_BLOCK shape it. */
int ts_plan_tiers(uint64_t n_local, int k, int num_sms, int *smem_per_thread, uint64_t *n_streamed);
int ts_get_tiers(const ts_engine *e, int *smem_per_thread, uint64_t *n_streamed);

/* ---- profiling hooks ------------------------------------------------------------------- */
/* Kernels launched by this engine since creation (for bench.py's gpu_launches). */
uint64_t ts_launch_count(const ts_engine *e);
/* Device time of the hot-path kernels between two marks, measured with CUDA events on the
 * engine's stream.  ts_timer_start records, ts_timer_stop records + syncs, returns ms. */
int ts_timer_start(ts_engine *e);
int ts_timer_stop(ts_engine *e, float *ms_out);
/* Developer aid: with TSGPU_TRACE=1 in the environment at ts_create, CTA 0 of the persistent kernel
 * stamps clock64() at its phase boundaries for the first 64 work items of a launch (64 x 128 slots) --
 * in the developer build of the library only (make -C terastructure_b200/csrc trace); the product
 * library carries no stamps and returns zeros. */
int ts_debug_trace(ts_engine *e, long long *out);

/* ---- host-side, RNG-exact initialisation (no GPU needed) ------------------------------
 * The SNP-sampling RNG stays on the host and must reproduce the reference's GSL stream
 * (gsl_rng_mt19937; snpsamplinge.cc:59-63) bit for bit. */
typedef struct ts_rng ts_rng;
/* gsl_rng_alloc(gsl_rng_default) then, if seed != 0, gsl_rng_set(r, (unsigned long)seed). */
ts_rng *ts_rng_create(double seed);
void ts_rng_destroy(ts_rng *r);
uint32_t ts_rng_get(ts_rng *r);                       /* gsl_rng_get */
uint32_t ts_rng_uniform_int(ts_rng *r, uint32_t n);   /* gsl_rng_uniform_int (cc:423) */
double ts_rng_gamma(ts_rng *r, double a, double b);   /* gsl_ran_gamma (cc:233) */
/* Draw n loci for ts_steps. */
void ts_rng_sample_locs(ts_rng *r, uint32_t l, uint32_t *out, uint64_t n);

/* set_validation_sample (snpsamplinge.cc:196-224) with identical draw and rejection order.
 * `bed` = whole-data-set SNP-major rows (for the is_missing test), or NULL when the data set
 * has no missing genotypes (device-generated synthetic data).  Outputs are malloc'ed
 * CSR arrays in ascending locus order, released with ts_free. */
int ts_sample_validation(ts_rng *r, uint64_t n, uint64_t l, const uint8_t *bed, uint64_t row_pitch,
                         uint64_t *nval_out, uint32_t **val_loc_out, uint64_t **val_off_out,
                         uint32_t **val_indiv_out);
/* init_gamma (snpsamplinge.cc:226-237): N*K sequential gsl_ran_gamma(100*v, 0.01) draws. */
void ts_init_gamma(ts_rng *r, uint64_t n, uint32_t k, double *gamma_out);
void ts_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* TSGPU_H */
