/* TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never linked into, imported by or
 * executed from the product path (terastructure_b200/, include/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Plain-C sequential restatement of the reference's SNPSamplingE hot path at
 * `-nthreads 1` (the only deterministic mode).  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/src).  Operation order
 * is kept identical to the reference so that results agree with the reference binary
 * (oracle/_ref) to the last printed digit.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port against
 *   (1) the reference's own fixture data/output_theta.txt (seed 1234, 1 thread), and
 *   (2) outputs of the reference binary built here from the unmodified sources
 *       (tests/golden/, made by tools/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_sf.h>

typedef struct {
  uint32_t n, l, k;
  const uint8_t *y;          /* SNP-major codes y[loc*n + indiv] in {0,1,2,3=missing} (snp.cc:203-216) */
  gsl_rng *r;
  /* env.hh:200-249 defaults */
  double alpha, eta0, eta1, nodetau0, nodekappa, meanchangethresh, stop_threshold;
  uint32_t online_iterations;
  /* model state (snpsamplinge.hh:210-234) */
  double *gamma, *Elogtheta, *Etheta;      /* n*k */
  double *lambda, *Elogbeta;                /* l*k*2 */
  double *Ebeta;                            /* l*k */
  uint32_t *c_indiv;                        /* n */
  double *phimom, *phidad;                  /* n*k, the single worker's buffers */
  /* validation set: per-locus membership bitmap, built in draw order */
  uint8_t *is_val_loc;                      /* l */
  uint32_t nval_loc, per_loc_h;
  uint32_t *val_loc;                        /* sorted ascending */
  uint8_t **val_mask;                       /* per validation locus (sorted order): n bytes */
  int32_t *val_slot;                        /* l: index into val_loc or -1 */
  /* loop state */
  uint32_t iter;
  int pending;                              /* a lazy gamma step is owed for pending_loc */
  uint32_t pending_loc;
  double prev_h, max_h;
  uint32_t nh;
  uint32_t last_rounds;
} tso;

/* kv_ok (snpsamplinge.hh:389-408): not held out and not missing */
static int kv_ok(const tso *o, uint32_t indiv, uint32_t loc) {
  int32_t s = o->val_slot[loc];
  if (s >= 0 && o->val_mask[s][indiv]) return 0;
  return o->y[(size_t)loc * o->n + indiv] != 3;
}

/* PopLib::set_dir_exp (lib.hh:19-35) for one row + estimate_all_theta (snpsamplinge.cc:595-609) */
static void refresh_theta_row(tso *o, uint32_t n) {
  uint32_t K = o->k;
  const double *g = o->gamma + (size_t)n * K;
  double s = .0;
  for (uint32_t k = 0; k < K; ++k) s += g[k];
  double psi_sum = gsl_sf_psi(s);
  for (uint32_t k = 0; k < K; ++k) {
    o->Etheta[(size_t)n * K + k] = g[k] / s;
    o->Elogtheta[(size_t)n * K + k] = gsl_sf_psi(g[k]) - psi_sum;
  }
}

/* set_validation_sample (snpsamplinge.cc:196-224): draw order and rejection rules kept */
static void set_validation_sample(tso *o, double validation_ratio) {
  uint32_t N = o->n, L = o->l;
  o->per_loc_h = N < 2000 ? (N / 10) : (N / 100);
  uint32_t nlocs = (uint32_t)(L * validation_ratio);
  o->is_val_loc = (uint8_t *)calloc(L, 1);
  o->val_slot = (int32_t *)malloc(sizeof(int32_t) * L);
  for (uint32_t i = 0; i < L; ++i) o->val_slot[i] = -1;
  uint32_t *draw_loc = (uint32_t *)malloc(sizeof(uint32_t) * (nlocs + 1));
  uint8_t **draw_mask = (uint8_t **)malloc(sizeof(uint8_t *) * (nlocs + 1));
  uint32_t cnt = 0;
  /* do { ... } while (lm.size() < nlocs): the body runs at least once */
  do {
    uint32_t loc = (uint32_t)gsl_rng_uniform_int(o->r, L);
    if (o->is_val_loc[loc]) continue;
    o->is_val_loc[loc] = 1;
    uint8_t *m = (uint8_t *)calloc(N, 1);
    draw_loc[cnt] = loc;
    draw_mask[cnt] = m;
    /* temporary slot so kv_ok sees entries already drawn for this locus */
    o->val_slot[loc] = (int32_t)cnt;
    o->val_mask = draw_mask;
    uint32_t c = 0;
    while (c < o->per_loc_h) {
      uint32_t indiv = (uint32_t)gsl_rng_uniform_int(o->r, N);
      if (kv_ok(o, indiv, loc)) { m[indiv] = 1; c++; }
    }
    cnt++;
  } while (cnt < nlocs);
  /* regroup ascending by locus (compute_likelihood iterates a std::map, snpsamplinge.cc:478-499) */
  o->nval_loc = cnt;
  o->val_loc = (uint32_t *)malloc(sizeof(uint32_t) * cnt);
  o->val_mask = (uint8_t **)malloc(sizeof(uint8_t *) * cnt);
  uint32_t j = 0;
  for (uint32_t loc = 0; loc < L; ++loc)
    if (o->is_val_loc[loc]) {
      uint32_t d = (uint32_t)o->val_slot[loc];
      o->val_loc[j] = loc;
      o->val_mask[j] = draw_mask[d];
      j++;
    }
  for (uint32_t i = 0; i < cnt; ++i) o->val_slot[o->val_loc[i]] = (int32_t)i;
  free(draw_loc);
  free(draw_mask);
}

/* init_gamma (snpsamplinge.cc:226-237) */
static void init_gamma(tso *o) {
  for (uint32_t i = 0; i < o->n; ++i)
    for (uint32_t j = 0; j < o->k; ++j) {
      double v = (o->k < 100) ? 1.0 : (double)100.0 / o->k;
      o->gamma[(size_t)i * o->k + j] = gsl_ran_gamma(o->r, 100 * v, 0.01);
    }
}

/* estimate_beta (snpsamplinge.cc:279-296) */
static void estimate_beta(tso *o, uint32_t loc) {
  uint32_t K = o->k;
  for (uint32_t k = 0; k < K; ++k) {
    const double *ld = o->lambda + ((size_t)loc * K + k) * 2;
    double s = .0;
    for (uint32_t t = 0; t < 2; ++t) s += ld[t];
    o->Ebeta[(size_t)loc * K + k] = ld[0] / s;
    double psi_sum = gsl_sf_psi(s);
    o->Elogbeta[((size_t)loc * K + k) * 2 + 0] = gsl_sf_psi(ld[0]) - psi_sum;
    o->Elogbeta[((size_t)loc * K + k) * 2 + 1] = gsl_sf_psi(ld[1]) - psi_sum;
  }
}

/* D1Array::logsum + lognormalize (matrix.hh:271-293): sequential pairwise log-sum-exp */
static void lognormalize(double *d, uint32_t n) {
  double r = d[0];
  if (n > 1)
    for (uint32_t i = 1; i < n; ++i)
      if (d[i] < r) r = r + log(1 + exp(d[i] - r));
      else r = d[i] + log(1 + exp(r - d[i]));
  for (uint32_t i = 0; i < n; ++i) d[i] = exp(d[i] - r);
}

/* PhiRunnerE::update_gamma + estimate_theta (snpsamplinge.cc:695-740), update_rho_indiv (:688-693) */
static void gamma_step(tso *o, uint32_t loc) {
  uint32_t N = o->n, K = o->k;
  double gamma_scale = o->l;
  for (uint32_t n = 0; n < N; ++n) {
    if (!kv_ok(o, n, loc)) continue;
    double rho = pow(o->nodetau0 + o->c_indiv[n], -1 * o->nodekappa);
    o->c_indiv[n]++;
    uint8_t y = o->y[(size_t)loc * N + n];
    double *gd = o->gamma + (size_t)n * K;
    for (uint32_t k = 0; k < K; ++k)
      gd[k] += rho * (o->alpha + (gamma_scale * (y * o->phimom[(size_t)n * K + k] +
                                                 (2 - y) * o->phidad[(size_t)n * K + k])) - gd[k]);
  }
  for (uint32_t n = 0; n < N; ++n) refresh_theta_row(o, n);
}

/* optimize_lambda (snpsamplinge.cc:320-366) with the single worker's process()
 * (snpsamplinge.hh:416-431, :276-300) and update_lambda_t (snpsamplinge.cc:742-759) inlined.
 * do_work (snpsamplinge.cc:649-686): on a new SNP the worker first applies the lazy gamma
 * step of the previous SNP unless that one ran in hol mode. */
static uint32_t optimize_lambda(tso *o, uint32_t loc, int hol_mode) {
  uint32_t N = o->n, K = o->k;
  if (o->pending) { gamma_step(o, o->pending_loc); o->pending = 0; }
  double *lt = (double *)malloc(sizeof(double) * K * 2);
  double *old = (double *)malloc(sizeof(double) * K * 2);
  double *ph = (double *)malloc(sizeof(double) * K);
  uint32_t x = 0;
  do {
    for (uint32_t n = 0; n < N; ++n) {
      if (!kv_ok(o, n, loc)) continue;
      for (uint32_t t = 0; t < 2; ++t) {
        for (uint32_t k = 0; k < K; ++k)
          ph[k] = o->Elogtheta[(size_t)n * K + k] + o->Elogbeta[((size_t)loc * K + k) * 2 + t];
        lognormalize(ph, K);
        memcpy((t == 0 ? o->phimom : o->phidad) + (size_t)n * K, ph, sizeof(double) * K);
      }
    }
    for (uint32_t k = 0; k < K; ++k) {
      double s0 = .0, s1 = .0;
      for (uint32_t n = 0; n < N; ++n) {
        if (!kv_ok(o, n, loc)) continue;
        uint8_t y = o->y[(size_t)loc * N + n];
        s0 += o->phimom[(size_t)n * K + k] * y;
        s1 += o->phidad[(size_t)n * K + k] * (2 - y);
      }
      /* main thread: _lambdat.zero(); _lambdat += t->lambdat() (cc:337-352) */
      lt[k * 2 + 0] = .0 + s0;
      lt[k * 2 + 1] = .0 + s1;
    }
    double *ld = o->lambda + (size_t)loc * K * 2;
    memcpy(old, ld, sizeof(double) * K * 2);
    for (uint32_t k = 0; k < K; ++k) {            /* update_lambda (cc:267-277) */
      ld[k * 2 + 0] = o->eta0 + lt[k * 2 + 0];
      ld[k * 2 + 1] = o->eta1 + lt[k * 2 + 1];
    }
    estimate_beta(o, loc);
    double s = .0;                                 /* sub + abs_mean (matrix.hh:873-893) */
    for (uint32_t i = 0; i < K * 2; ++i) s += fabs(ld[i] - old[i]);
    x++;
    if (s / (K * 2) < o->meanchangethresh) break;
  } while (x < o->online_iterations);
  free(lt); free(old); free(ph);
  if (!hol_mode) { o->pending = 1; o->pending_loc = loc; }
  o->last_rounds = x;
  return x;
}

/* ------------------------------------------------------------------ public API */

/* ctor path (snpsamplinge.cc:6-120) up to, not including, the initial likelihood.
 * `seeded` mirrors `if (env.seed) gsl_rng_set(...)`. */
tso *tso_create(uint32_t n, uint32_t l, uint32_t k, const uint8_t *y_snp_major, double seed,
                uint32_t online_iterations, int compute_beta_mode) {
  tso *o = (tso *)calloc(1, sizeof(tso));
  o->n = n; o->l = l; o->k = k; o->y = y_snp_major;
  o->alpha = (double)1.0 / k; o->eta0 = 1.0; o->eta1 = 1.0;
  o->nodetau0 = 1 + 1; o->nodekappa = 0.5;
  o->meanchangethresh = 0.001; o->stop_threshold = 1e-5;
  o->online_iterations = online_iterations;
  o->prev_h = -2147483647; o->max_h = -2147483647;
  gsl_rng_env_setup();
  o->r = gsl_rng_alloc(gsl_rng_default);
  if (seed) gsl_rng_set(o->r, (unsigned long)seed);
  size_t nk = (size_t)n * k, lk = (size_t)l * k;
  o->gamma = (double *)calloc(nk, sizeof(double));
  o->Elogtheta = (double *)calloc(nk, sizeof(double));
  o->Etheta = (double *)calloc(nk, sizeof(double));
  o->phimom = (double *)calloc(nk, sizeof(double));
  o->phidad = (double *)calloc(nk, sizeof(double));
  o->lambda = (double *)calloc(lk * 2, sizeof(double));
  o->Elogbeta = (double *)calloc(lk * 2, sizeof(double));
  o->Ebeta = (double *)calloc(lk, sizeof(double));
  o->c_indiv = (uint32_t *)calloc(n, sizeof(uint32_t));
  set_validation_sample(o, 0.005);
  if (!compute_beta_mode) {
    init_gamma(o);
    /* init_lambda (cc:239-250): lambda = eta; Elogbeta via set_dir_exp */
    for (size_t i = 0; i < lk; ++i) { o->lambda[2 * i] = o->eta0; o->lambda[2 * i + 1] = o->eta1; }
    for (uint32_t loc = 0; loc < l; ++loc)
      for (uint32_t kk = 0; kk < k; ++kk) {
        const double *ld = o->lambda + ((size_t)loc * k + kk) * 2;
        double s = .0; s += ld[0]; s += ld[1];
        double psi_sum = gsl_sf_psi(s);
        o->Elogbeta[((size_t)loc * k + kk) * 2 + 0] = gsl_sf_psi(ld[0]) - psi_sum;
        o->Elogbeta[((size_t)loc * k + kk) * 2 + 1] = gsl_sf_psi(ld[1]) - psi_sum;
      }
    for (uint32_t i = 0; i < n; ++i) refresh_theta_row(o, i);
  }
  /* compute-beta mode (cc:74-95): lambda/Elogbeta stay zero-filled (never initialised in the
   * reference; zero pages in practice); gamma comes from tso_set_gamma. */
  return o;
}

void tso_destroy(tso *o) {
  if (!o) return;
  free(o->gamma); free(o->Elogtheta); free(o->Etheta); free(o->phimom); free(o->phidad);
  free(o->lambda); free(o->Elogbeta); free(o->Ebeta); free(o->c_indiv);
  for (uint32_t i = 0; i < o->nval_loc; ++i) free(o->val_mask[i]);
  free(o->val_mask); free(o->val_loc); free(o->val_slot); free(o->is_val_loc);
  gsl_rng_free(o->r);
  free(o);
}

uint32_t tso_nval_loc(const tso *o) { return o->nval_loc; }
uint32_t tso_per_loc_h(const tso *o) { return o->per_loc_h; }
/* validation set as CSR: loci ascending, individuals ascending within a locus */
void tso_validation(const tso *o, uint32_t *loc_out, uint32_t *indiv_out) {
  size_t p = 0;
  for (uint32_t i = 0; i < o->nval_loc; ++i) {
    loc_out[i] = o->val_loc[i];
    for (uint32_t n = 0; n < o->n; ++n) if (o->val_mask[i][n]) indiv_out[p++] = n;
  }
}
void tso_get_gamma(const tso *o, double *out) { memcpy(out, o->gamma, sizeof(double) * o->n * o->k); }
void tso_get_theta(const tso *o, double *out) { memcpy(out, o->Etheta, sizeof(double) * o->n * o->k); }
void tso_get_elogtheta(const tso *o, double *out) { memcpy(out, o->Elogtheta, sizeof(double) * o->n * o->k); }
void tso_get_lambda(const tso *o, double *out) { memcpy(out, o->lambda, sizeof(double) * o->l * o->k * 2); }
void tso_get_beta(const tso *o, double *out) { memcpy(out, o->Ebeta, sizeof(double) * o->l * o->k); }
void tso_get_counts(const tso *o, uint32_t *out) { memcpy(out, o->c_indiv, sizeof(uint32_t) * o->n); }
/* load_gamma + estimate_all_theta (cc:800-837, :85-86) */
void tso_set_gamma(tso *o, const double *g) {
  memcpy(o->gamma, g, sizeof(double) * o->n * o->k);
  for (uint32_t i = 0; i < o->n; ++i) refresh_theta_row(o, i);
}
uint32_t tso_iter(const tso *o) { return o->iter; }
uint32_t tso_last_rounds(const tso *o) { return o->last_rounds; }

/* infer loop head (cc:423) */
uint32_t tso_sample_loc(tso *o) { return (uint32_t)gsl_rng_uniform_int(o->r, o->l); }

/* one training iteration on a given locus (cc:425-434) */
uint32_t tso_train_loc(tso *o, uint32_t loc) {
  uint32_t x = optimize_lambda(o, loc, 0);
  o->iter++;
  return x;
}

/* apply the owed lazy gamma step now (what the worker would do on its next SNP) */
void tso_flush(tso *o) {
  if (o->pending) { gamma_step(o, o->pending_loc); o->pending = 0; }
}

/* compute_likelihood(first, validation=true) (cc:461-544) + snp_likelihood (hh:322-361).
 * Returns 1 if the stopping rule fired.  mean_ll/count as written to validation.txt. */
int tso_heldout(tso *o, int first, double *mean_ll, uint32_t *count, double *per_locus_sum) {
  uint32_t K = o->k, N = o->n;
  double s = .0;
  uint32_t kcount = 0;
  for (uint32_t i = 0; i < o->nval_loc; ++i) {
    uint32_t loc = o->val_loc[i];
    if (first) estimate_beta(o, loc);
    else { optimize_lambda(o, loc, 1); o->iter++; }
    double lsum = .0;
    for (uint32_t n = 0; n < N; ++n) {
      if (!o->val_mask[i][n]) continue;
      uint8_t x = o->y[(size_t)loc * N + n];
      double q = .0;
      double v = gsl_sf_fact(2) / (gsl_sf_fact(x) * gsl_sf_fact(2 - x));
      for (uint32_t k = 0; k < K; ++k) q += o->Ebeta[(size_t)loc * K + k] * o->Etheta[(size_t)n * K + k];
      double sum = v * pow(q, x) * pow(1 - q, 2 - x);
      if (sum < 1e-30) sum = 1e-30;
      lsum += log(sum);
      kcount++;
    }
    if (per_locus_sum) per_locus_sum[i] = lsum;
    s += lsum;
  }
  double a = s / kcount;
  *mean_ll = a; *count = kcount;
  int stop = 0;
  if (o->iter > 2000) {
    if (a > o->prev_h && o->prev_h != 0 && fabs((a - o->prev_h) / o->prev_h) < o->stop_threshold) stop = 1;
    else if (a < o->prev_h) o->nh++;
    else if (a > o->prev_h) o->nh = 0;
    if (a > o->max_h) o->max_h = a;
    if (o->nh > 3) stop = 1;
  }
  o->prev_h = a;
  return stop;
}

/* The whole of infer() (cc:417-459) after the ctor's initial report: runs until the stopping
 * rule fires or max_iter training+validation iterations have elapsed.  Report rows
 * (iter, mean LL, count) are appended to the out arrays (capacity cap); the training loci
 * sampled are appended to locs_out (capacity lcap) when non-NULL.  Returns #report rows. */
uint32_t tso_infer(tso *o, uint32_t rfreq, uint32_t max_iter, uint32_t cap, uint32_t *rep_iter,
                   double *rep_ll, uint32_t *rep_count, uint32_t lcap, uint32_t *locs_out,
                   uint32_t *nlocs_out, int *stopped) {
  uint32_t nrep = 0, nl = 0;
  *stopped = 0;
  while (o->iter < max_iter) {
    uint32_t loc = tso_sample_loc(o);
    if (locs_out && nl < lcap) locs_out[nl] = loc;
    nl++;
    tso_train_loc(o, loc);
    if (o->iter % rfreq == 0) {
      double a; uint32_t c;
      int stop = tso_heldout(o, 0, &a, &c, NULL);
      if (nrep < cap) { rep_iter[nrep] = o->iter; rep_ll[nrep] = a; rep_count[nrep] = c; }
      nrep++;
      if (stop) { *stopped = 1; break; }
    }
  }
  if (nlocs_out) *nlocs_out = nl;
  return nrep;
}

/* -compute-beta sweep (compute_all_lambda cc:368-381 + estimate_all_beta cc:611-625):
 * every locus in order, lazy gamma steps keep being applied between loci. */
void tso_compute_all_lambda(tso *o) {
  for (uint32_t loc = 0; loc < o->l; ++loc) { optimize_lambda(o, loc, 0); o->iter++; }
  for (uint32_t loc = 0; loc < o->l; ++loc)
    for (uint32_t k = 0; k < o->k; ++k) {
      const double *ld = o->lambda + ((size_t)loc * o->k + k) * 2;
      double s = .0; s += ld[0]; s += ld[1];
      o->Ebeta[(size_t)loc * o->k + k] = ld[0] / s;
    }
}

/* PLINK .bed SNP-major 2-bit decode (snp.cc:186-229): 00->0, 10->1, 11->2, 01->3(missing) */
void tso_decode_bed(const uint8_t *bed, uint32_t n, uint32_t l, uint8_t *y_out) {
  static const uint8_t map[4] = {0, 3, 1, 2};
  size_t bps = (n + 3) / 4;
  for (uint32_t loc = 0; loc < l; ++loc)
    for (uint32_t i = 0; i < n; ++i)
      y_out[(size_t)loc * n + i] = map[(bed[loc * bps + (i >> 2)] >> (2 * (i & 3))) & 3];
}
