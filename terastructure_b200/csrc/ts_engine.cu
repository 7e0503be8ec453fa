// libtsgpu.so -- B200-native engine for the TeraStructure SNPSamplingE hot path.
//
// Replaces the PhiRunnerE thread pool / TSQueue hand-off of the reference
// (src/snpsamplinge.cc:320-366 <-> :649-759) with stream-ordered sm_100a kernels.
//
// Device-resident state of one engine (one shard of individuals on one GPU):
//   bed    [L][pitch]      PLINK 2-bit packed genotypes, SNP-major, this shard's bytes only
//   vcol   [nval][pitch]   copies of the validation loci's columns in which the held-out
//                          individuals carry the "missing" code, so the hot kernels need no mask
//   gamma  [K][npad]       fp64, population-major so that a warp reads 32 consecutive doubles
//   E      [K][npad]       exp(psi(gamma)); all the E-step needs (see ts_math.cuh)
//   cnt    [npad]          per-individual step counts (_c_indiv, snpsamplinge.cc:688-693)
//   lambda [L][K][2]       fp64; Ebeta/Elogbeta are derived from it on the fly
//
// Kernels (all on the engine's stream, no host round trip):
//   tsp::k_persist<K, I, TIER>   a whole batch of SVI iterations in one cooperative launch (ts_persist.cuh)
//   k_heldout<K>                 held-out log-likelihood of one validation locus          (hol mode)
// TEST-ONLY build (-DTS_STAGED_PATH, lib/libtsgpu_staged.so, never the product library): the first
// correct path, one launch per round, kept as an independent implementation of the same mathematics
// that tests/ cross-check the persistent kernel against (TSGPU_PATH=staged):
//   k_begin            1 CTA: advance the work cursor, b[k][t] = exp(Elogbeta[loc][k][t])
//   k_estep<K> x I     E-step + fp64 sufficient-statistic reduction + lambda update; later
//                      launches return immediately once the round loop has converged
//   k_gamma<K>         gamma natural-gradient step + exp(psi(gamma)) refresh   (training)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ts_device.cuh"
#include "ts_fixed.cuh"
#include "ts_symm.hpp"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int set_err(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                             \
  do {                                                                                       \
    cudaError_t err__ = (call);                                                              \
    if (err__ != cudaSuccess)                                                                \
      return set_err(TS_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,       \
                     cudaGetErrorString(err__));                                             \
  } while (0)

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
#ifdef TS_STAGED_PATH
constexpr int ESTEP_THREADS = 256;

// b[k][t] = exp(psi(lambda[k][t]) - psi(lambda[k][0]+lambda[k][1]))  (estimate_beta, cc:279-296)
__device__ __forceinline__ void elogbeta_exp(double l0, double l1, double &b0, double &b1) {
  const double ps = tsm::digamma(l0 + l1);
  b0 = exp(tsm::digamma(l0) - ps);
  b1 = exp(tsm::digamma(l1) - ps);
}

__global__ void k_begin(Params p, int K) {
  Ctl *c = p.ctl;
  __shared__ long long cur;
  if (threadIdx.x == 0) {
    cur = c->cursor + 1;
    c->cursor = cur;
  }
  __syncthreads();
  const WorkItem it = p.items[cur];
  if ((int)threadIdx.x < K) {
    const int k = threadIdx.x;
    const double *lam = p.lambda + ((size_t)it.loc * K + k) * 2;
    double b0, b1;
    elogbeta_exp(lam[0], lam[1], b0, b1);
    c->bcur[2 * k] = b0;
    c->bcur[2 * k + 1] = b1;
    c->bprev[2 * k] = b0;
    c->bprev[2 * k + 1] = b1;
  }
  if (threadIdx.x == 0) {
    c->x = 0;
    c->done = (it.flags & ITEM_FIRST) ? 1u : 0u;
    c->ticket = 0;
    if (it.flags & ITEM_FIRST) p.rounds[cur] = 0;
  }
}

// One round of optimize_lambda (cc:320-366): the workers' process()+update_lambda_t
// (hh:416-431, cc:742-759) for every individual of the shard, the sum over workers
// (cc:337-352), update_lambda (cc:267-277), estimate_beta (cc:279-296) and the convergence
// test (cc:359-364).
//
// phi_t[n][k] = softmax_k(Elogtheta[n][k] + Elogbeta[k][t]) = E[n][k] b[k][t] / s_t[n] with
// s_t[n] = sum_k E[n][k] b[k][t], so  S_t[k] = sum_n phi_t[n][k] w_t[n] = b[k][t] * sum_n E[n][k] w_t[n]/s_t[n]
// (w_0 = y, w_1 = 2-y): per individual 4K FMAs and two divisions, no transcendental.
template <int K>
__global__ void __launch_bounds__(ESTEP_THREADS) k_estep(Params p) {
  Ctl *c = p.ctl;
  if (c->done) return;
  const WorkItem it = p.items[c->cursor];
  const unsigned char *col = it.col;

  double b0[K], b1[K], a0[K], a1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    b0[k] = c->bcur[2 * k];
    b1[k] = c->bcur[2 * k + 1];
    a0[k] = 0.0;
    a1[k] = 0.0;
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < p.n_local; n += stride) {
    const int code = tsm::plink_code(col, n);
    if (code == 1) continue;  // missing or held out: kv_ok false (hh:389-408)
    const int y = tsm::code_to_y(code);
    double e[K];
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      e[k] = p.E[(size_t)k * p.npad + n];
      s0 = fma(e[k], b0[k], s0);
      s1 = fma(e[k], b1[k], s1);
    }
    const double r0 = (double)y / s0;
    const double r1 = (double)(2 - y) / s1;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      a0[k] = fma(e[k], r0, a0[k]);
      a1[k] = fma(e[k], r1, a1[k]);
    }
  }

  // CTA reduction in a fixed order: warp butterflies, then warps in index order.
  __shared__ double sm[ESTEP_THREADS / 32][2 * K];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double v0 = warp_sum(a0[k]);
    const double v1 = warp_sum(a1[k]);
    if (lane == 0) {
      sm[warp][2 * k] = v0;
      sm[warp][2 * k + 1] = v1;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * K) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < ESTEP_THREADS / 32; ++w) s += sm[w][threadIdx.x];
    p.partial[(size_t)blockIdx.x * (2 * K) + threadIdx.x] = s;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&c->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();

  // Last CTA: add the CTA partials (lanes stride over CTAs, then a butterfly: fixed order).
  __shared__ double tot[2 * K];
  for (int v = warp; v < 2 * K; v += ESTEP_THREADS / 32) {
    double s = 0.0;
    for (uint32_t g = lane; g < gridDim.x; g += 32) s += __ldcg(&p.partial[(size_t)g * (2 * K) + v]);
    s = warp_sum(s);
    if (lane == 0) tot[v] = s * c->bcur[v];  // S_t[k] = b[k][t] * sum
  }
  __syncthreads();

  // Multi-GPU: publish this shard's S into every peer's slot, then add slots in rank order
  // (the reference's main thread adding each worker's lambdat, cc:337-352).
  if (p.nranks > 1) {
    const uint32_t ep = c->epoch + 1;
    const int par = ep & 1;
    if (threadIdx.x < 2 * K)
      for (int r = 0; r < p.nranks; ++r) p.xpeer[r]->val[par][p.rank][threadIdx.x] = tot[threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p.nranks) {
      volatile unsigned long long *f = &p.xpeer[threadIdx.x]->flag[par][p.rank];
      *f = ep;
      volatile unsigned long long *mine = &p.xlocal->flag[par][threadIdx.x];
      long long spins = c->fault ? (1ll << 28) : 0;  // once a peer is lost, do not wait again
      while (*mine < ep) {
        if (++spins > (1ll << 28)) {  // a peer died: flag the fault and carry on
          c->fault = 1;
          break;
        }
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < 2 * K) {
      double s = 0.0;
      for (int r = 0; r < p.nranks; ++r) s += ((volatile double *)p.xlocal->val[par][r])[threadIdx.x];
      tot[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) c->epoch = ep;
    __syncthreads();
  }

  __shared__ double chg[K];
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    double *lam = p.lambda + ((size_t)it.loc * K + k) * 2;
    const double n0 = p.eta0 + tot[2 * k], n1 = p.eta1 + tot[2 * k + 1];
    chg[k] = fabs(n0 - lam[0]) + fabs(n1 - lam[1]);
    lam[0] = n0;
    lam[1] = n1;
    double nb0, nb1;
    elogbeta_exp(n0, n1, nb0, nb1);
    c->bprev[2 * k] = c->bcur[2 * k];
    c->bprev[2 * k + 1] = c->bcur[2 * k + 1];
    c->bcur[2 * k] = nb0;
    c->bcur[2 * k + 1] = nb1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += chg[k];
    const uint32_t x = c->x + 1;
    c->x = x;
    c->ticket = 0;
    if (s / (2 * K) < p.thresh || x >= p.max_rounds) {
      c->done = 1;
      p.rounds[c->cursor] = x;
    }
  }
}

// PhiRunnerE::update_gamma + estimate_theta (cc:695-740) with update_rho_indiv (cc:688-693):
// phi is the phi of the LAST E-step, i.e. built from bprev.
template <int K>
__global__ void __launch_bounds__(256) k_gamma(Params p) {
  const Ctl *c = p.ctl;
  const WorkItem it = p.items[c->cursor];
  if (it.flags & ITEM_HOL) return;
  const unsigned char *col = it.col;
  double b0[K], b1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    b0[k] = c->bprev[2 * k];
    b1[k] = c->bprev[2 * k + 1];
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < p.n_local; n += stride) {
    const int code = tsm::plink_code(col, n);
    if (code == 1) continue;
    const int y = tsm::code_to_y(code);
    double e[K];
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      e[k] = p.E[(size_t)k * p.npad + n];
      s0 = fma(e[k], b0[k], s0);
      s1 = fma(e[k], b1[k], s1);
    }
    const double r0 = (double)y / s0;
    const double r1 = (double)(2 - y) / s1;
    const uint32_t cn = p.cnt[n];
    const double base = p.nodetau0 + (double)cn;
    const double rho = (p.nodekappa == 0.5) ? rsqrt(base) : pow(base, -p.nodekappa);
    p.cnt[n] = cn + 1;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double g = p.gamma[(size_t)k * p.npad + n];
      // y*phimom + (2-y)*phidad = E[k] * (b0[k]*y/s0 + b1[k]*(2-y)/s1)
      const double w = e[k] * fma(b0[k], r0, b1[k] * r1);
      const double gn = g + rho * (p.alpha + p.lscale * w - g);
      p.gamma[(size_t)k * p.npad + n] = gn;
      p.E[(size_t)k * p.npad + n] = tsm::exp_digamma(gn);
    }
  }
}

#endif  // TS_STAGED_PATH

// snp_likelihood (hh:322-361): one CTA per work item (= validation locus).  gamma does not
// change during a held-out pass and every locus owns its lambda row, so the whole pass is
// scored by one launch after the hol-mode optimisations.
template <int K>
__global__ void __launch_bounds__(256) k_heldout(Params p) {
  const WorkItem it = p.items[blockIdx.x];
  if (it.vslot < 0 || !(it.flags & ITEM_HOL)) return;
  const unsigned char *col = p.bed + (size_t)it.loc * p.pitch;  // the unmasked column
  double beta[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double *lam = p.lambda + ((size_t)it.loc * K + k) * 2;
    double s = 0.0;
    s += lam[0];
    s += lam[1];
    beta[k] = lam[0] / s;
  }
  double acc = 0.0;
  const unsigned long long lo = p.voff[it.vslot], hi = p.voff[it.vslot + 1];
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint32_t n = p.vind[i];
    const int y = tsm::code_to_y(tsm::plink_code(col, n));
    double g[K], s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      g[k] = p.gamma[(size_t)k * p.npad + n];
      s += g[k];
    }
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) q += beta[k] * (g[k] / s);
    const double v = (y == 1) ? 2.0 : 1.0;  // C(2,y)
    const double pq = (y == 0) ? 1.0 : (y == 1 ? q : q * q);
    const double p1 = (y == 2) ? 1.0 : (y == 1 ? (1.0 - q) : (1.0 - q) * (1.0 - q));
    double t = v * pq * p1;
    if (t < 1e-30) t = 1e-30;
    acc += log(t);
  }
  __shared__ double sm[256 / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 256 / 32; ++w) s += sm[w];
    p.ll[it.vslot] = s;
  }
}

// E = exp(psi(gamma)) for the whole shard (after ts_set_gamma).
__global__ void k_refresh_E(const double *gamma, double *E, size_t npad, uint32_t n_local, int K) {
  const size_t total = (size_t)K * npad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i % npad;
    E[i] = (n < n_local) ? tsm::exp_digamma(gamma[i]) : 0.0;
  }
}

// out[n][k] (row-major) = mode 0: gamma/rowsum ; mode 1: psi(gamma) - psi(rowsum)
__global__ void k_theta(const double *gamma, size_t npad, uint32_t n_local, int K, int mode, double *out) {
  for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < n_local; n += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += gamma[(size_t)k * npad + n];
    const double ps = mode ? tsm::digamma(s) : 0.0;
    for (int k = 0; k < K; ++k) {
      const double g = gamma[(size_t)k * npad + n];
      out[(size_t)n * K + k] = mode ? tsm::digamma(g) - ps : g / s;
    }
  }
}

__global__ void k_fill_lambda(double *lambda, size_t count, double eta0, double eta1) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (size_t)gridDim.x * blockDim.x)
    lambda[i] = (i & 1) ? eta1 : eta0;
}

__global__ void k_beta(const double *lambda, size_t count, double *out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    s += lambda[2 * i];
    s += lambda[2 * i + 1];
    out[i] = lambda[2 * i] / s;
  }
}

// vcol[v] = bed[val_loc[v]] with the held-out individuals re-coded as missing (01).
__global__ void k_build_vcol(const unsigned char *bed, size_t pitch, const uint32_t *val_loc,
                             const unsigned long long *voff, const uint32_t *vind, unsigned char *vcol) {
  const int v = blockIdx.x;
  const uint4 *src = reinterpret_cast<const uint4 *>(bed + (size_t)val_loc[v] * pitch);
  uint4 *dst = reinterpret_cast<uint4 *>(vcol + (size_t)v * pitch);
  for (size_t i = threadIdx.x; i < pitch / 16; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  unsigned int *w = reinterpret_cast<unsigned int *>(vcol + (size_t)v * pitch);
  for (unsigned long long i = voff[v] + threadIdx.x; i < voff[v + 1]; i += blockDim.x) {
    const uint32_t n = vind[i];
    const unsigned sh = 2 * (n & 15);
    atomicAnd(&w[n >> 4], ~(2u << sh));
    atomicOr(&w[n >> 4], 1u << sh);
  }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Synthetic PSD genotypes, one output byte (4 individuals) per thread iteration.
__global__ void k_synth(unsigned char *bed, size_t pitch, uint64_t L, uint32_t n_local, uint64_t n_begin,
                        int K, const float *theta, const float *beta, unsigned long long seed,
                        float missing_rate) {
  const size_t bytes = (n_local + 3) / 4;
  const size_t total = (size_t)L * bytes;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t loc = i / bytes, byte = i % bytes;
    const float *bl = beta + loc * K;
    unsigned out = 0;
    for (int j = 0; j < 4; ++j) {
      const size_t n = byte * 4 + j;
      if (n >= n_local) break;
      float q = 0.f;
      for (int k = 0; k < K; ++k) q = fmaf(theta[n * K + k], __ldg(&bl[k]), q);
      const unsigned long long z =
          mix64(seed + loc * 0x9E3779B97F4A7C15ull + (n_begin + n) * 0xD1B54A32D192ED03ull);
      const float u1 = (float)(z >> 40) * (1.0f / 16777216.0f);
      const float u2 = (float)((z >> 16) & 0xFFFFFFull) * (1.0f / 16777216.0f);
      const int y = (u1 < q) + (u2 < q);
      unsigned code = y + (y > 0);  // 0->00, 1->10, 2->11
      if (missing_rate > 0.f) {
        const float u3 = (float)(mix64(z) >> 40) * (1.0f / 16777216.0f);
        if (u3 < missing_rate) code = 1;
      }
      out |= code << (2 * j);
    }
    bed[loc * pitch + byte] = (unsigned char)out;
  }
}

// ------------------------------------------------------------------------------------------
// host-side engine
// ------------------------------------------------------------------------------------------
struct ts_engine {
  ts_config cfg;
  int K = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int num_sms = 0;
  size_t pitch = 0, npad = 0, local_bytes = 0;
  unsigned char *bed = nullptr, *vcol = nullptr;
  double *gamma = nullptr, *E = nullptr, *lambda = nullptr, *partial = nullptr, *ll = nullptr;
  uint32_t *cnt = nullptr, *rounds = nullptr;
  Ctl *ctl = nullptr;
  WorkItem *items = nullptr;
  size_t items_cap = 0;
  Xchg *xchg = nullptr;      // the exchange buffer in use: xchg_own, or this rank's part of a symmetric group
  Xchg *xchg_own = nullptr;  // cudaMalloc'ed at creation (exported through CUDA IPC)
  Xchg *xchg_mc = nullptr;   // NVLS multicast alias of the symmetric buffer, or null
  Xbuf *xbuf = nullptr;      // &xchg->x
  int xmode = XMODE_GACC;
  unsigned long long mc_arrivals = 0;
  SymmGroup *symm = nullptr; // owned by rank 0's engine of a ts_comm_connect_local group
  bool staged = false;   // TSGPU_PATH=staged: one launch per round (debug cross-check path)
  int grid_persist = 1, block_persist = 32, ind_per_thread = 1;
  bool tier = false;      // TIER kernel: registers + shared memory + streaming (ts_persist.cuh)
  uint64_t n_stream = 0;  // individuals of this shard that stream from L2/HBM every round
  std::vector<void *> ipc_opened;
  // validation set (host copies)
  std::vector<uint32_t> val_loc;         // ascending
  std::vector<unsigned long long> voff;  // local CSR
  std::vector<uint32_t> vind;            // local ids
  unsigned long long *d_voff = nullptr;
  uint32_t *d_vind = nullptr;
  uint32_t *d_val_loc = nullptr;
  bool bed_loaded = false, gamma_set = false;
  int grid_estep = 1, grid_gamma = 1;
  uint64_t launches = 0;
  Params prm;
  std::vector<WorkItem> h_items;
};

template <typename T>
static cudaError_t dalloc(T **p, size_t count) {
  return cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
}

static int use_device(ts_engine *e) {
  CK(cudaSetDevice(e->cfg.device));
  // a stale non-sticky error of an earlier runtime call in this thread (ours, e.g. from a refused ts_create's
  // clean-up, or the host application's) must not be reported by the cudaGetLastError() after our next launch
  (void)cudaGetLastError();
  return TS_OK;
}

static void fill_params(ts_engine *e) {
  Params &p = e->prm;
  p.bed = e->bed;
  p.pitch = e->pitch;
  p.gamma = e->gamma;
  p.E = e->E;
  p.cnt = e->cnt;
  p.npad = e->npad;
  p.n_local = (uint32_t)e->cfg.n_local;
  p.lambda = e->lambda;
  p.ctl = e->ctl;
  p.items = e->items;
  p.partial = e->partial;
  p.rounds = e->rounds;
  p.voff = e->d_voff;
  p.vind = e->d_vind;
  p.ll = e->ll;
  p.alpha = e->cfg.alpha;
  p.eta0 = e->cfg.eta0;
  p.eta1 = e->cfg.eta1;
  p.nodetau0 = e->cfg.nodetau0;
  p.nodekappa = e->cfg.nodekappa;
  p.thresh = e->cfg.meanchangethresh;
  p.lscale = (double)e->cfg.l;
  p.max_rounds = e->cfg.online_iterations;
  p.rank = e->cfg.rank;
  p.nranks = e->cfg.nranks;
  p.xlocal = e->xbuf;
  p.pst = &e->xchg->ps;
  p.pst_mc = e->xchg_mc ? &e->xchg_mc->ps : nullptr;
  p.xmode = e->xmode;
  p.mc_arrivals = e->mc_arrivals;
}

// K dispatch ---------------------------------------------------------------------------------
#define TS_FOR_EACH_K(X)                                                                       \
  X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) \
  X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)

#ifdef TS_STAGED_PATH
static void launch_estep(ts_engine *e) {
  switch (e->K) {
#define X(k) \
  case k: k_estep<k><<<e->grid_estep, ESTEP_THREADS, 0, e->stream>>>(e->prm); break;
    TS_FOR_EACH_K(X)
#undef X
  }
}
static void launch_gamma(ts_engine *e) {
  switch (e->K) {
#define X(k) \
  case k: k_gamma<k><<<e->grid_gamma, 256, 0, e->stream>>>(e->prm); break;
    TS_FOR_EACH_K(X)
#undef X
  }
}
#endif
static void launch_heldout(ts_engine *e, unsigned n_items) {
  switch (e->K) {
#define X(k) \
  case k: k_heldout<k><<<n_items, 256, 0, e->stream>>>(e->prm); break;
    TS_FOR_EACH_K(X)
#undef X
  }
}

static cudaError_t launch_persist(ts_engine *e, uint32_t n_items) {
  return ts_launch_persist(e->K, e->ind_per_thread, e->tier, e->prm, n_items, e->grid_persist, e->block_persist, e->stream);
}

// How the ranks' totals of a round are exchanged (XMODE_*, ts_device.cuh): CTA 0 of every GPU adds the GPU's totals
// into an accumulator replicated on every rank -- with one multimem.red per word where an NVLS alias exists
// (XMODE_MCACC), with one NVLink red.add per peer and word otherwise (XMODE_GACC) -- and every CTA polls one local
// word pair.  Measured against the alternatives on 8 B200s (profiles/r2_summary.md section 4; cycles per round in
// isolation): slots written by every GPU and polled by every CTA 8 332, the same with one multimem.st 8 591,
// multimem.red from every CTA 18 673, this scheme 5 280 (unicast) / 4 947 (multicast); in the product on 2 B200s,
// us per SVI iteration: slots 60.1, gacc 51.2, mcacc 48.9.  The rejected schemes are no longer compiled in.
// TSGPU_XCHG = gacc forces the unicast form (mcacc, or nothing: multicast when available).
static int choose_xmode(bool have_mc) {
  const char *xm = getenv("TSGPU_XCHG");
  if (xm && (!strcmp(xm, "gacc") || !strcmp(xm, "slots") || !strcmp(xm, "ipc"))) return XMODE_GACC;
  return have_mc ? XMODE_MCACC : XMODE_GACC;
}

// Move the engine's exchange state into rank_ptrs[rank] (one buffer per rank, all reachable from this
// device) and pick the exchange mode.  The caller guarantees that no rank steps before all have attached.
static int attach_symmetric(ts_engine *e, void *const *rank_ptrs, void *mc, unsigned long long total_ctas) {
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaStreamSynchronize(e->stream));
  Xchg *mine = (Xchg *)rank_ptrs[e->cfg.rank];
  CK(cudaMemcpy(mine, e->xchg, sizeof(Xchg), cudaMemcpyDeviceToDevice));  // barrier history, round counter
  CK(cudaDeviceSynchronize());
  e->xchg = mine;
  e->xbuf = &mine->x;
  e->xchg_mc = (Xchg *)mc;
  for (int r = 0; r < e->cfg.nranks; ++r) {
    e->prm.xpeer[r] = &((Xchg *)rank_ptrs[r])->x;
    e->prm.pst_peer[r] = &((Xchg *)rank_ptrs[r])->ps;
  }
  e->mc_arrivals = total_ctas;
  e->xmode = choose_xmode(mc != nullptr);
  fill_params(e);
  return TS_OK;
}

extern "C" {

const char *ts_last_error(void) { return g_err.c_str(); }
int ts_abi_version(void) { return TSGPU_ABI_VERSION; }

int ts_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void ts_config_defaults(ts_config *cfg, uint64_t n, uint64_t l, uint32_t k) {
  memset(cfg, 0, sizeof *cfg);
  cfg->n_total = n;
  cfg->n_begin = 0;
  cfg->n_local = n;
  cfg->l = l;
  cfg->k = k;
  cfg->online_iterations = 10;
  cfg->alpha = k ? 1.0 / k : 0.0;
  cfg->eta0 = 1.0;
  cfg->eta1 = 1.0;
  cfg->nodetau0 = 2.0;
  cfg->nodekappa = 0.5;
  cfg->meanchangethresh = 1e-3;
  cfg->device = 0;
  cfg->rank = 0;
  cfg->nranks = 1;
}

// Launch geometry of the persistent kernel for a shard of n individuals.
// Register tier: I individuals per thread live in registers; among the I that fit, take the one with
// the fewest warps on the busiest of the SM's four schedulers (the warp-level reduction and the CTA
// sum are issue-bound per scheduler), not going below two; ties go to the smaller I (shorter
// dependent FP64 chains, fewer registers).  B200, K = 10, us per SVI iteration: 60K individuals
// I=1/13 warps 30.9, I=2/7 warps 28.3; 80K I=2/9 warps 30.9, I=3/6 warps 28.9; 100K I=2/11 warps 31.2,
// I=3/8 warps 29.7, I=4/6 warps 32.5.
// Shards beyond the register-resident capacity run the TIER kernel: itier(K) individuals per thread
// in registers, J more in shared memory (as many as fit), the rest streamed from L2/HBM every round.
struct ShardPlan {
  int ipt = 1, grid = 1, block = 32;
  bool tier = false;
  int tier_j = 0;
  uint64_t n_stream = 0;
};
constexpr size_t SMEM_OPTIN_B200 = 232448;  // 227 KB of dynamic shared memory per CTA (sm_100)
// pin < 0: choose; pin == 0: force the TIER kernel; pin >= 1: the first register-only I >= pin that fits.
// pin_j / pin_grid / pin_block (< 0 or 0: choose) shape the TIER kernel in tests.
static ShardPlan plan_shard(uint64_t n, int K, int num_sms, int pin, int pin_j = -1, int pin_grid = 0, int pin_block = 0,
                            size_t smem_limit = SMEM_OPTIN_B200) {
  ShardPlan pl;
  int I = 0, best = 1 << 30;
  for (int c = std::max(pin, 1); pin != 0 && c <= ts_persist_imax(K); ++c) {
    if ((uint64_t)num_sms * ts_persist_tmax(K, c) * c < n) continue;
    if (pin > 0) { I = c; break; }  // the knob pins the first I >= its value that fits
    const uint64_t threads = (n + c - 1) / c;
    const uint64_t grid = std::min<uint64_t>(num_sms, (threads + 63) / 64);
    const int warps = (int)(((threads + grid - 1) / grid + 31) / 32);
    const int score = std::max((warps + 3) / 4, 2);
    if (score < best) { best = score; I = c; }
  }
  if (I > 0) {
    const uint64_t threads = (n + I - 1) / I;
    pl.ipt = I;
    pl.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(num_sms, (threads + 63) / 64));  // all SMs once there are 2 warps each
    const uint64_t t = (threads + pl.grid - 1) / pl.grid;
    pl.block = (int)std::min<uint64_t>(ts_persist_tmax(K, I), std::max<uint64_t>(32, (t + 31) / 32 * 32));
    return pl;
  }
  pl.tier = true;
  pl.ipt = ts_persist_itier(K);
  pl.block = pin_block > 0 ? std::min(pin_block, ts_persist_tier_threads()) / 32 * 32 : ts_persist_tier_threads();
  pl.block = std::max(pl.block, 32);
  pl.grid = pin_grid > 0 ? std::min(pin_grid, num_sms) : num_sms;
  const uint64_t gt = (uint64_t)pl.grid * pl.block;
  const uint64_t rest = n > (uint64_t)pl.ipt * gt ? n - (uint64_t)pl.ipt * gt : 0;
  const size_t base = ts_persist_smem_base(K, ts_persist_tier_threads());
  const int jmax = (int)std::min<size_t>(16, smem_limit > base ? (smem_limit - base) / ts_persist_tier_slot_bytes(K) : 0);
  pl.tier_j = pin_j >= 0 ? std::min(pin_j, jmax) : (int)std::min<uint64_t>(jmax, (rest + gt - 1) / gt);
  const uint64_t held = (uint64_t)(pl.ipt + pl.tier_j) * gt;
  pl.n_stream = n > held ? n - held : 0;
  return pl;
}

int ts_plan_shard(uint64_t n_local, int k, int num_sms, int *ind_per_thread, int *grid, int *block) {
  if (k < 1 || k > TS_MAX_K || num_sms < 1 || !ind_per_thread || !grid || !block)
    return set_err(TS_ERR_ARG, "ts_plan_shard: K=%d outside 1..%d, num_sms=%d or null output", k, TS_MAX_K, num_sms);
  const ShardPlan pl = plan_shard(n_local, k, num_sms, -1);
  *ind_per_thread = pl.ipt;
  *grid = pl.grid;
  *block = pl.block;
  return TS_OK;
}

int ts_plan_tiers(uint64_t n_local, int k, int num_sms, int *smem_per_thread, uint64_t *n_streamed) {
  if (k < 1 || k > TS_MAX_K || num_sms < 1 || !smem_per_thread || !n_streamed)
    return set_err(TS_ERR_ARG, "ts_plan_tiers: K=%d outside 1..%d, num_sms=%d or null output", k, TS_MAX_K, num_sms);
  const ShardPlan pl = plan_shard(n_local, k, num_sms, -1);
  *smem_per_thread = pl.tier ? pl.tier_j : 0;
  *n_streamed = pl.n_stream;
  return TS_OK;
}

int ts_get_plan(const ts_engine *e, int *ind_per_thread, int *grid, int *block);

int ts_create(const ts_config *cfg, ts_engine **out) {
  if (!cfg || !out) return set_err(TS_ERR_ARG, "ts_create: null argument");
  *out = nullptr;
  if (cfg->k < 1 || cfg->k > TS_MAX_K)
    return set_err(TS_ERR_ARG, "ts_create: K=%u outside 1..%d", cfg->k, TS_MAX_K);
  if (cfg->n_local == 0 || cfg->l == 0 || cfg->n_local > 0xfffffff0ull || cfg->l > 0xffffffffull)
    return set_err(TS_ERR_ARG, "ts_create: bad shape n_local=%llu l=%llu",
                   (unsigned long long)cfg->n_local, (unsigned long long)cfg->l);
  if (cfg->n_begin % 4 != 0 || cfg->n_begin + cfg->n_local > cfg->n_total)
    return set_err(TS_ERR_ARG, "ts_create: shard [%llu,+%llu) must start on a multiple of 4 inside N=%llu",
                   (unsigned long long)cfg->n_begin, (unsigned long long)cfg->n_local,
                   (unsigned long long)cfg->n_total);
  if (cfg->nranks < 1 || cfg->nranks > MAXR || cfg->rank < 0 || cfg->rank >= cfg->nranks)
    return set_err(TS_ERR_ARG, "ts_create: bad rank %d of %d", cfg->rank, cfg->nranks);
  if (cfg->online_iterations < 1) return set_err(TS_ERR_ARG, "ts_create: online_iterations < 1");
  int ndev = ts_device_count();
  if (ndev <= 0) return set_err(TS_ERR_CUDA, "ts_create: no CUDA device (this library has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev)
    return set_err(TS_ERR_ARG, "ts_create: device %d of %d", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10)
    return set_err(TS_ERR_CUDA, "ts_create: device %d is sm_%d%d; this library is built for sm_100a only",
                   cfg->device, prop.major, prop.minor);

  ts_engine *e = new ts_engine;
  e->cfg = *cfg;
  e->K = (int)cfg->k;
  e->num_sms = prop.multiProcessorCount;
  e->local_bytes = (cfg->n_local + 3) / 4;
  e->pitch = (e->local_bytes + 15) / 16 * 16;
  e->npad = (cfg->n_local + 31) / 32 * 32;
  const size_t K = e->K;
#define CKE(call)                      \
  do {                                 \
    cudaError_t er_ = (call);          \
    if (er_ != cudaSuccess) {          \
      set_err(TS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(er_)); \
      ts_destroy(e);                   \
      return TS_ERR_CUDA;              \
    }                                  \
  } while (0)
  CKE(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CKE(cudaEventCreate(&e->ev0));
  CKE(cudaEventCreate(&e->ev1));
  CKE(dalloc(&e->bed, (size_t)cfg->l * e->pitch));
  CKE(dalloc(&e->gamma, K * e->npad));
  CKE(dalloc(&e->E, K * e->npad));
  CKE(dalloc(&e->cnt, e->npad));
  CKE(dalloc(&e->lambda, (size_t)cfg->l * K * 2));
  CKE(dalloc(&e->ctl, 1));
  CKE(dalloc(&e->xchg_own, 1));
  e->xchg = e->xchg_own;
  e->xbuf = &e->xchg->x;
  CKE(cudaMemsetAsync(e->gamma, 0, K * e->npad * sizeof(double), e->stream));
  CKE(cudaMemsetAsync(e->E, 0, K * e->npad * sizeof(double), e->stream));
  CKE(cudaMemsetAsync(e->cnt, 0, e->npad * sizeof(uint32_t), e->stream));
  CKE(cudaMemsetAsync(e->ctl, 0, sizeof(Ctl), e->stream));
  CKE(cudaMemsetAsync(e->xchg, 0, sizeof(Xchg), e->stream));
  // Grid sizing: whole CTAs of individuals, capped at two CTAs per SM.
#ifdef TS_STAGED_PATH
  const int need = (int)((cfg->n_local + ESTEP_THREADS - 1) / ESTEP_THREADS);
  e->grid_estep = std::max(1, std::min(need, e->num_sms));
  e->grid_gamma = std::max(1, std::min(need, 4 * e->num_sms));
#endif
  {
    // Persistent kernel: one CTA per SM at most, threads sized so that every thread owns the
    // same number of individuals (I = ceil(n / (SMs * TMAX))).
    const char *path = getenv("TSGPU_PATH");
    e->staged = path && !strcmp(path, "staged");
#ifndef TS_STAGED_PATH
    if (e->staged) {
      set_err(TS_ERR_STATE, "TSGPU_PATH=staged needs the test-only library lib/libtsgpu_staged.so (make staged); the product library has no staged path");
      ts_destroy(e);
      return TS_ERR_STATE;
    }
#endif
    // test/developer knobs: TSGPU_IPT pins the individuals per thread in registers (0 = the TIER kernel);
    // TSGPU_TIER_J / TSGPU_TIER_GRID / TSGPU_TIER_BLOCK shape the TIER kernel so that the parity tests
    // reach its shared-memory and streaming tiers at sizes the oracle finishes in seconds
    auto env_int = [](const char *name, int dflt) { const char *v = getenv(name); return v ? atoi(v) : dflt; };
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const ShardPlan pl = plan_shard(cfg->n_local, e->K, e->num_sms, std::max(-1, env_int("TSGPU_IPT", -1)), env_int("TSGPU_TIER_J", -1),
                                    env_int("TSGPU_TIER_GRID", 0), env_int("TSGPU_TIER_BLOCK", 0),
                                    optin > 0 ? (size_t)optin : SMEM_OPTIN_B200);
    e->ind_per_thread = pl.ipt;
    e->grid_persist = pl.grid;
    e->block_persist = pl.block;
    e->tier = pl.tier;
    e->prm.tier_j = (uint32_t)pl.tier_j;
    e->n_stream = pl.n_stream;
    const char *tmo = getenv("TSGPU_TIMEOUT_S");
    e->prm.timeout_ns = (unsigned long long)(1e9 * (tmo ? std::max(0.001, atof(tmo)) : 60.0));
    const int sh = tsfx::shift_for(cfg->n_total);  // per-warp sums stay below 2^52 (mantissa-trick conversion)
    e->prm.fx_scale = ldexp(1.0, sh);
    e->prm.fx_inv = ldexp(1.0, -sh);
    e->prm.trace = nullptr;
    e->prm.xflush = getenv("TSGPU_XFLUSH") ? atoi(getenv("TSGPU_XFLUSH")) : 0;
    if (getenv("TSGPU_TRACE")) {
      CKE(dalloc(&e->prm.trace, 64 * 128));
      CKE(cudaMemset(e->prm.trace, 0, 64 * 128 * sizeof(long long)));
    }
  }
  CKE(dalloc(&e->partial, (size_t)e->grid_estep * 2 * K));
  e->items_cap = 1 << 16;
  CKE(dalloc(&e->items, e->items_cap));
  CKE(dalloc(&e->rounds, e->items_cap));
  CKE(dalloc(&e->ll, 1));
#undef CKE
  fill_params(e);
  for (int r = 0; r < MAXR; ++r) {
    e->prm.xpeer[r] = e->xbuf;
    e->prm.pst_peer[r] = &e->xchg->ps;
  }
  const size_t cnt = (size_t)cfg->l * K * 2;
  k_fill_lambda<<<std::min<size_t>((cnt + 255) / 256, 4096), 256, 0, e->stream>>>(e->lambda, cnt, cfg->eta0, cfg->eta1);
  e->launches++;
  CK(cudaStreamSynchronize(e->stream));
  *out = e;
  return TS_OK;
}

int ts_destroy(ts_engine *e) {
  if (!e) return TS_OK;
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void *p : e->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(e->bed);
  cudaFree(e->vcol);
  cudaFree(e->gamma);
  cudaFree(e->E);
  cudaFree(e->cnt);
  cudaFree(e->lambda);
  cudaFree(e->partial);
  cudaFree(e->ll);
  cudaFree(e->rounds);
  cudaFree(e->ctl);
  cudaFree(e->items);
  cudaFree(e->xchg_own);
  symm_free(e->symm);
  cudaFree(e->prm.trace);
  cudaFree(e->d_voff);
  cudaFree(e->d_vind);
  cudaFree(e->d_val_loc);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  (void)cudaGetLastError();  // freeing what a partly built engine never allocated leaves an error behind
  delete e;
  return TS_OK;
}

int ts_load_bed(ts_engine *e, uint64_t loc_begin, uint64_t nloc, const uint8_t *rows, uint64_t row_pitch) {
  if (!e || !rows) return set_err(TS_ERR_ARG, "ts_load_bed: null argument");
  if (loc_begin + nloc > e->cfg.l) return set_err(TS_ERR_ARG, "ts_load_bed: loci out of range");
  const size_t full_bytes = (e->cfg.n_total + 3) / 4;
  if (row_pitch < full_bytes) return set_err(TS_ERR_ARG, "ts_load_bed: row_pitch < ceil(N/4)");
  if (use_device(e)) return TS_ERR_CUDA;
  if (nloc == 0) return TS_OK;
  if (e->pitch != e->local_bytes)  // zero the pad bytes once so vector loads see defined data
    CK(cudaMemset2DAsync(e->bed + loc_begin * e->pitch, e->pitch, 0, e->pitch, nloc, e->stream));
  CK(cudaMemcpy2DAsync(e->bed + loc_begin * e->pitch, e->pitch, rows + e->cfg.n_begin / 4, row_pitch,
                       e->local_bytes, nloc, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->bed_loaded = true;
  return TS_OK;
}

// One pass over the rows for ALL engines of a process (the CLI's -gpus N): rows are staged through two
// pinned host buffers (the source is usually a memory-mapped .bed: copying it into the staging buffer
// is what reads the file, done by a few host threads), and every engine pulls its byte range of the
// staged chunk with an asynchronous 2-D copy while the next chunk is being read.
int ts_load_bed_fanout(ts_engine **engines, int n, uint64_t loc_begin, uint64_t nloc, const uint8_t *rows,
                       uint64_t row_pitch) {
  if (!engines || n < 1 || !rows) return set_err(TS_ERR_ARG, "ts_load_bed_fanout: null argument");
  const ts_engine *e0 = engines[0];
  const size_t full_bytes = (e0->cfg.n_total + 3) / 4;
  if (row_pitch < full_bytes) return set_err(TS_ERR_ARG, "ts_load_bed_fanout: row_pitch < ceil(N/4)");
  for (int i = 0; i < n; ++i)
    if (!engines[i] || engines[i]->cfg.n_total != e0->cfg.n_total || engines[i]->cfg.l != e0->cfg.l || loc_begin + nloc > engines[i]->cfg.l)
      return set_err(TS_ERR_ARG, "ts_load_bed_fanout: engines disagree on the data set shape, or loci out of range");
  if (nloc == 0) return TS_OK;
  const size_t spitch = (full_bytes + 15) / 16 * 16;
  const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(nloc, (128ull << 20) / spitch));
  uint8_t *stage[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> done[2];
  int rc = TS_OK;
  auto fail = [&](const char *what, cudaError_t er) { rc = set_err(TS_ERR_CUDA, "ts_load_bed_fanout: %s: %s", what, cudaGetErrorString(er)); };
  for (int b = 0; b < 2 && rc == TS_OK; ++b) {
    cudaError_t er = cudaHostAlloc((void **)&stage[b], chunk * spitch, cudaHostAllocPortable);
    if (er != cudaSuccess) { fail("cudaHostAlloc", er); break; }
    done[b].resize(n);
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(engines[i]->cfg.device);
      if ((er = cudaEventCreateWithFlags(&done[b][i], cudaEventDisableTiming)) != cudaSuccess) { fail("cudaEventCreate", er); break; }
    }
  }
  const unsigned nthreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  uint64_t c = 0;
  for (uint64_t lo = 0; lo < nloc && rc == TS_OK; lo += chunk, ++c) {
    const uint64_t m = std::min<uint64_t>(chunk, nloc - lo);
    const int b = (int)(c & 1);
    if (c >= 2)
      for (int i = 0; i < n; ++i) cudaEventSynchronize(done[b][i]);  // the copies that read this buffer two chunks ago
    {  // file -> pinned staging, rows split over host threads
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
          for (uint64_t r = t; r < m; r += nthreads) memcpy(stage[b] + r * spitch, rows + (lo + r) * row_pitch, full_bytes);
        });
      for (auto &x : th) x.join();
    }
    for (int i = 0; i < n && rc == TS_OK; ++i) {
      ts_engine *e = engines[i];
      cudaError_t er = cudaSetDevice(e->cfg.device);
      unsigned char *dst = e->bed + (loc_begin + lo) * e->pitch;
      if (er == cudaSuccess && e->pitch != e->local_bytes) er = cudaMemset2DAsync(dst, e->pitch, 0, e->pitch, m, e->stream);
      if (er == cudaSuccess)
        er = cudaMemcpy2DAsync(dst, e->pitch, stage[b] + e->cfg.n_begin / 4, spitch, e->local_bytes, m, cudaMemcpyHostToDevice, e->stream);
      if (er == cudaSuccess) er = cudaEventRecord(done[b][i], e->stream);
      if (er != cudaSuccess) fail("upload", er);
    }
  }
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(engines[i]->cfg.device);
    cudaError_t er = cudaStreamSynchronize(engines[i]->stream);
    if (er != cudaSuccess && rc == TS_OK) fail("cudaStreamSynchronize", er);
    if (rc == TS_OK && loc_begin + nloc > 0) engines[i]->bed_loaded = true;
  }
  for (int b = 0; b < 2; ++b) {
    for (size_t i = 0; i < done[b].size(); ++i) cudaEventDestroy(done[b][i]);
    if (stage[b]) cudaFreeHost(stage[b]);
  }
  return rc;
}

int ts_synth_bed(ts_engine *e, uint64_t seed, const float *theta, const float *beta, double missing_rate) {
  if (!e || !theta || !beta) return set_err(TS_ERR_ARG, "ts_synth_bed: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  float *dt = nullptr, *db = nullptr;
  const size_t nt = (size_t)e->cfg.n_local * e->K, nb = (size_t)e->cfg.l * e->K;
  CK(dalloc(&dt, nt));
  CK(dalloc(&db, nb));
  CK(cudaMemcpyAsync(dt, theta, nt * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(db, beta, nb * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemsetAsync(e->bed, 0, (size_t)e->cfg.l * e->pitch, e->stream));
  k_synth<<<e->num_sms * 16, 256, 0, e->stream>>>(e->bed, e->pitch, e->cfg.l, (uint32_t)e->cfg.n_local,
                                                 e->cfg.n_begin, e->K, dt, db, seed, (float)missing_rate);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(dt);
  cudaFree(db);
  e->bed_loaded = true;
  return TS_OK;
}

int ts_get_bed_row(ts_engine *e, uint64_t loc, uint8_t *out) {
  if (!e || !out || loc >= e->cfg.l) return set_err(TS_ERR_ARG, "ts_get_bed_row: bad argument");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaMemcpyAsync(out, e->bed + loc * e->pitch, e->local_bytes, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TS_OK;
}

int ts_set_validation(ts_engine *e, uint64_t nval, const uint32_t *val_loc, const uint64_t *val_off,
                      const uint32_t *val_indiv) {
  if (!e || (nval && (!val_loc || !val_off || !val_indiv)))
    return set_err(TS_ERR_ARG, "ts_set_validation: null argument");
  if (!e->bed_loaded) return set_err(TS_ERR_STATE, "ts_set_validation: genotypes not resident yet");
  if (use_device(e)) return TS_ERR_CUDA;
  e->val_loc.assign(val_loc, val_loc + nval);
  e->voff.assign(nval + 1, 0);
  e->vind.clear();
  const uint64_t lo = e->cfg.n_begin, hi = e->cfg.n_begin + e->cfg.n_local;
  for (uint64_t v = 0; v < nval; ++v) {
    if (val_loc[v] >= e->cfg.l || (v && val_loc[v] <= val_loc[v - 1]))
      return set_err(TS_ERR_ARG, "ts_set_validation: val_loc must be ascending and < L");
    e->voff[v] = e->vind.size();
    for (uint64_t i = val_off[v]; i < val_off[v + 1]; ++i) {
      if (val_indiv[i] >= e->cfg.n_total) return set_err(TS_ERR_ARG, "ts_set_validation: individual id >= N");
      if (val_indiv[i] >= lo && val_indiv[i] < hi) e->vind.push_back((uint32_t)(val_indiv[i] - lo));
    }
  }
  e->voff[nval] = e->vind.size();
  cudaFree(e->vcol); e->vcol = nullptr;
  cudaFree(e->d_voff); e->d_voff = nullptr;
  cudaFree(e->d_vind); e->d_vind = nullptr;
  cudaFree(e->d_val_loc); e->d_val_loc = nullptr;
  cudaFree(e->ll); e->ll = nullptr;
  CK(dalloc(&e->vcol, (size_t)nval * e->pitch));
  CK(dalloc(&e->d_voff, nval + 1));
  CK(dalloc(&e->d_vind, e->vind.size()));
  CK(dalloc(&e->d_val_loc, nval));
  CK(dalloc(&e->ll, nval));
  CK(cudaMemcpyAsync(e->d_voff, e->voff.data(), (nval + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
  if (!e->vind.empty())
    CK(cudaMemcpyAsync(e->d_vind, e->vind.data(), e->vind.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  if (nval) {
    CK(cudaMemcpyAsync(e->d_val_loc, e->val_loc.data(), nval * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    k_build_vcol<<<(unsigned)nval, 256, 0, e->stream>>>(e->bed, e->pitch, e->d_val_loc, e->d_voff, e->d_vind, e->vcol);
    e->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(e->stream));
  fill_params(e);
  return TS_OK;
}

int ts_set_gamma(ts_engine *e, const double *rows) {
  if (!e || !rows) return set_err(TS_ERR_ARG, "ts_set_gamma: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  const size_t K = e->K, n = e->cfg.n_local;
  std::vector<double> soa(K * e->npad, 1.0);
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < K; ++k) {
      const double g = rows[i * K + k];
      if (!(g > 0.0)) return set_err(TS_ERR_ARG, "ts_set_gamma: gamma[%zu][%zu] = %g is not positive", i, k, g);
      soa[k * e->npad + i] = g;
    }
  CK(cudaMemcpyAsync(e->gamma, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  k_refresh_E<<<e->num_sms * 4, 256, 0, e->stream>>>(e->gamma, e->E, e->npad, (uint32_t)n, e->K);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  e->gamma_set = true;
  return TS_OK;
}

int ts_reset_lambda(ts_engine *e) {
  if (!e) return set_err(TS_ERR_ARG, "ts_reset_lambda: null engine");
  if (use_device(e)) return TS_ERR_CUDA;
  const size_t cnt = (size_t)e->cfg.l * e->K * 2;
  k_fill_lambda<<<std::min<size_t>((cnt + 255) / 256, 4096), 256, 0, e->stream>>>(e->lambda, cnt, e->cfg.eta0, e->cfg.eta1);
  e->launches++;
  CK(cudaGetLastError());
  return TS_OK;
}

int ts_reset_counts(ts_engine *e) {
  if (!e) return set_err(TS_ERR_ARG, "ts_reset_counts: null engine");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaMemsetAsync(e->cnt, 0, e->npad * sizeof(uint32_t), e->stream));
  return TS_OK;
}

// Enqueue a batch of work items (<= items_cap) and the kernels that process them.
static int check_fault(ts_engine *e) {
  uint32_t fault[2] = {0, 0};
  CK(cudaMemcpy(&fault[0], &e->ctl->fault, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&fault[1], &e->xchg->ps.fault, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (fault[0] || fault[1])
    return set_err(TS_ERR_CUDA, "grid/peer barrier timed out (a CTA or a rank stopped responding)");
  return TS_OK;
}

// Enqueue a batch of work items (<= items_cap) and the kernels that process them.
static int run_items(ts_engine *e, const std::vector<WorkItem> &items, uint32_t *rounds_out) {
  const size_t n = items.size();
  if (n == 0) return TS_OK;
  // The previous batch may still be reading items[]: drain before overwriting.
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpyAsync(e->items, items.data(), n * sizeof(WorkItem), cudaMemcpyHostToDevice, e->stream));
  const bool first = (items[0].flags & ITEM_FIRST) != 0;  // a batch is all-first or none
  const bool hol = (items[0].flags & ITEM_HOL) != 0;      // and all-hol or none
  if (first) {
    CK(cudaMemsetAsync(e->rounds, 0, n * sizeof(uint32_t), e->stream));
  } else if (!e->staged) {
    CK(launch_persist(e, (uint32_t)n));
    e->launches++;
  }
#ifdef TS_STAGED_PATH
  else {
    const long long minus1 = -1;
    CK(cudaMemcpyAsync(&e->ctl->cursor, &minus1, sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    for (size_t i = 0; i < n; ++i) {
      k_begin<<<1, 32, 0, e->stream>>>(e->prm, e->K);
      e->launches++;
      for (uint32_t x = 0; x < e->cfg.online_iterations; ++x) {
        launch_estep(e);
        e->launches++;
      }
      if (!hol) {
        launch_gamma(e);
        e->launches++;
      }
    }
  }
#endif
  if (hol) {
    launch_heldout(e, (unsigned)n);
    e->launches++;
  }
  CK(cudaGetLastError());
  if (rounds_out) {
    CK(cudaMemcpyAsync(rounds_out, e->rounds, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return check_fault(e);
  }
  return TS_OK;
}

static int find_vslot(const ts_engine *e, uint32_t loc) {
  auto it = std::lower_bound(e->val_loc.begin(), e->val_loc.end(), loc);
  if (it != e->val_loc.end() && *it == loc) return (int)(it - e->val_loc.begin());
  return -1;
}

static WorkItem make_item(const ts_engine *e, uint32_t loc, uint32_t flags) {
  WorkItem w;
  w.loc = loc;
  w.vslot = find_vslot(e, loc);
  w.flags = flags;
  w.pad = 0;
  w.col = (w.vslot >= 0) ? e->vcol + (size_t)w.vslot * e->pitch : e->bed + (size_t)loc * e->pitch;
  return w;
}

int ts_steps(ts_engine *e, const uint32_t *locs, uint64_t n, int hol_mode, uint32_t *rounds_out) {
  if (!e || (n && !locs)) return set_err(TS_ERR_ARG, "ts_steps: null argument");
  if (!e->bed_loaded || !e->gamma_set) return set_err(TS_ERR_STATE, "ts_steps: genotypes and gamma must be set first");
  if (use_device(e)) return TS_ERR_CUDA;
  for (uint64_t i = 0; i < n; ++i)
    if (locs[i] >= e->cfg.l) return set_err(TS_ERR_ARG, "ts_steps: locus %u >= L", locs[i]);
  for (uint64_t off = 0; off < n; off += e->items_cap) {
    const uint64_t m = std::min<uint64_t>(e->items_cap, n - off);
    e->h_items.resize(m);
    for (uint64_t i = 0; i < m; ++i) e->h_items[i] = make_item(e, locs[off + i], hol_mode ? ITEM_HOL : 0u);
    int rc = run_items(e, e->h_items, rounds_out ? rounds_out + off : nullptr);
    if (rc) return rc;
  }
  return TS_OK;
}

int ts_step(ts_engine *e, uint32_t loc, int hol_mode, int *rounds_out) {
  uint32_t r = 0;
  int rc = ts_steps(e, &loc, 1, hol_mode, rounds_out ? &r : nullptr);
  if (rc == TS_OK && rounds_out) *rounds_out = (int)r;
  return rc;
}

int ts_heldout_ll(ts_engine *e, int first, double *sum, uint64_t *count, double *per_locus_sum) {
  if (!e || !sum || !count) return set_err(TS_ERR_ARG, "ts_heldout_ll: null argument");
  if (!e->bed_loaded || !e->gamma_set) return set_err(TS_ERR_STATE, "ts_heldout_ll: genotypes and gamma must be set first");
  if (use_device(e)) return TS_ERR_CUDA;
  const size_t nval = e->val_loc.size();
  std::vector<double> ll(nval, 0.0);
  for (size_t off = 0; off < nval; off += e->items_cap) {
    const size_t m = std::min(e->items_cap, nval - off);
    e->h_items.resize(m);
    for (size_t i = 0; i < m; ++i)
      e->h_items[i] = make_item(e, e->val_loc[off + i], ITEM_HOL | (first ? ITEM_FIRST : 0u));
    int rc = run_items(e, e->h_items, nullptr);
    if (rc) return rc;
  }
  if (nval) CK(cudaMemcpyAsync(ll.data(), e->ll, nval * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (int rc = check_fault(e)) return rc;
  double s = 0.0;
  for (size_t v = 0; v < nval; ++v) s += ll[v];
  if (per_locus_sum) memcpy(per_locus_sum, ll.data(), nval * sizeof(double));
  *sum = s;
  *count = e->vind.size();
  return TS_OK;
}

static int get_rows(ts_engine *e, int mode, double *out) {
  if (!e || !out) return set_err(TS_ERR_ARG, "ts_get_*: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  const size_t n = e->cfg.n_local, K = e->K;
  if (mode == 2) {  // raw gamma: transpose on the host
    std::vector<double> soa(K * e->npad);
    CK(cudaMemcpyAsync(soa.data(), e->gamma, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < n; ++i)
      for (size_t k = 0; k < K; ++k) out[i * K + k] = soa[k * e->npad + i];
    return TS_OK;
  }
  double *d = nullptr;
  CK(dalloc(&d, n * K));
  k_theta<<<e->num_sms * 2, 256, 0, e->stream>>>(e->gamma, e->npad, (uint32_t)n, e->K, mode, d);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d, n * K * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(d);
  return TS_OK;
}

int ts_get_gamma(ts_engine *e, double *out) { return get_rows(e, 2, out); }
int ts_get_theta(ts_engine *e, double *out) { return get_rows(e, 0, out); }
int ts_get_elogtheta(ts_engine *e, double *out) { return get_rows(e, 1, out); }

int ts_get_counts(ts_engine *e, uint32_t *out) {
  if (!e || !out) return set_err(TS_ERR_ARG, "ts_get_counts: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaMemcpyAsync(out, e->cnt, e->cfg.n_local * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TS_OK;
}

int ts_get_lambda(ts_engine *e, uint64_t loc_begin, uint64_t nloc, double *out) {
  if (!e || !out || loc_begin + nloc > e->cfg.l) return set_err(TS_ERR_ARG, "ts_get_lambda: bad argument");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaMemcpyAsync(out, e->lambda + loc_begin * e->K * 2, nloc * e->K * 2 * sizeof(double),
                     cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TS_OK;
}

int ts_get_beta(ts_engine *e, uint64_t loc_begin, uint64_t nloc, double *out) {
  if (!e || !out || loc_begin + nloc > e->cfg.l) return set_err(TS_ERR_ARG, "ts_get_beta: bad argument");
  if (use_device(e)) return TS_ERR_CUDA;
  if (nloc == 0) return TS_OK;
  double *d = nullptr;
  const size_t cnt = nloc * e->K;
  CK(dalloc(&d, cnt));
  k_beta<<<std::min<size_t>((cnt + 255) / 256, 4096), 256, 0, e->stream>>>(e->lambda + loc_begin * e->K * 2, cnt, d);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d, cnt * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(d);
  return TS_OK;
}

int ts_sync(ts_engine *e) {
  if (!e) return set_err(TS_ERR_ARG, "ts_sync: null engine");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaStreamSynchronize(e->stream));
  return check_fault(e);  // asynchronous ts_steps batches report a lost CTA / rank here
}

// ---- exchange -------------------------------------------------------------------------------
int ts_comm_export(ts_engine *e, void *handle_out) {
  if (!e || !handle_out) return set_err(TS_ERR_ARG, "ts_comm_export: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  static_assert(sizeof(cudaIpcMemHandle_t) == TS_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (e->xchg != e->xchg_own) return set_err(TS_ERR_STATE, "ts_comm_export: a symmetric buffer is already attached");
  CK(cudaIpcGetMemHandle(&h, e->xchg_own));
  memcpy(handle_out, &h, sizeof h);
  return TS_OK;
}

int ts_comm_connect(ts_engine *e, const void *all_handles) {
  if (!e || !all_handles) return set_err(TS_ERR_ARG, "ts_comm_connect: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaStreamSynchronize(e->stream));
  for (int r = 0; r < e->cfg.nranks; ++r) {
    if (r == e->cfg.rank) {
      e->prm.xpeer[r] = e->xbuf;
      e->prm.pst_peer[r] = &e->xchg->ps;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_handles + (size_t)r * TS_COMM_HANDLE_BYTES, sizeof h);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->ipc_opened.push_back(p);
    e->prm.xpeer[r] = &((Xchg *)p)->x;
    e->prm.pst_peer[r] = &((Xchg *)p)->ps;
  }
  e->xmode = choose_xmode(false);
  fill_params(e);
  return TS_OK;
}

int ts_comm_connect_local(ts_engine **engines, int n) {
  if (!engines || n < 1 || n > MAXR) return set_err(TS_ERR_ARG, "ts_comm_connect_local: bad argument");
  for (int i = 0; i < n; ++i) {
    ts_engine *e = engines[i];
    if (!e || e->cfg.nranks != n || e->cfg.rank != i)
      return set_err(TS_ERR_ARG, "ts_comm_connect_local: engine %d is not rank %d of %d", i, i, n);
    CK(cudaSetDevice(e->cfg.device));
    for (int j = 0; j < n; ++j) {
      if (j != i && engines[j]->cfg.device != e->cfg.device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, e->cfg.device, engines[j]->cfg.device));
        if (!can) return set_err(TS_ERR_CUDA, "device %d cannot access device %d", e->cfg.device, engines[j]->cfg.device);
        cudaError_t er = cudaDeviceEnablePeerAccess(engines[j]->cfg.device, 0);
        if (er != cudaSuccess && er != cudaErrorPeerAccessAlreadyEnabled)
          return set_err(TS_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(er));
        cudaGetLastError();
      }
      e->prm.xpeer[j] = engines[j]->xbuf;
      e->prm.pst_peer[j] = &engines[j]->xchg->ps;
    }
    e->xmode = choose_xmode(false);
    fill_params(e);
  }
  // Distinct devices: move the exchange state into symmetric buffers with an NVLS multicast alias when the
  // fabric offers one (one multimem.red per word instead of one red.add per peer); TSGPU_XCHG=gacc|slots|ipc
  // keep the plain peer mappings set up above.
  bool distinct = n > 1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (engines[i]->cfg.device == engines[j]->cfg.device) distinct = false;
  const char *xm = getenv("TSGPU_XCHG");
  if (distinct && !(xm && (!strcmp(xm, "gacc") || !strcmp(xm, "slots") || !strcmp(xm, "ipc"))) && !engines[0]->symm) {
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) devs[i] = engines[i]->cfg.device;
    std::string err;
    SymmGroup *g = symm_alloc_local(devs.data(), n, sizeof(Xchg), true, &err);
    if (g && g->has_multicast()) {
      unsigned long long ctas = 0;
      for (int i = 0; i < n; ++i) ctas += (unsigned long long)engines[i]->grid_persist;
      for (int i = 0; i < n; ++i) {
        int rc = attach_symmetric(engines[i], g->uc.data(), g->mc[i], ctas);
        if (rc) { symm_free(g); return rc; }
      }
      engines[0]->symm = g;
    } else {
      symm_free(g);  // no NVLS here: the peer-store exchange set up above stays
    }
  }
  return TS_OK;
}

uint64_t ts_comm_state_bytes(void) { return sizeof(Xchg); }
int ts_comm_mode(const ts_engine *e) { return e ? e->xmode : -1; }

int ts_comm_attach_symmetric(ts_engine *e, const void *const *rank_ptrs, void *multicast_ptr, uint64_t bytes, uint32_t total_ctas) {
  if (!e || !rank_ptrs) return set_err(TS_ERR_ARG, "ts_comm_attach_symmetric: null argument");
  if (bytes < sizeof(Xchg)) return set_err(TS_ERR_ARG, "ts_comm_attach_symmetric: %llu bytes, need %zu", (unsigned long long)bytes, sizeof(Xchg));
  for (int r = 0; r < e->cfg.nranks; ++r)
    if (!rank_ptrs[r] || ((uintptr_t)rank_ptrs[r] & 15)) return set_err(TS_ERR_ARG, "ts_comm_attach_symmetric: rank %d pointer null or not 16-byte aligned", r);
  if ((uintptr_t)multicast_ptr & 15) return set_err(TS_ERR_ARG, "ts_comm_attach_symmetric: multicast pointer not 16-byte aligned");
  return attach_symmetric(e, (void *const *)rank_ptrs, multicast_ptr,
                          total_ctas ? total_ctas : (unsigned long long)e->grid_persist * e->cfg.nranks);
}

uint64_t ts_launch_count(const ts_engine *e) { return e ? e->launches : 0; }

int ts_get_plan(const ts_engine *e, int *ind_per_thread, int *grid, int *block) {
  if (!e || !ind_per_thread || !grid || !block) return set_err(TS_ERR_ARG, "ts_get_plan: null argument");
  *ind_per_thread = e->ind_per_thread;
  *grid = e->grid_persist;
  *block = e->block_persist;
  return TS_OK;
}

int ts_get_tiers(const ts_engine *e, int *smem_per_thread, uint64_t *n_streamed) {
  if (!e || !smem_per_thread || !n_streamed) return set_err(TS_ERR_ARG, "ts_get_tiers: null argument");
  *smem_per_thread = e->tier ? (int)e->prm.tier_j : -1;
  *n_streamed = e->n_stream;
  return TS_OK;
}

int ts_debug_trace(ts_engine *e, long long *out /* 64 x 128 */) {
  if (!e || !out || !e->prm.trace) return set_err(TS_ERR_STATE, "ts_debug_trace: tracing is off (TSGPU_TRACE=1)");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(out, e->prm.trace, 64 * 128 * sizeof(long long), cudaMemcpyDeviceToHost));
  return TS_OK;
}

int ts_timer_start(ts_engine *e) {
  if (!e) return set_err(TS_ERR_ARG, "ts_timer_start: null engine");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaEventRecord(e->ev0, e->stream));
  return TS_OK;
}

int ts_timer_stop(ts_engine *e, float *ms_out) {
  if (!e || !ms_out) return set_err(TS_ERR_ARG, "ts_timer_stop: null argument");
  if (use_device(e)) return TS_ERR_CUDA;
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaEventSynchronize(e->ev1));
  CK(cudaEventElapsedTime(ms_out, e->ev0, e->ev1));
  return TS_OK;
}

}  // extern "C"
