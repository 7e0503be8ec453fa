// TEST INFRASTRUCTURE ONLY (never linked into the product): a sequential host model of the
// persistent kernel's ALGORITHM -- the reformulation that the CUDA code implements -- built from
// the kernel's own host-compilable arithmetic (ts_expsi.cuh: f = exp o digamma, fast_rcp; ts_ftab.cuh: the control path's f, 1/f;
// ts_fixed.cuh: fixed-point words), so that the reformulation can be checked against the oracle
// without a GPU:
//   E = f(gamma) instead of Elogtheta; b = f(lambda_t) / f(lambda_0 + lambda_1) instead of Elogbeta;
//   s_t = sum_k E_k b_kt;  S_t[k] = b_kt * sum_n E_nk w_tn / s_tn  (w_0 = y, w_1 = 2 - y);
//   per-"warp" (32 * ipt individuals) double sums -> fixed-point words -> integer totals -> double;
//   lambda = eta + S; convergence on mean |delta lambda|; gamma natural-gradient step with the phi of
//   the last E-step; E refreshed for the updated individuals.
// Follows ts_persist.cuh (k_persist) step by step; the reference it must agree with is
// snpsamplinge.cc:320-366, :695-740 (through oracle/ts_oracle.c).
#include <cmath>
#include <cstdint>
#include <vector>

#define TS_FTAB_HOST_TABLE
#include "ts_expsi.cuh"
#include "ts_fixed.cuh"
#include "ts_ftab.cuh"

extern "C" int km_train(uint32_t n, uint32_t l, uint32_t k, const uint8_t *y /* [l][n]: 0,1,2; 3 = missing or held out */,
                        double *gamma /* [n][k] */, uint32_t *cnt /* [n] */, double *lambda /* [l][k][2] */,
                        const uint32_t *locs, uint32_t nlocs, uint32_t max_rounds, double thresh, double alpha,
                        double eta, double tau0, int ipt, uint32_t *rounds_out) {
  const uint32_t V = 2 * k;
  const int sh = tsfx::shift_for(n);
  const double fx_scale = std::ldexp(1.0, sh);
  const tsfx::Unscale fxu = tsfx::unscale(std::ldexp(1.0, -sh));
  std::vector<double> E((size_t)n * k), b(V), bn(V), lam(V), q0(n), q1(n), vv(V);
  for (size_t i = 0; i < (size_t)n * k; ++i) E[i] = tsp::f_expsi(gamma[i]);
  const uint32_t warp_span = 32u * (uint32_t)(ipt > 0 ? ipt : 1);
  // estimate_beta as the control warp does it (ts_ftab.cuh): the table-driven f and 1/f while every statistic
  // of the row (and every pair sum) lies in the table's domain, f_expsi / fast_rcp for the whole row otherwise
  auto b_from = [&](const double *row, double *dst) {
    bool all_in = true;
    for (uint32_t v = 0; v < V; ++v) all_in = all_in && tsp::ftab_covers(row[v], row[v & ~1u] + row[v | 1u]);
    for (uint32_t v = 0; v < V; ++v) dst[v] = tsp::beta_ratio(tsp::h_ftab, all_in, row[v], row[v & ~1u] + row[v | 1u]);
  };
  for (uint32_t it = 0; it < nlocs; ++it) {
    const uint32_t loc = locs[it];
    const uint8_t *col = y + (size_t)loc * n;
    double *row = lambda + (size_t)loc * V;
    for (uint32_t v = 0; v < V; ++v) lam[v] = row[v];
    b_from(lam.data(), b.data());
    uint32_t x = 0;
    for (;;) {
      long long hi_tot[64] = {0}, lo_tot[64] = {0};
      for (uint32_t w0 = 0; w0 < n; w0 += warp_span) {  // one "warp": double sums, then fixed point
        for (uint32_t v = 0; v < V; ++v) vv[v] = 0.0;
        for (uint32_t i = w0; i < n && i < w0 + warp_span; ++i) {
          const int c = col[i];
          const double wt0 = c == 3 ? 0.0 : (double)c, wt1 = c == 3 ? 0.0 : (double)(2 - c);
          double s0 = 0.0, s1 = 0.0;
          for (uint32_t kk = 0; kk < k; ++kk) {
            s0 = std::fma(E[(size_t)i * k + kk], b[2 * kk], s0);
            s1 = std::fma(E[(size_t)i * k + kk], b[2 * kk + 1], s1);
          }
          q0[i] = wt0 * tsp::fast_rcp1(s0);  // one Newton step, as in the kernel's E-step
          q1[i] = wt1 * tsp::fast_rcp1(s1);
          for (uint32_t kk = 0; kk < k; ++kk) {
            vv[2 * kk] = std::fma(E[(size_t)i * k + kk], q0[i], vv[2 * kk]);
            vv[2 * kk + 1] = std::fma(E[(size_t)i * k + kk], q1[i], vv[2 * kk + 1]);
          }
        }
        for (uint32_t v = 0; v < V; ++v) {
          long long h, lo;
          tsfx::split((b[v] * vv[v]) * fx_scale, h, lo);
          hi_tot[v] += h;
          lo_tot[v] += lo;
        }
      }
      double chg = 0.0;
      for (uint32_t v = 0; v < V; ++v) {
        tsfx::normalize(hi_tot[v], lo_tot[v]);
        const double own = tsfx::to_double_plus((unsigned long long)hi_tot[v], (unsigned long long)lo_tot[v], fxu, eta);  // as the kernel
        chg += std::fabs(own - lam[v]);
        lam[v] = own;
      }
      ++x;
      const bool done = (chg / (double)V < thresh) || x >= max_rounds;
      if (done) break;
      b_from(lam.data(), bn.data());
      b.swap(bn);
    }
    for (uint32_t v = 0; v < V; ++v) row[v] = lam[v];
    if (rounds_out) rounds_out[it] = x;
    // gamma step with the phi of the last E-step (b still holds the b that E-step used)
    for (uint32_t i = 0; i < n; ++i) {
      if (col[i] == 3) continue;
      const double rho = 1.0 / std::sqrt(tau0 + (double)cnt[i]);
      cnt[i]++;
      for (uint32_t kk = 0; kk < k; ++kk) {
        const double w = E[(size_t)i * k + kk] * std::fma(b[2 * kk], q0[i], b[2 * kk + 1] * q1[i]);
        const double g = gamma[(size_t)i * k + kk];
        const double gn = g + rho * (alpha + (double)l * w - g);
        gamma[(size_t)i * k + kk] = gn;
        E[(size_t)i * k + kk] = tsp::f_expsi(gn);
      }
    }
  }
  return 0;
}
