// Host-side, RNG-exact initialisation for the TeraStructure hot path (part of libtsgpu.so).
//
// The SNP-sampling RNG stays on the host and has to reproduce the reference's GSL stream bit
// for bit: gsl_rng_default (MT19937) allocated at snpsamplinge.cc:59-63, consumed by
// set_validation_sample (cc:196-224), init_gamma (cc:226-237) and the infer loop (cc:423).
// GSL is not available in this image, so its published algorithms are implemented here:
// MT19937 with Knuth seeding (seed 0 -> 4357), the scale/reject uniform_int, Marsaglia-Tsang
// gamma on Voss' 128-level ziggurat (tables: tools/gen_zig_tables.py).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "tsgpu.h"

namespace {
#include "zig_tables.inc"
}

struct ts_rng {
  uint32_t mt[624];
  int idx;

  void seed(uint32_t s) {
    if (s == 0) s = 4357u;
    mt[0] = s;
    for (uint32_t i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i;
    idx = 624;
  }
  void refill() {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
      mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    idx = 0;
  }
  uint32_t get() {
    if (idx >= 624) refill();
    uint32_t k = mt[idx++];
    k ^= k >> 11;
    k ^= (k << 7) & 0x9d2c5680u;
    k ^= (k << 15) & 0xefc60000u;
    k ^= k >> 18;
    return k;
  }
  double uniform() { return get() / 4294967296.0; }
  double uniform_pos() {
    double x;
    do x = uniform(); while (x == 0.0);
    return x;
  }
  uint32_t uniform_int(uint32_t n) {
    const uint32_t scale = 0xffffffffu / n;
    uint32_t k;
    do k = get() / scale; while (k >= n);
    return k;
  }
  double gauss_zig() {
    for (;;) {
      const uint32_t k = get();
      uint32_t i = k & 0xFF;
      const uint32_t j = (k >> 8) & 0xFFFFFF;
      const double sign = (i & 0x80) ? 1.0 : -1.0;
      i &= 0x7f;
      double x = j * tszig_wtab[i];
      if (j < tszig_ktab[i]) return sign * x;
      double y;
      if (i < 127) {
        const double y0 = tszig_ytab[i], y1 = tszig_ytab[i + 1];
        y = y1 + (y0 - y1) * uniform();
      } else {
        const double u1 = 1.0 - uniform();
        const double u2 = uniform();
        x = TSZIG_PARAM_R - std::log(u1) / TSZIG_PARAM_R;
        y = std::exp(-TSZIG_PARAM_R * (x - 0.5 * TSZIG_PARAM_R)) * u2;
      }
      if (y < std::exp(-0.5 * x * x)) return sign * x;
    }
  }
  double gamma(double a, double b) {
    if (a < 1.0) {
      const double u = uniform_pos();
      return gamma(1.0 + a, b) * std::pow(u, 1.0 / a);
    }
    const double d = a - 1.0 / 3.0;
    const double c = (1.0 / 3.0) / std::sqrt(d);
    double x, v;
    for (;;) {
      do {
        x = gauss_zig();
        v = 1.0 + c * x;
      } while (v <= 0.0);
      v = v * v * v;
      const double u = uniform_pos();
      if (u < 1.0 - 0.0331 * x * x * x * x) break;
      if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) break;
    }
    return b * d * v;
  }
};

extern "C" {

ts_rng *ts_rng_create(double seed) {
  ts_rng *r = new ts_rng;
  r->seed(0);  // gsl_rng_alloc seeds with gsl_rng_default_seed = 0
  if (seed != 0.0) r->seed((uint32_t)(unsigned long)seed);
  return r;
}
void ts_rng_destroy(ts_rng *r) { delete r; }
uint32_t ts_rng_get(ts_rng *r) { return r->get(); }
uint32_t ts_rng_uniform_int(ts_rng *r, uint32_t n) { return r->uniform_int(n); }
double ts_rng_gamma(ts_rng *r, double a, double b) { return r->gamma(a, b); }
void ts_rng_sample_locs(ts_rng *r, uint32_t l, uint32_t *out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) out[i] = r->uniform_int(l);
}

// init_gamma (snpsamplinge.cc:226-237)
void ts_init_gamma(ts_rng *r, uint64_t n, uint32_t k, double *gamma_out) {
  const double v = (k < 100) ? 1.0 : 100.0 / k;
  for (uint64_t i = 0; i < n * k; ++i) gamma_out[i] = r->gamma(100 * v, 0.01);
}

// set_validation_sample (snpsamplinge.cc:196-224).  The reference keeps a
// std::map<pair<indiv,loc>,bool>; membership is all that matters, so each drawn locus gets a
// bitmap over individuals and the result is emitted as CSR in ascending locus order.
int ts_sample_validation(ts_rng *r, uint64_t n, uint64_t l, const uint8_t *bed, uint64_t row_pitch,
                         uint64_t *nval_out, uint32_t **val_loc_out, uint64_t **val_off_out,
                         uint32_t **val_indiv_out) {
  if (!r || n == 0 || l == 0 || n > 0xffffffffull || l > 0xffffffffull) return TS_ERR_ARG;
  const uint32_t per_loc_h = (uint32_t)(n < 2000 ? n / 10 : n / 100);
  const uint64_t nlocs = (uint64_t)(l * 0.005);
  // One N-bit bitmap, reused locus by locus (the reference draws a locus, then all of its held-out
  // individuals, before the next locus); every locus holds exactly per_loc_h individuals, so the
  // ascending lists go straight into the output array in DRAW order and the blocks are then
  // permuted in place into ascending-locus order.  Peak memory = the output itself
  // (200 MB at 1M x 1M) + N/8 bytes.
  const size_t h = per_loc_h;
  const size_t cap = (size_t)(nlocs ? nlocs : 1);
  std::vector<uint8_t> taken(l, 0);
  std::vector<uint32_t> drawn;
  drawn.reserve(cap);
  const size_t words = (n + 63) / 64;
  std::vector<uint64_t> m(words, 0ull);
  uint32_t *vi = (uint32_t *)malloc(sizeof(uint32_t) * (cap * h + 1));
  if (!vi) return TS_ERR_ARG;
  do {
    const uint32_t loc = r->uniform_int((uint32_t)l);
    if (taken[loc]) continue;
    taken[loc] = 1;
    const uint8_t *row = bed ? bed + (size_t)loc * row_pitch : nullptr;
    uint32_t c = 0;
    while (c < per_loc_h) {
      const uint32_t indiv = r->uniform_int((uint32_t)n);
      const bool held = (m[indiv >> 6] >> (indiv & 63)) & 1ull;
      const bool missing = row && ((row[indiv >> 2] >> (2 * (indiv & 3))) & 3) == 1;
      if (!held && !missing) {  // kv_ok (snpsamplinge.hh:389-408)
        m[indiv >> 6] |= 1ull << (indiv & 63);
        c++;
      }
    }
    uint32_t *blk = vi + drawn.size() * h;
    size_t q = 0;
    for (size_t w = 0; w < words && q < h; ++w) {
      uint64_t bits = m[w];
      m[w] = 0;
      while (bits) {
        blk[q++] = (uint32_t)(w * 64 + __builtin_ctzll(bits));
        bits &= bits - 1;
      }
    }
    drawn.push_back(loc);
  } while (drawn.size() < nlocs);

  const size_t nv = drawn.size();
  std::vector<size_t> order(nv);
  for (size_t i = 0; i < nv; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return drawn[a] < drawn[b]; });
  uint32_t *vl = (uint32_t *)malloc(sizeof(uint32_t) * (nv ? nv : 1));
  uint64_t *vo = (uint64_t *)malloc(sizeof(uint64_t) * (nv + 1));
  if (!vl || !vo) { free(vl); free(vo); free(vi); return TS_ERR_ARG; }
  for (size_t i = 0; i < nv; ++i) {
    vl[i] = drawn[order[i]];
    vo[i] = i * h;
  }
  // gather the blocks in place: block i of the result is block order[i] of the draw order
  {
    std::vector<uint32_t> tmp(h ? h : 1);
    std::vector<uint8_t> done(nv, 0);
    for (size_t s0 = 0; s0 < nv && h; ++s0) {
      if (done[s0]) continue;
      if (order[s0] == s0) { done[s0] = 1; continue; }
      memcpy(tmp.data(), vi + s0 * h, h * sizeof(uint32_t));
      size_t j = s0;
      for (;;) {
        const size_t src = order[j];
        done[j] = 1;
        if (src == s0) { memcpy(vi + j * h, tmp.data(), h * sizeof(uint32_t)); break; }
        memcpy(vi + j * h, vi + src * h, h * sizeof(uint32_t));
        j = src;
      }
    }
  }
  const size_t p = nv * h;
  vo[nv] = p;
  *nval_out = nv;
  *val_loc_out = vl;
  *val_off_out = vo;
  *val_indiv_out = vi;
  return TS_OK;
}

void ts_free(void *p) { free(p); }

}  // extern "C"
