// Host check of the lane -> statistic mapping (tsfx::tr_slot, ts_fixed.cuh) against a lock-step
// emulation of the transposed warp reduction of ts_persist.cuh (tr_reduce / tr_level): after the five
// levels, lane L must hold, in its first `len` registers, the sums over all 32 lanes of statistics
// start .. start+len-1, and every statistic must be held by exactly one lane.  V = 2K, K = 1..32.
#include <cstdio>
#include <vector>

#include "ts_fixed.cuh"

static bool fail = false;

// the network as the kernel runs it: at a level with n live values a lane whose `bit` is clear keeps
// values [0, ceil(n/2)) and receives its partner's copies of them; the other lane keeps
// [ceil(n/2), n) (zero-padded to ceil(n/2)) and receives the partner's copies of those
static void emulate(std::vector<std::vector<double>> &v, int V) {
  int n = V;
  for (int bit = 16; bit > 0; bit >>= 1) {
    const int LO = (n + 1) / 2, HI = n / 2;
    std::vector<std::vector<double>> nv(32, std::vector<double>(V, 0.0));
    for (int lane = 0; lane < 32; ++lane) {
      const bool up = lane & bit;
      const int partner = lane ^ bit;
      const bool pup = partner & bit;
      for (int i = 0; i < LO; ++i) {
        const double lo = v[lane][i], hi = i < HI ? v[lane][LO + i] : 0.0;
        const double plo = v[partner][i], phi = i < HI ? v[partner][LO + i] : 0.0;
        const double recv = pup ? plo : phi;  // what the partner sends: the half it does NOT keep
        nv[lane][i] = (up ? hi : lo) + recv;
      }
    }
    v = nv;
    n = LO;
  }
}

template <int V>
static void check() {
  std::vector<std::vector<double>> v(32, std::vector<double>(V));
  std::vector<double> want(V, 0.0);
  for (int lane = 0; lane < 32; ++lane)
    for (int i = 0; i < V; ++i) {
      v[lane][i] = (double)((lane * 131 + i * 17 + V) % 1009 + 1);
      want[i] += v[lane][i];
    }
  emulate(v, V);
  std::vector<int> owners(V, 0);
  for (int lane = 0; lane < 32; ++lane) {
    int start, len;
    tsfx::tr_slot<V>(lane, start, len);
    // lanes with len == 0 hold nothing
    if (len < 0 || start < 0 || (len > 0 && start + len > V)) { printf("V=%d lane %d: slot [%d,+%d) out of range\n", V, lane, start, len); fail = true; continue; }
    for (int q = 0; q < len; ++q) {
      owners[start + q]++;
      if (v[lane][q] != want[start + q]) { printf("V=%d lane %d q=%d: %g != %g\n", V, lane, q, v[lane][q], want[start + q]); fail = true; }
    }
  }
  for (int i = 0; i < V; ++i)
    if (owners[i] != 1) { printf("V=%d: statistic %d held by %d lanes\n", V, i, owners[i]); fail = true; }
  if constexpr (V > 2) check<V - 2>();
}

int main() {
  check<64>();
  if (!fail) printf("ok\n");
  return fail;
}
