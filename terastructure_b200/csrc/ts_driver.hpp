// Host front end of the drop-in `terastructure` CLI: the reference's Env (src/env.hh),
// SNP::read (src/snp.cc) and SNPSamplingE driver (src/snpsamplinge.cc) re-expressed on top of
// the C ABI (include/tsgpu.h).  Same flags, same output directory and files; the per-SNP work
// runs on the GPU(s).  Header-only, used by main.cpp.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <array>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "tsgpu.h"

namespace tsd {

// ---------------------------------------------------------------------------------------------
// Env: options, run directory, param.txt, infer.log (src/env.hh:50-313, src/log.cc)
// ---------------------------------------------------------------------------------------------
struct Env {
  uint32_t n = 0, k = 0, l = 0;
  uint32_t nthreads = 6;  // accepted for compatibility; the GPU replaces the PhiRunnerE pool
  uint32_t reportfreq = 100000;
  double seed = 0;
  double stop_threshold = 1e-5;
  std::string datfname = "network.dat", label, eta_type = "default", locations_file, idfile;
  bool force = false, file_suffix = false, save_beta = false, adagrad = false, use_test_set = false;
  bool compute_beta = false, logl = false, loadcmp = false;
  int ngpus = 1, device = 0;  // extensions (not in the reference)
  // fixed defaults (env.hh:200-249)
  double meanchangethresh = 0.001, validation_ratio = 0.005, heldout_indiv_ratio = 0.001, test_ratio = 0.005;
  double eta0 = 1.0, eta1 = 1.0, tau0 = 1, kappa = 0.5, nodetau0 = 1, nodekappa = 0.5;
  uint32_t online_iterations = 10;
  volatile bool terminate = false;
  std::string prefix;
  FILE *plogf = nullptr, *logf = nullptr;

  std::string file_str(const std::string &f) const { return prefix + f; }
  double alpha() const { return 1.0 / k; }

  void plog(const std::string &s, double v) { fprintf(plogf, "%s: %.9f\n", s.c_str(), v); fflush(plogf); }
  void plog(const std::string &s, bool v) { fprintf(plogf, "%s: %s\n", s.c_str(), v ? "True" : "False"); fflush(plogf); }
  void plog(const std::string &s, uint32_t v) { fprintf(plogf, "%s: %d\n", s.c_str(), v); fflush(plogf); }
  void plog(const std::string &s, int v) { fprintf(plogf, "%s: %d\n", s.c_str(), v); fflush(plogf); }
  void plog(const std::string &s, uint64_t v) { fprintf(plogf, "%s: %lu\n", s.c_str(), (unsigned long)v); fflush(plogf); }

  // Logger::xlog at level ERROR, the only live level (log.hh:46-50)
  void lerr(const char *fmt, ...) {
    if (!logf) return;
    char ts[64];
    time_t now = time(0);
    struct tm p;
    localtime_r(&now, &p);
    strftime(ts, sizeof ts, "%b %e %T", &p);
    fprintf(logf, "[%s] [%d] [ERR] ", ts, getpid());
    va_list ap;
    va_start(ap, fmt);
    vfprintf(logf, fmt, ap);
    va_end(ap);
    fprintf(logf, "\n\n");
    fflush(logf);
  }

  // env.hh:251-312 + Logger::setup_log_dir (log.cc:97-118)
  int init_dir() {
    std::ostringstream sa;
    sa << "n" << n << "-" << "k" << k << "-" << "l" << l;
    if (label != "") sa << "-" << label;
    else if (datfname.length() > 3) {
      std::string q = datfname.substr(0, 2);
      if (q == "..") q = "xx";
      sa << "-" << q;
    }
    if (seed != 0) sa << "-" << "seed" << seed;
    prefix = sa.str();
    fprintf(stdout, "+ Creating directory %s\n", prefix.c_str());
    struct stat st;
    if (stat(prefix.c_str(), &st) != 0) {
      if (errno != ENOENT || (mkdir(prefix.c_str(), 0775), stat(prefix.c_str(), &st) != 0)) {
        fprintf(stderr, "Warning: could not create dir %s\n", prefix.c_str());
        return -1;
      }
    } else if (!force) {
      fprintf(stderr, "Error: dir %s already exists\n", prefix.c_str());
      return -1;
    }
    logf = fopen(file_str("/infer.log").c_str(), "w");
    if (logf) fprintf(stderr, "+ Writing log to %s\n", file_str("/infer.log").c_str());
    fflush(stdout);
    plogf = fopen(file_str("/param.txt").c_str(), "w");
    if (!plogf) {
      printf("cannot open param file:%s\n", strerror(errno));
      return -1;
    }
    const uint32_t blocks = n > 10000 ? 100 : 10;
    plog("n", n); plog("k", k); plog("t", (uint32_t)2); plog("l", l); plog("nthreads", nthreads);
    plog("tau0", tau0); plog("nodetau0", nodetau0); plog("kappa", kappa); plog("nodekappa", nodekappa);
    plog("alpha", alpha()); plog("heldout_indiv_ratio", heldout_indiv_ratio);
    plog("validation_ratio", validation_ratio); plog("online_iterations", online_iterations);
    plog("GSL seed", seed); plog("file suffix", file_suffix); plog("save beta", save_beta);
    plog("adagrad", adagrad); plog("indiv sample size", n / blocks); plog("blocks", blocks);
    plog("compute_beta", compute_beta); plog("stop_threshold", stop_threshold);
    const std::string nd = file_str("/network.dat");
    unlink(nd.c_str());
    if (symlink(datfname.c_str(), nd.c_str()) < 0) return -1;
    fprintf(stderr, "+ done initializing env\n");
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------
// SNP: ingest that keeps genotypes 2-bit packed (src/snp.cc)
// ---------------------------------------------------------------------------------------------
struct SNP {
  Env &env;
  // SNP-major PLINK codes, bps bytes per locus.  A .bed file is memory-mapped (a 1M x 1M data
  // set is 250 GB: it is never copied into host RAM as a whole, and never unpacked);
  // a .012 text file is packed into `owned`.
  const uint8_t *rows = nullptr;
  std::vector<uint8_t> owned;
  void *map_base = nullptr;
  size_t map_len = 0;
  size_t bps = 0;
  std::vector<std::string> labels;
  explicit SNP(Env &e) : env(e) {}
  ~SNP() { if (map_base) munmap(map_base, map_len); }

  static long count_lines(const std::string &path) {
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    long nl = 0;
    std::vector<char> buf(20480);
    while (fgets(buf.data(), (int)buf.size(), f)) nl++;
    fclose(f);
    return nl;
  }

  // "missing snps", "0s/1s/2s snps" lines of param.txt (snp.cc:243-247): one pass over the packed rows,
  // 32 genotypes per 64-bit word (three popcounts), rows split over host threads -- at 250 GB the
  // reference's per-genotype loop would take longer than the inference.
  void count_codes() {
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::array<uint64_t, 4>> part(nt, std::array<uint64_t, 4>{0, 0, 0, 0});
    const size_t full = env.n / 4, tail = env.n % 4;
    auto work = [&](unsigned t) {
      const uint64_t M = 0x5555555555555555ull;
      uint64_t c1 = 0, c2 = 0, c3 = 0, tot = 0, cz[4] = {0, 0, 0, 0};
      for (uint32_t loc = t; loc < env.l; loc += nt) {
        const uint8_t *r = rows + (size_t)loc * bps;
        size_t i = 0;
        for (; i + 8 <= full; i += 8) {
          uint64_t x;
          memcpy(&x, r + i, 8);
          const uint64_t lo = x & M, hi = (x >> 1) & M;
          c3 += __builtin_popcountll(lo & hi);
          c1 += __builtin_popcountll(lo & ~hi);
          c2 += __builtin_popcountll(hi & ~lo);
          tot += 32;
        }
        for (; i < full; ++i)
          for (int j = 0; j < 4; ++j) cz[(r[i] >> (2 * j)) & 3]++;
        for (size_t j = 0; j < tail; ++j) cz[(r[full] >> (2 * j)) & 3]++;
      }
      part[t] = {cz[0] + (tot - c1 - c2 - c3), cz[1] + c1, cz[2] + c2, cz[3] + c3};
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
    uint64_t c[4] = {0, 0, 0, 0};
    for (auto &pt : part)
      for (int k = 0; k < 4; ++k) c[k] += pt[k];
    env.plog("missing snps", (uint32_t)c[1]);
    env.plog("0s snps", c[3]);  // the reference counts y=2 under "0s" (snp.cc:207-209)
    env.plog("1s snps", c[2]);
    env.plog("2s snps", c[0]);
  }

  int read_bed(const std::string &s) {  // snp.cc:95-253
    const std::string prefix = s.substr(0, s.length() - 4);
    long l = count_lines(prefix + ".bim");
    if (l < 0) { env.lerr("cannot open file %s.bim:%s", prefix.c_str(), strerror(errno)); return -1; }
    printf("+ bim file tells us %ld SNPs\n", l);
    if ((long)env.l != l) { env.lerr("-l input doesn't match SNPs in bim file\n"); return -1; }
    long n = count_lines(prefix + ".fam");
    if (n < 0) { env.lerr("cannot open file %s.fam:%s", prefix.c_str(), strerror(errno)); return -1; }
    printf("+ fam file tells us %ld individuals\n", n);
    if ((long)env.n != n) { env.lerr("-n input doesn't match individuals in fam file\n"); return -1; }
    bps = (env.n + 3) / 4;
    const int fd = open(s.c_str(), O_RDONLY);
    if (fd < 0) { env.lerr("cannot open file %s:%s", s.c_str(), strerror(errno)); return -1; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 3) { env.lerr("%s magic number incorrect\n", s.c_str()); close(fd); return -1; }
    map_len = (size_t)st.st_size;
    map_base = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map_base == MAP_FAILED) { map_base = nullptr; env.lerr("cannot map %s:%s", s.c_str(), strerror(errno)); return -1; }
    const unsigned char *h = (const unsigned char *)map_base;
    if (h[0] != 108 || h[1] != 27) { env.lerr("%s magic number incorrect\n", s.c_str()); return -1; }
    if (h[2] == 0) { env.lerr("individual major mode not supported yet!\n"); return -1; }
    if (h[2] != 1) { env.lerr("mode problem in %s\n", s.c_str()); return -1; }
    if (map_len - 3 < (size_t)env.l * bps) { env.lerr("%s is shorter than -n/-l imply\n", s.c_str()); return -1; }
    rows = h + 3;
    count_codes();
    return 0;
  }

  int read_012(const std::string &s) {  // snp.cc:6-93: one line of N characters per locus
    printf("+ reading (%d,%d) snps from %s\n", env.n, env.l, s.c_str());
    bps = (env.n + 3) / 4;
    owned.assign((size_t)env.l * bps, 0);
    std::ifstream in(s);
    if (!in) { env.lerr("cannot open file %s:%s", s.c_str(), strerror(errno)); return -1; }
    std::string line;
    uint32_t loc = 0;
    static const uint8_t code_of_y[3] = {0, 2, 3};
    while (loc < env.l && (in >> line)) {
      if (line.size() < env.n) { printf("Error: unexpected lines in file\n"); return -1; }
      for (uint32_t i = 0; i < env.n; ++i) {
        const uint8_t code = (line[i] == '-') ? 1 : code_of_y[(line[i] - '0') % 3];
        owned[loc * bps + (i >> 2)] |= code << (2 * (i & 3));
      }
      loc++;
    }
    rows = owned.data();
    count_codes();
    return 0;
  }

  int read(const std::string &s) {
    if (s.length() < 4) { env.lerr("unrecognized file extension"); return -1; }
    const std::string ext = s.substr(s.length() - 4, 4);
    if (ext == ".bed") { printf("+ bed format detected\n"); return read_bed(s); }
    if (ext == ".012") { printf("+ .012 detected"); return read_012(s); }
    env.lerr("unrecognized file extension");
    return -1;
  }

  int read_idfile(const std::string &s) {  // snp.cc:255-276
    std::ifstream in(s);
    if (!in) { env.lerr("cannot open file %s:%s", s.c_str(), strerror(errno)); return -1; }
    std::string t;
    while (in >> t) labels.push_back(t);
    return 0;
  }
  std::string label(uint32_t i) const { return i < labels.size() ? labels[i] : ""; }
};

// ---------------------------------------------------------------------------------------------
// SNPSamplingE: the driver (src/snpsamplinge.cc)
// ---------------------------------------------------------------------------------------------
#define TSD_CHECK(call)                                                        \
  do {                                                                         \
    if ((call) != 0) {                                                         \
      fprintf(stderr, "tsgpu: %s\n  in %s\n", ts_last_error(), #call);         \
      exit(-1);                                                                \
    }                                                                          \
  } while (0)

class SNPSamplingE {
 public:
  SNPSamplingE(Env &env, SNP &snp) : _env(env), _snp(snp), _n(env.n), _k(env.k), _l(env.l), _start_time(time(0)) {
    printf("+ initialization begin\n");
    fflush(stdout);
    _env.plog("individuals n", _n);
    _env.plog("locations l", _l);
    _env.plog("populations k", _k);
    _r = ts_rng_create(_env.seed);  // gsl_rng_alloc + gsl_rng_set if seed (cc:59-63)
    unlink(_env.file_str("/likelihood-analysis.txt").c_str());
    _vf = fopen(_env.file_str("/validation.txt").c_str(), "w");
    if (!_vf) { printf("cannot open heldout file:%s\n", strerror(errno)); exit(-1); }
    if (_env.compute_beta) _env.online_iterations = 100;  // cc:75

    init_heldout_sets();
    create_engines();

    if (_env.compute_beta) {  // cc:74-95
      _env.lerr("done starting threads");
      load_gamma();
      _env.lerr("done estimating all theta");
      if (_env.locations_file == "") {
        compute_all_lambda();
        save_beta(nullptr);
      } else
        compute_and_save_beta();
      exit(0);
    }

    init_gamma();
    printf("+ computing initial heldout likelihood\n");
    compute_likelihood(true);
    save_gamma();
    printf("\n+ computing initial training likelihood\n");
    printf("+ done..\n");
    printf("+ initialization end\n");
    fflush(stdout);
  }

  ~SNPSamplingE() {
    for (ts_engine *e : _eng) ts_destroy(e);
    if (_vf) fclose(_vf);
    ts_free(_val_loc); ts_free(_val_off); ts_free(_val_indiv);
    ts_rng_destroy(_r);
  }

  // cc:417-459.  SNP indices are drawn on the host from the reference's RNG stream, in
  // batches that end at the next report; the engine runs a batch without host round trips.
  // The reference prints a progress line every 100 iterations (cc:436-439); the lines of a batch
  // are printed when the batch has completed (same text, the seconds are the batch's).
  // SIGTERM (main.cc:28-39) is honoured at batch boundaries (at most `chunk_max` iterations later;
  // the reference: after the current iteration, before the pending lazy gamma step -- here the
  // gamma step of the last SNP has been applied, which SURVEY App. A note 3 allows).
  void infer() {
    const uint32_t chunk_max = 1u << 15;
    std::vector<uint32_t> locs;
    while (1) {
      const uint32_t to_report = _env.reportfreq - _iter % _env.reportfreq;
      const uint32_t m = std::min(to_report, chunk_max);
      locs.resize(m);
      ts_rng_sample_locs(_r, _l, locs.data(), m);
      for_each_engine([&](size_t r) {
        TSD_CHECK(ts_steps(_eng[r], locs.data(), m, 0, nullptr));
        TSD_CHECK(ts_sync(_eng[r]));
      });
      const uint32_t before = _iter;
      _iter += m;
      for (uint32_t it = (before / 100 + 1) * 100; it <= _iter; it += 100) printf("\riteration = %d took %d secs", it, duration());
      fflush(stdout);
      if (_iter % _env.reportfreq == 0) {
        printf("iteration = %d took %d secs\n", _iter, duration());
        _env.lerr("iteration = %d took %d secs\n", _iter, duration());
        _env.lerr("computing heldout likelihood @ %d secs", duration());
        compute_likelihood(false);
        _env.lerr("saving theta @ %d secs", duration());
        save_model();
        _env.lerr("done @ %d secs", duration());
      }
      if (_env.terminate) {
        save_model();
        exit(0);
      }
    }
  }

 private:
  uint32_t duration() const { return (uint32_t)(time(0) - _start_time); }

  void init_heldout_sets() {  // cc:131-140, :196-224
    TSD_CHECK(ts_sample_validation(_r, _n, _l, _snp.rows, _snp.bps, &_nval, &_val_loc, &_val_off, &_val_indiv));
    const uint32_t per_loc_h = _n < 2000 ? (_n / 10) : (_n / 100);
    const uint32_t nlocs = (uint32_t)(_l * _env.validation_ratio);
    _env.plog("validation snps per location", per_loc_h);
    _env.plog("validation locations", nlocs);
    _env.plog("total validation snps", per_loc_h * nlocs);
    _env.plog("(VAL1) total validation snps (check)", (uint32_t)_val_off[_nval]);
    _env.plog("test ratio", _env.test_ratio);
    _env.plog("validation ratio", _env.validation_ratio);
  }

  void create_engines() {
    int ng = _env.ngpus;
    auto per_of = [&](int g) { return (((uint64_t)_n + g - 1) / g + 3) / 4 * 4; };
    // every shard needs at least one individual (shard boundaries are multiples of 4)
    while (ng > 1 && (uint64_t)(ng - 1) * per_of(ng) >= _n) ng--;
    if (ng != _env.ngpus) {
      fprintf(stderr, "+ -gpus %d leaves a GPU without individuals at n = %u; using %d\n", _env.ngpus, _n, ng);
      _env.ngpus = ng;
    }
    const uint64_t per = per_of(ng);
    for (int r = 0; r < ng; ++r) {
      ts_config cfg;
      ts_config_defaults(&cfg, _n, _l, _k);
      cfg.online_iterations = _env.online_iterations;
      cfg.meanchangethresh = _env.meanchangethresh;
      cfg.eta0 = _env.eta0; cfg.eta1 = _env.eta1;
      cfg.nodetau0 = _env.nodetau0 + 1; cfg.nodekappa = _env.nodekappa;
      cfg.device = _env.device + r; cfg.rank = r; cfg.nranks = ng;
      cfg.n_begin = std::min<uint64_t>(r * per, _n);
      cfg.n_local = std::min<uint64_t>(per, _n - cfg.n_begin);
      ts_engine *e = nullptr;
      TSD_CHECK(ts_create(&cfg, &e));
      _eng.push_back(e);
      _begin.push_back(cfg.n_begin);
      _local.push_back(cfg.n_local);
    }
    // read_bed's payload loop (snp.cc:186-229): ONE pass over the (memory-mapped) rows, every GPU takes its bytes
    const auto t0 = std::chrono::steady_clock::now();
    TSD_CHECK(ts_load_bed_fanout(_eng.data(), ng, 0, _l, _snp.rows, _snp.bps));
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "+ genotypes resident on %d GPU(s): %.3f GB in %.2f s (%.2f GB/s)\n", ng, 1e-9 * (double)_l * _snp.bps, secs,
            1e-9 * (double)_l * _snp.bps / std::max(secs, 1e-9));
    for (int r = 0; r < ng; ++r) TSD_CHECK(ts_set_validation(_eng[r], _nval, _val_loc, _val_off, _val_indiv));
    if (ng > 1) TSD_CHECK(ts_comm_connect_local(_eng.data(), ng));
  }

  void set_gamma_all(const std::vector<double> &g) {
    for (size_t r = 0; r < _eng.size(); ++r) TSD_CHECK(ts_set_gamma(_eng[r], g.data() + _begin[r] * _k));
  }

  void init_gamma() {  // cc:226-237 (+ estimate_all_theta cc:101 on the device)
    std::vector<double> g((size_t)_n * _k);
    ts_init_gamma(_r, _n, _k, g.data());
    set_gamma_all(g);
  }

  std::vector<double> gather(int (*getter)(ts_engine *, double *)) {
    std::vector<double> out((size_t)_n * _k);
    for (size_t r = 0; r < _eng.size(); ++r) TSD_CHECK(getter(_eng[r], out.data() + _begin[r] * _k));
    return out;
  }

  // compute_likelihood(first, validation=true) (cc:461-544)
  void compute_likelihood(bool first) {
    std::vector<double> per(_nval, 0.0), tmp(_nval);
    uint64_t k = 0;
    // all shards must run the hol-mode optimisation together (they exchange every round)
    std::vector<double> sums(_eng.size());
    std::vector<uint64_t> cnts(_eng.size());
    if (_eng.size() == 1) {
      TSD_CHECK(ts_heldout_ll(_eng[0], first, &sums[0], &cnts[0], per.data()));
      k = cnts[0];
    } else {
      run_heldout_multi(first, per, k);
    }
    if (!first) _iter += (uint32_t)_nval;  // snp_likelihood: _iter++ per locus (hh:333)
    for (uint64_t v = 0; v < _nval; ++v) printf("\rdone:%.2f%%", ((double)v / _nval) * 100);  // cc:494
    double s = 0.0;
    for (uint64_t v = 0; v < _nval; ++v) s += per[v];
    const double a = s / k;
    fprintf(_vf, "%d\t%d\t%.9f\t%d\t%f\n", _iter, duration(), a, (int)k, exp(a));
    fflush(_vf);
    bool stop = false;
    if (_iter > 2000) {
      if (a > _prev_h && _prev_h != 0 && fabs((a - _prev_h) / _prev_h) < _env.stop_threshold) stop = true;
      else if (a < _prev_h) _nh++;
      else if (a > _prev_h) _nh = 0;
      if (a > _max_h) _max_h = a;
      if (_nh > 3) stop = true;
    }
    _prev_h = a;
    if (stop) {  // use_validation_stop is always true (env.hh:241)
      save_model();
      exit(0);
    }
  }

  // Shards exchange partial sums every round, so calls that end in a stream sync must be
  // issued to all engines concurrently: one host thread per engine.
  void for_each_engine(const std::function<void(size_t)> &fn) {
    if (_eng.size() == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (size_t r = 0; r < _eng.size(); ++r) th.emplace_back(fn, r);
    for (auto &t : th) t.join();
  }

  void run_heldout_multi(bool first, std::vector<double> &per, uint64_t &k) {
    std::vector<std::vector<double>> parts(_eng.size(), std::vector<double>(_nval, 0.0));
    std::vector<uint64_t> cnts(_eng.size(), 0);
    for_each_engine([&](size_t r) {
      double s;
      TSD_CHECK(ts_heldout_ll(_eng[r], first, &s, &cnts[r], parts[r].data()));
    });
    k = 0;
    for (size_t r = 0; r < _eng.size(); ++r) {  // rank order
      k += cnts[r];
      for (uint64_t v = 0; v < _nval; ++v) per[v] += parts[r][v];
    }
  }

  std::string add_iter_suffix(const char *c) {  // cc:578-587
    std::ostringstream sa;
    if (_env.file_suffix) sa << c << "_" << _iter << ".txt";
    else sa << c << ".txt";
    return _env.file_str(sa.str());
  }

  void save_gamma() {  // cc:546-576
    FILE *f = fopen(add_iter_suffix("/gamma").c_str(), "w");
    FILE *g = fopen(add_iter_suffix("/theta").c_str(), "w");
    if (!f || !g) { _env.lerr("cannot open gamma/theta file:%s\n", strerror(errno)); exit(-1); }
    const std::vector<double> gd = gather(ts_get_gamma), td = gather(ts_get_theta);
    for (uint32_t n = 0; n < _n; ++n) {
      for (uint32_t k = 0; k < _k; ++k) {
        fprintf(f, "%.8f\t", gd[(size_t)n * _k + k]);
        fprintf(g, "%.8f\t", td[(size_t)n * _k + k]);
      }
      fprintf(f, "\n");
      fprintf(g, "\n");
    }
    fclose(f);
    fclose(g);
  }
  void save_model() { save_gamma(); }

  void load_gamma() {  // cc:800-862: reads ./gamma.txt (CWD), writes gammasave.txt
    FILE *gammaf = fopen("gamma.txt", "r");
    if (!gammaf) { _env.lerr("cannot open gamma file:%s\n", strerror(errno)); exit(-1); }
    std::vector<double> g((size_t)_n * _k, 0.0);
    const int sz = 128 * _k;
    std::vector<char> line(sz);
    uint32_t n = 0;
    while (n < _n && fgets(line.data(), sz, gammaf)) {
      char *p = line.data();
      for (uint32_t k = 0; k < _k; ++k) {
        char *q = nullptr;
        const double d = strtod(p, &q);
        if (p == q) { fprintf(stderr, "error parsing gamma file\n"); abort(); }
        g[(size_t)n * _k + k] = d;
        p = q;
      }
      n++;
    }
    fclose(gammaf);
    FILE *f = fopen(_env.file_str("/gammasave.txt").c_str(), "w");
    if (!f) { _env.lerr("cannot open gammasave file:%s\n", strerror(errno)); exit(-1); }
    for (uint32_t i = 0; i < _n; ++i) {
      std::string s = _snp.label(i);
      if (s == "") s = "unknown";
      fprintf(f, "%d\t%s\t", i, s.c_str());
      double max = .0;
      uint32_t max_k = 0;
      for (uint32_t k = 0; k < _k; ++k) {
        fprintf(f, "%.8f\t", g[(size_t)i * _k + k]);
        if (g[(size_t)i * _k + k] > max) { max = g[(size_t)i * _k + k]; max_k = k; }
      }
      fprintf(f, "%d\n", max_k);
    }
    fclose(f);
    set_gamma_all(g);
  }

  void sweep(const std::vector<uint32_t> &locs) {  // compute_all_lambda loop body (cc:372-380)
    const size_t chunk = 1u << 14;
    for (size_t off = 0; off < locs.size(); off += chunk) {
      const size_t m = std::min(chunk, locs.size() - off);
      for_each_engine([&](size_t r) {
        TSD_CHECK(ts_steps(_eng[r], locs.data() + off, m, 0, nullptr));
        TSD_CHECK(ts_sync(_eng[r]));
      });
      for (size_t j = 0; j < m; ++j) {  // cc:376-379 / :405-408
        _iter++;
        if (locs[off + j] % 100 == 0) printf("\rloc = %d took %d secs", _iter, duration());
      }
      fflush(stdout);
    }
  }

  void compute_all_lambda() {  // cc:368-381
    std::vector<uint32_t> locs(_l);
    for (uint32_t i = 0; i < _l; ++i) locs[i] = i;
    sweep(locs);
  }

  void compute_and_save_beta() {  // cc:384-413
    _env.lerr("within compute_and_save_beta()");
    // File grammar of the reference (cc:389-397): records "<loc><TAB><rest of line>", read with
    // fscanf("%d\t%*[^\n]s\n").  The conversion is kept as it is there, quirks included: the
    // white-space directive after %d also eats a newline, so in a file with bare numbers (no second
    // column) every other line is swallowed as the "rest of line" -- the reference behaves the same.
    FILE *lf = fopen(_env.locations_file.c_str(), "r");
    if (!lf) { _env.lerr("cannot open locations file:%s\n", strerror(errno)); exit(-1); }
    std::vector<uint32_t> locs;
    int v = 0;
    while (!feof(lf)) {
      if (fscanf(lf, "%d\t%*[^\n]s\n", &v) >= 0) {
        if (v < 0 || (uint32_t)v >= _l) { fprintf(stderr, "bad location %d in %s\n", v, _env.locations_file.c_str()); exit(-1); }
        _env.lerr("loc = %d", v);
        locs.push_back((uint32_t)v);
      }
    }
    fclose(lf);
    _env.lerr("locs size = %d", (int)locs.size());
    sweep(locs);
    save_beta(&locs);
  }

  void save_beta(const std::vector<uint32_t> *locs) {  // cc:761-798
    FILE *f = fopen(add_iter_suffix("/beta").c_str(), "w");
    if (!f) { _env.lerr("cannot open beta or lambda file:%s\n", strerror(errno)); exit(-1); }
    std::vector<double> beta((size_t)_l * _k);
    TSD_CHECK(ts_get_beta(_eng[0], 0, _l, beta.data()));  // lambda is replicated on every shard
    const size_t cnt = locs ? locs->size() : _l;
    for (size_t i = 0; i < cnt; ++i) {
      const uint32_t loc = locs ? (*locs)[i] : (uint32_t)i;
      fprintf(f, "%d\t", loc);
      for (uint32_t k = 0; k < _k; ++k) fprintf(f, "%.8f\t", beta[(size_t)loc * _k + k]);
      fprintf(f, "\n");
    }
    fclose(f);
  }

  Env &_env;
  SNP &_snp;
  uint32_t _n, _k, _l;
  uint32_t _iter = 0;
  time_t _start_time;
  ts_rng *_r = nullptr;
  FILE *_vf = nullptr;
  uint64_t _nval = 0;
  uint32_t *_val_loc = nullptr, *_val_indiv = nullptr;
  uint64_t *_val_off = nullptr;
  std::vector<ts_engine *> _eng;
  std::vector<uint64_t> _begin, _local;
  double _prev_h = -2147483647, _max_h = -2147483647;
  uint32_t _nh = 0;
};

}  // namespace tsd
