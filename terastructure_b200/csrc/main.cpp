// terastructure -- drop-in CLI for the reference's `terastructure` binary (src/main.cc:41-303):
// same flags, same output directory (n{N}-k{K}-l{L}[-label][-seed{S}]) and files (param.txt,
// infer.log, validation.txt, gamma.txt, theta.txt, beta.txt, gammasave.txt).  The SVI hot path
// runs on B200 GPU(s) through libtsgpu.so; there is no CPU path.
// Extensions (not in the reference): -gpus <n> shards the individuals over n GPUs of this
// node, -device <d> picks the first device ordinal.
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ts_driver.hpp"

static tsd::Env *env_global = nullptr;

static void term_handler(int sig) {  // main.cc:28-39
  if (env_global) {
    printf("Got termination signal. Saving model state and quitting.\n");
    fflush(stdout);
    env_global->terminate = true;
  } else {
    signal(sig, SIG_DFL);
    raise(sig);
  }
}

static void usage() {  // main.cc:305-323
  fprintf(stdout,
          "Population inference software for SNP data.\n"
          "terastructure [OPTIONS]\n"
          "\t-help\t\tusage\n"
          "\t-file <name>\t location by individuals matrix of SNP values: PLINK .bed or .012 text\n"
          "\t-n <N>\t\t number of individuals\n"
          "\t-l <L>\t\t number of locations\n"
          "\t-k <K>\t\t number of populations\n"
          "\t-label\t\t descriptive tag for the output directory\n"
          "\t-force\t\t overwrite existing output directory\n"
          "\t-rfreq <val>\t checks for convergence and logs output every <val> iterations\n"
          "\t-idfile\t\t file containing individual name/meta-data, one per line\n"
          "\t-seed <val>\t RNG seed (GSL mt19937 stream of the reference)\n"
          "\t-compute-beta\t second pass: reload ./gamma.txt, sweep all loci, write beta.txt\n"
          "\t-gpus <n>\t shard individuals over n GPUs (extension)\n");
  fflush(stdout);
}

int main(int argc, char **argv) {
  signal(SIGTERM, term_handler);
  tsd::Env env;
  bool rfreq_set = false;
  uint32_t rfreq = 10000;
  if (argc == 1) {
    usage();
    exit(-1);
  }
  auto need = [&](int i) {
    if (i + 1 > argc - 1) {
      fprintf(stderr, "+ insufficient arguments!\n");
      exit(-1);
    }
  };
  for (int i = 1; i < argc; ++i) {
    const char *a = argv[i];
    if (!strcmp(a, "-help")) { usage(); exit(0); }
    else if (!strcmp(a, "-force")) { fprintf(stdout, "+ overwrite option set\n"); env.force = true; }
    else if (!strcmp(a, "-online")) fprintf(stdout, "+ online option set\n");
    else if (!strcmp(a, "-file") || !strcmp(a, "-bed")) { need(i); env.datfname = argv[++i]; fprintf(stdout, "+ using file %s\n", env.datfname.c_str()); }
    else if (!strcmp(a, "-batch")) { fprintf(stdout, "batch option currently not available"); exit(0); }
    else if (!strcmp(a, "-n")) { need(i); env.n = atoi(argv[++i]); fprintf(stdout, "+ n = %d\n", env.n); }
    else if (!strcmp(a, "-k")) { need(i); env.k = atoi(argv[++i]); fprintf(stdout, "+ K = %d\n", env.k); }
    else if (!strcmp(a, "-l")) { need(i); env.l = atoi(argv[++i]); fprintf(stdout, "+ L = %d\n", env.l); }
    else if (!strcmp(a, "-label")) { need(i); env.label = argv[++i]; }
    else if (!strcmp(a, "-eta-type")) { need(i); env.eta_type = argv[++i]; fprintf(stdout, "+ eta-type = %s\n", env.eta_type.c_str()); }
    else if (!strcmp(a, "-rfreq")) { need(i); rfreq = atoi(argv[++i]); fprintf(stdout, "+ rfreq = %d\n", rfreq); rfreq_set = true; }
    else if (!strcmp(a, "-logl")) { env.logl = true; fprintf(stdout, "+ logl option set\n"); }
    else if (!strcmp(a, "-idfile")) { need(i); env.idfile = argv[++i]; fprintf(stdout, "+ idfile = %s\n", env.idfile.c_str()); }
    else if (!strcmp(a, "-loadcmp")) { env.loadcmp = true; fprintf(stdout, "+ loadcmp option set\n"); }
    else if (!strcmp(a, "-E")) fprintf(stdout, "+ algorithm E option set\n");
    else if (!strcmp(a, "-stochastic")) fprintf(stdout, "+ stochastic option set\n");
    else if (!strcmp(a, "-seed")) { need(i); env.seed = atof(argv[++i]); fprintf(stdout, "+ random seed set to %.5f\n", env.seed); }
    else if (!strcmp(a, "-file-suffix")) env.file_suffix = true;
    else if (!strcmp(a, "-save-beta")) env.save_beta = true;
    else if (!strcmp(a, "-adagrad")) env.adagrad = true;
    else if (!strcmp(a, "-nthreads")) { need(i); env.nthreads = atoi(argv[++i]); }
    else if (!strcmp(a, "-use-test-set")) env.use_test_set = true;
    else if (!strcmp(a, "-locations-file")) { need(i); env.locations_file = argv[++i]; }
    else if (!strcmp(a, "-compute-beta")) env.compute_beta = true;
    else if (!strcmp(a, "-stop-threshold")) { need(i); env.stop_threshold = atof(argv[++i]); }
    else if (!strcmp(a, "-gpus")) { need(i); env.ngpus = atoi(argv[++i]); }
    else if (!strcmp(a, "-device")) { need(i); env.device = atoi(argv[++i]); }
    else {
      fprintf(stdout, "error: unknown option %s\n", a);  // the reference: assert(0)
      fflush(stdout);
      abort();
    }
  }
  env.reportfreq = rfreq_set ? rfreq : 100000;  // main.cc:189-190
  if (env.use_test_set) {
    fprintf(stderr, "-use-test-set writes to a never-opened file in the reference and crashes there; not supported\n");
    return -1;
  }
  if (env.n == 0 || env.l == 0 || env.k == 0 || env.k > TS_MAX_K || env.ngpus < 1 || env.ngpus > 16) {
    fprintf(stderr, "error: need -n, -l, -k (K <= %d) and 1 <= -gpus <= 16\n", TS_MAX_K);
    return -1;
  }
  if (ts_device_count() < env.device + env.ngpus) {
    fprintf(stderr, "error: %d CUDA device(s) visible, need %d (this build has no CPU path)\n", ts_device_count(),
            env.device + env.ngpus);
    return -1;
  }
  if (env.init_dir() < 0) abort();
  env_global = &env;

  tsd::SNP snp(env);
  if (snp.read(env.datfname) < 0) {
    fprintf(stderr, "error reading %s; quitting\n", env.datfname.c_str());
    return -1;
  }
  if (env.idfile != "" && snp.read_idfile(env.idfile) < 0)
    fprintf(stderr, "error reading %s; quitting\n", env.idfile.c_str());

  if (!env.loadcmp) {
    tsd::SNPSamplingE e(env, snp);
    e.infer();
  }
  return 0;
}
