/*
 * sipnet_gpu_main.c -- `sipnet_gpu`: drop-in for the reference's `sipnet` driver.
 *
 * Same command line, same sipnet.in / .param / .clim / events.in inputs, same
 * <prefix>.out / events.out / <prefix>.config outputs and exit codes as
 * reference src/sipnet/frontend.c:130-253 -- but the body of runModelOutput()
 * (reference src/sipnet/sipnet.c:1954-1990) runs on a B200 through the C ABI
 * include/sipnet_gpu.h.
 * Extensions (one launch for many members / sites):
 *   --ensemble-params FILE  one member per listed .param file on the same site; member k's outputs go to
 *                           <prefix>.out.<k> / <events-prefix>.out.<k>
 *   --site-list FILE        one site per line: <file-prefix> [<events-prefix> [<member-list>]] -- each site has its
 *                           own .clim / .param (or member list) / events files and gets its own output files; all
 *                           sites run under this invocation's flags
 *   --devices N             GPUs used by those launches (0 = all visible): one host thread drives them through
 *                           sipnet_gpu_multi_* -- whole sites per GPU, or an even share of one site's members
 * --debug-log <prefix> writes the reference's three per-step debug logs from the device's validation dump.
 * --restart-in / --restart-out read and write the reference's checkpoint format (sip_restart.c), so a segmented
 * run can alternate between this binary and the reference's.
 * --do-single-outputs writes <prefix>.NEE / .NEE_cum / .GPP / .GPP_cum like outputItems.c (one "%f " per step).
 */
#define _GNU_SOURCE
#include <libgen.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "sip_host.h"

static int die(int code, const char *what) {
  printf("[ERROR  ] %s\n", what);
  return code;
}

/* one site of the launch: its files, its members, its parsed inputs */
typedef struct {
  char prefix[SIP_NAME_MAX];       /* <prefix>.clim / .param / .out */
  char eventsPrefix[SIP_NAME_MAX]; /* <eventsPrefix>.in / .out */
  char memberList[SIP_NAME_MAX];   /* "" => one member from <prefix>.param */
  char **paramFiles;
  int64_t nmembers, member0;
  sip_site_data data;
} site_job;

static int read_member_list(site_job *job) {
  if (!job->memberList[0]) {
    job->paramFiles = (char **)malloc(sizeof *job->paramFiles);
    char name[SIP_NAME_MAX + 16];
    snprintf(name, sizeof name, "%s.param", job->prefix);
    job->paramFiles[0] = strdup(name);
    job->nmembers = 1;
    return 0;
  }
  FILE *lf = fopen(job->memberList, "r");
  if (!lf) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open the member (.param) list");
  char line[1024];
  while (fgets(line, sizeof line, lf)) {
    line[strcspn(line, "\r\n")] = '\0';
    if (!line[0]) continue;
    job->paramFiles = (char **)realloc(job->paramFiles, (size_t)(job->nmembers + 1) * sizeof *job->paramFiles);
    job->paramFiles[job->nmembers++] = strdup(line);
  }
  fclose(lf);
  if (job->nmembers == 0) return die(SIPNET_GPU_ERR_INPUT_FILE, "the member (.param) list is empty");
  return 0;
}

/* --site-list: "<file-prefix> [<events-prefix> [<member-list>]]" per line, '#' comments */
static int read_site_list(const char *path, site_job **jobs, int64_t *njobs) {
  FILE *f = fopen(path, "r");
  if (!f) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open the --site-list file");
  char line[3 * SIP_NAME_MAX + 16];
  while (fgets(line, sizeof line, f)) {
    char a[SIP_NAME_MAX] = "", b[SIP_NAME_MAX] = "", c[SIP_NAME_MAX] = "";
    char *hash = strchr(line, '#');
    if (hash) *hash = '\0';
    const int n = sscanf(line, " %255s %255s %255s", a, b, c);
    if (n <= 0) continue;
    *jobs = (site_job *)realloc(*jobs, (size_t)(*njobs + 1) * sizeof **jobs);
    site_job *job = &(*jobs)[(*njobs)++];
    memset(job, 0, sizeof *job);
    snprintf(job->prefix, sizeof job->prefix, "%s", a);
    if (n >= 2) {
      snprintf(job->eventsPrefix, sizeof job->eventsPrefix, "%s", b);
    } else { /* events.in / events.out next to the site's files */
      char tmp[SIP_NAME_MAX];
      snprintf(tmp, sizeof tmp, "%s", a);
      snprintf(job->eventsPrefix, sizeof job->eventsPrefix, "%.240s/events", dirname(tmp));
    }
    if (n >= 3) snprintf(job->memberList, sizeof job->memberList, "%s", c);
  }
  fclose(f);
  if (*njobs == 0) return die(SIPNET_GPU_ERR_INPUT_FILE, "the --site-list file names no site");
  return 0;
}

/* ---- inputs of a many-site launch on all host cores: initModel() + initEvents() per site ----
 * (10 000 ten-year sites are 73 million .clim lines: two minutes on one thread next to a one-second run.)
 * Sites are handed out in order; the first failing site IN ORDER decides the exit code, like a sequential pass:
 * after a failure no NEW site is started, the ones before it were all started already. */
typedef struct {
  site_job *jobs;
  int64_t njobs, M;
  const sip_context *ctx;
  double *params; /* [80][M] */
  int32_t *memberSite;
  sipnet_gpu_site *views;
  int *rcs;            /* [njobs] */
  char (*msgs)[512];   /* [njobs] */
  int64_t next;
  int failed;
  pthread_mutex_t lock;
} loader_pool;

static int load_site(loader_pool *lp, int64_t s) {
  site_job *job = &lp->jobs[s];
  const sip_context *ctx = lp->ctx;
  const int64_t M = lp->M;
  char name[SIP_NAME_MAX + 16];
  int rc;
  for (int64_t k = 0; k < job->nmembers; ++k) {
    double one[SIPNET_GPU_NPARAMS];
    const int64_t m = job->member0 + k;
    if ((rc = sip_read_params(job->paramFiles[k], &ctx->flags, ctx->quiet || m > 0, one))) return rc;
    for (int p = 0; p < SIPNET_GPU_NPARAMS; ++p) lp->params[(size_t)p * (size_t)M + (size_t)m] = one[p];
    lp->memberSite[m] = (int32_t)s;
  }
  snprintf(name, sizeof name, "%s.clim", job->prefix);
  if ((rc = sip_read_clim(name, ctx->flags.gdd, ctx->quiet || s > 0, &job->data))) return rc;
  if (ctx->flags.events) {
    double first[SIPNET_GPU_NPARAMS];
    for (int p = 0; p < SIPNET_GPU_NPARAMS; ++p) first[p] = lp->params[(size_t)p * (size_t)M + (size_t)job->member0];
    snprintf(name, sizeof name, "%s.in", job->eventsPrefix);
    if ((rc = sip_read_events(name, &ctx->flags, first, ctx->quiet || s > 0, &job->data))) return rc;
  }
  sip_site_view(&job->data, &lp->views[s]);
  return 0;
}

static void *loader_main(void *argp) {
  loader_pool *lp = (loader_pool *)argp;
  for (;;) {
    pthread_mutex_lock(&lp->lock);
    const int64_t s = (lp->failed || lp->next >= lp->njobs) ? -1 : lp->next++;
    pthread_mutex_unlock(&lp->lock);
    if (s < 0) break;
    const int rc = load_site(lp, s);
    if (rc) {
      snprintf(lp->msgs[s], sizeof lp->msgs[s], "%s", sip_host_error()); /* (this thread's message) */
      pthread_mutex_lock(&lp->lock);
      lp->rcs[s] = rc;
      lp->failed = 1;
      pthread_mutex_unlock(&lp->lock);
    }
  }
  return NULL;
}

/* returns 0, or the exit code of the first failing site with its message in *msg (static storage) */
static int load_sites(site_job *jobs, int64_t njobs, int64_t M, const sip_context *ctx, double *params, int32_t *memberSite,
                      sipnet_gpu_site *views, const char **msg) {
  static char firstMsg[512];
  loader_pool lp;
  memset(&lp, 0, sizeof lp);
  lp.jobs = jobs, lp.njobs = njobs, lp.M = M, lp.ctx = ctx, lp.params = params, lp.memberSite = memberSite, lp.views = views;
  lp.rcs = (int *)calloc((size_t)njobs, sizeof *lp.rcs);
  lp.msgs = (char(*)[512])calloc((size_t)njobs, sizeof *lp.msgs);
  if (!lp.rcs || !lp.msgs) {
    *msg = "memory allocation failure";
    return SIPNET_GPU_ERR_INTERNAL;
  }
  pthread_mutex_init(&lp.lock, NULL);
  long want = sysconf(_SC_NPROCESSORS_ONLN);
  const char *env = getenv("SIPNET_GPU_READER_THREADS");
  if (env && atoi(env) > 0) want = atoi(env);
  if (want > 64) want = 64;
  if ((int64_t)want > njobs) want = (long)njobs;
  pthread_t threads[64];
  int started = 0;
  for (long i = 1; i < want; ++i) { /* this thread is a worker too; a thread that cannot be started is simply absent */
    if (pthread_create(&threads[started], NULL, loader_main, &lp) != 0) break;
    ++started;
  }
  loader_main(&lp);
  for (int i = 0; i < started; ++i) pthread_join(threads[i], NULL);
  pthread_mutex_destroy(&lp.lock);
  int rc = 0;
  for (int64_t s = 0; s < njobs && rc == 0; ++s)
    if (lp.rcs[s]) {
      rc = lp.rcs[s];
      snprintf(firstMsg, sizeof firstMsg, "%s", lp.msgs[s]);
      *msg = firstMsg;
    }
  free(lp.rcs);
  free(lp.msgs);
  return rc;
}

int main(int argc, char **argv) {
  sip_context ctx;
  sip_context_init(&ctx);
  int rc = sip_parse_cli(&ctx, argc, argv);
  if (rc) return die(rc, sip_host_error());
  if (ctx.helpOrVersion) return 0;
  if ((rc = sip_read_input_file(&ctx))) return die(rc, sip_host_error());
  if ((rc = sip_validate_context(&ctx))) return die(rc, sip_host_error());
  const int useRestart = ctx.restartIn[0] || ctx.restartOut[0];
  const int many = ctx.ensembleParamList[0] || ctx.siteList[0];
  if (useRestart && many)
    return die(SIPNET_GPU_ERR_BAD_CLI,
               "a restart checkpoint holds one member: not available with --ensemble-params / --site-list");
  if (ctx.debugLogPrefix[0] && many) return die(SIPNET_GPU_ERR_BAD_CLI, "--debug-log is a single-member feature");
  if (ctx.ensembleParamList[0] && ctx.siteList[0])
    return die(SIPNET_GPU_ERR_BAD_CLI, "--ensemble-params and --site-list exclude each other (a site list names member lists)");
  if ((rc = sip_derive_file_names(&ctx))) return die(rc, sip_host_error());

  /* the reference opens its main output before reading anything (frontend.c:211-215) */
  FILE *out = NULL;
  if (ctx.doMainOutput && !many) {
    out = fopen(ctx.outFile, "w");
    if (!out) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open main output file");
  }
  if (ctx.dumpConfig) {
    FILE *cf = fopen(ctx.outConfigFile, "w");
    if (!cf) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open config dump file");
    char stamp[100];
    time_t now = time(NULL);
    strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S UTC", gmtime(&now));
    sip_print_config(&ctx, cf, stamp);
    fclose(cf);
  }

  /* ---- the sites of this launch ---- */
  site_job *jobs = NULL;
  int64_t njobs = 0;
  if (ctx.siteList[0]) {
    if ((rc = read_site_list(ctx.siteList, &jobs, &njobs))) return rc;
  } else {
    jobs = (site_job *)calloc(1, sizeof *jobs);
    njobs = 1;
    snprintf(jobs[0].prefix, sizeof jobs[0].prefix, "%s", ctx.filePrefix);
    snprintf(jobs[0].eventsPrefix, sizeof jobs[0].eventsPrefix, "%s", ctx.eventsPrefix);
    snprintf(jobs[0].memberList, sizeof jobs[0].memberList, "%s", ctx.ensembleParamList);
  }

  /* ---- inputs: initModel() + initEvents() per site, sipnet.c:2001-2009, events.c:427-433 ---- */
  int64_t M = 0, Tmax = 0, maxEvents = 0, maxYears = 0;
  for (int64_t s = 0; s < njobs; ++s) {
    if ((rc = read_member_list(&jobs[s]))) return rc;
    jobs[s].member0 = M;
    M += jobs[s].nmembers;
  }
  double *params = (double *)calloc((size_t)SIPNET_GPU_NPARAMS * (size_t)M, sizeof(double)); /* [80][M] */
  int32_t *memberSite = (int32_t *)malloc((size_t)M * sizeof *memberSite);
  sipnet_gpu_site *views = (sipnet_gpu_site *)calloc((size_t)njobs, sizeof *views);
  {
    const char *msg = "";
    if ((rc = load_sites(jobs, njobs, M, &ctx, params, memberSite, views, &msg))) return die(rc, msg);
  }
  if (getenv("SIPNET_GPU_TRACE_INPUTS")) { /* FNV-1a over everything the loaders produced (stderr; tests compare thread counts) */
    uint64_t hsh = 1469598103934665603ull;
#define SIP_MIX(ptr, nbytes)                                                   \
    for (size_t i_ = 0; i_ < (size_t)(nbytes); ++i_) hsh = (hsh ^ ((const unsigned char *)(ptr))[i_]) * 1099511628211ull
    SIP_MIX(params, sizeof(double) * SIPNET_GPU_NPARAMS * (size_t)M);
    SIP_MIX(memberSite, sizeof(int32_t) * (size_t)M);
    for (int64_t s = 0; s < njobs; ++s) {
      const sip_site_data *d = &jobs[s].data;
      const double *cols[] = {d->time, d->length, d->tair, d->tsoil, d->par, d->precip, d->vpd, d->vpdSoil, d->vPress, d->wspd, d->gdd};
      SIP_MIX(&d->nsteps, sizeof d->nsteps);
      SIP_MIX(d->year, sizeof(int32_t) * (size_t)d->nsteps);
      SIP_MIX(d->day, sizeof(int32_t) * (size_t)d->nsteps);
      for (size_t c = 0; c < sizeof cols / sizeof cols[0]; ++c)
        if (cols[c]) SIP_MIX(cols[c], sizeof(double) * (size_t)d->nsteps);
      SIP_MIX(&d->nevents, sizeof d->nevents);
      for (int64_t e = 0; e < d->nevents; ++e) { /* field by field: the struct has padding */
        SIP_MIX(&d->events[e].year, sizeof d->events[e].year);
        SIP_MIX(&d->events[e].day, sizeof d->events[e].day);
        SIP_MIX(&d->events[e].type, sizeof d->events[e].type);
        SIP_MIX(&d->events[e].method, sizeof d->events[e].method);
        SIP_MIX(d->events[e].p, sizeof d->events[e].p);
      }
    }
#undef SIP_MIX
    fprintf(stderr, "[TRACE  ] inputs: %lld site(s), %lld member(s), checksum %016llx\n", (long long)njobs, (long long)M,
            (unsigned long long)hsh);
  }
  for (int64_t s = 0; s < njobs; ++s) {
    const site_job *job = &jobs[s];
    if (job->data.nsteps > Tmax) Tmax = job->data.nsteps;
    if (job->data.nevents > maxEvents) maxEvents = job->data.nevents;
    if (job->data.nsteps > 0) {
      const int64_t years = (int64_t)job->data.year[job->data.nsteps - 1] - (int64_t)job->data.year[0] + 1;
      if (years > maxYears) maxYears = years;
    }
  }

  /* ---- the run: setupModel() + the updateState() loop, on the device ---- */
  sipnet_gpu_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = SIPNET_GPU_ABI_VERSION;
  cfg.device = 0;
  cfg.flags = ctx.flags;
  cfg.nsites = njobs;
  cfg.sites = views;
  cfg.nmembers = M;
  cfg.member_site = memberSite;
  cfg.params = params;
  cfg.params_ld = M;
  cfg.outputs = ((ctx.doMainOutput || ctx.doSingleOutputs) ? SIPNET_GPU_OUT_FULL : 0) | (ctx.flags.events ? SIPNET_GPU_OUT_EVENTS : 0) |
                ((ctx.debugLogPrefix[0] || ctx.restartOut[0]) ? SIPNET_GPU_OUT_DEBUG : 0);
  /* checkpoints carry the reference's 250-slot mean-NPP ring slot for slot (restart.c:799-806) */
  cfg.ring_slots = useRestart ? SIPNET_GPU_RING_SLOTS_REFERENCE : 0;
  cfg.math = ctx.validationMath ? SIPNET_GPU_MATH_VALIDATION : SIPNET_GPU_MATH_FAST;
  /* events.out rows per member: the file's events (a harvest also logs nothing extra; leaf-on/off rows are computed)
   * plus the computed rows -- leaf on, leaf off, plant death / re-emergence -- budgeted per YEAR of the record,
   * whatever the step length */
  cfg.max_event_records = ctx.flags.events ? (int32_t)(2 * maxEvents + 6 * (maxYears + 1) + 16) : 0;
  /* many members: every visible GPU (or --devices N) through the multi-GPU entry points; the single-member
   * features (restart, debug log) stay on one handle */
  sipnet_gpu_handle *h = NULL;
  sipnet_gpu_multi *mh = NULL;
  if (many) {
    if ((rc = sipnet_gpu_multi_init(&cfg, ctx.devices, NULL, &mh))) return die(rc, sipnet_gpu_last_error());
    if (!ctx.quiet) printf("[INFO   ] %lld member(s) of %lld site(s) on %d GPU(s)\n", (long long)M, (long long)njobs,
                           (int)sipnet_gpu_multi_ndevices(mh));
  } else if ((rc = sipnet_gpu_init(&cfg, &h))) {
    return die(rc, sipnet_gpu_last_error());
  }
#define SIP_GATHER(what, dst, bytes) (mh ? sipnet_gpu_multi_gather(mh, what, dst, bytes) : sipnet_gpu_gather(h, what, dst, bytes))
  const int64_t T = Tmax;
  const sip_site_data *site0 = &jobs[0].data; /* the single-member features below have exactly one site */
  long long processedBefore = 0; /* meta_info.processed_steps keeps counting across segments (restart.c:160, 905) */
  double totGppBefore = 0.0;     /* trackers.totGpp carried in from a checkpoint (GPP_cum single output) */
  if (ctx.restartIn[0]) { /* restartLoadCheckpoint() after setupModel()+setupEvents(), sipnet.c:1963-1967 */
    sip_restart *rs = (sip_restart *)malloc(sizeof *rs);
    double state[SIPNET_GPU_NSTATE], ringV[SIP_RESTART_RING], ringW[SIP_RESTART_RING];
    if ((rc = sip_read_restart(ctx.restartIn, rs))) return die(rc, sip_host_error());
    if ((rc = sip_check_restart(ctx.restartIn, rs, &ctx, site0))) return die(rc, sip_host_error());
    processedBefore = rs->processedSteps;
    totGppBefore = rs->trackers[20]; /* trackers.totGpp */
    sip_restart_to_state(rs, state, 1, ringV, ringW, 1);
    if ((rc = sipnet_gpu_set_state(h, state, 1, ringV, ringW, 1, 0))) return die(rc, sipnet_gpu_last_error());
    free(rs);
  }
  if ((rc = mh ? sipnet_gpu_multi_run(mh, 0, T) : sipnet_gpu_run(h, 0, T))) return die(rc, sipnet_gpu_last_error());

  uint32_t *status = (uint32_t *)malloc((size_t)M * sizeof *status);
  if ((rc = SIP_GATHER(SIPNET_GPU_GATHER_STATUS, status, (size_t)M * sizeof *status)))
    return die(rc, sipnet_gpu_last_error());
  for (int64_t m = 0; m < M; ++m) {
    if (status[m] & SIPNET_GPU_ST_BAD_ALLOCATION) /* ensureAllocation(), sipnet.c:1117-1122 */
      return die(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE,
                 "NPP allocation params must be less than one individually and add to less than one");
    if (status[m] & SIPNET_GPU_ST_RING_OVERFLOW) /* sipnet.c:1562-1569 */
      return die(SIPNET_GPU_ERR_INTERNAL, "Error while trying to add value to NPP mean tracker");
  }

  /* ---- writers: outputState() per step, events.out rows ---- */
  if (ctx.doMainOutput || ctx.doSingleOutputs) {
    const size_t n = (size_t)SIPNET_GPU_NOUT * (size_t)T * (size_t)M;
    double *buf = (double *)malloc(n * sizeof(double));
    if (!buf) return die(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure");
    if ((rc = SIP_GATHER(SIPNET_GPU_GATHER_FULL, buf, n * sizeof(double)))) return die(rc, sipnet_gpu_last_error());
    for (int64_t s = 0; s < njobs; ++s) {
      const site_job *job = &jobs[s];
      for (int64_t k = 0; k < job->nmembers; ++k) {
        const int64_t m = job->member0 + k;
        if (ctx.doSingleOutputs) { /* setupOutputItems() + writeOutputItemValues(), sipnet.c:1993-1998, outputItems.c:126-148 */
          static const char *const kItems[4] = {"NEE", "NEE_cum", "GPP", "GPP_cum"};
          static const int kCols[4] = {SIPNET_O_nee, SIPNET_O_cumNEE, SIPNET_O_gpp, -1};
          for (int it = 0; it < 4; ++it) {
            char name[SIP_NAME_MAX + 48];
            if (job->memberList[0])
              snprintf(name, sizeof name, "%s.%s.%lld", job->prefix, kItems[it], (long long)k);
            else
              snprintf(name, sizeof name, "%s.%s", job->prefix, kItems[it]);
            FILE *sf = fopen(name, "w");
            if (!sf) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open a single-variable output file");
            double totGpp = totGppBefore; /* trackers.totGpp += trackers.gpp, the reference's own accumulation order */
            for (int64_t t = 0; t < job->data.nsteps; ++t) {
              double v;
              if (kCols[it] >= 0) {
                v = buf[((size_t)kCols[it] * (size_t)T + (size_t)t) * (size_t)M + (size_t)m];
              } else {
                totGpp += buf[((size_t)SIPNET_O_gpp * (size_t)T + (size_t)t) * (size_t)M + (size_t)m];
                v = totGpp;
              }
              fprintf(sf, "%f ", v);
            }
            fprintf(sf, "\n");
            fclose(sf);
          }
        }
      }
    }
    if (ctx.doMainOutput && out) { /* one member, the file the reference opens up front */
      if (ctx.printHeader) sip_write_header(out);
      for (int64_t t = 0; t < jobs[0].data.nsteps; ++t)
        sip_write_state_row(out, jobs[0].data.year[t], jobs[0].data.day[t], jobs[0].data.time[t], buf + (size_t)t * (size_t)M,
                            (int64_t)T * M);
      fclose(out);
    } else if (ctx.doMainOutput) {
      /* every member's <prefix>.out[.k], formatted on all host cores (sip_write_state_files) */
      char *paths = (char *)malloc((size_t)M * SIP_STATE_PATH_MAX);
      int64_t *nsteps = (int64_t *)malloc((size_t)M * sizeof *nsteps);
      const int32_t **years = (const int32_t **)malloc((size_t)M * sizeof *years);
      const int32_t **days = (const int32_t **)malloc((size_t)M * sizeof *days);
      const double **times = (const double **)malloc((size_t)M * sizeof *times);
      if (!paths || !nsteps || !years || !days || !times) return die(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure");
      for (int64_t s = 0; s < njobs; ++s) {
        const site_job *job = &jobs[s];
        for (int64_t k = 0; k < job->nmembers; ++k) {
          const int64_t m = job->member0 + k;
          char *name = paths + (size_t)m * SIP_STATE_PATH_MAX;
          if (job->memberList[0])
            snprintf(name, SIP_STATE_PATH_MAX, "%s.out.%lld", job->prefix, (long long)k);
          else
            snprintf(name, SIP_STATE_PATH_MAX, "%s.out", job->prefix);
          nsteps[m] = job->data.nsteps;
          years[m] = job->data.year;
          days[m] = job->data.day;
          times[m] = job->data.time;
        }
      }
      if ((rc = sip_write_state_files(paths, M, nsteps, years, days, times, T, buf, ctx.printHeader, 0)))
        return die(rc, sip_host_error());
      free(paths);
      free(nsteps);
      free(years);
      free(days);
      free(times);
    }
    free(buf);
  }
  if (ctx.debugLogPrefix[0]) { /* outputDebugState() per step, debug_log.c:277-312 */
    const size_t n = (size_t)SIPNET_GPU_NDEBUG * (size_t)T;
    double *dbg = (double *)malloc(n * sizeof(double));
    if (!dbg) return die(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure");
    if ((rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_DEBUG, dbg, n * sizeof(double)))) return die(rc, sipnet_gpu_last_error());
    char name[SIP_NAME_MAX + 32];
    FILE *f[3];
    const char *suffix[3] = {"_envi.log", "_fluxes.log", "_trackers.log"};
    for (int k = 0; k < 3; ++k) {
      snprintf(name, sizeof name, "%s%s", ctx.debugLogPrefix, suffix[k]);
      f[k] = fopen(name, "w");
      if (!f[k]) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open debug log file");
    }
    if (ctx.printHeader) sip_write_debug_headers(f[0], f[1], f[2]);
    for (int64_t t = 0; t < T; ++t)
      sip_write_debug_rows(f[0], f[1], f[2], site0->year[t], site0->day[t], site0->time[t], dbg + t, T);
    for (int k = 0; k < 3; ++k) fclose(f[k]);
    free(dbg);
  }
  if (ctx.restartOut[0]) { /* restartWriteCheckpoint(), restart.c:908-961 */
    if (T <= 0) return die(SIPNET_GPU_ERR_BAD_RESTART, "Cannot write restart checkpoint: no timestep processed");
    sip_restart *rs = (sip_restart *)malloc(sizeof *rs);
    double state[SIPNET_GPU_NSTATE], ringV[SIP_RESTART_RING], ringW[SIP_RESTART_RING];
    const size_t n = (size_t)SIPNET_GPU_NDEBUG * (size_t)T;
    double *dbg = (double *)malloc(n * sizeof(double));
    if (!rs || !dbg) return die(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure");
    if ((rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_STATE, state, sizeof state)) ||
        (rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_RING_VALUES, ringV, sizeof ringV)) ||
        (rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_RING_WEIGHTS, ringW, sizeof ringW)) ||
        (rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_DEBUG, dbg, n * sizeof(double))))
      return die(rc, sipnet_gpu_last_error());
    sip_restart_from_device(rs, &ctx, site0, processedBefore + (long long)T, (long long)time(NULL), state, 1, dbg + (T - 1), T,
                            ringV, ringW, 1);
    if ((rc = sip_check_restart_boundary_for_write(ctx.restartOut, rs, ctx.quiet))) return die(rc, sip_host_error());
    if ((rc = sip_write_restart(ctx.restartOut, rs))) return die(rc, sip_host_error());
    free(dbg);
    free(rs);
  }
  if (ctx.flags.events) {
    const size_t nrec = (size_t)M * (size_t)cfg.max_event_records;
    sipnet_gpu_event_record *recs = (sipnet_gpu_event_record *)malloc((nrec ? nrec : 1) * sizeof *recs);
    int32_t *counts = (int32_t *)malloc((size_t)M * sizeof *counts);
    if ((rc = SIP_GATHER(SIPNET_GPU_GATHER_EVENT_COUNTS, counts, (size_t)M * sizeof *counts)))
      return die(rc, sipnet_gpu_last_error());
    if (nrec && (rc = SIP_GATHER(SIPNET_GPU_GATHER_EVENT_RECORDS, recs, nrec * sizeof *recs)))
      return die(rc, sipnet_gpu_last_error());
    for (int64_t s = 0; s < njobs; ++s) {
      const site_job *job = &jobs[s];
      for (int64_t k = 0; k < job->nmembers; ++k) {
        const int64_t m = job->member0 + k;
        char name[SIP_NAME_MAX + 32];
        if (job->memberList[0])
          snprintf(name, sizeof name, "%s.out.%lld", job->eventsPrefix, (long long)k);
        else
          snprintf(name, sizeof name, "%s.out", job->eventsPrefix);
        FILE *eo = fopen(name, "w");
        if (!eo) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open events output file");
        if (ctx.printHeader) sip_write_events_header(eo);
        if (counts[m] > cfg.max_event_records) return die(SIPNET_GPU_ERR_INTERNAL, "event record buffer too small");
        for (int32_t r = 0; r < counts[m]; ++r) {
          const sipnet_gpu_event_record *rec = &recs[(size_t)m * (size_t)cfg.max_event_records + (size_t)r];
          if ((rc = sip_write_event_row(eo, job->data.year[rec->step], job->data.day[rec->step], rec)))
            return die(rc, sip_host_error());
        }
        fclose(eo);
      }
    }
    free(recs);
    free(counts);
  }
  if (mh) sipnet_gpu_multi_destroy(mh);
  else sipnet_gpu_destroy(h);
  for (int64_t s = 0; s < njobs; ++s) {
    sip_site_free(&jobs[s].data);
    for (int64_t k = 0; k < jobs[s].nmembers; ++k) free(jobs[s].paramFiles[k]);
    free(jobs[s].paramFiles);
  }
  free(jobs);
  free(views);
  free(memberSite);
  free(params);
  free(status);
  return 0;
}
