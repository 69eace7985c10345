/*
 * sipnet_gpu_main.c -- `sipnet_gpu`: drop-in for the reference's `sipnet` driver.
 *
 * Same command line, same sipnet.in / .param / .clim / events.in inputs, same
 * <prefix>.out / events.out / <prefix>.config outputs and exit codes as
 * reference src/sipnet/frontend.c:130-253 -- but the body of runModelOutput()
 * (reference src/sipnet/sipnet.c:1954-1990) runs on a B200 through the C ABI
 * include/sipnet_gpu.h.  Extension: --ensemble-params FILE runs one member per
 * listed .param file on the same site in one launch; member k's outputs go to
 * <prefix>.out.<k> / events.out.<k>.
 * --debug-log <prefix> writes the reference's three per-step debug logs from the device's validation dump.
 * --restart-in / --restart-out read and write the reference's checkpoint format (sip_restart.c), so a segmented
 * run can alternate between this binary and the reference's.
 * Not supported (outside the hot-path scope): --do-single-outputs (exit 8).
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sip_host.h"

static int die(int code, const char *what) {
  printf("[ERROR  ] %s\n", what);
  return code;
}

int main(int argc, char **argv) {
  sip_context ctx;
  sip_context_init(&ctx);
  int rc = sip_parse_cli(&ctx, argc, argv);
  if (rc) return die(rc, sip_host_error());
  if (ctx.helpOrVersion) return 0;
  if ((rc = sip_read_input_file(&ctx))) return die(rc, sip_host_error());
  if ((rc = sip_validate_context(&ctx))) return die(rc, sip_host_error());
  if (ctx.doSingleOutputs)
    return die(SIPNET_GPU_ERR_BAD_CLI,
               "single-variable outputs are not part of the GPU hot path; use the reference binary for those");
  const int useRestart = ctx.restartIn[0] || ctx.restartOut[0];
  if (useRestart && ctx.ensembleParamList[0])
    return die(SIPNET_GPU_ERR_BAD_CLI, "a restart checkpoint holds one member: not available with --ensemble-params");
  if (ctx.debugLogPrefix[0] && ctx.ensembleParamList[0])
    return die(SIPNET_GPU_ERR_BAD_CLI, "--debug-log is a single-member feature");
  if ((rc = sip_derive_file_names(&ctx))) return die(rc, sip_host_error());

  FILE *out = NULL;
  if (ctx.doMainOutput && !ctx.ensembleParamList[0]) {
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

  /* ---- inputs: initModel() + initEvents(), sipnet.c:2001-2009, events.c:427-433 ---- */
  char **paramFiles = NULL;
  int64_t M = 0;
  if (ctx.ensembleParamList[0]) {
    FILE *lf = fopen(ctx.ensembleParamList, "r");
    if (!lf) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open the --ensemble-params list");
    char line[1024];
    while (fgets(line, sizeof line, lf)) {
      line[strcspn(line, "\r\n")] = '\0';
      if (!line[0]) continue;
      paramFiles = (char **)realloc(paramFiles, (size_t)(M + 1) * sizeof *paramFiles);
      paramFiles[M++] = strdup(line);
    }
    fclose(lf);
    if (M == 0) return die(SIPNET_GPU_ERR_INPUT_FILE, "the --ensemble-params list is empty");
  } else {
    paramFiles = (char **)malloc(sizeof *paramFiles);
    paramFiles[0] = strdup(ctx.paramFile);
    M = 1;
  }
  double *params = (double *)calloc((size_t)SIPNET_GPU_NPARAMS * (size_t)M, sizeof(double)); /* [80][M] */
  for (int64_t m = 0; m < M; ++m) {
    double one[SIPNET_GPU_NPARAMS];
    if ((rc = sip_read_params(paramFiles[m], &ctx.flags, ctx.quiet || m > 0, one))) return die(rc, sip_host_error());
    for (int k = 0; k < SIPNET_GPU_NPARAMS; ++k) params[(size_t)k * (size_t)M + (size_t)m] = one[k];
  }
  sip_site_data site;
  if ((rc = sip_read_clim(ctx.climFile, ctx.flags.gdd, ctx.quiet, &site))) return die(rc, sip_host_error());
  if (ctx.flags.events) {
    double first[SIPNET_GPU_NPARAMS];
    for (int k = 0; k < SIPNET_GPU_NPARAMS; ++k) first[k] = params[(size_t)k * (size_t)M];
    if ((rc = sip_read_events(ctx.eventsInFile, &ctx.flags, first, ctx.quiet, &site))) return die(rc, sip_host_error());
  }

  /* ---- the run: setupModel() + the updateState() loop, on the device ---- */
  sipnet_gpu_site view;
  sip_site_view(&site, &view);
  sipnet_gpu_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = SIPNET_GPU_ABI_VERSION;
  cfg.device = 0;
  cfg.flags = ctx.flags;
  cfg.nsites = 1;
  cfg.sites = &view;
  cfg.nmembers = M;
  cfg.params = params;
  cfg.params_ld = M;
  cfg.outputs = (ctx.doMainOutput ? SIPNET_GPU_OUT_FULL : 0) | (ctx.flags.events ? SIPNET_GPU_OUT_EVENTS : 0) |
                ((ctx.debugLogPrefix[0] || ctx.restartOut[0]) ? SIPNET_GPU_OUT_DEBUG : 0);
  /* checkpoints carry the reference's 250-slot mean-NPP ring slot for slot (restart.c:799-806) */
  cfg.ring_slots = useRestart ? SIPNET_GPU_RING_SLOTS_REFERENCE : 0;
  cfg.math = ctx.validationMath ? SIPNET_GPU_MATH_VALIDATION : SIPNET_GPU_MATH_FAST;
  cfg.max_event_records = ctx.flags.events ? (int32_t)(site.nevents + 4 * (site.nsteps / 300 + 8)) : 0;
  sipnet_gpu_handle *h = NULL;
  if ((rc = sipnet_gpu_init(&cfg, &h))) return die(rc, sipnet_gpu_last_error());
  const int64_t T = site.nsteps;
  long long processedBefore = 0; /* meta_info.processed_steps keeps counting across segments (restart.c:160, 905) */
  if (ctx.restartIn[0]) { /* restartLoadCheckpoint() after setupModel()+setupEvents(), sipnet.c:1963-1967 */
    sip_restart *rs = (sip_restart *)malloc(sizeof *rs);
    double state[SIPNET_GPU_NSTATE], ringV[SIP_RESTART_RING], ringW[SIP_RESTART_RING];
    if ((rc = sip_read_restart(ctx.restartIn, rs))) return die(rc, sip_host_error());
    if ((rc = sip_check_restart(ctx.restartIn, rs, &ctx, &site))) return die(rc, sip_host_error());
    processedBefore = rs->processedSteps;
    sip_restart_to_state(rs, state, 1, ringV, ringW, 1);
    if ((rc = sipnet_gpu_set_state(h, state, 1, ringV, ringW, 1, 0))) return die(rc, sipnet_gpu_last_error());
    free(rs);
  }
  if ((rc = sipnet_gpu_run(h, 0, T))) return die(rc, sipnet_gpu_last_error());

  uint32_t *status = (uint32_t *)malloc((size_t)M * sizeof *status);
  if ((rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_STATUS, status, (size_t)M * sizeof *status)))
    return die(rc, sipnet_gpu_last_error());
  for (int64_t m = 0; m < M; ++m) {
    if (status[m] & SIPNET_GPU_ST_BAD_ALLOCATION) /* ensureAllocation(), sipnet.c:1117-1122 */
      return die(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE,
                 "NPP allocation params must be less than one individually and add to less than one");
    if (status[m] & SIPNET_GPU_ST_RING_OVERFLOW) /* sipnet.c:1562-1569 */
      return die(SIPNET_GPU_ERR_INTERNAL, "Error while trying to add value to NPP mean tracker");
  }

  /* ---- writers: outputState() per step, events.out rows ---- */
  if (ctx.doMainOutput) {
    const size_t n = (size_t)SIPNET_GPU_NOUT * (size_t)T * (size_t)M;
    double *buf = (double *)malloc(n * sizeof(double));
    if (!buf) return die(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure");
    if ((rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_FULL, buf, n * sizeof(double)))) return die(rc, sipnet_gpu_last_error());
    for (int64_t m = 0; m < M; ++m) {
      FILE *o = out;
      if (!o) {
        char name[SIP_NAME_MAX + 32];
        snprintf(name, sizeof name, "%s.%lld", ctx.outFile, (long long)m);
        o = fopen(name, "w");
        if (!o) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open ensemble output file");
      }
      if (ctx.printHeader) sip_write_header(o);
      for (int64_t t = 0; t < T; ++t)
        sip_write_state_row(o, site.year[t], site.day[t], site.time[t], buf + (size_t)t * (size_t)M + (size_t)m,
                            (int64_t)T * M);
      fclose(o);
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
      sip_write_debug_rows(f[0], f[1], f[2], site.year[t], site.day[t], site.time[t], dbg + t, T);
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
    sip_restart_from_device(rs, &ctx, &site, processedBefore + (long long)T, (long long)time(NULL), state, 1, dbg + (T - 1), T, ringV, ringW, 1);
    if ((rc = sip_check_restart_boundary_for_write(ctx.restartOut, rs, ctx.quiet))) return die(rc, sip_host_error());
    if ((rc = sip_write_restart(ctx.restartOut, rs))) return die(rc, sip_host_error());
    free(dbg);
    free(rs);
  }
  if (ctx.flags.events) {
    const size_t nrec = (size_t)M * (size_t)cfg.max_event_records;
    sipnet_gpu_event_record *recs = (sipnet_gpu_event_record *)malloc((nrec ? nrec : 1) * sizeof *recs);
    int32_t *counts = (int32_t *)malloc((size_t)M * sizeof *counts);
    if ((rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_EVENT_COUNTS, counts, (size_t)M * sizeof *counts)))
      return die(rc, sipnet_gpu_last_error());
    if (nrec && (rc = sipnet_gpu_gather(h, SIPNET_GPU_GATHER_EVENT_RECORDS, recs, nrec * sizeof *recs)))
      return die(rc, sipnet_gpu_last_error());
    for (int64_t m = 0; m < M; ++m) {
      char name[SIP_NAME_MAX + 32];
      if (ctx.ensembleParamList[0])
        snprintf(name, sizeof name, "%s.%lld", ctx.eventsOutFile, (long long)m);
      else
        snprintf(name, sizeof name, "%s", ctx.eventsOutFile);
      FILE *eo = fopen(name, "w");
      if (!eo) return die(SIPNET_GPU_ERR_FILE_OPEN, "cannot open events output file");
      if (ctx.printHeader) sip_write_events_header(eo);
      if (counts[m] > cfg.max_event_records) return die(SIPNET_GPU_ERR_INTERNAL, "event record buffer too small");
      for (int32_t k = 0; k < counts[m]; ++k) {
        const sipnet_gpu_event_record *r = &recs[(size_t)m * (size_t)cfg.max_event_records + (size_t)k];
        if ((rc = sip_write_event_row(eo, site.year[r->step], site.day[r->step], r))) return die(rc, sip_host_error());
      }
      fclose(eo);
    }
    free(recs);
    free(counts);
  }
  sipnet_gpu_destroy(h);
  sip_site_free(&site);
  free(params);
  free(status);
  for (int64_t m = 0; m < M; ++m) free(paramFiles[m]);
  free(paramFiles);
  return 0;
}
