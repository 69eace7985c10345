/*
 * sip_host.h -- host side of the drop-in (plain C11).
 *
 * Restates the semantics of the reference's input readers and text writers so
 * that `sipnet_gpu -i sipnet.in` is a drop-in for `sipnet -i sipnet.in`:
 *
 *   configuration   sipnet.in + command line        reference src/sipnet/frontend.c:35-128,
 *                                                    src/sipnet/cli.c:144-233, src/common/context.c
 *   <prefix>.param  name/value parameter file       src/sipnet/sipnet.c:290-427, src/common/modelParams.c:136-230
 *   <prefix>.clim   12- or legacy 14-column forcing  src/sipnet/sipnet.c:128-277
 *   events.in       agronomic events                 src/sipnet/events.c:39-367
 *   <prefix>.out    per-step text rows               src/sipnet/sipnet.c:434-473
 *   events.out      applied / computed events        src/sipnet/events.c:369-418
 *   <prefix>.config final configuration dump         src/common/context.c:225-267
 *   restart files   --restart-in / --restart-out      src/sipnet/restart.c (sip_restart.c)
 *
 * No function here calls exit(): each returns 0 or the reference's exit code
 * (src/common/exitCodes.h:16-27) and leaves a message in sip_host_error().
 * Nothing here touches the GPU; the driver (sipnet_gpu_main.c) hands the parsed
 * inputs to the device through include/sipnet_gpu.h.
 */
#ifndef SIP_HOST_H
#define SIP_HOST_H

#include <stdint.h>
#include <stdio.h>

#include "../../include/sipnet_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SIP_NAME_MAX 256 /* CONTEXT_CHAR_MAXLEN, context.h:8 */

/* where a setting came from; higher wins (context.h:18-26) */
enum sip_source { SIP_SRC_DEFAULT = 0, SIP_SRC_INPUT_FILE = 1, SIP_SRC_COMMAND_LINE = 2, SIP_SRC_CALCULATED = 3 };

/* the reference's struct Context (context.h:42-92), without the hash map */
typedef struct sip_context {
  sipnet_gpu_flags flags; /* the 12 model flags */
  int32_t doMainOutput, doSingleOutputs, dumpConfig, printHeader, quiet;
  char paramFile[SIP_NAME_MAX], climFile[SIP_NAME_MAX], outFile[SIP_NAME_MAX], outConfigFile[SIP_NAME_MAX];
  char eventsPrefix[SIP_NAME_MAX], eventsInFile[SIP_NAME_MAX], eventsOutFile[SIP_NAME_MAX];
  char inputFile[SIP_NAME_MAX], restartIn[SIP_NAME_MAX], restartOut[SIP_NAME_MAX], debugLogPrefix[SIP_NAME_MAX];
  char filePrefix[SIP_NAME_MAX];
  /* provenance per setting, indexed like sip_setting_name() */
  int32_t source[32];
  /* extensions of this implementation (not in the reference) */
  char ensembleParamList[SIP_NAME_MAX]; /* --ensemble-params FILE: one .param path per line => one member each */
  int32_t validationMath;               /* --validation-math: run the general kernel */
  int32_t helpOrVersion;                /* 1 = --help printed, 2 = --version printed (caller exits 0) */
  char siteList[SIP_NAME_MAX];          /* --site-list FILE: one site per line, all sites in one launch (see usage) */
  int32_t devices;                      /* --devices N: GPUs for many-member launches (0 = all visible) */
} sip_context;

const char *sip_host_error(void);

/* ---- configuration ------------------------------------------------------------------------ */
void sip_context_init(sip_context *ctx);                               /* initContext(), context.c:26-67 */
int sip_parse_cli(sip_context *ctx, int argc, char **argv);           /* parseCommandLineArgs(), cli.c:144-233 */
int sip_read_input_file(sip_context *ctx);                             /* readInputFile(), frontend.c:35-128 */
int sip_validate_context(const sip_context *ctx);                      /* validateContext(), context.c:195-223 */
int sip_derive_file_names(sip_context *ctx);                           /* frontend.c:164-209 */
int sip_print_config(const sip_context *ctx, FILE *out, const char *timestamp); /* printConfig(), context.c:225-267 */

/* ---- inputs --------------------------------------------------------------------------------- */
/* One site's forcing as left by readClimData() plus its events as left by readEventData(). */
typedef struct sip_site_data {
  int64_t nsteps;
  int32_t *year, *day;
  double *time, *length, *tair, *tsoil, *par, *precip, *vpd, *vpdSoil, *vPress, *wspd, *gdd;
  int64_t nevents;
  sipnet_gpu_event *events;
} sip_site_data;

int sip_read_params(const char *path, const sipnet_gpu_flags *flags, int quiet, double out[SIPNET_GPU_NPARAMS]);
int sip_read_clim(const char *path, int gddFlag, int quiet, sip_site_data *site);
int sip_read_events(const char *path, const sipnet_gpu_flags *flags, const double params[SIPNET_GPU_NPARAMS],
                    int quiet, sip_site_data *site);
void sip_site_free(sip_site_data *site);
void sip_site_view(const sip_site_data *site, sipnet_gpu_site *view); /* borrow as the ABI's site struct */

/* ---- outputs ---------------------------------------------------------------------------------- */
void sip_write_header(FILE *out);                                                         /* outputHeader() */
void sip_write_state_row(FILE *out, int year, int day, double time, const double *out32, int64_t stride);
                                                                                          /* outputState(); out32[c*stride] */
/* The same row into memory, without printf for the usual values (byte-identical; anything near a rounding boundary,
 * huge or non-finite goes through snprintf), and the plain printf statements it must reproduce (tests). */
#define SIP_STATE_ROW_MAX 12288 /* 35 fields of at most ~320 characters ("%f" of DBL_MAX) */
size_t sip_format_state_row(char *dst, int year, int day, double time, const double *out32, int64_t stride);
size_t sip_format_state_row_printf(char *dst, int year, int day, double time, const double *out32, int64_t stride);
/* the rows of up to SIP_STATE_BLOCK_MAX members that are neighbours in the gathered [col][step][member] array */
#define SIP_STATE_BLOCK_MAX 8
int sip_write_state_block(FILE *const *files, int count, const int64_t *nsteps, const int32_t *const *year,
                          const int32_t *const *day, const double *const *time, const double *out32, int64_t colStride,
                          int64_t stepStride);
/* The main output files of ALL members of a launch from the gathered [col][T][M] array: member m's file is
 * paths + m * SIP_STATE_PATH_MAX (NUL-terminated), it gets nsteps[m] rows dated year[m][t] / day[m][t] / time[m][t].
 * Blocks of members are formatted on nthreads host threads (0: SIPNET_GPU_WRITER_THREADS or every online core). */
#define SIP_STATE_PATH_MAX (SIP_NAME_MAX + 32)
#define SIP_STATE_THREADS_MAX 64
int sip_write_state_files(const char *paths, int64_t M, const int64_t *nsteps, const int32_t *const *year,
                          const int32_t *const *day, const double *const *time, int64_t T, const double *out32,
                          int printHeader, int nthreads);
void sip_write_events_header(FILE *out);                                                  /* openEventOutFile() header */
int sip_write_event_row(FILE *out, int year, int day, const sipnet_gpu_event_record *rec); /* doWriteEventOut() */
const char *sip_event_type_name(int type);                                                /* eventTypeToString() */
/* --debug-log <prefix>: <prefix>_envi.log, _fluxes.log, _trackers.log (debug_log.c:196-312); dbg[k * stride] is
 * field k of SIPNET_GPU_GATHER_DEBUG for this member-step */
void sip_write_debug_headers(FILE *envi, FILE *fluxes, FILE *trackers);
void sip_write_debug_rows(FILE *envi, FILE *fluxes, FILE *trackers, int year, int day, double time, const double *dbg,
                          int64_t stride);

/* ---- restart checkpoints (sip_restart.c; reference src/sipnet/restart.c) --------------------------- */
#define SIP_MODEL_VERSION "2.1.0"             /* NUMERIC_VERSION, version.h:4 -- checkpoints of another version are rejected */
#define SIP_BUILD_INFO "2.1.0_(sipnet-b200)"  /* sanitizeBuildInfo(VERSION_STRING); a mismatch is only reported */
#define SIP_RESTART_NENVI 13
#define SIP_RESTART_NTRACKERS 33
#define SIP_RESTART_RING SIPNET_GPU_RING_SLOTS_REFERENCE

/* one member's checkpoint: the key/value payload of the reference's file (restart.c:150-308) */
typedef struct sip_restart {
  char modelVersion[32], buildInfo[96];
  long long checkpointUtcEpoch, processedSteps;
  sipnet_gpu_flags flags;
  int boundaryYear, boundaryDay; /* the last processed climate record */
  double boundaryTime, boundaryLength;
  int meanLength, meanStart, meanLast; /* MeanTracker cursors (runmean.h) */
  double meanTotWeight, meanSum;
  double envi[SIP_RESTART_NENVI];         /* struct Environment order */
  double trackers[SIP_RESTART_NTRACKERS]; /* struct TrackerVars order; [26] = lastYear */
  int didLeafGrowth, didLeafFall, phenLastYear, isAlive;
  double dTillMod, harvestFracRemoved, harvestFracTransferred;
  double values[SIP_RESTART_RING], weights[SIP_RESTART_RING];
} sip_restart;

int sip_write_restart(const char *path, const sip_restart *r);  /* writeRestartState(), restart.c:784-827 */
int sip_read_restart(const char *path, sip_restart *r);         /* readRestartState(), restart.c:593-757 */
/* the checks of restartLoadCheckpoint() (restart.c:963-983) against the run's flags and first climate record */
int sip_check_restart(const char *path, const sip_restart *r, const sip_context *ctx, const sip_site_data *site);
int sip_check_restart_boundary_for_write(const char *path, const sip_restart *r, int quiet);
/* checkpoint -> one member's column of sipnet_gpu_set_state() inputs (state[k * stride], ring[i * ringStride]) */
void sip_restart_to_state(const sip_restart *r, double *state, int64_t stride, double *ringV, double *ringW,
                          int64_t ringStride);
/* gathered device results of one member -> checkpoint: SIPNET_GPU_GATHER_STATE column, the LAST step's row of
 * SIPNET_GPU_GATHER_DEBUG (dbgLast[k * dbgStride]), SIPNET_GPU_GATHER_RING_VALUES / _WEIGHTS columns */
void sip_restart_from_device(sip_restart *r, const sip_context *ctx, const sip_site_data *site, long long processedSteps,
                             long long utcEpoch, const double *state, int64_t stride, const double *dbgLast,
                             int64_t dbgStride, const double *ringV, const double *ringW, int64_t ringStride);

#ifdef __cplusplus
}
#endif
#endif
