/*
 * sip_restart.c -- restart checkpoints, the reference's own text format, so that a run can be
 * segmented and handed between `sipnet_gpu` and the reference binary in either direction.
 *
 * Restates the behaviour of reference src/sipnet/restart.c:
 *   file layout and number formats       writeRestartState(), restart.c:784-827 (%.17g doubles, %d ints)
 *   key set                              initResetState(), restart.c:150-308
 *   line grammar, strict number parsing, readRestartState(), restart.c:593-740; parse*Strict(), :422-450
 *   duplicate / unknown / missing keys
 *   checks on load                       restartLoadCheckpoint(), restart.c:963-983 and the validate*() helpers
 * Every failure returns the reference's exit code (9 = EXIT_CODE_BAD_RESTART_PARAMETER, 6 = cannot open,
 * 5 = no climate) with the message in sip_host_error(); nothing here exits.
 *
 * The checkpoint holds one member.  Device-side the payload is the state rows + the mean-NPP ring of a handle
 * created with ring_slots = SIPNET_GPU_RING_SLOTS_REFERENCE (include/sipnet_gpu.h: sipnet_gpu_set_state,
 * SIPNET_GPU_GATHER_STATE / RING_VALUES / RING_WEIGHTS) and, for the per-step tracker values the file also
 * carries, the last step's row of the validation dump (SIPNET_GPU_GATHER_DEBUG).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "sip_host.h"

int sip_fail(int code, const char *fmt, ...);
void sip_info(int quiet, const char *fmt, ...);

#define RESTART_EPS 1e-8 /* RESTART_FLOAT_EPSILON, restart.c:19 */
#define BAD SIPNET_GPU_ERR_BAD_RESTART

static const char *const kEnviKeys[SIP_RESTART_NENVI] = {
    "plantWoodC", "plantLeafC", "soilC",    "soilWater", "litterC",       "snow",                 "coarseRootC",
    "fineRootC",  "minN",       "soilOrgN", "litterN",   "plantStorageN", "plantCAccountingDelta"};
static const char *const kTrackerKeys[SIP_RESTART_NTRACKERS] = {
    "gpp",       "rtot",       "ra",         "rh",        "rRoot",        "rSoil",     "rAboveground",
    "npp",       "nee",        "woodCreation", "gdd",     "evapotranspiration", "soilWetnessFrac", "yearlyGpp",
    "yearlyRtot", "yearlyRa",  "yearlyRh",   "yearlyNpp", "yearlyNee",    "yearlyLitter", "totGpp",
    "totRtot",   "totRa",      "totRh",      "totNpp",    "totNee",       "lastYear",  "methane",
    "n2o",       "nLeaching",  "nFixation",  "nUptake",   "meanNPP"};
#define TRACKER_LASTYEAR 26
static const char *const kFlagKeys[12] = {"events",     "gdd",        "growthResp",    "leafWater", "litterPool", "snow",
                                          "soilPhenol", "waterHResp", "nitrogenCycle", "anaerobic", "flooding",
                                          "carbonSaturation"};
/* schema_layout.*: the byte sizes of the reference's structs (restart.c:46-55) */
static const char *const kSchemaKeys[5] = {"envi_size", "trackers_size", "phenology_trackers_size",
                                           "survival_trackers_size", "event_trackers_size"};
static const int kSchemaValues[5] = {8 * SIP_RESTART_NENVI, 8 * SIP_RESTART_NTRACKERS, 4 * 3, 4 * 1, 8 * 3};

static int32_t *flag_slot(sipnet_gpu_flags *f, int k) { return &((int32_t *)f)[k]; }

/* ---- writer ---------------------------------------------------------------------------------- */
int sip_write_restart(const char *path, const sip_restart *r) {
  FILE *out = fopen(path, "w");
  if (!out) return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "Error opening %s for w", path);
  fprintf(out, "SIPNET_RESTART\n");
  fprintf(out, "meta_info.model_version %s\n", r->modelVersion);
  fprintf(out, "meta_info.build_info %s\n", r->buildInfo);
  fprintf(out, "meta_info.checkpoint_utc_epoch %lld\n", r->checkpointUtcEpoch);
  fprintf(out, "meta_info.processed_steps %lld\n", r->processedSteps);
  for (int k = 0; k < 5; ++k) fprintf(out, "schema_layout.%s %d\n", kSchemaKeys[k], kSchemaValues[k]);
  fprintf(out, "\n");
  for (int k = 0; k < 12; ++k) fprintf(out, "flags.%s %d\n", kFlagKeys[k], *flag_slot((sipnet_gpu_flags *)&r->flags, k));
  fprintf(out, "\n");
  fprintf(out, "boundary.year %d\nboundary.day %d\nboundary.time %.17g\nboundary.length %.17g\n", r->boundaryYear,
          r->boundaryDay, r->boundaryTime, r->boundaryLength);
  fprintf(out, "\n");
  for (int k = 0; k < SIP_RESTART_NENVI; ++k) fprintf(out, "envi.%s %.17g\n", kEnviKeys[k], r->envi[k]);
  fprintf(out, "\n");
  for (int k = 0; k < SIP_RESTART_NTRACKERS; ++k) {
    if (k == TRACKER_LASTYEAR)
      fprintf(out, "trackers.%s %d\n", kTrackerKeys[k], (int)r->trackers[k]);
    else
      fprintf(out, "trackers.%s %.17g\n", kTrackerKeys[k], r->trackers[k]);
  }
  fprintf(out, "\n");
  fprintf(out, "phenology.didLeafGrowth %d\nphenology.didLeafFall %d\nphenology.lastYear %d\n", r->didLeafGrowth,
          r->didLeafFall, r->phenLastYear);
  fprintf(out, "\n");
  fprintf(out, "survival.isAlive %d\n", r->isAlive);
  fprintf(out, "\n");
  fprintf(out, "event_trackers.d_till_mod %.17g\nevent_trackers.harvestFracRemoved %.17g\n"
               "event_trackers.harvestFracTransferred %.17g\n",
          r->dTillMod, r->harvestFracRemoved, r->harvestFracTransferred);
  fprintf(out, "\n");
  fprintf(out, "mean.npp.length %d\nmean.npp.totWeight %.17g\nmean.npp.start %d\nmean.npp.last %d\nmean.npp.sum %.17g\n",
          r->meanLength, r->meanTotWeight, r->meanStart, r->meanLast, r->meanSum);
  fprintf(out, "\n");
  for (int i = 0; i < r->meanLength; ++i) fprintf(out, "mean.npp.values.%d %.17g\n", i, r->values[i]);
  fprintf(out, "\n");
  for (int i = 0; i < r->meanLength; ++i) fprintf(out, "mean.npp.weights.%d %.17g\n", i, r->weights[i]);
  fprintf(out, "\n");
  fprintf(out, "end_restart 1\n");
  if (fclose(out) != 0) return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "Error writing %s", path);
  return 0;
}

/* ---- reader ---------------------------------------------------------------------------------- */
typedef struct {
  const char *path;
  int rc; /* first failure */
} parse_ctx;

static int bad_value(parse_ctx *pc, const char *key, const char *value) {
  if (!pc->rc) pc->rc = sip_fail(BAD, "Restart parse error in %s: invalid value '%s' for key '%s'", pc->path, value, key);
  return 0;
}
static long long parse_ll(parse_ctx *pc, const char *key, const char *value) { /* parseLongLongStrict */
  char *end = NULL;
  errno = 0;
  const long long v = strtoll(value, &end, 10);
  if (end == value || *end != '\0' || errno == ERANGE) return bad_value(pc, key, value);
  return v;
}
static int parse_int(parse_ctx *pc, const char *key, const char *value) { /* parseIntStrict */
  const long long v = parse_ll(pc, key, value);
  if (v < INT_MIN || v > INT_MAX) return bad_value(pc, key, value);
  return (int)v;
}
static double parse_double(parse_ctx *pc, const char *key, const char *value) { /* parseDoubleStrict */
  char *end = NULL;
  const double v = strtod(value, &end);
  if (end == value || *end != '\0' || !isfinite(v)) return bad_value(pc, key, value);
  return v;
}

/* bookkeeping of which keys have been seen: index spaces below */
enum { SEEN_META = 0, SEEN_SCHEMA = 4, SEEN_FLAGS = 9, SEEN_BOUNDARY = 21, SEEN_NPP = 25, SEEN_ENVI = 30,
       SEEN_TRACKERS = 43, SEEN_PHEN = 76, SEEN_SURVIVAL = 79, SEEN_EVENT = 80, SEEN_END = 83, SEEN_COUNT = 84 };

static int mark(parse_ctx *pc, unsigned char *seen, int idx, const char *key) {
  if (seen[idx]) {
    if (!pc->rc) pc->rc = sip_fail(BAD, "Restart parse error in %s: duplicate key '%s'", pc->path, key);
    return 0;
  }
  seen[idx] = 1;
  return 1;
}

static int find_key(const char *const *keys, int n, const char *name) {
  for (int k = 0; k < n; ++k)
    if (strcmp(keys[k], name) == 0) return k;
  return -1;
}

int sip_read_restart(const char *path, sip_restart *r) {
  FILE *in = fopen(path, "r");
  if (!in) return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "Error opening %s for r", path);
  memset(r, 0, sizeof *r);
  parse_ctx pc = {path, 0};
  unsigned char seen[SEEN_COUNT] = {0};
  unsigned char seenV[SIP_RESTART_RING] = {0}, seenW[SIP_RESTART_RING] = {0};
  int meanLengthRead = SIP_RESTART_RING;

  char first[256];
  if (!fgets(first, sizeof first, in)) {
    fclose(in);
    return sip_fail(BAD, "Restart parse error in %s: missing header line", path);
  }
  size_t len = strlen(first);
  if (len > 0 && first[len - 1] != '\n' && !feof(in)) {
    fclose(in);
    return sip_fail(BAD, "Restart parse error in %s: line too long or truncated", path);
  }
  first[strcspn(first, "\r\n")] = '\0';
  if (strcmp(first, "SIPNET_RESTART") != 0) {
    fclose(in);
    return sip_fail(BAD, "Restart file %s has invalid magic header", path);
  }

  static const char *const metaKeys[4] = {"model_version", "build_info", "checkpoint_utc_epoch", "processed_steps"};
  static const char *const boundaryKeys[4] = {"year", "day", "time", "length"};
  static const char *const nppKeys[5] = {"length", "totWeight", "start", "last", "sum"};
  static const char *const phenKeys[3] = {"didLeafGrowth", "didLeafFall", "lastYear"};
  static const char *const eventKeys[3] = {"d_till_mod", "harvestFracRemoved", "harvestFracTransferred"};

  char line[4096], key[128], value[2048], extra[32];
  while (!pc.rc && fgets(line, sizeof line, in)) {
    len = strlen(line);
    if (len > 0 && line[len - 1] != '\n' && !feof(in)) {
      pc.rc = sip_fail(BAD, "Restart parse error in %s: line too long or truncated", path);
      break;
    }
    if (seen[SEEN_END]) {
      sip_info(0, "Ignoring extra lines after end_restart in %s\n", path);
      break;
    }
    const int n = sscanf(line, " %127s %2047s %31s", key, value, extra);
    if (n <= 0) continue;
    if (n != 2) {
      pc.rc = sip_fail(BAD, "Restart parse error in %s: line must contain exactly '<key> <value>'", path);
      break;
    }
    int k;
    const char *dot;
    /* the reference matches the group prefix first, then the full key inside the group; a key with a known
       prefix but an unknown tail falls through to "unknown key" (restart.c:639-720) */
    if (strncmp(key, "meta_info.", 10) == 0 && (k = find_key(metaKeys, 4, key + 10)) >= 0) {
      if (!mark(&pc, seen, SEEN_META + k, key)) break;
      if (k == 0) {
        snprintf(r->modelVersion, sizeof r->modelVersion, "%.31s", value);
      } else if (k == 1) {
        snprintf(r->buildInfo, sizeof r->buildInfo, "%.95s", value);
      } else if (k == 2) {
        r->checkpointUtcEpoch = parse_ll(&pc, key, value);
      } else {
        r->processedSteps = parse_ll(&pc, key, value);
      }
    } else if (strncmp(key, "schema_layout.", 14) == 0 && (k = find_key(kSchemaKeys, 5, key + 14)) >= 0) {
      if (!mark(&pc, seen, SEEN_SCHEMA + k, key)) break;
      const long long found = parse_int(&pc, key, value);
      if (!pc.rc && found != kSchemaValues[k])
        pc.rc = sip_fail(BAD, "Restart schema layout mismatch in %s: key=%s found=%lld expected=%lld", path, key, found,
                         (long long)kSchemaValues[k]);
    } else if (strncmp(key, "flags.", 6) == 0 && (k = find_key(kFlagKeys, 12, key + 6)) >= 0) {
      if (!mark(&pc, seen, SEEN_FLAGS + k, key)) break;
      *flag_slot(&r->flags, k) = parse_int(&pc, key, value);
    } else if (strncmp(key, "boundary.", 9) == 0 && (k = find_key(boundaryKeys, 4, key + 9)) >= 0) {
      if (!mark(&pc, seen, SEEN_BOUNDARY + k, key)) break;
      if (k == 0) r->boundaryYear = parse_int(&pc, key, value);
      if (k == 1) r->boundaryDay = parse_int(&pc, key, value);
      if (k == 2) r->boundaryTime = parse_double(&pc, key, value);
      if (k == 3) r->boundaryLength = parse_double(&pc, key, value);
    } else if (strncmp(key, "mean.npp.", 9) == 0 && (k = find_key(nppKeys, 5, key + 9)) >= 0) {
      if (!mark(&pc, seen, SEEN_NPP + k, key)) break;
      if (k == 0) meanLengthRead = parse_int(&pc, key, value);
      if (k == 1) r->meanTotWeight = parse_double(&pc, key, value);
      if (k == 2) r->meanStart = parse_int(&pc, key, value);
      if (k == 3) r->meanLast = parse_int(&pc, key, value);
      if (k == 4) r->meanSum = parse_double(&pc, key, value);
    } else if (strncmp(key, "envi.", 5) == 0 && (k = find_key(kEnviKeys, SIP_RESTART_NENVI, key + 5)) >= 0) {
      if (!mark(&pc, seen, SEEN_ENVI + k, key)) break;
      r->envi[k] = parse_double(&pc, key, value);
    } else if (strncmp(key, "trackers.", 9) == 0 && (k = find_key(kTrackerKeys, SIP_RESTART_NTRACKERS, key + 9)) >= 0) {
      if (!mark(&pc, seen, SEEN_TRACKERS + k, key)) break;
      r->trackers[k] = (k == TRACKER_LASTYEAR) ? (double)parse_int(&pc, key, value) : parse_double(&pc, key, value);
    } else if (strncmp(key, "phenology.", 10) == 0 && (k = find_key(phenKeys, 3, key + 10)) >= 0) {
      if (!mark(&pc, seen, SEEN_PHEN + k, key)) break;
      const int v = parse_int(&pc, key, value);
      if (k == 0) r->didLeafGrowth = v;
      if (k == 1) r->didLeafFall = v;
      if (k == 2) r->phenLastYear = v;
    } else if (strcmp(key, "survival.isAlive") == 0) {
      if (!mark(&pc, seen, SEEN_SURVIVAL, key)) break;
      r->isAlive = parse_int(&pc, key, value);
    } else if (strncmp(key, "event_trackers.", 15) == 0 && (k = find_key(eventKeys, 3, key + 15)) >= 0) {
      if (!mark(&pc, seen, SEEN_EVENT + k, key)) break;
      const double v = parse_double(&pc, key, value);
      if (k == 0) r->dTillMod = v;
      if (k == 1) r->harvestFracRemoved = v;
      if (k == 2) r->harvestFracTransferred = v;
    } else if (strcmp(key, "end_restart") == 0) {
      if (!mark(&pc, seen, SEEN_END, key)) break;
      (void)parse_int(&pc, key, value);
    } else if (strncmp(key, "mean.npp.values.", 16) == 0 || strncmp(key, "mean.npp.weights.", 17) == 0) {
      const int isV = key[9] == 'v';
      dot = key + (isV ? 16 : 17);
      const int idx = parse_int(&pc, key, dot);
      if (pc.rc) break;
      if (idx < 0 || idx >= SIP_RESTART_RING) {
        pc.rc = sip_fail(BAD, "Restart parse error in %s: mean.npp.%s index out of range (%s)", path,
                         isV ? "values" : "weights", key);
        break;
      }
      if (!mark(&pc, isV ? seenV : seenW, idx, key)) break;
      (isV ? r->values : r->weights)[idx] = parse_double(&pc, key, value);
    } else {
      pc.rc = sip_fail(BAD, "Restart parse error in %s: unknown key '%s'", path, key);
    }
  }
  fclose(in);
  if (pc.rc) return pc.rc;

  /* the checkpoint may not resize the ring (restart.c:725-731) */
  if (meanLengthRead != SIP_RESTART_RING)
    return sip_fail(BAD, "Restart schema mismatch in %s: mean.npp.length (%d) does not match the compiled model length (%d)",
                    path, meanLengthRead, SIP_RESTART_RING);
  r->meanLength = SIP_RESTART_RING;

  /* required keys, in the reference's order of complaint (restart.c:733-744) */
  struct {
    int base, n;
    const char *prefix;
    const char *const *keys;
  } groups[] = {{SEEN_META, 4, "meta_info.", metaKeys},        {SEEN_SCHEMA, 5, "schema_layout.", kSchemaKeys},
                {SEEN_FLAGS, 12, "flags.", kFlagKeys},          {SEEN_BOUNDARY, 4, "boundary.", boundaryKeys},
                {SEEN_NPP, 5, "mean.npp.", nppKeys},            {SEEN_ENVI, SIP_RESTART_NENVI, "envi.", kEnviKeys},
                {SEEN_TRACKERS, SIP_RESTART_NTRACKERS, "trackers.", kTrackerKeys},
                {SEEN_PHEN, 3, "phenology.", phenKeys}};
  for (size_t g = 0; g < sizeof groups / sizeof groups[0]; ++g)
    for (int k = 0; k < groups[g].n; ++k)
      if (!seen[groups[g].base + k])
        return sip_fail(BAD, "Restart parse error in %s: missing required key (%s%s)", path, groups[g].prefix,
                        groups[g].keys[k]);
  if (!seen[SEEN_SURVIVAL]) return sip_fail(BAD, "Restart parse error in %s: missing required key (survival.isAlive)", path);
  for (int k = 0; k < 3; ++k)
    if (!seen[SEEN_EVENT + k])
      return sip_fail(BAD, "Restart parse error in %s: missing required key (event_trackers.%s)", path, eventKeys[k]);
  if (!seen[SEEN_END]) return sip_fail(BAD, "Restart parse error in %s: missing required key (end_restart)", path);
  for (int i = 0; i < SIP_RESTART_RING; ++i)
    if (!seenV[i]) return sip_fail(BAD, "Restart parse error in %s: mean.npp.values array is incomplete", path);
  for (int i = 0; i < SIP_RESTART_RING; ++i)
    if (!seenW[i]) return sip_fail(BAD, "Restart parse error in %s: mean.npp.weights array is incomplete", path);
  return 0;
}

/* ---- checks on load (restartLoadCheckpoint, restart.c:963-983) --------------------------------- */
static int is_leap(int y) { return ((y % 4 == 0) && (y % 100 != 0)) || (y % 400 == 0); }

static void warn(int quiet, const char *fmt, ...) {
  if (quiet) return;
  va_list ap;
  va_start(ap, fmt);
  fputs("[WARNING] ", stdout);
  vprintf(fmt, ap);
  va_end(ap);
}

int sip_check_restart(const char *path, const sip_restart *r, const sip_context *ctx, const sip_site_data *site) {
  const int quiet = ctx->quiet;
  /* validateCheckpointBoundaryForLoad */
  const double stepHours = r->boundaryLength * 24.0;
  if (stepHours <= RESTART_EPS)
    return sip_fail(BAD,
                    "Restart boundary mismatch in %s: checkpoint boundary has non-positive timestep length (year=%d day=%d "
                    "time=%.8f length=%.8f)",
                    path, r->boundaryYear, r->boundaryDay, r->boundaryTime, r->boundaryLength);
  if (24.0 - r->boundaryTime > stepHours + RESTART_EPS) {
    warn(quiet, "Restart checkpoint boundary in %s is more than one timestep before midnight; there is a time gap on resume.\n",
         path);
    warn(quiet, "Checkpoint boundary: year=%d day=%d time=%.8f length=%.8f\n", r->boundaryYear, r->boundaryDay,
         r->boundaryTime, r->boundaryLength);
  }
  /* checkRestartContextCompatibility */
  for (int k = 0; k < 12; ++k)
    if (*flag_slot((sipnet_gpu_flags *)&ctx->flags, k) != *flag_slot((sipnet_gpu_flags *)&r->flags, k))
      return sip_fail(BAD, "Restart context mismatch: model flags must match checkpoint exactly");
  /* validateRestartModelBuild */
  if (strcmp(r->modelVersion, SIP_MODEL_VERSION) != 0)
    return sip_fail(BAD, "Restart model version mismatch: checkpoint=%s current=%s", r->modelVersion, SIP_MODEL_VERSION);
  if (strcmp(r->buildInfo, SIP_BUILD_INFO) != 0)
    sip_info(quiet, "Restart build info mismatch: checkpoint=%s current=%s\n", r->buildInfo, SIP_BUILD_INFO);
  /* validateRestartBoundary */
  if (site->nsteps <= 0) return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "Cannot restart: climate forcing has no records");
  const int y = site->year[0], d = site->day[0];
  const double tm = site->time[0], ln = site->length[0];
  const int after = (y != r->boundaryYear) ? (y > r->boundaryYear)
                    : (d != r->boundaryDay) ? (d > r->boundaryDay)
                                            : (tm > r->boundaryTime + RESTART_EPS);
  if (!after)
    return sip_fail(BAD,
                    "Restart boundary mismatch: first climate timestamp does not follow checkpoint boundary timestamp "
                    "(checkpoint year=%d day=%d time=%.8f; found year=%d day=%d time=%.8f)",
                    r->boundaryYear, r->boundaryDay, r->boundaryTime, y, d, tm);
  const double firstHours = ln * 24.0;
  if (firstHours <= RESTART_EPS)
    return sip_fail(BAD, "Cannot restart: first climate timestep length is non-positive (year=%d day=%d time=%.8f length=%.8f)",
                    y, d, tm, ln);
  int ey = r->boundaryYear, ed = r->boundaryDay + 1;
  if (ed > (is_leap(ey) ? 366 : 365)) {
    ed = 1;
    ++ey;
  }
  if (y != ey || d != ed || tm > firstHours + RESTART_EPS) {
    warn(quiet, "Restart resumed segment starts more than one timestep after midnight checkpoint boundary; there is a time gap\n");
    warn(quiet, "Expected start on year=%d day=%d with time<=%.8f; found year=%d day=%d time=%.8f length=%.8f\n", ey, ed,
         firstHours, y, d, tm, ln);
  }
  if (r->meanStart < 0 || r->meanStart >= r->meanLength || r->meanLast < 0 || r->meanLast >= r->meanLength)
    return sip_fail(BAD, "Restart mean-tracker cursor out of range in %s", path);
  return 0;
}

/* ---- device state <-> checkpoint ---------------------------------------------------------------- */
static const int kEnviRows[SIP_RESTART_NENVI] = {
    SIPNET_S_plantWoodC, SIPNET_S_plantLeafC, SIPNET_S_soilC,    SIPNET_S_soilWater, SIPNET_S_litterC,
    SIPNET_S_snow,       SIPNET_S_coarseRootC, SIPNET_S_fineRootC, SIPNET_S_minN,     SIPNET_S_soilOrgN,
    SIPNET_S_litterN,    SIPNET_S_plantStorageN, SIPNET_S_plantCAccountingDelta};
/* tracker index -> state row, for the trackers that are carried between steps (-1: per-step value, recomputed
   before it is read: updateTrackers(), sipnet.c:1433-1487) */
static const int kTrackerRows[SIP_RESTART_NTRACKERS] = {
    -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
    SIPNET_S_gdd, -1, SIPNET_S_soilWetnessFrac, SIPNET_S_yearlyGpp, SIPNET_S_yearlyRtot, SIPNET_S_yearlyRa,
    SIPNET_S_yearlyRh, SIPNET_S_yearlyNpp, SIPNET_S_yearlyNee, SIPNET_S_yearlyLitter, SIPNET_S_totGpp,
    SIPNET_S_totRtot, SIPNET_S_totRa, SIPNET_S_totRh, SIPNET_S_totNpp, SIPNET_S_totNee, SIPNET_S_trackersLastYear,
    -1, -1, -1, -1, -1, -1};

void sip_restart_to_state(const sip_restart *r, double *state, int64_t stride, double *ringV, double *ringW,
                          int64_t ringStride) {
  for (int k = 0; k < SIPNET_GPU_NSTATE; ++k) state[(int64_t)k * stride] = 0.0;
  for (int k = 0; k < SIP_RESTART_NENVI; ++k) state[(int64_t)kEnviRows[k] * stride] = r->envi[k];
  for (int k = 0; k < SIP_RESTART_NTRACKERS; ++k)
    if (kTrackerRows[k] >= 0) state[(int64_t)kTrackerRows[k] * stride] = r->trackers[k];
  state[(int64_t)SIPNET_S_didLeafGrowth * stride] = r->didLeafGrowth;
  state[(int64_t)SIPNET_S_didLeafFall * stride] = r->didLeafFall;
  state[(int64_t)SIPNET_S_phenLastYear * stride] = r->phenLastYear;
  state[(int64_t)SIPNET_S_dTillMod * stride] = r->dTillMod;
  state[(int64_t)SIPNET_S_harvestFracRemoved * stride] = r->harvestFracRemoved;
  state[(int64_t)SIPNET_S_harvestFracTransferred * stride] = r->harvestFracTransferred;
  state[(int64_t)SIPNET_S_meanSum * stride] = r->meanSum;
  state[(int64_t)SIPNET_S_meanStart * stride] = r->meanStart;
  state[(int64_t)SIPNET_S_meanLast * stride] = r->meanLast;
  for (int i = 0; i < SIP_RESTART_RING; ++i) {
    ringV[(int64_t)i * ringStride] = r->values[i];
    ringW[(int64_t)i * ringStride] = r->weights[i];
  }
}

void sip_restart_from_device(sip_restart *r, const sip_context *ctx, const sip_site_data *site, long long processedSteps,
                             long long utcEpoch, const double *state, int64_t stride, const double *dbgLast,
                             int64_t dbgStride, const double *ringV, const double *ringW, int64_t ringStride) {
  memset(r, 0, sizeof *r);
  strncpy(r->modelVersion, SIP_MODEL_VERSION, sizeof r->modelVersion - 1);
  strncpy(r->buildInfo, SIP_BUILD_INFO, sizeof r->buildInfo - 1);
  r->checkpointUtcEpoch = utcEpoch;
  r->processedSteps = processedSteps;
  r->flags = ctx->flags;
  const int64_t last = site->nsteps - 1; /* lastProcessedClimateStep, restart.c:919 */
  r->boundaryYear = site->year[last];
  r->boundaryDay = site->day[last];
  r->boundaryTime = site->time[last];
  r->boundaryLength = site->length[last];
  for (int k = 0; k < SIP_RESTART_NENVI; ++k) r->envi[k] = state[(int64_t)kEnviRows[k] * stride];
  /* all 33 trackers as of the last step: the validation dump's tracker block (include/sipnet_gpu.h: 13 envi,
     56 fluxes, then the trackers in struct order) */
  const int t0 = SIPNET_GPU_NDEBUG_ENVI + SIPNET_GPU_NDEBUG_FLUX;
  for (int k = 0; k < SIP_RESTART_NTRACKERS; ++k) r->trackers[k] = dbgLast[(int64_t)(t0 + k) * dbgStride];
  r->didLeafGrowth = (int)state[(int64_t)SIPNET_S_didLeafGrowth * stride];
  r->didLeafFall = (int)state[(int64_t)SIPNET_S_didLeafFall * stride];
  r->phenLastYear = (int)state[(int64_t)SIPNET_S_phenLastYear * stride];
  r->isAlive = (int)dbgLast[(int64_t)(SIPNET_GPU_NDEBUG - 1) * dbgStride];
  r->dTillMod = state[(int64_t)SIPNET_S_dTillMod * stride];
  r->harvestFracRemoved = state[(int64_t)SIPNET_S_harvestFracRemoved * stride];
  r->harvestFracTransferred = state[(int64_t)SIPNET_S_harvestFracTransferred * stride];
  r->meanLength = SIP_RESTART_RING;
  r->meanTotWeight = 5.0; /* MEAN_NPP_DAYS, sipnet.c:39 */
  r->meanStart = (int)state[(int64_t)SIPNET_S_meanStart * stride];
  r->meanLast = (int)state[(int64_t)SIPNET_S_meanLast * stride];
  r->meanSum = state[(int64_t)SIPNET_S_meanSum * stride];
  for (int i = 0; i < SIP_RESTART_RING; ++i) {
    r->values[i] = ringV[(int64_t)i * ringStride];
    r->weights[i] = ringW[(int64_t)i * ringStride];
  }
}

/* validateCheckpointBoundaryForWrite(), restart.c:357-378 */
int sip_check_restart_boundary_for_write(const char *path, const sip_restart *r, int quiet) {
  const double stepHours = r->boundaryLength * 24.0;
  if (stepHours <= RESTART_EPS)
    return sip_fail(BAD,
                    "Cannot write restart checkpoint %s: non-positive timestep length at boundary (year=%d day=%d time=%.8f "
                    "length=%.8f)",
                    path, r->boundaryYear, r->boundaryDay, r->boundaryTime, r->boundaryLength);
  if (24.0 - r->boundaryTime > stepHours + RESTART_EPS) {
    warn(quiet, "Restart checkpoint %s ends more than one timestep before midnight; there will be a time gap if this file is "
                "used to resume.\n",
         path);
    warn(quiet, "Boundary timestep: year=%d day=%d time=%.8f length=%.8f\n", r->boundaryYear, r->boundaryDay, r->boundaryTime,
         r->boundaryLength);
  }
  return 0;
}
