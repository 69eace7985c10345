/*
 * sip_output.c -- text writers with the reference's exact row formats.
 *
 *   <prefix>.out   outputHeader()/outputState(), reference src/sipnet/sipnet.c:434-473
 *   events.out     openEventOutFile()/doWriteEventOut(), reference src/sipnet/events.c:369-402
 *
 * The device hands back numbers (sipnet_gpu_gather); only the formatting lives here, so a
 * single-member run reproduces the reference's files byte for byte.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "sip_host.h"

int sip_fail(int code, const char *fmt, ...);

void sip_write_header(FILE *out) { /* column titles of sipnet.c:435-443 */
  fputs("year day  time plantWoodC plantLeafC woodCreation     ", out);
  fputs("soil coarseRootC fineRootC   ", out);
  fputs("litter  soilWater soilWetnessFrac     snow      ", out);
  fputs("npp      nee   cumNEE      gpp rAboveground    rSoil    rRoot       ra       rh     rtot evapotranspiration ", out);
  fputs("fluxestranspiration     minN  soilOrgN    litterN  plantStorageN       n2o nLeaching  nFixation  nUptake      ch4  "
        "nppStorage\n",
        out);
}

/* ---- "%W.Pf" without printf ----------------------------------------------------------------------------------
 * A many-member run writes tens of millions of rows of 33 fixed-point fields; glibc's printf spends ~100 ns on each
 * field.  fmt_fixed() produces the SAME bytes for the usual values and hands everything else to snprintf:
 *   printf rounds the EXACT binary value to P decimals.  y = |x| * 10^P (one rounded product) differs from the exact
 *   product by at most y * 2^-53, so unless y sits within that distance of a rounding boundary k + 1/2, rounding y
 *   is rounding the exact value.  Near a boundary (this includes every true tie), at or above 2^52 (no fraction bits
 *   left), and for NaN / infinity, snprintf decides.  The sign is the sign BIT, as in printf ("-0.00"). */
static const double kPow10d[10] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};
static const uint64_t kPow10u[10] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
                                     100000000ull, 1000000000ull};

static char *fmt_fixed(char *p, char *end, double x, int width, int prec) {
  uint64_t bits;
  memcpy(&bits, &x, sizeof bits);
  const int neg = (int)(bits >> 63);
  const double ax = neg ? -x : x;
  if (ax < 1e15 && prec >= 0 && prec <= 9) { /* (false for NaN) */
    const double y = ax * kPow10d[prec];
    if (y < 4e15) {
      const uint64_t fi = (uint64_t)y; /* y >= 0: truncation is floor */
      const double fl = (double)fi;
      const double d = y - fl; /* exact below 2^52 */
      const double margin = y * 2.3e-16;
      const double dist = d > 0.5 ? d - 0.5 : 0.5 - d;
      if (dist > margin) {
        const uint64_t n = fi + (d > 0.5 ? 1u : 0u);
        uint64_t ip; /* integer part; the split by a CONSTANT power of ten compiles to a multiply and a shift */
        uint32_t frac;
        switch (prec) {
          case 2: ip = n / 100u, frac = (uint32_t)(n % 100u); break;
          case 3: ip = n / 1000u, frac = (uint32_t)(n % 1000u); break;
          case 4: ip = n / 10000u, frac = (uint32_t)(n % 10000u); break;
          case 6: ip = n / 1000000u, frac = (uint32_t)(n % 1000000u); break;
          case 8: ip = n / 100000000u, frac = (uint32_t)(n % 100000000u); break;
          default: ip = n / kPow10u[prec], frac = (uint32_t)(n % kPow10u[prec]); break;
        }
        char tmp[40]; /* the field backwards */
        int len = 0;
        for (int i = 0; i < prec; ++i) {
          tmp[len++] = (char)('0' + (int)(frac % 10u));
          frac /= 10u;
        }
        if (prec > 0) tmp[len++] = '.';
        if (ip < 10u) { /* (most fluxes) */
          tmp[len++] = (char)('0' + (int)ip);
        } else if (ip <= 0xffffffffu) {
          uint32_t q = (uint32_t)ip;
          do {
            tmp[len++] = (char)('0' + (int)(q % 10u));
            q /= 10u;
          } while (q != 0);
        } else {
          do {
            tmp[len++] = (char)('0' + (int)(ip % 10u));
            ip /= 10u;
          } while (ip != 0);
        }
        if (neg) tmp[len++] = '-';
        if ((int64_t)(end - p) > (int64_t)(len > width ? len : width)) {
          for (int i = len; i < width; ++i) *p++ = ' ';
          while (len > 0) *p++ = tmp[--len];
          return p;
        }
      }
    }
  }
  const int room = (int)(end - p);
  const int w = room > 0 ? snprintf(p, (size_t)room, "%*.*f", width, prec, x) : 0;
  return p + (w < room ? w : (room > 0 ? room - 1 : 0));
}

static char *fmt_int(char *p, char *end, int v, int width) { /* "%Wd" */
  if (v >= 0 && (end - p) > 16) {
    char tmp[16];
    int len = 0;
    unsigned u = (unsigned)v;
    do {
      tmp[len++] = (char)('0' + (int)(u % 10u));
      u /= 10u;
    } while (u != 0);
    for (int i = len; i < width; ++i) *p++ = ' ';
    while (len > 0) *p++ = tmp[--len];
    return p;
  }
  const int room = (int)(end - p);
  const int w = room > 0 ? snprintf(p, (size_t)room, "%*d", width, v) : 0;
  return p + (w < room ? w : (room > 0 ? room - 1 : 0));
}

/* the fields of one outputState() row after "year day time", in print order (sipnet.c:455-472): column, width,
 * precision, and the byte that follows the field (0 = nothing: "%8.4f%12.4f\n" has no blank between its last two) */
static const struct {
  int col, width, prec;
  char after;
} kRowFields[SIPNET_GPU_NOUT] = {
    {SIPNET_O_plantWoodC, 10, 2, ' '},   {SIPNET_O_plantLeafC, 10, 2, ' '},    {SIPNET_O_woodCreation, 12, 2, ' '},
    {SIPNET_O_soilC, 8, 2, ' '},         {SIPNET_O_coarseRootC, 11, 2, ' '},   {SIPNET_O_fineRootC, 9, 2, ' '},
    {SIPNET_O_litterC, 8, 2, ' '},       {SIPNET_O_soilWater, 10, 3, ' '},     {SIPNET_O_soilWetnessFrac, 15, 3, ' '},
    {SIPNET_O_snow, 8, 2, ' '},          {SIPNET_O_npp, 8, 3, ' '},            {SIPNET_O_nee, 8, 3, ' '},
    {SIPNET_O_cumNEE, 8, 3, ' '},        {SIPNET_O_gpp, 8, 3, ' '},            {SIPNET_O_rAboveground, 12, 3, ' '},
    {SIPNET_O_rSoil, 8, 3, ' '},         {SIPNET_O_rRoot, 8, 3, ' '},          {SIPNET_O_ra, 8, 3, ' '},
    {SIPNET_O_rh, 8, 3, ' '},            {SIPNET_O_rtot, 8, 3, ' '},           {SIPNET_O_evapotranspiration, 18, 8, ' '},
    {SIPNET_O_fluxestranspiration, 19, 4, ' '}, {SIPNET_O_minN, 8, 4, ' '},    {SIPNET_O_soilOrgN, 9, 4, ' '},
    {SIPNET_O_litterN, 10, 4, ' '},      {SIPNET_O_plantStorageN, 14, 4, ' '}, {SIPNET_O_n2o, 9, 6, ' '},
    {SIPNET_O_nLeaching, 9, 4, ' '},     {SIPNET_O_nFixation, 10, 4, ' '},     {SIPNET_O_nUptake, 8, 4, ' '},
    {SIPNET_O_ch4, 8, 4, 0},             {SIPNET_O_nppStorage, 12, 4, '\n'}};

/* one row into memory (dst holds SIP_STATE_ROW_MAX bytes; no terminating NUL is counted); column c of this
 * member-step is o[c * stride] (the device layout is [col][step][member]) */
size_t sip_format_state_row(char *dst, int year, int day, double time, const double *o, int64_t stride) {
  char *p = dst, *const end = dst + SIP_STATE_ROW_MAX;
  p = fmt_int(p, end, year, 4);
  *p++ = ' ';
  p = fmt_int(p, end, day, 3);
  *p++ = ' ';
  p = fmt_fixed(p, end, time, 5, 2);
  *p++ = ' ';
  for (int f = 0; f < SIPNET_GPU_NOUT; ++f) {
    p = fmt_fixed(p, end - 2, o[(int64_t)kRowFields[f].col * stride], kRowFields[f].width, kRowFields[f].prec);
    if (kRowFields[f].after) *p++ = kRowFields[f].after;
  }
  return (size_t)(p - dst);
}

void sip_write_state_row(FILE *out, int year, int day, double time, const double *o, int64_t stride) {
  char row[SIP_STATE_ROW_MAX];
  const size_t n = sip_format_state_row(row, year, day, time, o, stride);
  fwrite(row, 1, n, out);
}

/* the row as printf writes it (the reference's own statements): what sip_format_state_row() must reproduce */
size_t sip_format_state_row_printf(char *dst, int year, int day, double time, const double *o, int64_t stride) {
#define COL(name) o[(int64_t)SIPNET_O_##name * stride]
  char *p = dst;
  char *const end = dst + SIP_STATE_ROW_MAX;
#define EMIT(...) p += snprintf(p, (size_t)(end - p), __VA_ARGS__)
  EMIT("%4d %3d %5.2f %10.2f %10.2f %12.2f ", year, day, time, COL(plantWoodC), COL(plantLeafC), COL(woodCreation));
  EMIT("%8.2f ", COL(soilC));
  EMIT("%11.2f %9.2f ", COL(coarseRootC), COL(fineRootC));
  EMIT("%8.2f %10.3f %15.3f %8.2f ", COL(litterC), COL(soilWater), COL(soilWetnessFrac), COL(snow));
  EMIT("%8.3f %8.3f %8.3f %8.3f %12.3f %8.3f %8.3f %8.3f %8.3f %8.3f %18.8f ", COL(npp), COL(nee), COL(cumNEE), COL(gpp),
       COL(rAboveground), COL(rSoil), COL(rRoot), COL(ra), COL(rh), COL(rtot), COL(evapotranspiration));
  EMIT("%19.4f %8.4f %9.4f %10.4f %14.4f ", COL(fluxestranspiration), COL(minN), COL(soilOrgN), COL(litterN),
       COL(plantStorageN));
  EMIT("%9.6f %9.4f %10.4f %8.4f %8.4f", COL(n2o), COL(nLeaching), COL(nFixation), COL(nUptake), COL(ch4));
  EMIT("%12.4f\n", COL(nppStorage));
#undef EMIT
#undef COL
  return (size_t)(p - dst);
}

/* Many members at once: the main output files of `count` members whose columns are neighbours in the gathered
 * [col][step][member] array (out32 points at the first one; colStride = steps * members, stepStride = members).
 * Reading a (column, step) pair brings in the cache line that holds all of them, so the block is transposed chunk by
 * chunk into member-major tiles and every member's rows are formatted from its tile and written with one fwrite per
 * chunk.  files[k] may be NULL (member skipped); nsteps[k] rows are written for member k. */
int sip_write_state_block(FILE *const *files, int count, const int64_t *nsteps, const int32_t *const *year,
                          const int32_t *const *day, const double *const *time, const double *out32, int64_t colStride,
                          int64_t stepStride) {
  enum { CHUNK = 64 };
  if (count < 1 || count > SIP_STATE_BLOCK_MAX) return sip_fail(SIPNET_GPU_ERR_INTERNAL, "bad member block");
  int64_t tmax = 0;
  for (int k = 0; k < count; ++k)
    if (files[k] && nsteps[k] > tmax) tmax = nsteps[k];
  double *tile = (double *)malloc((size_t)count * CHUNK * SIPNET_GPU_NOUT * sizeof(double)); /* [member][step][col] */
  char *text = (char *)malloc((size_t)CHUNK * SIP_STATE_ROW_MAX);
  if (!tile || !text) {
    free(tile);
    free(text);
    return sip_fail(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure");
  }
  int rc = 0;
  for (int64_t t0 = 0; t0 < tmax && rc == 0; t0 += CHUNK) {
    const int64_t nt = tmax - t0 < CHUNK ? tmax - t0 : CHUNK;
    for (int c = 0; c < SIPNET_GPU_NOUT; ++c)
      for (int64_t t = 0; t < nt; ++t) {
        const double *src = out32 + (int64_t)c * colStride + (t0 + t) * stepStride;
        for (int k = 0; k < count; ++k) tile[((size_t)k * CHUNK + (size_t)t) * SIPNET_GPU_NOUT + (size_t)c] = src[k];
      }
    for (int k = 0; k < count && rc == 0; ++k) {
      if (!files[k]) continue;
      size_t len = 0;
      for (int64_t t = 0; t < nt && t0 + t < nsteps[k]; ++t)
        len += sip_format_state_row(text + len, year[k][t0 + t], day[k][t0 + t], time[k][t0 + t],
                                    tile + ((size_t)k * CHUNK + (size_t)t) * SIPNET_GPU_NOUT, 1);
      if (len && fwrite(text, 1, len, files[k]) != len) rc = sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "short write to a main output file");
    }
  }
  free(tile);
  free(text);
  return rc;
}

/* ---- main output of a many-member launch: blocks of SIP_STATE_BLOCK_MAX neighbouring members, spread over the
 * host's cores (the reference formats one member per process; 4096 members x 7306 steps are 30 million rows) ---- */
typedef struct {
  const char *paths;
  int64_t M, T;
  const int64_t *nsteps;
  const int32_t *const *year, *const *day;
  const double *const *time;
  const double *buf; /* SIPNET_GPU_GATHER_FULL: [col][T][M] */
  int printHeader;
  int nworkers;
  volatile int rc; /* first failure */
  char msg[SIP_STATE_PATH_MAX + 64];
  pthread_mutex_t lock;
} writer_pool;

typedef struct {
  writer_pool *pool;
  int id;
} writer_arg;

static void *writer_main(void *argp) {
  writer_arg *wa = (writer_arg *)argp;
  writer_pool *wp = wa->pool;
  const int64_t nblocks = (wp->M + SIP_STATE_BLOCK_MAX - 1) / SIP_STATE_BLOCK_MAX;
  for (int64_t b = wa->id; b < nblocks && wp->rc == 0; b += wp->nworkers) {
    const int64_t m0 = b * SIP_STATE_BLOCK_MAX;
    const int count = (int)(wp->M - m0 < SIP_STATE_BLOCK_MAX ? wp->M - m0 : SIP_STATE_BLOCK_MAX);
    FILE *files[SIP_STATE_BLOCK_MAX] = {0};
    int rc = 0;
    const char *bad = NULL;
    for (int k = 0; k < count; ++k) {
      const char *name = wp->paths + (size_t)(m0 + k) * SIP_STATE_PATH_MAX;
      files[k] = fopen(name, "w");
      if (!files[k]) {
        rc = SIPNET_GPU_ERR_FILE_OPEN;
        bad = name;
        break;
      }
      if (wp->printHeader) sip_write_header(files[k]);
    }
    if (rc == 0) {
      rc = sip_write_state_block(files, count, wp->nsteps + m0, wp->year + m0, wp->day + m0, wp->time + m0, wp->buf + m0,
                                 wp->T * wp->M, wp->M);
      if (rc) bad = wp->paths + (size_t)m0 * SIP_STATE_PATH_MAX; /* (the block's first member names it) */
    }
    for (int k = 0; k < count; ++k)
      if (files[k] && fclose(files[k]) != 0 && rc == 0) {
        rc = SIPNET_GPU_ERR_FILE_OPEN;
        bad = wp->paths + (size_t)(m0 + k) * SIP_STATE_PATH_MAX;
      }
    if (rc) {
      pthread_mutex_lock(&wp->lock);
      if (wp->rc == 0) {
        wp->rc = rc;
        snprintf(wp->msg, sizeof wp->msg, "cannot open or write main output file %s", bad ? bad : "?");
      }
      pthread_mutex_unlock(&wp->lock);
    }
  }
  return NULL;
}

int sip_write_state_files(const char *paths, int64_t M, const int64_t *nsteps, const int32_t *const *year,
                          const int32_t *const *day, const double *const *time, int64_t T, const double *out32,
                          int printHeader, int nthreads) {
  if (M <= 0) return 0;
  const int64_t nblocks = (M + SIP_STATE_BLOCK_MAX - 1) / SIP_STATE_BLOCK_MAX;
  long want = nthreads;
  if (want <= 0) {
    const char *env = getenv("SIPNET_GPU_WRITER_THREADS");
    want = (env && atoi(env) > 0) ? atoi(env) : sysconf(_SC_NPROCESSORS_ONLN);
  }
  int nworkers = (int)(want < 1 ? 1 : (want > SIP_STATE_THREADS_MAX ? SIP_STATE_THREADS_MAX : want));
  if ((int64_t)nworkers > nblocks) nworkers = (int)nblocks;
  writer_pool wp;
  memset(&wp, 0, sizeof wp);
  wp.paths = paths, wp.M = M, wp.T = T, wp.nsteps = nsteps, wp.year = year, wp.day = day, wp.time = time, wp.buf = out32;
  wp.printHeader = printHeader, wp.nworkers = nworkers;
  pthread_mutex_init(&wp.lock, NULL);
  writer_arg args[SIP_STATE_THREADS_MAX];
  pthread_t threads[SIP_STATE_THREADS_MAX];
  int started = 0;
  for (int i = 1; i < nworkers; ++i) { /* worker 0 is this thread */
    args[i].pool = &wp;
    args[i].id = i;
    if (pthread_create(&threads[i], NULL, writer_main, &args[i]) != 0) break;
    started = i;
  }
  if (started + 1 < nworkers) { /* could not start them all: the blocks of the missing workers would be skipped, */
    for (int i = 1; i <= started; ++i) pthread_join(threads[i], NULL);
    wp.nworkers = 1; /* so everything is (re)written by this thread; the started workers have finished by now */
    wp.rc = 0;
    started = 0;
  }
  args[0].pool = &wp;
  args[0].id = 0;
  writer_main(&args[0]);
  for (int i = 1; i <= started; ++i) pthread_join(threads[i], NULL);
  pthread_mutex_destroy(&wp.lock);
  if (wp.rc) return sip_fail(wp.rc, "%s", wp.msg);
  return 0;
}

void sip_write_events_header(FILE *out) { /* events.c:374-375 */
  fprintf(out, "%4s  %3s  %-7s  %s", "year", "day", "type", "param_name=delta[,param_name=delta,...]\n");
}

/* the (name, value) lists each event kind prints, in print order */
static const char *const kIrrig[] = {"eventSoilWater", "eventEvap"};                                  /* events.c:504 */
static const char *const kPlant[] = {"eventLeafC",       "eventWoodC",  "eventFineRootC",
                                     "eventCoarseRootC", "eventInputC", "eventInputN"};               /* :534-540 */
static const char *const kHarv[] = {"eventSoilC",     "eventLitterC",     "eventLeafC",    "eventWoodC",  "eventFineRootC",
                                    "eventCoarseRootC", "eventSoilOrgN",  "eventLitterN",  "eventOutputC", "eventOutputN"}; /* :622-633 */
static const char *const kTill[] = {"eventTrackers.d_till_mod"};                                      /* :644 */
static const char *const kFert[] = {"eventLitterC", "eventSoilC", "eventMinN", "eventLitterN", "eventInputC", "eventInputN"}; /* :677-683 */
static const char *const kLeafOffEvent[] = {"eventLeafOffLitter", "eventLeafOffNResorption", "eventLitterN"}; /* :723-726 */
static const char *const kLeafOffComputed[] = {"leafLitter"};                                         /* sipnet.c:837-839 */
static const char *const kLeafOnComputed[] = {"leafOnCreation", "leafOnCreationFromWood"};            /* sipnet.c:1234-1237 */
static const char *const kLeafOnEvent[] = {"eventLeafOnCreation", "eventLeafOnCreationFromWood"};     /* sipnet.c:1242-1245 */
static const char *const kDeath[] = {"harvestFracRemoved", "harvestFracTransferred", "totalWoodC", "totalRootC"}; /* sipnet.c:1760-1764 */

int sip_write_event_row(FILE *out, int year, int day, const sipnet_gpu_event_record *rec) {
  const char *const *names = NULL;
  int n = 0;
  switch (rec->type) {
    case SIPNET_EV_IRRIGATION: names = kIrrig, n = 2; break;
    case SIPNET_EV_PLANTING: names = kPlant, n = 6; break;
    case SIPNET_EV_HARVEST: names = kHarv, n = 10; break;
    case SIPNET_EV_TILLAGE: names = kTill, n = 1; break;
    case SIPNET_EV_FERTILIZATION: names = kFert, n = 6; break;
    case SIPNET_EV_LEAFOFF:
      if (rec->variant) names = kLeafOffEvent, n = 3; else names = kLeafOffComputed, n = 1;
      break;
    case SIPNET_EV_LEAFON:
      if (rec->variant) names = kLeafOnEvent, n = 2; else names = kLeafOnComputed, n = 2;
      break;
    case SIPNET_EV_PLANTDEATH: names = kDeath, n = 4; break;
    default: return sip_fail(SIPNET_GPU_ERR_UNKNOWN_EVENT, "unknown event type in event record (%d)", rec->type);
  }
  if (n != rec->nval) return sip_fail(SIPNET_GPU_ERR_INTERNAL, "event record of type %d has %d values, expected %d", rec->type, rec->nval, n);
  fprintf(out, "%4d  %3d  %-7s  ", year, day, sip_event_type_name(rec->type)); /* events.c:387 */
  for (int k = 0; k < n - 1; ++k) fprintf(out, "%s=%-.2f,", names[k], rec->val[k]);
  fprintf(out, "%s=%-.2f\n", names[n - 1], rec->val[n - 1]);
  return 0;
}

/* ---- debug logs: field order of debug_log.c:51-170 ---------------------------------------------------------- */
static const char *const kEnviNames[] = {"plantWoodC", "plantLeafC", "soilC", "soilWater", "litterC", "snow", "coarseRootC", "fineRootC", "minN", "soilOrgN", "litterN", "plantStorageN", "plantCAccountingDelta"};
static const char *const kFluxNames[] = {"photosynthesis", "leafLitter", "woodLitter", "rVeg", "rSoil", "rain", "transpiration", "drainage", "litterToSoil", "rLitter", "snowFall", "snowMelt", "sublimation", "immedEvap", "fastFlow", "evaporation", "fineRootLoss", "coarseRootLoss", "fineRootCreation", "coarseRootCreation", "rCoarseRoot", "rFineRoot", "leafCreation", "woodCreation", "leafOnCreation", "leafOnCreationFromWood", "nVolatilization", "nLeaching", "nOrgSoil", "nOrgLitter", "nMin", "nFixation", "nUptake", "leafOffNResorption", "reductionNResorption", "eventLeafC", "eventWoodC", "eventFineRootC", "eventCoarseRootC", "eventEvap", "eventSoilWater", "eventSoilC", "eventLitterC", "eventMinN", "eventSoilOrgN", "eventLitterN", "eventInputC", "eventOutputC", "eventInputN", "eventOutputN", "eventLeafOnCreation", "eventLeafOnCreationFromWood", "eventLeafOffLitter", "eventLeafOffNResorption", "soilMethane", "litterMethane"};
static const char *const kTrackerNames[] = {"gpp", "rtot", "ra", "rh", "rRoot", "rSoil", "rAboveground", "npp", "nee", "woodCreation", "gdd", "evapotranspiration", "soilWetnessFrac", "yearlyGpp", "yearlyRtot", "yearlyRa", "yearlyRh", "yearlyNpp", "yearlyNee", "yearlyLitter", "totGpp", "totRtot", "totRa", "totRh", "totNpp", "totNee", "lastYear", "methane", "n2o", "nLeaching", "nFixation", "nUptake", "meanNPP"};

void sip_write_debug_headers(FILE *envi, FILE *fluxes, FILE *trackers) { /* outputDebugHeaders(), debug_log.c:250-275 */
  fprintf(envi, "year day time");
  for (int k = 0; k < SIPNET_GPU_NDEBUG_ENVI; ++k) fprintf(envi, " %s", kEnviNames[k]);
  fprintf(envi, "\n");
  fprintf(fluxes, "year day time");
  for (int k = 0; k < SIPNET_GPU_NDEBUG_FLUX; ++k) fprintf(fluxes, " %s", kFluxNames[k]);
  fprintf(fluxes, "\n");
  fprintf(trackers, "year day time");
  for (int k = 0; k < SIPNET_GPU_NDEBUG_TRACK; ++k) fprintf(trackers, " t.%s", kTrackerNames[k]);
  fprintf(trackers, " pt.didLeafGrowth pt.didLeafFall pt.lastYear s.isAlive\n");
}

void sip_write_debug_rows(FILE *envi, FILE *fluxes, FILE *trackers, int year, int day, double time, const double *d,
                          int64_t stride) { /* outputDebugState(), debug_log.c:277-312: "%4d %3d %5.2f" then " %.15g" / " %d" */
  int k = 0;
  fprintf(envi, "%4d %3d %5.2f", year, day, time);
  for (int i = 0; i < SIPNET_GPU_NDEBUG_ENVI; ++i, ++k) fprintf(envi, " %.15g", d[(int64_t)k * stride]);
  fprintf(envi, "\n");
  fprintf(fluxes, "%4d %3d %5.2f", year, day, time);
  for (int i = 0; i < SIPNET_GPU_NDEBUG_FLUX; ++i, ++k) fprintf(fluxes, " %.15g", d[(int64_t)k * stride]);
  fprintf(fluxes, "\n");
  fprintf(trackers, "%4d %3d %5.2f", year, day, time);
  for (int i = 0; i < SIPNET_GPU_NDEBUG_TRACK; ++i, ++k) {
    if (i == 26) /* trackers.lastYear is an int field */
      fprintf(trackers, " %d", (int)d[(int64_t)k * stride]);
    else
      fprintf(trackers, " %.15g", d[(int64_t)k * stride]);
  }
  for (int i = 0; i < 4; ++i, ++k) fprintf(trackers, " %d", (int)d[(int64_t)k * stride]);
  fprintf(trackers, "\n");
}
