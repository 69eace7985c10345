/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * A thin in-process harness around the UNMODIFIED reference sources, compiled
 * from where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libsipnet_refshim.so.  It follows the reference's own unit-test
 * idiom (tests/sipnet/test_modeling/testFluxCalculations.c:2-3 #include the
 * production .c files to reach `static` state; tests/utils/exitHandler.c:40-49
 * stubs exit() with setjmp/longjmp) so that one call can
 *   - load a parameter vector / flag set / climate list / event list from memory
 *     (no text round trip),
 *   - run setupModel() + the updateState() loop of runModelOutput()
 *     (sipnet.c:1963-1982), and
 *   - hand back every Envi / Fluxes / Trackers field per step as raw doubles
 *     (the same fields, in the same order, that --debug-log prints with %.15g,
 *     debug_log.c:51-170) plus the outputState() columns (sipnet.c:453-473).
 *
 * Nothing here is product code and nothing of the reference is copied: the
 * reference translation units are #included / linked from /root/reference.
 * The Makefile compiles every reference file with -Dexit=sipref_exit_ so the
 * reference's exit(code) calls land in the longjmp stub below.
 */
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sipnet/sipnet.c" /* from -I/root/reference/src */
#include "sipnet/events.c"

#include "../include/sipnet_gpu.h"

static jmp_buf sipref_jmp;
static int sipref_jmp_armed = 0;
static int sipref_code = 0;

/* every exit() in the reference is redirected here by -Dexit=sipref_exit_ */
void sipref_exit_(int code) {
  sipref_code = code;
  if (sipref_jmp_armed) {
    longjmp(sipref_jmp, 1);
  }
  _Exit(code);
}

int sipref_num_params(void) { return (int)NUM_PARAMS; }
int sipref_sizeof_envi(void) { return (int)sizeof(Envi); }
int sipref_sizeof_fluxes(void) { return (int)sizeof(Fluxes); }
int sipref_sizeof_trackers(void) { return (int)sizeof(Trackers); }

/* byte offset of a Params member by name (to check the ABI's enum order) */
int sipref_param_offset(const char *name) {
#define SIPNET_GPU_X(n)                                                        \
  if (strcmp(name, #n) == 0)                                                   \
    return (int)offsetof(Params, n);
  SIPNET_GPU_PARAM_LIST(SIPNET_GPU_X)
#undef SIPNET_GPU_X
  return -1;
}

static void free_climate(ClimateNode *head) {
  while (head) {
    ClimateNode *n = head->nextClim;
    free(head);
    head = n;
  }
}

static EventNode *build_events(int64_t nev, const sipnet_gpu_event *ev) {
  EventNode *head = NULL, *tail = NULL;
  for (int64_t i = 0; i < nev; ++i) {
    EventNode *n = (EventNode *)calloc(1, sizeof(EventNode));
    n->year = ev[i].year;
    n->day = ev[i].day;
    n->type = (event_type_t)ev[i].type;
    switch (ev[i].type) {
      case HARVEST: {
        HarvestParams *p = (HarvestParams *)calloc(1, sizeof *p);
        p->fractionRemovedAbove = ev[i].p[0];
        p->fractionRemovedBelow = ev[i].p[1];
        p->fractionTransferredAbove = ev[i].p[2];
        p->fractionTransferredBelow = ev[i].p[3];
        n->eventParams = p;
      } break;
      case IRRIGATION: {
        IrrigationParams *p = (IrrigationParams *)calloc(1, sizeof *p);
        p->amountAdded = ev[i].p[0];
        p->method = (irrigation_method_t)ev[i].method;
        n->eventParams = p;
      } break;
      case FERTILIZATION: {
        FertilizationParams *p = (FertilizationParams *)calloc(1, sizeof *p);
        p->orgN = ev[i].p[0];
        p->orgC = ev[i].p[1];
        p->minN = ev[i].p[2];
        n->eventParams = p;
      } break;
      case PLANTING: {
        PlantingParams *p = (PlantingParams *)calloc(1, sizeof *p);
        p->leafC = ev[i].p[0];
        p->woodC = ev[i].p[1];
        p->fineRootC = ev[i].p[2];
        p->coarseRootC = ev[i].p[3];
        n->eventParams = p;
      } break;
      case TILLAGE: {
        TillageParams *p = (TillageParams *)calloc(1, sizeof *p);
        p->tillageEffect = ev[i].p[0];
        n->eventParams = p;
      } break;
      default:
        n->eventParams = calloc(1, sizeof(double));
        break;
    }
    if (tail) {
      tail->nextEvent = n;
    } else {
      head = n;
    }
    tail = n;
  }
  return head;
}

static void dump_debug(double *d) {
  int k = 0;
  const double *e = (const double *)&envi;
  for (size_t i = 0; i < sizeof(Envi) / sizeof(double); ++i) d[k++] = e[i];
  const double *f = (const double *)&fluxes;
  for (size_t i = 0; i < sizeof(Fluxes) / sizeof(double); ++i) d[k++] = f[i];
  /* trackers in debug_log.c:126-158 order (struct order; lastYear is an int) */
  d[k++] = trackers.gpp;
  d[k++] = trackers.rtot;
  d[k++] = trackers.ra;
  d[k++] = trackers.rh;
  d[k++] = trackers.rRoot;
  d[k++] = trackers.rSoil;
  d[k++] = trackers.rAboveground;
  d[k++] = trackers.npp;
  d[k++] = trackers.nee;
  d[k++] = trackers.woodCreation;
  d[k++] = trackers.gdd;
  d[k++] = trackers.evapotranspiration;
  d[k++] = trackers.soilWetnessFrac;
  d[k++] = trackers.yearlyGpp;
  d[k++] = trackers.yearlyRtot;
  d[k++] = trackers.yearlyRa;
  d[k++] = trackers.yearlyRh;
  d[k++] = trackers.yearlyNpp;
  d[k++] = trackers.yearlyNee;
  d[k++] = trackers.yearlyLitter;
  d[k++] = trackers.totGpp;
  d[k++] = trackers.totRtot;
  d[k++] = trackers.totRa;
  d[k++] = trackers.totRh;
  d[k++] = trackers.totNpp;
  d[k++] = trackers.totNee;
  d[k++] = (double)trackers.lastYear;
  d[k++] = trackers.methane;
  d[k++] = trackers.n2o;
  d[k++] = trackers.nLeaching;
  d[k++] = trackers.nFixation;
  d[k++] = trackers.nUptake;
  d[k++] = trackers.meanNPP;
  d[k++] = (double)phenologyTrackers.didLeafGrowth;
  d[k++] = (double)phenologyTrackers.didLeafFall;
  d[k++] = (double)phenologyTrackers.lastYear;
  d[k++] = (double)plantSurvivalTracker.isAlive;
}

static void dump_out(double *o) {
  /* the value list of outputState(), sipnet.c:455-472 */
  o[0] = getTotalWoodC();
  o[1] = envi.plantLeafC;
  o[2] = trackers.woodCreation;
  o[3] = envi.soilC;
  o[4] = envi.coarseRootC;
  o[5] = envi.fineRootC;
  o[6] = envi.litterC;
  o[7] = envi.soilWater;
  o[8] = trackers.soilWetnessFrac;
  o[9] = envi.snow;
  o[10] = trackers.npp;
  o[11] = trackers.nee;
  o[12] = trackers.totNee;
  o[13] = trackers.gpp;
  o[14] = trackers.rAboveground;
  o[15] = trackers.rSoil;
  o[16] = trackers.rRoot;
  o[17] = trackers.ra;
  o[18] = trackers.rh;
  o[19] = trackers.rtot;
  o[20] = trackers.evapotranspiration;
  o[21] = fluxes.transpiration;
  o[22] = envi.minN;
  o[23] = envi.soilOrgN;
  o[24] = envi.litterN;
  o[25] = envi.plantStorageN;
  o[26] = trackers.n2o;
  o[27] = trackers.nLeaching;
  o[28] = trackers.nFixation;
  o[29] = trackers.nUptake;
  o[30] = trackers.methane;
  o[31] = envi.plantCAccountingDelta;
}

/* the reference's context (flag defaults) is initialised once for all entry points: a per-function guard would let a
 * later first call of another entry point reset flags the caller has just set through sipref_read_params() */
static void ensure_context(void) {
  static int ready = 0;
  if (!ready) {
    initContext();
    ready = 1;
  }
}

/* Where the next sipref_run() stores balanceTracker.deltaC / deltaN per step ([cap][2]); NULL = nowhere. */
static double *sipref_balance_out = NULL;
static int64_t sipref_balance_cap = 0;
void sipref_set_balance_out(double *buf, int64_t cap) {
  sipref_balance_out = buf;
  sipref_balance_cap = cap;
}

/*
 * Run one member through the reference.  Arrays are per step, already in the
 * units readClimData() leaves them in.  out32 is [T][32], dbg is [T][106]
 * (either may be NULL).  main_out_path, if not NULL, receives the reference's
 * own sipnet.out text (outputHeader/outputState) for the writer tests.
 * Returns 0 or the reference exit code; *steps_done counts finished steps.
 */
int sipref_run(const int32_t *flags, const double *params_in, int64_t T,
               const int32_t *year, const int32_t *day, const double *time,
               const double *length, const double *tair, const double *tsoil,
               const double *par, const double *precip, const double *vpd,
               const double *vpdSoil, const double *vPress, const double *wspd,
               const double *gdd, int64_t nev, const sipnet_gpu_event *ev,
               const char *events_out_path, const char *main_out_path,
               int print_header, double *out32, double *dbg,
               int64_t *steps_done) {
  ensure_context();
  ctx.events = flags[0];
  ctx.gdd = flags[1];
  ctx.growthResp = flags[2];
  ctx.leafWater = flags[3];
  ctx.litterPool = flags[4];
  ctx.snow = flags[5];
  ctx.soilPhenol = flags[6];
  ctx.waterHResp = flags[7];
  ctx.nitrogenCycle = flags[8];
  ctx.anaerobic = flags[9];
  ctx.flooding = flags[10];
  ctx.carbonSaturation = flags[11];
  ctx.quiet = 1;
  ctx.restartIn[0] = '\0';
  ctx.restartOut[0] = '\0';
  ctx.debugLogPrefix[0] = '\0';

  /* emulate a fresh process: the reference's globals start zeroed */
  memset(&envi, 0, sizeof envi);
  memset(&fluxes, 0, sizeof fluxes);
  memset(&trackers, 0, sizeof trackers);
  memset(&phenologyTrackers, 0, sizeof phenologyTrackers);
  memset(&plantSurvivalTracker, 0, sizeof plantSurvivalTracker);
  memset(&eventTrackers, 0, sizeof eventTrackers);
  memcpy(&params, params_in, sizeof(Params));

  ClimateNode *head = NULL, *tail = NULL;
  for (int64_t t = 0; t < T; ++t) {
    ClimateNode *n = (ClimateNode *)calloc(1, sizeof(ClimateNode));
    n->year = year[t];
    n->day = day[t];
    n->time = time[t];
    n->length = length[t];
    n->tair = tair[t];
    n->tsoil = tsoil[t];
    n->par = par[t];
    n->precip = precip[t];
    n->vpd = vpd[t];
    n->vpdSoil = vpdSoil[t];
    n->vPress = vPress[t];
    n->wspd = wspd[t];
    n->gdd = gdd[t];
    if (tail) {
      tail->nextClim = n;
    } else {
      head = n;
    }
    tail = n;
  }
  firstClimate = head;

  if (meanNPP == NULL) {
    meanNPP = newMeanTracker(0, MEAN_NPP_DAYS, MEAN_NPP_MAX_ENTRIES);
  }

  gEvents = NULL;
  FILE *mainOut = NULL;
  int64_t done = 0;
  sipref_code = 0;
  sipref_jmp_armed = 1;
  if (setjmp(sipref_jmp) == 0) {
    if (ctx.events) {
      gEvents = build_events(nev, ev);
      openEventOutFile(events_out_path ? events_out_path : "/dev/null",
                       print_header);
      if (isFirstEventBefore(firstClimate->year, firstClimate->day)) {
        sipref_exit_(EXIT_CODE_INPUT_FILE_ERROR); /* frontend.c:217-222 */
      }
    }
    if (main_out_path) {
      mainOut = fopen(main_out_path, "w");
      if (mainOut && print_header) {
        outputHeader(mainOut);
      }
    }
    setupModel();
    setupEvents();
    while (climate != NULL) {
      updateState();
      if (mainOut) {
        outputState(mainOut, climate->year, climate->day, climate->time);
      }
      if (out32) dump_out(out32 + done * SIPNET_GPU_NOUT);
      if (dbg) dump_debug(dbg + done * SIPNET_GPU_NDEBUG);
      if (sipref_balance_out && done < sipref_balance_cap) { /* checkBalance()'s results, balance.c:129-148 */
        sipref_balance_out[2 * done] = balanceTracker.deltaC;
        sipref_balance_out[2 * done + 1] = balanceTracker.deltaN;
      }
      ++done;
      climate = climate->nextClim;
    }
  }
  sipref_jmp_armed = 0;
  if (mainOut) fclose(mainOut);
  if (ctx.events) {
    closeEventOutFile();
    freeEventList();
    gEvents = gEvent = NULL;
  }
  free_climate(head);
  firstClimate = climate = NULL;
  if (steps_done) *steps_done = done;
  return sipref_code;
}

/* The reference's climate reader (sipnet.c:128-277) into flat arrays, so the
 * host-reader tests can be checked against it.  Returns number of steps or
 * -exitcode. */
int64_t sipref_read_clim(const char *path, int gddFlag, int64_t cap,
                         int32_t *year, int32_t *day, double *cols /*[11][cap]*/) {
  ensure_context();
  ctx.gdd = gddFlag;
  ctx.quiet = 1;
  sipref_code = 0;
  sipref_jmp_armed = 1;
  int64_t n = 0;
  if (setjmp(sipref_jmp) == 0) {
    readClimData(path);
    for (ClimateNode *c = firstClimate; c && n < cap; c = c->nextClim, ++n) {
      year[n] = c->year;
      day[n] = c->day;
      cols[0 * cap + n] = c->time;
      cols[1 * cap + n] = c->length;
      cols[2 * cap + n] = c->tair;
      cols[3 * cap + n] = c->tsoil;
      cols[4 * cap + n] = c->par;
      cols[5 * cap + n] = c->precip;
      cols[6 * cap + n] = c->vpd;
      cols[7 * cap + n] = c->vpdSoil;
      cols[8 * cap + n] = c->vPress;
      cols[9 * cap + n] = c->wspd;
      cols[10 * cap + n] = c->gdd;
    }
    freeClimateList();
    firstClimate = climate = NULL;
  }
  sipref_jmp_armed = 0;
  return sipref_code ? -(int64_t)sipref_code : n;
}

/* The reference's parameter reader (sipnet.c:290-427) into a flat vector. */
int sipref_read_params(const char *path, const int32_t *flags, double *out80) {
  ensure_context();
  ctx.events = flags[0];
  ctx.gdd = flags[1];
  ctx.growthResp = flags[2];
  ctx.leafWater = flags[3];
  ctx.litterPool = flags[4];
  ctx.snow = flags[5];
  ctx.soilPhenol = flags[6];
  ctx.waterHResp = flags[7];
  ctx.nitrogenCycle = flags[8];
  ctx.anaerobic = flags[9];
  ctx.flooding = flags[10];
  ctx.carbonSaturation = flags[11];
  ctx.quiet = 1;
  memset(&params, 0, sizeof(Params));
  sipref_code = 0;
  sipref_jmp_armed = 1;
  if (setjmp(sipref_jmp) == 0) {
    ModelParams *mp = NULL;
    readParamData(&mp, path);
    deleteModelParams(mp);
    memcpy(out80, &params, sizeof(Params));
  }
  sipref_jmp_armed = 0;
  return sipref_code;
}

/* The reference's event reader (events.c:263-367). Returns count or -exitcode. */
int64_t sipref_read_events(const char *path, int64_t cap, sipnet_gpu_event *out) {
  ensure_context();
  ctx.quiet = 1;
  sipref_code = 0;
  sipref_jmp_armed = 1;
  int64_t n = 0;
  if (setjmp(sipref_jmp) == 0) {
    EventNode *list = readEventData(path);
    for (EventNode *e = list; e && n < cap; e = e->nextEvent, ++n) {
      memset(&out[n], 0, sizeof out[n]);
      out[n].year = e->year;
      out[n].day = e->day;
      out[n].type = (int32_t)e->type;
      switch (e->type) {
        case HARVEST: {
          HarvestParams *p = (HarvestParams *)e->eventParams;
          out[n].p[0] = p->fractionRemovedAbove;
          out[n].p[1] = p->fractionRemovedBelow;
          out[n].p[2] = p->fractionTransferredAbove;
          out[n].p[3] = p->fractionTransferredBelow;
        } break;
        case IRRIGATION: {
          IrrigationParams *p = (IrrigationParams *)e->eventParams;
          out[n].p[0] = p->amountAdded;
          out[n].method = (int32_t)p->method;
        } break;
        case FERTILIZATION: {
          FertilizationParams *p = (FertilizationParams *)e->eventParams;
          out[n].p[0] = p->orgN;
          out[n].p[1] = p->orgC;
          out[n].p[2] = p->minN;
        } break;
        case PLANTING: {
          PlantingParams *p = (PlantingParams *)e->eventParams;
          out[n].p[0] = p->leafC;
          out[n].p[1] = p->woodC;
          out[n].p[2] = p->fineRootC;
          out[n].p[3] = p->coarseRootC;
        } break;
        case TILLAGE: {
          TillageParams *p = (TillageParams *)e->eventParams;
          out[n].p[0] = p->tillageEffect;
        } break;
        default:
          break;
      }
    }
    gEvents = list;
    freeEventList();
    gEvents = gEvent = NULL;
  }
  sipref_jmp_armed = 0;
  return sipref_code ? -(int64_t)sipref_code : n;
}
