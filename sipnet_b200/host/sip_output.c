/*
 * sip_output.c -- text writers with the reference's exact row formats.
 *
 *   <prefix>.out   outputHeader()/outputState(), reference src/sipnet/sipnet.c:434-473
 *   events.out     openEventOutFile()/doWriteEventOut(), reference src/sipnet/events.c:369-402
 *
 * The device hands back numbers (sipnet_gpu_gather); only the formatting lives here, so a
 * single-member run reproduces the reference's files byte for byte.
 */
#include <string.h>

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

/* one row; column c of this member-step is out32[c * stride] (the device layout is [col][step][member]) */
void sip_write_state_row(FILE *out, int year, int day, double time, const double *o, int64_t stride) {
#define COL(name) o[(int64_t)SIPNET_O_##name * stride]
  fprintf(out, "%4d %3d %5.2f %10.2f %10.2f %12.2f ", year, day, time, COL(plantWoodC), COL(plantLeafC), COL(woodCreation));
  fprintf(out, "%8.2f ", COL(soilC));
  fprintf(out, "%11.2f %9.2f ", COL(coarseRootC), COL(fineRootC));
  fprintf(out, "%8.2f %10.3f %15.3f %8.2f ", COL(litterC), COL(soilWater), COL(soilWetnessFrac), COL(snow));
  fprintf(out, "%8.3f %8.3f %8.3f %8.3f %12.3f %8.3f %8.3f %8.3f %8.3f %8.3f %18.8f ", COL(npp), COL(nee), COL(cumNEE),
          COL(gpp), COL(rAboveground), COL(rSoil), COL(rRoot), COL(ra), COL(rh), COL(rtot), COL(evapotranspiration));
  fprintf(out, "%19.4f %8.4f %9.4f %10.4f %14.4f ", COL(fluxestranspiration), COL(minN), COL(soilOrgN), COL(litterN),
          COL(plantStorageN));
  fprintf(out, "%9.6f %9.4f %10.4f %8.4f %8.4f", COL(n2o), COL(nLeaching), COL(nFixation), COL(nUptake), COL(ch4));
  fprintf(out, "%12.4f\n", COL(nppStorage));
#undef COL
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
