/*
 * sipnet_oracle.c -- TEST INFRASTRUCTURE ONLY (see sipnet_oracle.h).
 *
 * A from-scratch CPU restatement of the reference's hot path
 * (updateState(), reference src/sipnet/sipnet.c:1818-1855 and everything it
 * calls) written as one re-entrant simulator object instead of the reference's
 * process-global state.  Every routine cites the reference lines it follows.
 * Expression shapes (association order, `x / 365.0`, `-1.0 *`, constant chains)
 * are kept literally because the parity bar is bit-exactness against the
 * reference's `gcc -O0` x86-64 build: plain IEEE double operations, no FMA
 * contraction, glibc pow/exp.  Compile with -ffp-contract=off.
 *
 * Parity pinned: see the header comment of sipnet_oracle.h.
 */
#define _POSIX_C_SOURCE 200809L
#include "sipnet_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define OR_TINY 0.000001 /* common/util.h:14 */
#define OR_EPS 1e-8      /* balance.h:6 */
#define OR_RING 250      /* MEAN_NPP_MAX_ENTRIES, sipnet.c:39-40 */
#define OR_MEAN_DAYS 5.0 /* MEAN_NPP_DAYS, sipnet.c:39 */

/* pools, state.h:416-463 (same order as the debug log) */
typedef struct {
  double plantWoodC, plantLeafC, soilC, soilWater, litterC, snow, coarseRootC,
      fineRootC, minN, soilOrgN, litterN, plantStorageN, plantCAccountingDelta;
} OrPools;

/* per-day rates, state.h:469-645 (same order as the debug log) */
typedef struct {
  double photosynthesis, leafLitter, woodLitter, rVeg, rSoil, rain,
      transpiration, drainage, litterToSoil, rLitter, snowFall, snowMelt,
      sublimation, immedEvap, fastFlow, evaporation, fineRootLoss,
      coarseRootLoss, fineRootCreation, coarseRootCreation, rCoarseRoot,
      rFineRoot, leafCreation, woodCreation, leafOnCreation,
      leafOnCreationFromWood, nVolatilization, nLeaching, nOrgSoil, nOrgLitter,
      nMin, nFixation, nUptake, leafOffNResorption, reductionNResorption,
      eventLeafC, eventWoodC, eventFineRootC, eventCoarseRootC, eventEvap,
      eventSoilWater, eventSoilC, eventLitterC, eventMinN, eventSoilOrgN,
      eventLitterN, eventInputC, eventOutputC, eventInputN, eventOutputN,
      eventLeafOnCreation, eventLeafOnCreationFromWood, eventLeafOffLitter,
      eventLeafOffNResorption, soilMethane, litterMethane;
} OrRates;

/* trackers, state.h:650-726 */
typedef struct {
  double gpp, rtot, ra, rh, rRoot, rSoil, rAboveground, npp, nee, woodCreation,
      gdd, evapotranspiration, soilWetnessFrac, yearlyGpp, yearlyRtot, yearlyRa,
      yearlyRh, yearlyNpp, yearlyNee, yearlyLitter, totGpp, totRtot, totRa,
      totRh, totNpp, totNee;
  int lastYear;
  double methane, n2o, nLeaching, nFixation, nUptake, meanNPP;
} OrTrack;

/* weighted running mean, runmean.h / runmean.c */
typedef struct {
  double values[OR_RING], weights[OR_RING];
  int start, last;
  double sum, totWeight;
} OrRing;

typedef struct {
  double time, length, tair, tsoil, par, precip, vpd, vpdSoil, vPress, wspd, gdd;
  int year, day;
} OrClim;

typedef struct {
  /* configuration */
  sipnet_gpu_flags f;
  double p[SIPNET_GPU_NPARAMS];
  /* state */
  OrPools e;
  OrRates r;
  OrTrack t;
  OrRing ring;
  int didLeafGrowth, didLeafFall, phenLastYear; /* state.h:731-745 */
  int isAlive;                                  /* state.h:749-752 */
  double dTillMod, harvRemoved, harvTransferred; /* events.h:213-221 */
  /* current forcing */
  OrClim c;
  /* event cursor (gEvent of events.c:34) */
  const sipnet_gpu_event *ev;
  int64_t nev, evpos;
  /* events.out records */
  sipnet_gpu_event_record *recs;
  int32_t max_recs, nrec;
  int64_t step;
  double balDeltaC, balDeltaN; /* balanceTracker.deltaC / deltaN of the last step, balance.h:46-47 */
  uint32_t info;               /* SIPNET_GPU_ST_* bits of the reference's informational messages, whole run */
  uint32_t counts[SIPNET_GPU_NCOUNTERS]; /* how often each of them occurred (SIPNET_GPU_CNT_*) */
  int exit_code;
} OrSim;

#define P(name) (s->p[SIPNET_P_##name])

/* ---- numeric leaves ------------------------------------------------------- */
static double or_clip01(double x) { return fmin(fmax(x, 0.0), 1.0); } /* util.h:38 */
static double or_ratio(double num, double den) {                      /* util.c:72-75 */
  const double d = den < OR_TINY ? OR_TINY : den;
  return num / d;
}
static double or_wood_total(const OrSim *s) { /* getTotalWoodC, state.c */
  return s->e.plantWoodC + s->e.plantCAccountingDelta;
}

/* ---- running mean: runmean.c:44-121 ---------------------------------------- */
static void ring_reset(OrRing *g, double v) { /* runmean.c:44-51 */
  g->start = g->last = 0;
  g->values[0] = v;
  g->weights[0] = g->totWeight;
  g->sum = v * g->totWeight;
}
static int ring_push(OrRing *g, double value, double weight) { /* runmean.c:61-115 */
  if (weight <= 0) {
    return -1;
  }
  if (weight >= g->totWeight) {
    ring_reset(g, value);
    return 0;
  }
  double left = weight;
  int i = g->start;
  while (left > 0) {
    if (g->weights[i] > left) {
      g->weights[i] -= left;
      g->sum -= left * g->values[i];
      left = 0;
    } else {
      g->sum -= g->weights[i] * g->values[i];
      left -= g->weights[i];
      i = (i + 1) % OR_RING;
    }
  }
  g->start = i;
  i = (g->last + 1) % OR_RING;
  if (i == g->start) {
    g->weights[i] += weight;
    g->sum += weight * g->values[i];
    return -2;
  }
  g->last = i;
  g->values[i] = value;
  g->weights[i] = weight;
  g->sum += value * weight;
  return 0;
}
static double ring_mean(const OrRing *g) { return g->sum / g->totWeight; } /* runmean.c:118 */

/* ---- events.out records ---------------------------------------------------- */
static void rec_add(OrSim *s, int type, int variant, int nval, const double *v) {
  if (s->recs && s->nrec < s->max_recs) {
    sipnet_gpu_event_record *r = &s->recs[s->nrec];
    memset(r, 0, sizeof *r);
    r->step = (int32_t)s->step;
    r->type = type;
    r->variant = variant;
    r->nval = nval;
    for (int k = 0; k < nval; ++k) r->val[k] = v[k];
  }
  s->nrec++;
}

/* ---- dependency functions: depeffects.c ------------------------------------ */
static double dep_water_frac(double water, double whc) { return or_clip01(water / whc); } /* :11 */
static double dep_anaerobic_index(const OrSim *s, double water, double whc) { /* :15-22 */
  double f = dep_water_frac(water, whc);
  double fa = P(fAnoxia);
  return or_clip01((f - fa) / (1 - fa));
}
static double dep_resp_moist(const OrSim *s, double water, double whc) { /* :24-63 */
  if (!s->f.waterHResp || s->c.tsoil < 0) {
    return 1.0;
  }
  double f = dep_water_frac(water, whc);
  if (!s->f.anaerobic) {
    return pow(f, P(soilRespMoistEffect));
  }
  double dAer = or_clip01(f / P(fAnoxia));
  double a = dep_anaerobic_index(s, water, whc);
  return (1 - a) * dAer + P(anaerobicDecompRate) * a;
}
static double dep_methane_moist(const OrSim *s, double water, double whc) { /* :65-70 */
  return pow(dep_anaerobic_index(s, water, whc), P(anaerobicTransExp));
}
static double dep_temp(const OrSim *s, double tsoil) { return pow(P(soilRespQ10), tsoil / 10); } /* :72-75 */
static double dep_tillage(const OrSim *s) { return 1 + s->dTillMod; }                            /* :77 */
static double dep_cn(const OrSim *s, double kCN, double c, double n) {                          /* :79-88 */
  if (!s->f.nitrogenCycle) {
    return 1.0;
  }
  return kCN / (kCN + or_ratio(c, n));
}
static double dep_vol_moist(const OrSim *s, double water, double whc) { /* :90-96 */
  double a = dep_anaerobic_index(s, water, whc);
  return 0.05 + 3.8 * a * (1 - a);
}

/* ---- nitrogen helpers: nitrogen.c ------------------------------------------ */
static double n_leafon_from_c(const OrSim *s, double c) { /* nitrogen.c:86-88 */
  return fmax(0.0, c / P(leafCN) - c / P(woodCN));
}
static double n_demand(const OrSim *s) { /* nitrogen.c:91-106 */
  if (!s->f.nitrogenCycle) {
    return 0.0;
  }
  double d = s->r.woodCreation / P(woodCN) + s->r.leafCreation / P(leafCN) +
             s->r.fineRootCreation / P(fineRootCN) + s->r.coarseRootCreation / P(woodCN);
  return fmax(0.0, d);
}
static double n_non_uptake(const OrSim *s) { /* nitrogen.c:124-126 */
  return s->r.nMin - s->r.nVolatilization - s->r.nLeaching;
}
static double n_unclaimed_storage(const OrSim *s) { /* nitrogen.c:129-136 */
  double cflux = s->r.leafOnCreation + s->r.eventLeafOnCreation;
  double nflux = n_leafon_from_c(s, cflux);
  return fmax(0.0, s->e.plantStorageN - nflux * s->c.length);
}
static double n_fix_frac(const OrSim *s) { /* nitrogen.c:139-153 */
  double inhib;
  double denom = P(halfNFixationMax) + s->e.minN;
  if (denom < OR_TINY) {
    inhib = 1;
  } else {
    inhib = P(halfNFixationMax) / denom;
  }
  return P(nFixationFracMax) * inhib;
}
static void n_fix_and_uptake(OrSim *s) { /* nitrogen.c:156-168 */
  double demand = n_demand(s);
  double storage = n_unclaimed_storage(s) / s->c.length;
  double rem = fmax(0.0, demand - storage);
  double ff = n_fix_frac(s);
  s->r.nFixation = ff * rem;
  s->r.nUptake = (1 - ff) * rem;
}

/* ---- limitations.c:13-64 ---------------------------------------------------- */
static void limit_leaf_on(OrSim *s, double *flux) {
  double demandC = *flux * s->c.length;
  if (demandC < OR_TINY) {
    return;
  }
  double availC = (s->e.plantWoodC + s->e.coarseRootC) * P(leafOnReallocFrac);
  double cLim = availC / demandC;
  double nLim = 1.0;
  if (s->f.nitrogenCycle) {
    double demandN = n_leafon_from_c(s, demandC);
    if (demandN > OR_TINY) {
      nLim = s->e.plantStorageN / demandN;
    }
  }
  double lim = or_clip01(fmin(cLim, nLim));
  if (lim < 1) {
    *flux *= lim;
    s->counts[SIPNET_GPU_CNT_LEAFON_LIMITED]++;
    s->info |= SIPNET_GPU_ST_LEAFON_LIMITED; /* logInfo("Leaf on creation ... exceeds available ..."), limitations.c:48-61 */
  }
}

/* ---- events.c:449-742 processEvents ----------------------------------------- */
static void step_events(OrSim *s) {
  const int cy = s->c.year, cd = s->c.day;
  const double len = s->c.length;
  if (len <= 0) { /* events.c:460-465 */
    s->exit_code = SIPNET_GPU_ERR_BAD_PARAMETER_VALUE;
    return;
  }
  s->harvRemoved = 0;
  s->harvTransferred = 0;
  while (s->evpos < s->nev && s->ev[s->evpos].year <= cy && s->ev[s->evpos].day <= cd) { /* :471 */
    const sipnet_gpu_event *ev = &s->ev[s->evpos];
    if (ev->year < cy || ev->day < cd) { /* :476-481 */
      s->exit_code = SIPNET_GPU_ERR_INPUT_FILE;
      return;
    }
    switch (ev->type) {
      case SIPNET_EV_IRRIGATION: { /* :484-506 */
        const double amount = ev->p[0];
        double soilAmt, evapAmt;
        if (ev->method == 0) {
          evapAmt = P(immedEvapFrac) * amount;
          soilAmt = amount - evapAmt;
        } else if (ev->method == 1) {
          evapAmt = 0.0;
          soilAmt = amount;
        } else {
          s->exit_code = SIPNET_GPU_ERR_UNKNOWN_EVENT;
          return;
        }
        s->r.eventEvap += evapAmt / len;
        s->r.eventSoilWater += soilAmt / len;
        double v[2] = {soilAmt, evapAmt};
        rec_add(s, ev->type, 0, 2, v);
      } break;
      case SIPNET_EV_PLANTING: { /* :507-542 */
        const double leafC = ev->p[0], woodC = ev->p[1], fineC = ev->p[2], coarseC = ev->p[3];
        s->r.eventLeafC += leafC / len;
        s->r.eventWoodC += woodC / len;
        s->r.eventFineRootC += fineC / len;
        s->r.eventCoarseRootC += coarseC / len;
        const double inC = leafC + woodC + fineC + coarseC;
        double inN = 0.0;
        s->r.eventInputC += inC / len;
        if (s->f.nitrogenCycle) {
          inN = leafC / P(leafCN) + woodC / P(woodCN) + fineC / P(fineRootCN) + coarseC / P(woodCN);
          s->r.eventInputN += inN / len;
        }
        double v[6] = {leafC, woodC, fineC, coarseC, inC, inN};
        rec_add(s, ev->type, 0, 6, v);
      } break;
      case SIPNET_EV_HARVEST: { /* :543-635 */
        const double fRA = ev->p[0], fRB = ev->p[1], fTA = ev->p[2], fTB = ev->p[3];
        const double woodC = s->e.plantWoodC + s->e.plantCAccountingDelta;
        double above = woodC + s->e.plantLeafC;
        double below = s->e.fineRootC + s->e.coarseRootC;
        double total = above + below;
        if (total > OR_TINY) {
          double removed = fRA * above + fRB * below;
          double moved = fTA * above + fTB * below;
          s->harvRemoved += removed / total;
          s->harvTransferred += moved / total;
        }
        double litterAdd = fTA * (s->e.plantLeafC + woodC);
        double soilAdd = fTB * (s->e.fineRootC + s->e.coarseRootC);
        const double dLeaf = -s->e.plantLeafC * (fRA + fTA);
        const double dWood = -woodC * (fRA + fTA);
        const double dFine = -s->e.fineRootC * (fRB + fTB);
        const double dCoarse = -s->e.coarseRootC * (fRB + fTB);
        if (!s->f.litterPool) {
          soilAdd += litterAdd;
          litterAdd = 0.0;
        }
        s->r.eventLitterC += litterAdd / len;
        s->r.eventSoilC += soilAdd / len;
        s->r.eventLeafC += dLeaf / len;
        s->r.eventWoodC += dWood / len;
        s->r.eventFineRootC += dFine / len;
        s->r.eventCoarseRootC += dCoarse / len;
        double litterNAdd = 0.0, soilNAdd = 0.0;
        if (s->f.nitrogenCycle) {
          const double nAbove = (s->e.plantLeafC / P(leafCN)) + (s->e.plantWoodC / P(woodCN));
          const double nBelow = (s->e.fineRootC / P(fineRootCN)) + (s->e.coarseRootC / P(woodCN));
          litterNAdd = fTA * nAbove;
          soilNAdd = fTB * nBelow;
          s->r.eventSoilOrgN += soilNAdd / len;
          s->r.eventLitterN += litterNAdd / len;
        }
        const double outC = ((woodC + s->e.plantLeafC) * fRA + (s->e.fineRootC + s->e.coarseRootC) * fRB);
        double outN = 0.0;
        s->r.eventOutputC += outC / len;
        if (s->f.nitrogenCycle) {
          outN = (s->e.plantWoodC / P(woodCN) + s->e.plantLeafC / P(leafCN)) * fRA +
                 (s->e.fineRootC / P(fineRootCN) + s->e.coarseRootC / P(woodCN)) * fRB;
          s->r.eventOutputN += outN / len;
        }
        double v[10] = {soilAdd, litterAdd, dLeaf, dWood, dFine, dCoarse, soilNAdd, litterNAdd, outC, outN};
        rec_add(s, ev->type, 0, 10, v);
      } break;
      case SIPNET_EV_TILLAGE: { /* :636-646 */
        s->dTillMod += ev->p[0];
        double v[1] = {ev->p[0]};
        rec_add(s, ev->type, 0, 1, v);
      } break;
      case SIPNET_EV_FERTILIZATION: { /* :647-685 */
        const double orgC = ev->p[1];
        double orgN = 0.0, minN = 0.0;
        if (s->f.nitrogenCycle) {
          orgN = ev->p[0];
          minN = ev->p[2];
        }
        if (s->f.litterPool) {
          s->r.eventLitterC += orgC / len;
        } else {
          s->r.eventSoilC += orgC / len;
        }
        if (s->f.nitrogenCycle) {
          s->r.eventLitterN += orgN / len;
          s->r.eventMinN += minN / len;
        }
        s->r.eventInputC += orgC / len;
        if (s->f.nitrogenCycle) {
          s->r.eventInputN += (orgN + minN) / len;
        }
        double v[6] = {s->f.litterPool ? orgC : 0.0, s->f.litterPool ? 0.0 : orgC, minN, orgN, orgC, (orgN + minN)};
        rec_add(s, ev->type, 0, 6, v);
      } break;
      case SIPNET_EV_LEAFON: { /* :686-705 */
        double flux = P(leafGrowth) / len;
        limit_leaf_on(s, &flux);
        s->r.eventLeafOnCreation += flux;
        double src = s->e.plantWoodC + s->e.coarseRootC;
        if (src > OR_TINY) {
          s->r.eventLeafOnCreationFromWood += flux * s->e.plantWoodC / src;
        }
      } break;
      case SIPNET_EV_LEAFOFF: { /* :706-728 */
        double leafOff = s->e.plantLeafC * P(fracLeafFall);
        s->r.eventLeafOffLitter += leafOff / len;
        double litterNAdd = 0.0, resorb = 0.0;
        if (s->f.nitrogenCycle) {
          double leafN = leafOff / P(leafCN);
          resorb = leafN * P(leafNResorptionFrac);
          litterNAdd = leafN - resorb;
          s->r.eventLeafOffNResorption += resorb / len;
          s->r.eventLitterN += litterNAdd / len;
        }
        double v[3] = {leafOff, resorb, litterNAdd};
        rec_add(s, ev->type, 1, 3, v);
      } break;
      case SIPNET_EV_PLANTDEATH: /* :729-734, ignored with a warning */
        break;
      default: /* :735-737 */
        s->exit_code = SIPNET_GPU_ERR_UNKNOWN_EVENT;
        return;
    }
    s->evpos++;
  }
}

/* ---- canopy light integral: sipnet.c:517-570 -------------------------------- */
static double light_effect(const OrSim *s, double lai, double par) {
  if (!(lai > 0 && par > 0)) {
    return 0;
  }
  const int layers = 6;
  double cum = 0.0, cur = 0.0;
  int coeff = 1;
  for (int layer = 0; layer <= layers; ++layer) {
    double cumLai = lai * ((double)layer / layers);
    double inten = par * exp(-1.0 * P(attenuation) * cumLai);
    cur = (1 - pow(2, (-1.0 * inten / P(halfSatPar))));
    cum += coeff * cur;
    coeff = 2 * (1 + (layer + 1) % 2);
  }
  cum -= cur;
  return cum / (3.0 * layers);
}

/* ---- phenology triggers: sipnet.c:705-742 ----------------------------------- */
static int past_leaf_growth(const OrSim *s) {
  if (s->f.gdd) {
    double g = s->c.gdd;
    if (s->c.year == s->t.lastYear) {
      g += s->t.gdd;
    }
    return g >= P(gddLeafOn);
  }
  if (s->f.soilPhenol) {
    return s->c.tsoil >= P(soilTempLeafOn);
  }
  if (P(leafOnDay) > 0) {
    double now = (double)s->c.day + s->c.time / 24.0;
    return now >= P(leafOnDay);
  }
  return 0;
}
static int past_leaf_fall(const OrSim *s) {
  if (P(leafOffDay) > 0) {
    return (s->c.day + s->c.time / 24.0) >= P(leafOffDay);
  }
  return 0;
}

/* ---- calculateFluxes: sipnet.c:1256-1336 ------------------------------------- */
static void step_fluxes(OrSim *s) {
  OrRates *r = &s->r;
  const OrClim *c = &s->c;
  const double len = c->length;

  double lai = s->e.plantLeafC / P(leafCSpWt); /* :1274 */

  /* potPsn, :590-641 */
  double respPerGram = P(baseFolRespFrac) * P(aMax);
  double grossAMax = P(aMax) * P(aMaxFrac) + respPerGram;
  double dTemp = (P(psnTMax) - c->tair) * (c->tair - P(psnTMin)) / pow((P(psnTMax) - P(psnTMin)) / 2.0, 2);
  dTemp = fmax(dTemp, 0.0);
  double dVpd = 1.0 - P(dVpdSlope) * pow(c->vpd, P(dVpdExp));
  dVpd = fmax(dVpd, 0.0);
  double dLight = light_effect(s, lai, c->par);
  double conv = 12.0 * (1.0 / 1000000000.0) * (P(leafCSpWt) / P(cFracLeaf)) * lai * 86400.0;
  double potPsn = grossAMax * dTemp * dVpd * dLight * conv;
  double baseFolResp = respPerGram * conv;

  /* moisture, :656-699 */
  double dWater;
  if (potPsn < OR_TINY) {
    r->transpiration = 0.0;
    dWater = 1;
  } else {
    double wue = P(wueConst) / c->vpd;
    double potTrans = potPsn / wue * 1000.0 * (44.0 / 12.0) * (1.0 / 10000.0);
    double removable = fmin(s->e.soilWater, P(soilWHC)) * P(waterRemoveFrac);
    if (c->tsoil < P(frozenSoilThreshold)) {
      removable *= P(frozenSoilEff);
    }
    r->transpiration = fmin(removable, potTrans);
    dWater = r->transpiration / potTrans;
  }

  /* calcPrecip, :848-882 */
  if (c->tair <= 0) {
    r->snowFall = c->precip / len;
    r->rain = 0;
  } else {
    r->snowFall = 0;
    r->rain = c->precip / len;
  }
  if (s->f.leafWater) {
    double maxPool = lai * P(leafPoolDepth);
    r->immedEvap = r->rain * P(immedEvapFrac);
    if (r->immedEvap > maxPool) {
      r->immedEvap = maxPool;
    }
  } else {
    r->immedEvap = r->rain * P(immedEvapFrac);
  }
  double netRain = r->rain - r->immedEvap; /* :1281 */

  /* snowPack, :888-946 */
  {
    const double k = (1.3 * 1005.) / 66. * (1. / 2835000.) * 1000. * 1000. * (1. / 10000) * 86400.0;
    if (s->e.snow <= 0) {
      r->snowMelt = 0;
      r->sublimation = 0;
    } else {
      double rd = P(rdConst) / c->wspd;
      r->sublimation = k * (0.6 - c->vPress) / rd;
      double left = s->e.snow + (r->snowFall * len);
      if (r->sublimation < 0) {
        r->sublimation = 0;
      }
      if (left - (r->sublimation * len) < 0) {
        r->sublimation = left / len;
        left = 0;
      } else {
        left -= (r->sublimation * len);
      }
      if (c->tair <= 0) {
        r->snowMelt = 0;
      } else {
        r->snowMelt = P(snowMelt) * c->tair;
        if (left - (r->snowMelt * len) < 0) {
          r->snowMelt = left / len;
        }
      }
    }
  }

  /* calcSoilWaterFluxes, :963-1031 */
  {
    const double k = (1.3 * 1005.) / 66. * (1. / 2501000.) * 1000. * 1000. * (1. / 10000) * 86400.0;
    const double water = s->e.soilWater;
    double netIn = netRain + r->snowMelt;
    r->fastFlow = netIn * P(fastFlowFrac);
    netIn -= r->fastFlow;
    double left = water + netIn * len - r->transpiration * len;
    if (s->e.snow > 0) {
      r->evaporation = 0;
    } else {
      double wf = dep_water_frac(water, P(soilWHC));
      double rd = P(rdConst) / c->wspd;
      double rsoil = exp(P(rSoilConst1) - P(rSoilConst2) * wf);
      r->evaporation = k * c->vpdSoil / (rd + rsoil);
      if (r->evaporation < 0) {
        r->evaporation = 0;
      }
      if (left - (r->evaporation * len) < OR_TINY) {
        r->evaporation = (left - OR_TINY) / len;
        left = 0;
      } else {
        left -= (r->evaporation * len);
      }
    }
    if (left > P(soilWHC)) {
      double excess = left - P(soilWHC);
      if (s->f.flooding) {
        r->drainage = fmin(excess * P(waterDrainFrac), excess / len);
      } else {
        r->drainage = excess / len;
      }
    } else {
      r->drainage = 0;
    }
  }

  r->photosynthesis = potPsn * dWater; /* getGpp :1034 */

  /* vegResp / vegResp2, :1051-1103 */
  {
    double fol = baseFolResp * pow(P(vegRespQ10), (c->tair - P(psnTOpt)) / 10.0);
    if (c->tsoil < P(frozenSoilThreshold)) {
      fol *= P(frozenSoilFolREff);
    }
    double wood = P(baseVegResp) * or_wood_total(s) * pow(P(vegRespQ10), c->tair / 10.0);
    if (s->f.growthResp) {
      double growth = P(growthRespFrac) * ring_mean(&s->ring);
      if (growth < 0) {
        growth = 0;
      }
      r->rVeg = fol + wood + growth;
    } else {
      r->rVeg = fol + wood;
    }
  }

  /* calcWoodAndLeafFluxes, :756-782 */
  {
    r->woodLitter += or_wood_total(s) * P(woodTurnoverRate);
    double ll = s->e.plantLeafC * P(leafTurnoverRate);
    r->leafLitter += ll;
    double npp = ring_mean(&s->ring);
    double lc = npp * P(leafAllocation);
    double wc = npp * P(woodAllocation);
    r->leafCreation += lc;
    r->woodCreation += wc;
  }

  /* calcLeafOnOffFluxes, :800-842 */
  {
    if (c->year > s->phenLastYear) {
      s->didLeafGrowth = 0;
      s->didLeafFall = 0;
      s->phenLastYear = c->year;
    }
    if (!s->didLeafGrowth && past_leaf_growth(s)) {
      double on = P(leafGrowth) / len;
      limit_leaf_on(s, &on);
      r->leafOnCreation += on;
      double src = s->e.plantWoodC + s->e.coarseRootC;
      if (src > OR_TINY) {
        r->leafOnCreationFromWood += on * s->e.plantWoodC / src;
      }
      s->didLeafGrowth = 1;
    }
    if (!s->didLeafFall && past_leaf_fall(s)) {
      double off = (s->e.plantLeafC * P(fracLeafFall)) / len;
      r->leafLitter += off;
      s->didLeafFall = 1;
      if (off > OR_TINY && s->f.events) {
        double v[1] = {off * len};
        rec_add(s, SIPNET_EV_LEAFOFF, 0, 1, v);
      }
    }
  }

  /* calcLitterFluxes, :1150-1171 */
  if (s->f.litterPool) {
    double te = dep_temp(s, c->tsoil);
    double me = dep_resp_moist(s, s->e.soilWater, P(soilWHC));
    double ti = dep_tillage(s);
    double cn = dep_cn(s, P(kCN), s->e.litterC, s->e.litterN);
    double breakdown = s->e.litterC * P(litterBreakdownRate) * te * me * ti * cn;
    r->rLitter = breakdown * P(fracLitterRespired);
    r->litterToSoil = breakdown * (1.0 - P(fracLitterRespired));
  } else {
    r->rLitter = 0;
    r->litterToSoil = 0;
  }

  /* calcRootFluxes, :1176-1196 */
  {
    r->coarseRootLoss += P(coarseRootTurnoverRate) * s->e.coarseRootC;
    r->fineRootLoss += P(fineRootTurnoverRate) * s->e.fineRootC;
    double npp = ring_mean(&s->ring);
    double cc = P(coarseRootAllocation) * npp;
    double fc = P(fineRootAllocation) * npp;
    r->coarseRootCreation += cc;
    r->fineRootCreation += fc;
    r->rCoarseRoot = P(baseCoarseRootResp) * s->e.coarseRootC * pow(P(coarseRootQ10), c->tsoil / 10.0);
    r->rFineRoot = P(baseFineRootResp) * s->e.fineRootC * pow(P(fineRootQ10), c->tsoil / 10.0);
  }

  /* calcSoilRespiration, :1132-1148 */
  {
    double me = dep_resp_moist(s, s->e.soilWater, P(soilWHC));
    double te = dep_temp(s, c->tsoil);
    double ti = dep_tillage(s);
    double cn = dep_cn(s, P(kCN), s->e.soilC, s->e.soilOrgN);
    r->rSoil = s->e.soilC * P(baseSoilResp) * me * te * ti * cn;
  }

  /* calcMethaneFlux, :1201-1214 */
  if (s->f.anaerobic) {
    double te = dep_temp(s, c->tsoil);
    double me = dep_methane_moist(s, s->e.soilWater, P(soilWHC));
    r->soilMethane = P(soilMethaneRate) * s->e.soilC * te * me;
    if (s->f.litterPool) {
      r->litterMethane = P(litterMethaneRate) * s->e.litterC * te * me;
    } else {
      r->litterMethane = 0.0;
    }
  }

  /* checkNegativeCreation, limitations.c:146-182 */
  {
    double turnover = s->e.plantLeafC * P(leafTurnoverRate);
    double leafDef = s->e.plantLeafC / len + r->leafCreation - turnover;
    if (leafDef < 0) {
      r->woodCreation += leafDef;
      r->leafCreation -= leafDef;
    }
    double fineDef = s->e.fineRootC / len + r->fineRootCreation - r->fineRootLoss;
    double coarseDef = s->e.coarseRootC / len + r->coarseRootCreation - r->coarseRootLoss;
    if ((fineDef < 0.0) != (coarseDef < 0.0)) {
      if (fineDef < 0.0) {
        r->coarseRootCreation += fineDef;
        r->fineRootCreation -= fineDef;
      }
      if (coarseDef < 0.0) {
        r->fineRootCreation += coarseDef;
        r->coarseRootCreation -= coarseDef;
      }
    }
  }

  if (s->f.nitrogenCycle) {
    /* calcNResorptionFluxes, nitrogen.c:170-196 */
    if (r->woodCreation + r->leafCreation + r->fineRootCreation + r->coarseRootCreation < 0.0) {
      r->reductionNResorption -= (r->leafCreation / P(leafCN) + r->woodCreation / P(woodCN) +
                                  r->coarseRootCreation / P(woodCN) + r->fineRootCreation / P(fineRootCN));
    }
    double resorb = P(leafNResorptionFrac) * r->leafLitter / P(leafCN);
    r->leafOffNResorption += resorb;
    /* calcNVolatilizationFlux, nitrogen.c:15-25 */
    {
      double dt = dep_temp(s, c->tsoil);
      double dw = dep_vol_moist(s, s->e.soilWater, P(soilWHC));
      r->nVolatilization = P(nVolatilizationFrac) * s->e.minN * dt * dw;
    }
    /* calcNLeachingFlux, nitrogen.c:30-40 */
    {
      double phi;
      if ((r->drainage / P(soilWHC)) < 1) {
        phi = r->drainage / P(soilWHC);
      } else {
        phi = 1;
      }
      r->nLeaching = s->e.minN * phi * P(nLeachingFrac);
    }
    /* calcNPoolFluxes, nitrogen.c:45-83 */
    {
      double litterCN = or_ratio(s->e.litterC, s->e.litterN);
      double soilCN = or_ratio(s->e.soilC, s->e.soilOrgN);
      double litterMin = r->rLitter / litterCN;
      double soilMin = r->rSoil / soilCN;
      double inputs = r->litterToSoil / litterCN + r->fineRootLoss / P(fineRootCN) + r->coarseRootLoss / P(woodCN);
      double sat = s->f.carbonSaturation ? or_clip01(s->e.soilC / P(soilCSaturation)) : 0.0;
      r->nOrgLitter = r->leafLitter / P(leafCN) - r->leafOffNResorption + r->woodLitter / P(woodCN) - litterMin -
                      r->litterToSoil / litterCN + (inputs * sat);
      r->nOrgSoil = inputs * (1 - sat) - soilMin;
      r->nMin = litterMin + soilMin;
    }
    n_fix_and_uptake(s); /* nitrogen.c:156-168 */

    /* checkMineralNLimitation, limitations.c:119-130 */
    {
      double pool = s->e.minN + (r->nMin + r->eventMinN) * len;
      double loss = (r->nLeaching + r->nVolatilization) * len;
      if (loss > OR_TINY && loss > pool) {
        double red = pool / loss;
        r->nLeaching *= red;
        r->nVolatilization *= red;
        s->counts[SIPNET_GPU_CNT_MINN_LIMITED]++;
        s->info |= SIPNET_GPU_ST_MINN_LIMITED;
      }
    }
    /* checkNitrogenLimitation, limitations.c:69-114 */
    {
      double uptakeDemand = r->nUptake * len;
      double nonUptake = n_non_uptake(s) * len;
      double avail = s->e.minN + nonUptake;
      if (uptakeDemand > OR_TINY && uptakeDemand > avail) {
        double unclaimed = n_unclaimed_storage(s);
        double demand = n_demand(s) * len;
        double uptakeFrac = 1 - n_fix_frac(s);
        double red = (avail / uptakeFrac + unclaimed) / demand;
        s->counts[SIPNET_GPU_CNT_N_LIMITED]++;
        s->info |= SIPNET_GPU_ST_N_LIMITED; /* logInfo("N limitation: ..."), limitations.c:98-102 */
        r->woodCreation *= red;
        r->leafCreation *= red;
        r->fineRootCreation *= red;
        r->coarseRootCreation *= red;
        n_fix_and_uptake(s);
      }
    }
  }

  /* writeLeafOnEventIfNeeded, sipnet.c:1230-1247 */
  if (r->leafOnCreation > OR_TINY && s->f.events) {
    double v[2] = {r->leafOnCreation * len, r->leafOnCreationFromWood * len};
    rec_add(s, SIPNET_EV_LEAFON, 0, 2, v);
  }
  if (r->eventLeafOnCreation > OR_TINY && s->f.events) {
    double v[2] = {r->eventLeafOnCreation * len, r->eventLeafOnCreationFromWood * len};
    rec_add(s, SIPNET_EV_LEAFON, 1, 2, v);
  }
}

/* ---- biomass test: sipnet.c:1530-1536 --------------------------------------- */
static int enough_biomass(const OrSim *s) {
  double wood = or_wood_total(s);
  double root = s->e.fineRootC + s->e.coarseRootC;
  return s->e.plantWoodC > OR_TINY && wood > OR_TINY && root > OR_TINY;
}

static void clamp_stock(OrSim *s, double *v, double floor_) { /* ensureNonNegative, sipnet.c:1346-1356 */
  if (*v < floor_) {
    if (fabs(*v) > 1e-8) s->counts[SIPNET_GPU_CNT_CLAMPED]++; /* the logWarning (EPS, balance.h:6) */
    *v = 0.;
  }
}

/* getMassTotals, balance.c:13-33 */
static void mass_totals(const OrSim *s, double *carbon, double *nitrogen) {
  const OrPools *e = &s->e;
  *carbon = (e->plantWoodC + e->plantCAccountingDelta) + e->plantLeafC + e->fineRootC + e->coarseRootC + e->soilC;
  if (s->f.litterPool) {
    *carbon += e->litterC;
  }
  if (s->f.nitrogenCycle) {
    *nitrogen = e->plantWoodC / P(woodCN) + e->plantLeafC / P(leafCN) + e->fineRootC / P(fineRootCN) +
                e->coarseRootC / P(woodCN) + e->soilOrgN + e->litterN + e->minN + e->plantStorageN;
  } else {
    *nitrogen = 0.0;
  }
}

/* ---- updatePoolsAndBalance: sipnet.c:1769-1806.  The balance tracker (balance.c) is diagnostic -- nothing in it
 * feeds back into state -- and is restated for its two check values, deltaC and deltaN. ------------------------- */
static void step_pools(OrSim *s) {
  OrPools *e = &s->e;
  const OrRates *r = &s->r;
  const double len = s->c.length;
  double preC, preN, postC, postN, finalC, finalN;
  mass_totals(s, &preC, &preN); /* updateBalanceTrackerPreUpdate, balance.c:35-38 */

  /* updatePoolsForEvents, events.c:744-790 */
  e->plantWoodC += r->eventWoodC * len;
  e->plantLeafC += r->eventLeafC * len;
  e->soilC += r->eventSoilC * len;
  if (s->f.litterPool) {
    e->litterC += r->eventLitterC * len;
  }
  e->plantWoodC -= r->eventLeafOnCreationFromWood * len;
  double evFromRoot = r->eventLeafOnCreation - r->eventLeafOnCreationFromWood;
  e->coarseRootC -= evFromRoot * len;
  e->plantLeafC += (r->eventLeafOnCreation - r->eventLeafOffLitter) * len;
  if (s->f.litterPool) {
    e->litterC += r->eventLeafOffLitter * len;
  } else {
    e->soilC += r->eventLeafOffLitter * len;
  }
  e->coarseRootC += r->eventCoarseRootC * len;
  e->fineRootC += r->eventFineRootC * len;
  e->soilWater += r->eventSoilWater * len;
  if (s->f.nitrogenCycle) {
    e->minN += r->eventMinN * len;
    e->soilOrgN += r->eventSoilOrgN * len;
    e->litterN += r->eventLitterN * len;
    double onN = n_leafon_from_c(s, r->eventLeafOnCreation);
    e->plantStorageN += (r->eventLeafOffNResorption - onN) * len;
  }

  /* updateMainPools, sipnet.c:1579-1626 */
  {
    double ra = r->rVeg + r->rFineRoot + r->rCoarseRoot;
    double alloc = r->leafCreation + r->woodCreation + r->fineRootCreation + r->coarseRootCreation;
    e->plantCAccountingDelta += ((r->photosynthesis - ra) - alloc) * len;
    e->plantWoodC += (r->woodCreation - r->woodLitter - r->leafOnCreationFromWood) * len;
    e->plantLeafC += (r->leafCreation + r->leafOnCreation - r->leafLitter) * len;
    e->soilWater += (r->rain + r->snowMelt - r->immedEvap - r->fastFlow - r->evaporation - r->transpiration -
                     r->drainage) *
                    len;
    e->snow += (r->snowFall - r->snowMelt - r->sublimation) * len;
  }

  /* updatePoolsForSoil, sipnet.c:1634-1680 */
  if (s->f.litterPool) {
    double inputs = r->coarseRootLoss + r->fineRootLoss + r->litterToSoil;
    double sat = s->f.carbonSaturation ? or_clip01(e->soilC / P(soilCSaturation)) : 0.0;
    e->litterC +=
        (r->woodLitter + r->leafLitter + (inputs * sat) - r->litterToSoil - r->rLitter - r->litterMethane) * len;
    e->soilC += (inputs * (1 - sat) - r->rSoil - r->soilMethane) * len;
  } else {
    e->soilC +=
        (r->coarseRootLoss + r->fineRootLoss + r->woodLitter + r->leafLitter - r->rSoil - r->soilMethane) * len;
  }
  {
    double fromRoot = r->leafOnCreation - r->leafOnCreationFromWood;
    e->coarseRootC += (r->coarseRootCreation - r->coarseRootLoss - fromRoot) * len;
    e->fineRootC += (r->fineRootCreation - r->fineRootLoss) * len;
  }

  /* updateNitrogenPools, nitrogen.c:210-239 */
  if (s->f.nitrogenCycle) {
    double demand = n_demand(s);
    double fromStorage = demand - r->nUptake - r->nFixation;
    double onN = n_leafon_from_c(s, r->leafOnCreation);
    e->plantStorageN += (r->leafOffNResorption + r->reductionNResorption - fromStorage - onN) * len;
    double nonUptake = n_non_uptake(s);
    e->minN += (nonUptake - r->nUptake) * len;
    e->soilOrgN += r->nOrgSoil * len;
    e->litterN += r->nOrgLitter * len;
  }

  mass_totals(s, &postC, &postN); /* updateBalanceTrackerPostUpdate, balance.c:40-43 */

  /* checkForMortality, sipnet.c:1688-1767 */
  if (!s->isAlive) {
    if (enough_biomass(s)) {
      s->isAlive = 1;
    }
  } else if (!enough_biomass(s)) {
    s->isAlive = 0;
    double wood = or_wood_total(s);
    double root = e->fineRootC + e->coarseRootC;
    e->soilC += root;
    if (s->f.litterPool) {
      e->litterC += e->plantWoodC + e->plantLeafC + e->plantCAccountingDelta;
    } else {
      e->soilC += e->plantWoodC + e->plantLeafC + e->plantCAccountingDelta;
    }
    if (s->f.nitrogenCycle) {
      e->soilOrgN += e->fineRootC / P(fineRootCN) + e->coarseRootC / P(woodCN);
      e->litterN += e->plantWoodC / P(woodCN) + e->plantLeafC / P(leafCN) + e->plantStorageN;
    }
    e->plantWoodC = 0.0;
    e->plantLeafC = 0.0;
    e->coarseRootC = 0.0;
    e->fineRootC = 0.0;
    e->plantCAccountingDelta = 0.0;
    if (s->f.nitrogenCycle) {
      e->plantStorageN = 0.0;
    }
    ring_reset(&s->ring, 0.0);
    if (s->f.events) {
      double v[4] = {s->harvRemoved, s->harvTransferred, wood, root};
      rec_add(s, SIPNET_EV_PLANTDEATH, 0, 4, v);
    }
  }

  /* ensureNonNegativeStocks, sipnet.c:1368-1397 */
  clamp_stock(s, &e->plantWoodC, 0);
  clamp_stock(s, &e->plantLeafC, 0);
  if (s->f.litterPool) {
    clamp_stock(s, &e->litterC, 0);
  }
  clamp_stock(s, &e->soilC, 0);
  clamp_stock(s, &e->coarseRootC, 0);
  clamp_stock(s, &e->fineRootC, 0);
  clamp_stock(s, &e->soilWater, 0);
  clamp_stock(s, &e->snow, OR_TINY);
  clamp_stock(s, &e->minN, 0);
  clamp_stock(s, &e->soilOrgN, 0);
  clamp_stock(s, &e->litterN, 0);
  clamp_stock(s, &e->plantStorageN, 0);

  /* updateBalanceTrackerPostClamp, balance.c:45-104 */
  mass_totals(s, &finalC, &finalN);
  double clampedC = finalC - postC;
  if (clampedC < OR_EPS) {
    clampedC = 0;
  }
  double clampedN = finalN - postN;
  if (clampedN < OR_EPS) {
    clampedN = 0;
  }
  double inputsC = r->photosynthesis + r->eventInputC;
  double outputsC = r->rVeg + r->rFineRoot + r->rCoarseRoot + r->rSoil + r->soilMethane + r->eventOutputC;
  if (s->f.litterPool) {
    outputsC += r->rLitter + r->litterMethane;
  }
  inputsC *= len;
  outputsC *= len;
  double inputsN = 0.0, outputsN = 0.0; /* initBalanceTracker, balance.c:106-127 (never written without nitrogen) */
  if (s->f.nitrogenCycle) {
    inputsN = r->nFixation + r->eventInputN;
    outputsN = r->nLeaching + r->nVolatilization + r->eventOutputN;
    inputsN *= len;
    outputsN *= len;
  }
  inputsC += clampedC;
  if (s->f.nitrogenCycle) {
    inputsN += clampedN;
  }
  /* checkBalance, balance.c:129-176 */
  double poolCDelta = finalC - preC;
  double systemCDelta = inputsC - outputsC;
  s->balDeltaC = poolCDelta - systemCDelta;
  double poolNDelta = finalN - preN;
  double systemNDelta = outputsN - inputsN;
  s->balDeltaN = poolNDelta + systemNDelta;
  if (fabs(s->balDeltaC) < OR_EPS) {
    s->balDeltaC = 0.0;
  }
  if (fabs(s->balDeltaN) < OR_EPS) {
    s->balDeltaN = 0.0;
  }
}

/* ---- updateTrackers: sipnet.c:1420-1496 -------------------------------------- */
static void step_trackers(OrSim *s, double oldSoilWater) {
  OrTrack *t = &s->t;
  const OrRates *r = &s->r;
  const double len = s->c.length;
  if (s->c.year != t->lastYear) {
    t->yearlyGpp = 0.0;
    t->yearlyRtot = 0.0;
    t->yearlyRa = 0.0;
    t->yearlyRh = 0.0;
    t->yearlyNpp = 0.0;
    t->yearlyNee = 0.0;
    t->gdd = 0.0;
    t->lastYear = s->c.year;
  }
  t->gpp = r->photosynthesis * len;
  t->rh = (r->rLitter + r->rSoil) * len;
  t->rAboveground = (r->rVeg) * len;
  t->rRoot = (r->rCoarseRoot + r->rFineRoot) * len;
  t->rSoil = t->rRoot + t->rh;
  t->ra = t->rRoot + t->rAboveground;
  t->rtot = t->ra + t->rh;
  t->npp = t->gpp - t->ra;
  t->nee = -1.0 * (t->npp - t->rh);
  t->yearlyGpp += t->gpp;
  t->yearlyRa += t->ra;
  t->yearlyRh += t->rh;
  t->yearlyRtot += t->rtot;
  t->yearlyNpp += t->npp;
  t->yearlyNee += t->nee;
  t->totGpp += t->gpp;
  t->totRa += t->ra;
  t->totRh += t->rh;
  t->totRtot += t->rtot;
  t->totNpp += t->npp;
  t->totNee += t->nee;
  t->woodCreation = r->woodCreation * len;
  t->methane = (r->soilMethane + r->litterMethane) * len;
  t->evapotranspiration =
      (r->transpiration + r->immedEvap + r->evaporation + r->sublimation + r->eventEvap) * len;
  t->soilWetnessFrac = (oldSoilWater + s->e.soilWater) / (2.0 * P(soilWHC));
  t->yearlyLitter += r->leafLitter + r->eventLeafOffLitter;
  if (s->f.gdd) {
    t->gdd += s->c.gdd;
  } else {
    t->gdd = 0.0;
  }
  t->meanNPP = ring_mean(&s->ring);
  if (s->f.nitrogenCycle) {
    t->n2o = r->nVolatilization * len;
    t->nLeaching = r->nLeaching * len;
    t->nFixation = r->nFixation * len;
    t->nUptake = r->nUptake * len;
  }
}

/* ---- updateState: sipnet.c:1818-1855 ----------------------------------------- */
static void step(OrSim *s) {
  double oldSoilWater = s->e.soilWater;
  memset(&s->r, 0, sizeof s->r);  /* resetFluxes :1222 */
  s->isAlive = enough_biomass(s); /* initPlantSurvivalTracker :1538 */
  step_events(s);
  if (s->exit_code) {
    return;
  }
  step_fluxes(s);
  step_pools(s);
  step_trackers(s, oldSoilWater);
  /* updateMeanTrackers :1546-1570 */
  if (s->isAlive) {
    double npp = s->r.photosynthesis - s->r.rVeg - s->r.rCoarseRoot - s->r.rFineRoot;
    if (ring_push(&s->ring, npp, s->c.length) != 0) {
      s->exit_code = SIPNET_GPU_ERR_INTERNAL;
      return;
    }
  }
  /* updateEventTrackers, events.c:811-822 */
  if (s->dTillMod > 0) {
    s->dTillMod *= exp(-s->c.length * (1 / 30.0));
    if (s->dTillMod < 0.01) {
      s->dTillMod = 0.0;
    }
  }
}

/* ---- setupModel: sipnet.c:1858-1951 (+ ensureAllocation :1111-1123) ----------- */
static int setup(OrSim *s) {
  P(coarseRootAllocation) = 1 - P(leafAllocation) - P(woodAllocation) - P(fineRootAllocation);
  if ((P(leafAllocation) >= 1.0) || (P(woodAllocation) >= 1.0) || (P(fineRootAllocation) >= 1.0) ||
      (P(coarseRootAllocation) < 0)) {
    return SIPNET_GPU_ERR_BAD_PARAMETER_VALUE;
  }
  P(baseVegResp) /= 365.0;
  P(litterBreakdownRate) /= 365.0;
  P(baseSoilResp) /= 365.0;
  P(woodTurnoverRate) /= 365.0;
  P(leafTurnoverRate) /= 365.0;
  P(psnTMax) = P(psnTOpt) + (P(psnTOpt) - P(psnTMin));
  s->e.plantWoodC = (1 - P(coarseRootFrac) - P(fineRootFrac)) * P(plantWoodInit);
  s->e.plantCAccountingDelta = 0.0;
  s->e.plantLeafC = P(laiInit) * P(leafCSpWt);
  s->e.litterC = s->f.litterPool ? P(litterInit) : 0.0;
  s->e.soilC = P(soilInit);
  P(fineRootTurnoverRate) /= 365.0;
  P(coarseRootTurnoverRate) /= 365.0;
  P(baseCoarseRootResp) /= 365.0;
  P(baseFineRootResp) /= 365.0;
  if (P(fAnoxia) <= 0.0) {
    P(fAnoxia) = OR_TINY;
  } else if (P(fAnoxia) >= 1.0) {
    P(fAnoxia) = 1.0 - OR_TINY;
  }
  if (P(anaerobicDecompRate) <= 0.0) {
    P(anaerobicDecompRate) = OR_TINY;
  } else if (P(anaerobicDecompRate) > 1.0) {
    P(anaerobicDecompRate) = 1.0;
  }
  s->e.coarseRootC = P(coarseRootFrac) * P(plantWoodInit);
  s->e.fineRootC = P(fineRootFrac) * P(plantWoodInit);
  s->e.soilWater = P(soilWFracInit) * P(soilWHC);
  if (s->e.soilWater < 0) {
    s->e.soilWater = 0;
  }
  s->e.snow = P(snowInit);
  if (s->f.nitrogenCycle) {
    s->e.minN = P(minNInit);
    s->e.soilOrgN = P(soilOrgNInit);
    s->e.litterN = P(litterOrgNInit);
    s->e.plantStorageN = P(plantStorageNInit);
  } else {
    s->e.minN = 0.0;
    s->e.soilOrgN = 0.0;
    s->e.litterN = 0.0;
    /* plantStorageN is left at its zero-initialised global value */
    s->e.plantStorageN = 0.0;
  }
  /* initTrackers, :1406-1413 */
  memset(&s->t, 0, sizeof s->t);
  s->t.soilWetnessFrac = s->e.soilWater / P(soilWHC);
  s->t.lastYear = -1;
  /* initPhenologyTrackers, :1501-1527 (uses the FIRST climate record) */
  s->didLeafGrowth = past_leaf_growth(s);
  s->didLeafFall = past_leaf_fall(s);
  if (s->didLeafFall && !s->didLeafGrowth) {
    s->didLeafGrowth = 1;
  }
  s->phenLastYear = s->c.year;
  s->dTillMod = 0.0; /* initEventTrackers, events.c:809 */
  s->harvRemoved = s->harvTransferred = 0.0;
  s->ring.totWeight = OR_MEAN_DAYS;
  ring_reset(&s->ring, 0);
  s->isAlive = 0;
  return 0;
}

static void write_out32(const OrSim *s, double *o) { /* outputState, sipnet.c:455-472 */
  o[0] = or_wood_total(s);
  o[1] = s->e.plantLeafC;
  o[2] = s->t.woodCreation;
  o[3] = s->e.soilC;
  o[4] = s->e.coarseRootC;
  o[5] = s->e.fineRootC;
  o[6] = s->e.litterC;
  o[7] = s->e.soilWater;
  o[8] = s->t.soilWetnessFrac;
  o[9] = s->e.snow;
  o[10] = s->t.npp;
  o[11] = s->t.nee;
  o[12] = s->t.totNee;
  o[13] = s->t.gpp;
  o[14] = s->t.rAboveground;
  o[15] = s->t.rSoil;
  o[16] = s->t.rRoot;
  o[17] = s->t.ra;
  o[18] = s->t.rh;
  o[19] = s->t.rtot;
  o[20] = s->t.evapotranspiration;
  o[21] = s->r.transpiration;
  o[22] = s->e.minN;
  o[23] = s->e.soilOrgN;
  o[24] = s->e.litterN;
  o[25] = s->e.plantStorageN;
  o[26] = s->t.n2o;
  o[27] = s->t.nLeaching;
  o[28] = s->t.nFixation;
  o[29] = s->t.nUptake;
  o[30] = s->t.methane;
  o[31] = s->e.plantCAccountingDelta;
}

static void write_debug(const OrSim *s, double *d) { /* debug_log.c:51-170 order */
  int k = 0;
  memcpy(d + k, &s->e, sizeof s->e);
  k += (int)(sizeof s->e / sizeof(double));
  memcpy(d + k, &s->r, sizeof s->r);
  k += (int)(sizeof s->r / sizeof(double));
  const OrTrack *t = &s->t;
  const double tv[33] = {t->gpp,        t->rtot,       t->ra,           t->rh,         t->rRoot,
                         t->rSoil,      t->rAboveground, t->npp,        t->nee,        t->woodCreation,
                         t->gdd,        t->evapotranspiration, t->soilWetnessFrac, t->yearlyGpp, t->yearlyRtot,
                         t->yearlyRa,   t->yearlyRh,   t->yearlyNpp,    t->yearlyNee,  t->yearlyLitter,
                         t->totGpp,     t->totRtot,    t->totRa,        t->totRh,      t->totNpp,
                         t->totNee,     (double)t->lastYear, t->methane, t->n2o,       t->nLeaching,
                         t->nFixation,  t->nUptake,    t->meanNPP};
  memcpy(d + k, tv, sizeof tv);
  k += 33;
  d[k++] = (double)s->didLeafGrowth;
  d[k++] = (double)s->didLeafFall;
  d[k++] = (double)s->phenLastYear;
  d[k++] = (double)s->isAlive;
}

static void load_clim(OrSim *s, int64_t t, const int32_t *year, const int32_t *day, const double *const *cl) {
  s->c.year = year[t];
  s->c.day = day[t];
  s->c.time = cl[0][t];
  s->c.length = cl[1][t];
  s->c.tair = cl[2][t];
  s->c.tsoil = cl[3][t];
  s->c.par = cl[4][t];
  s->c.precip = cl[5][t];
  s->c.vpd = cl[6][t];
  s->c.vpdSoil = cl[7][t];
  s->c.vPress = cl[8][t];
  s->c.wspd = cl[9][t];
  s->c.gdd = cl[10][t];
}

static int run_member(OrSim *s, const int32_t *flags, const double *params, int64_t pstride, int64_t T,
                      const int32_t *year, const int32_t *day, const double *const *cl, int64_t nev,
                      const sipnet_gpu_event *ev, double *out32, double *dbg, int64_t *steps_done,
                      double *final_out32, sipnet_gpu_event_record *recs, int32_t max_recs, double *balance) {
  memset(s, 0, sizeof *s);
  s->recs = recs;
  s->max_recs = max_recs;
  memcpy(&s->f, flags, sizeof s->f);
  for (int k = 0; k < SIPNET_GPU_NPARAMS; ++k) s->p[k] = params[k * pstride];
  s->ev = ev;
  s->nev = s->f.events ? nev : 0; /* initEvents only reads events when ctx.events, events.c:429 */
  int64_t done = 0;
  if (T > 0) {
    load_clim(s, 0, year, day, cl);
    /* frontend.c:217-222 -> isFirstEventBefore, events.c:437-447 */
    if (s->nev > 0) {
      const sipnet_gpu_event *e0 = &s->ev[0];
      int before = (e0->year != s->c.year) ? (e0->year < s->c.year) : (e0->day < s->c.day);
      if (before) {
        if (steps_done) *steps_done = 0;
        return SIPNET_GPU_ERR_INPUT_FILE;
      }
    }
    int rc = setup(s);
    if (rc) {
      if (steps_done) *steps_done = 0;
      return rc;
    }
  }
  for (int64_t t = 0; t < T; ++t) {
    load_clim(s, t, year, day, cl);
    s->step = t;
    step(s);
    if (s->exit_code) {
      break;
    }
    if (out32) write_out32(s, out32 + t * SIPNET_GPU_NOUT);
    if (dbg) write_debug(s, dbg + t * SIPNET_GPU_NDEBUG);
    if (balance) {
      balance[2 * t] = s->balDeltaC;
      balance[2 * t + 1] = s->balDeltaN;
    }
    ++done;
  }
  if (final_out32 && done > 0) write_out32(s, final_out32);
  if (steps_done) *steps_done = done;
  return s->exit_code;
}

int sipnet_oracle_run(const int32_t *flags, const double *params, int64_t T, const int32_t *year,
                      const int32_t *day, const double *time, const double *length, const double *tair,
                      const double *tsoil, const double *par, const double *precip, const double *vpd,
                      const double *vpdSoil, const double *vPress, const double *wspd, const double *gdd,
                      int64_t nev, const sipnet_gpu_event *ev, double *out32, double *dbg, int64_t *steps_done,
                      sipnet_gpu_event_record *recs, int32_t max_recs, int32_t *nrec) {
  const double *cl[11] = {time, length, tair, tsoil, par, precip, vpd, vpdSoil, vPress, wspd, gdd};
  OrSim *s = (OrSim *)malloc(sizeof(OrSim));
  if (!s) return SIPNET_GPU_ERR_INTERNAL;
  int rc = run_member(s, flags, params, 1, T, year, day, cl, nev, ev, out32, dbg, steps_done, NULL, recs, max_recs, NULL);
  if (nrec) *nrec = s->nrec;
  free(s);
  return rc;
}

/* One member's run with the balance tracker's check values: balance[t] = {deltaC, deltaN} (balance.c:129-148). */
int sipnet_oracle_run_balance(const int32_t *flags, const double *params, int64_t T, const int32_t *year,
                              const int32_t *day, const double *const *clim11, int64_t nev, const sipnet_gpu_event *ev,
                              double *balance, int64_t *steps_done) {
  return sipnet_oracle_run_diag(flags, params, T, year, day, clim11, nev, ev, balance, steps_done, NULL);
}

static __thread uint32_t g_last_counts[SIPNET_GPU_NCOUNTERS];
/* occurrence counts of the last sipnet_oracle_run_diag / _run_balance call on this thread (SIPNET_GPU_CNT_* order) */
void sipnet_oracle_last_counts(uint32_t *out) { memcpy(out, g_last_counts, sizeof(g_last_counts)); }

/* The same with the informational status bits of the whole run (SIPNET_GPU_ST_LEAFON_LIMITED, _N_LIMITED,
 * _MINN_LIMITED: the places where the reference prints a message or caps a flux, limitations.c). */
int sipnet_oracle_run_diag(const int32_t *flags, const double *params, int64_t T, const int32_t *year,
                           const int32_t *day, const double *const *clim11, int64_t nev, const sipnet_gpu_event *ev,
                           double *balance, int64_t *steps_done, uint32_t *info) {
  OrSim *s = (OrSim *)malloc(sizeof(OrSim));
  if (!s) return SIPNET_GPU_ERR_INTERNAL;
  int rc = run_member(s, flags, params, 1, T, year, day, clim11, nev, ev, NULL, NULL, steps_done, NULL, NULL, 0, balance);
  if (info) *info = s->info;
  memcpy(g_last_counts, s->counts, sizeof(g_last_counts));
  free(s);
  return rc;
}

/* ---- threaded ensemble (CPU baseline) ---------------------------------------- */
typedef struct {
  const int32_t *flags;
  const double *params;
  int64_t ld, m0, m1, T;
  const int32_t *year, *day;
  const double *const *cl;
  int64_t nev;
  const sipnet_gpu_event *ev;
  double *final_out32;
  int rc;
} OrJob;

static void *ensemble_worker(void *arg) {
  OrJob *j = (OrJob *)arg;
  OrSim *s = (OrSim *)malloc(sizeof(OrSim));
  j->rc = 0;
  for (int64_t m = j->m0; m < j->m1; ++m) {
    int64_t done = 0;
    int rc = run_member(s, j->flags, j->params + m, j->ld, j->T, j->year, j->day, j->cl, j->nev, j->ev, NULL,
                        NULL, &done, j->final_out32 + m * SIPNET_GPU_NOUT, NULL, 0, NULL);
    if (rc && !j->rc) j->rc = rc;
  }
  free(s);
  return NULL;
}

int sipnet_oracle_run_ensemble(const int32_t *flags, const double *params_soa, int64_t ld, int64_t nmembers,
                               int64_t T, const int32_t *year, const int32_t *day, const double *const *clim11,
                               int64_t nev, const sipnet_gpu_event *ev, int nthreads, double *final_out32) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nmembers) nthreads = (int)nmembers;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  OrJob *jobs = (OrJob *)malloc(sizeof(OrJob) * (size_t)nthreads);
  for (int i = 0; i < nthreads; ++i) {
    jobs[i] = (OrJob){flags, params_soa, ld, nmembers * i / nthreads, nmembers * (i + 1) / nthreads, T, year,
                      day,   clim11,     nev, ev, final_out32, 0};
    pthread_create(&th[i], NULL, ensemble_worker, &jobs[i]);
  }
  int rc = 0;
  for (int i = 0; i < nthreads; ++i) {
    pthread_join(th[i], NULL);
    if (jobs[i].rc && !rc) rc = jobs[i].rc;
  }
  free(th);
  free(jobs);
  return rc;
}
