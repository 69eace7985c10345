// sip_step.cuh -- one SIPNET timestep for one ensemble member, as device code.
//
// This is the device restatement of updateState() (reference
// src/sipnet/sipnet.c:1818-1855) and everything below it:
//   processEvents            events.c:449-742
//   calculateFluxes          sipnet.c:1256-1336 (+ depeffects.c, nitrogen.c, limitations.c)
//   updatePoolsAndBalance    sipnet.c:1769-1806 (+ events.c:744-790, nitrogen.c:210-239)
//   updateTrackers           sipnet.c:1420-1496
//   updateMeanTrackers       sipnet.c:1546-1570 (+ runmean.c:61-121)
//   updateEventTrackers      events.c:811-822
// One thread owns one member; pools and carried trackers live in registers
// (struct Member) across a chunk of steps; parameters are read from a
// per-block shared-memory tile; forcing comes from the block's staged ClimRec.
//
// Design differences from the reference (results identical):
//   * no global state; flags are a compile-time or kernel-uniform policy
//   * the mass-balance tracker (balance.c) is diagnostic only -- nothing in it
//     feeds back into state -- and is not evaluated
//   * pure functions evaluated several times with identical arguments in the
//     reference (calcTempEffect x4, calcRespMoistEffect x2, calcAnaerobicIndex
//     x4, getMeanTrackerMean x3) are evaluated once
//   * exit() paths become per-member status bits
// Expression shapes (association, `x / 365.0`, `-1.0 *`, constant chains) are
// kept literally: the validation build must do the same IEEE operations as the
// reference's gcc -O0 binary.
#pragma once
#include "sip_math.cuh"
#include "sip_num.cuh"
#include "sip_types.cuh"

namespace sip {

// ---- flag policies -----------------------------------------------------------
template <uint32_t MASK>
struct StaticFlags {
  __device__ __forceinline__ explicit StaticFlags(uint32_t) {}
  __device__ __forceinline__ bool on(uint32_t bit) const { return (MASK & bit) != 0; }
};
struct RuntimeFlags {
  uint32_t mask;
  __device__ __forceinline__ explicit RuntimeFlags(uint32_t m) : mask(m) {}
  __device__ __forceinline__ bool on(uint32_t bit) const { return (mask & bit) != 0; }
};

// ---- parameter tile in shared memory: row K of this thread at tile[slot(K) * stride] -----------------------------
// K is a template argument, so the slot folds to a constant and a read of a row the tile does not stage is a
// compile-time error.
struct DirectTile {
  const double *base;  // already offset by threadIdx.x
  int stride;
  template <int K>
  __device__ __forceinline__ double at() const {
    constexpr int slot = tile_slot(K);
    static_assert(slot >= 0, "the step reads a parameter row that is not staged (sip_types.cuh SIP_TILE_SKIP_LIST)");
    return base[slot * stride];
  }
};
#define SIP_P(name) prm.template at<SIPNET_P_##name>()
#define SIP_K(k) prm.template at<(k)>()

// Loads / stores of carried per-member data (ring slots, event counters, state rows).  COHERENT = true goes
// through L2 (ld.cg / st.cg): with dynamic scheduling consecutive sub-ranges of a member may run on different SMs
// within ONE launch, so a stale L1 line must not be served.  The static schedule keeps a member on one SM for the
// whole launch and uses ordinary accesses (the ring load sits on the step's critical path: 0.5 ms of 31 on C2).
template <bool COHERENT, class T>
__device__ __forceinline__ T carried_load(const T *p) {
  if constexpr (COHERENT) return __ldcg(p);
  else return *p;
}
template <bool COHERENT, class T>
__device__ __forceinline__ void carried_store(T *p, T v) {
  if constexpr (COHERENT) __stcg(p, v);
  else *p = v;
}

// ---- per-member register state -------------------------------------------------
struct Member {
  // Envi, state.h:416-463
  double wood, leaf, soil, water, litter, snow, coarse, fine, minN, orgN, litN, storN, delta;
  // carried trackers (state.h:650-726): only what feeds back or is output
  double gdd, totNee, wetFrac;
  double dTill;    // eventTrackers.d_till_mod, events.h:213-221
  double ringSum;  // MeanTracker.sum, runmean.h
  int ringStart, ringLast;
  int trkLastYear;   // trackers.lastYear
  int phenLastYear;  // phenologyTrackers.lastYear
  int didGrowth, didFall;
  uint32_t status;
};

// extended trackers, only carried by the DEBUG instantiation (restart/debug-log rows)
struct MemberExt {
  double yGpp, yRtot, yRa, yRh, yNpp, yNee, yLitter, tGpp, tRtot, tRa, tRh, tNpp;
  double harvRemoved, harvTransferred;  // eventTrackers of the last step (checkpoint payload only)
  uint32_t cnt[SIPNET_GPU_NCOUNTERS];   // occurrences of the reference's informational messages (SIPNET_GPU_CNT_*)
};

// ---- mean-NPP ring in HBM: slot s of member m at v[s * ld + m] -----------------
template <bool COHERENT>
struct RingRefT {
  double *v;
  double *w;
  unsigned ld;  // slots are ld doubles apart; cap * ld < 2^32 (checked at init), so a slot's offset is one 32-bit product
  int cap;
  __device__ __forceinline__ size_t off(int s) const { return (size_t)((unsigned)s * ld); }
  // L2-coherent accesses: with dynamic scheduling consecutive sub-ranges of a member may run on different SMs
  // within ONE launch, so ring slots must not be served from a stale L1 line
  __device__ __forceinline__ double val(int s) const { return carried_load<COHERENT>(v + off(s)); }
  __device__ __forceinline__ double wgt(int s) const { return carried_load<COHERENT>(w + off(s)); }
  __device__ __forceinline__ void set_val(int s, double x) const { carried_store<COHERENT>(v + off(s), x); }
  __device__ __forceinline__ void set_wgt(int s, double x) const { carried_store<COHERENT>(w + off(s), x); }
};

// The ring's oldest entry (what the next push evicts first), loaded AHEAD of ring_push: the ring lives in HBM and its
// loads are L2 round trips; issued inside ring_push, the eviction loop's exit test waits for them with nothing else
// of the step left to issue.  The step loads it after the water limitation of photosynthesis (register pressure has
// dropped, half of the step still lies ahead) and ring_push uses it for its first eviction.
struct RingHead {
  double w, v;
};
template <class RG>
__device__ __forceinline__ RingHead ring_head(const Member &mb, const RG &rg) {
  return RingHead{rg.wgt(mb.ringStart), rg.val(mb.ringStart)};
}

template <class RG>
__device__ __forceinline__ void ring_reset(Member &mb, const RG &rg, double v, RingHead &head) {  // runmean.c:44-51
  mb.ringStart = mb.ringLast = 0;
  rg.set_val(0, v);
  rg.set_wgt(0, kMeanNppDays);
  mb.ringSum = v * kMeanNppDays;
  head = RingHead{kMeanNppDays, v};  // slot 0 is the oldest entry now
}

// The usual push -- equal step lengths, a living plant: evict exactly the head entry, append one -- behind ONE test
// (the general routine below makes four on that path); the same operations on the same operands.
template <class RG>
__device__ __forceinline__ bool ring_push_usual(Member &mb, const RG &rg, double value, double weight, const RingHead &head,
                                                bool alive) {
  if (!(alive & (weight < kMeanNppDays) & (head.w == weight) & (mb.ringLast != mb.ringStart))) return false;
  double sum = mb.ringSum;
  sum -= head.w * head.v;
  mb.ringStart = (mb.ringStart + 1 == rg.cap) ? 0 : mb.ringStart + 1;
  const int i = (mb.ringLast + 1 == rg.cap) ? 0 : mb.ringLast + 1;
  mb.ringLast = i;
  rg.set_val(i, value);
  rg.set_wgt(i, weight);
  sum += value * weight;
  mb.ringSum = sum;
  return true;
}

template <class RG>
__device__ __forceinline__ void ring_push(Member &mb, const RG &rg, double value, double weight, RingHead head) {
  // addValueToMeanTracker, runmean.c:61-115 (weight <= 0 is rejected at init: events.c:460)
  if (weight >= kMeanNppDays) {
    ring_reset(mb, rg, value, head);
    return;
  }
  double left = weight;
  int i = mb.ringStart;
  double sum = mb.ringSum;
  // the usual push evicts exactly the head entry (equal step lengths): that case without entering the loop
  bool first = !(head.w == left);
  if (!first) {
    sum -= head.w * head.v;
    left = 0;
    i = (i + 1 == rg.cap) ? 0 : i + 1;
  }
  while (left > 0) {
    const double wi = first ? head.w : rg.wgt(i);
    const double vi = first ? head.v : rg.val(i);
    first = false;
    if (wi > left) {
      rg.set_wgt(i, wi - left);
      sum -= left * vi;
      left = 0;
    } else {
      sum -= wi * vi;
      left -= wi;
      i = (i + 1 == rg.cap) ? 0 : i + 1;
    }
  }
  mb.ringStart = i;
  i = (mb.ringLast + 1 == rg.cap) ? 0 : mb.ringLast + 1;
  if (i == mb.ringStart) {  // out of space: reference restores and exits 7 (sipnet.c:1562-1569)
    rg.set_wgt(i, rg.wgt(i) + weight);
    sum += weight * rg.val(i);
    mb.ringSum = sum;
    mb.status |= SIPNET_GPU_ST_RING_OVERFLOW;
    return;
  }
  mb.ringLast = i;
  rg.set_val(i, value);
  rg.set_wgt(i, weight);
  sum += value * weight;
  mb.ringSum = sum;
}

// products and sums that must not be contracted into an FMA in ANY build (the throughput policy's translation units
// are compiled with -fmad=true): the soil-water balance, where a half-ulp difference is amplified by cancellation
__device__ __forceinline__ double nc_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double nc_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double nc_sub(double a, double b) { return __dsub_rn(a, b); }

// ---- small helpers ---------------------------------------------------------------
__device__ __forceinline__ bool has_biomass(const Member &mb) {  // hasSufficientBiomass, sipnet.c:1530-1536
  const double totWood = mb.wood + mb.delta;
  const double totRoot = mb.fine + mb.coarse;
  return (mb.wood > kTiny) & (totWood > kTiny) & (totRoot > kTiny);  // three chained compares, no branches
}

// ensureNonNegative, sipnet.c:1346-1356: a stock below its floor is set to zero; a clamped amount above EPS is what
// the reference warns about (informational status bit, collected in one predicate and folded into the status once)
__device__ __forceinline__ void clamp_stock(double &v, bool &clamped) {  // floor 0: v < 0 and |v| > EPS <=> v < -EPS
  clamped = clamped | (v < -kEps);
  v = v < 0.0 ? 0.0 : v;
}
__device__ __forceinline__ void clamp_stock(double &v, double floorv, bool &clamped) {
  const bool low = v < floorv;
  clamped = clamped | (low & (fabs(v) > kEps));
  v = low ? 0.0 : v;
}

// record sink for events.out rows (events.c:379-402)
template <bool COHERENT>
struct RecSinkT {
  sipnet_gpu_event_record *recs;  // this member's slots, or null
  int32_t *count;                 // this member's counter, or null
  int32_t maxRecs;
  int32_t step;
  __device__ __forceinline__ void add(Member &mb, int type, int variant, int nval, const double *v) const {
    if (count == nullptr) return;
    const int n = carried_load<COHERENT>(count);
    if (recs != nullptr && n < maxRecs) {
      sipnet_gpu_event_record &r = recs[n];
      r.step = step;
      r.type = type;
      r.nval = nval;
      r.variant = variant;
      for (int k = 0; k < SIPNET_GPU_EVREC_NVAL; ++k) r.val[k] = (k < nval) ? v[k] : 0.0;
    } else if (recs != nullptr) {
      mb.status |= SIPNET_GPU_ST_EVREC_OVERFLOW;
    }
    carried_store<COHERENT>(count, n + 1);
  }
};

// all 56 per-day rates of struct FluxVars (state.h:469-645); the non-debug
// instantiations only keep the ones that are live
struct Rates {
  double photosynthesis, leafLitter, woodLitter, rVeg, rSoil, rain, transpiration, drainage, litterToSoil,
      rLitter, snowFall, snowMelt, sublimation, immedEvap, fastFlow, evaporation, fineRootLoss, coarseRootLoss,
      fineRootCreation, coarseRootCreation, rCoarseRoot, rFineRoot, leafCreation, woodCreation, leafOnCreation,
      leafOnCreationFromWood, nVolatilization, nLeaching, nOrgSoil, nOrgLitter, nMin, nFixation, nUptake,
      leafOffNResorption, reductionNResorption, eventLeafC, eventWoodC, eventFineRootC, eventCoarseRootC,
      eventEvap, eventSoilWater, eventSoilC, eventLitterC, eventMinN, eventSoilOrgN, eventLitterN, eventInputC,
      eventOutputC, eventInputN, eventOutputN, eventLeafOnCreation, eventLeafOnCreationFromWood,
      eventLeafOffLitter, eventLeafOffNResorption, soilMethane, litterMethane;
};

// per-step tracker values that are outputs only (not carried)
struct StepTrack {
  double gpp, rtot, ra, rh, rRoot, rSoil, rAboveground, npp, nee, woodCreation, evapotranspiration, methane, n2o,
      nLeaching, nFixation, nUptake, meanNPP;
};

// per-step divisor context: x / length and the member-constant C:N divisors
template <class NM, class PT>
struct Div {
  NM &nm;
  const PT &prm;
  double len, seedLen, invLenPow2;
  // x / climate->length
  __device__ __forceinline__ double byLen(double a) const {
    if (NM::kFast && invLenPow2 != 0.0) return a * invLenPow2;  // exact: length is a power of two
    return nm.divs(a, len, seedLen);
  }
  // the soil-water balance's divisions (ThroughNum keeps them correctly rounded)
  __device__ __forceinline__ double byLenW(double a) const {
    if (NM::kFast && invLenPow2 != 0.0) return a * invLenPow2;  // exact: length is a power of two
    return nm.divsw(a, len, seedLen);
  }
  __device__ __forceinline__ double byWhcW(double a) const { return nm.divsw(a, SIP_P(soilWHC), SIP_K(kSeedWhc)); }
  __device__ __forceinline__ double byLeafCN(double a) const { return nm.divs(a, SIP_P(leafCN), SIP_K(kSeedLeafCN)); }
  __device__ __forceinline__ double byWoodCN(double a) const { return nm.divs(a, SIP_P(woodCN), SIP_K(kSeedWoodCN)); }
  __device__ __forceinline__ double byFineCN(double a) const {
    return nm.divs(a, SIP_P(fineRootCN), SIP_K(kSeedFineRootCN));
  }
  __device__ __forceinline__ double byWhc(double a) const { return nm.divs(a, SIP_P(soilWHC), SIP_K(kSeedWhc)); }
};

// ---- nitrogen helpers, nitrogen.c ----------------------------------------------------
template <class DV>
__device__ __forceinline__ double n_leafon_from_c(const DV &dv, double c) {  // nitrogen.c:86-88
  // c is the leaf-on flux: zero on all but one step a year.  fmax(0.0, (+-0)/a - (+-0)/b) is +0 exactly,
  // so the two divisions are skipped (same bits).
  if (c == 0.0) return 0.0;
  return max0(dv.byLeafCN(c) - dv.byWoodCN(c));
}
template <class FL, class DV>
__device__ __forceinline__ double n_demand(const FL &fl, const DV &dv, const Rates &r) {  // nitrogen.c:91-106
  if (!fl.on(F_NITROGEN)) return 0.0;
  const double d = dv.byWoodCN(r.woodCreation) + dv.byLeafCN(r.leafCreation) + dv.byFineCN(r.fineRootCreation) +
                   dv.byWoodCN(r.coarseRootCreation);
  return max0(d);
}
__device__ __forceinline__ double n_non_uptake(const Rates &r) {  // nitrogen.c:124-126
  return r.nMin - r.nVolatilization - r.nLeaching;
}
template <class DV>
__device__ __forceinline__ double n_unclaimed_storage(const DV &dv, const Member &mb, const Rates &r,
                                                      double len) {  // nitrogen.c:129-136
  const double cflux = r.leafOnCreation + r.eventLeafOnCreation;
  const double nflux = n_leafon_from_c(dv, cflux);
  return max0(mb.storN - nflux * len);
}
template <class DV>
__device__ __forceinline__ double n_fix_frac(const DV &dv, const Member &mb) {  // nitrogen.c:139-153
  const auto &prm = dv.prm;
  double inhib;
  const double denom = SIP_P(halfNFixationMax) + mb.minN;
  if (denom < kTiny) {
    inhib = 1;
  } else {
    inhib = dv.nm.div(SIP_P(halfNFixationMax), denom);
  }
  return SIP_P(nFixationFracMax) * inhib;
}
// returns calcPlantNDemandFlux() of the current creation fluxes (reused by updateNitrogenPools: same inputs)
template <class FL, class DV>
__device__ __forceinline__ double n_fix_and_uptake(const FL &fl, const DV &dv, const Member &mb, Rates &r,
                                                   double len) {  // nitrogen.c:156-168
  const double demand = n_demand(fl, dv, r);
  const double storage = dv.byLen(n_unclaimed_storage(dv, mb, r, len));
  const double rem = max0(demand - storage);
  const double ff = n_fix_frac(dv, mb);
  r.nFixation = ff * rem;
  r.nUptake = (1 - ff) * rem;
  return demand;
}

// checkLeafOnLimitation, limitations.c:13-64
template <class FL, class DV>
__device__ __forceinline__ void limit_leaf_on(const FL &fl, const DV &dv, Member &mb, double len, double &flux,
                                              uint32_t *count) {  // count: the validation dump's message counter, or null
  const auto &prm = dv.prm;
  const double demandC = flux * len;
  if (demandC < kTiny) return;
  const double availC = (mb.wood + mb.coarse) * SIP_P(leafOnReallocFrac);
  const double cLim = dv.nm.div(availC, demandC);
  double nLim = 1.0;
  if (fl.on(F_NITROGEN)) {
    const double demandN = n_leafon_from_c(dv, demandC);
    if (demandN > kTiny) nLim = dv.nm.div(mb.storN, demandN);
  }
  const double lim = clip01(fmin(cLim, nLim));
  if (lim < 1) {
    flux *= lim;
    mb.status |= SIPNET_GPU_ST_LEAFON_LIMITED;  // the reference's logInfo, limitations.c:48-61
    if (count != nullptr) ++*count;
  }
}

// calcRatio, common/util.c:72-75
template <class NM>
__device__ __forceinline__ double ratio(NM &nm, double num, double den) {
  const double d = den < kTiny ? kTiny : den;
  return nm.div(num, d);
}

// getMassTotals, balance.c:13-33 (validation dump only)
template <class FL, class NM, class PT>
__device__ __forceinline__ void mass_totals(const FL &fl, NM &nm, const PT &prm, const Member &mb, double &carbon,
                                            double &nitrogen) {
  carbon = (mb.wood + mb.delta) + mb.leaf + mb.fine + mb.coarse + mb.soil;
  if (fl.on(F_LITTER_POOL)) carbon += mb.litter;
  if (fl.on(F_NITROGEN)) {
    nitrogen = nm.div(mb.wood, SIP_P(woodCN)) + nm.div(mb.leaf, SIP_P(leafCN)) + nm.div(mb.fine, SIP_P(fineRootCN)) +
               nm.div(mb.coarse, SIP_P(woodCN)) + mb.orgN + mb.litN + mb.minN + mb.storN;
  } else {
    nitrogen = 0.0;
  }
}

// ---- the step ----------------------------------------------------------------------
// Emit is a functor: emit.outputs(column) stores the outputState() columns it keeps and, in the
// DEBUG instantiation, emit.dbg(index, value) for the debug-log fields.
// NM is the numerics policy (sip_num.cuh): ExactNum or FastNum -- same bits.
template <class FL, bool DEBUG, class NM, class PT, class RG, class RS, class Emit>
__device__ __forceinline__ void step(const FL &fl, NM &nm, const PT &prm, const ClimRec &c, const EventDev *events,
                                     Member &mb, MemberExt &ext, const RG &rg, const RS &rec, Emit &emit,
                                     const StepConsts &kc) {
  const double len = c.length;
  Div<NM, PT> dv{nm, prm, len, 0.0, c.invLenPow2};  // (the reciprocal seed of an odd step length is set below)
  const double oldSoilWater = mb.water;  // sipnet.c:1821
  Rates r = {};                          // resetFluxes, sipnet.c:1222
  bool alive = has_biomass(mb);          // initPlantSurvivalTracker, sipnet.c:1538
  double harvRemoved = 0, harvTransferred = 0;  // events.c:468-469

  // ---------------- processEvents, events.c:471-741 (events pre-bound to this step) ----
  // a step length that is not a power of two (its reciprocal seed is needed) or events on this step: both are
  // block-uniform and unusual, one test covers them
  if ((NM::kFast & (c.invLenPow2 == 0.0)) | (c.evBegin < c.evEnd)) {
  if (NM::kFast && c.invLenPow2 == 0.0) {
    dv.seedLen = nm.seed(len);
    nm.divisor_check(len);
  }
  for (int e = c.evBegin; e < c.evEnd; ++e) {
    const EventDev ev = events[e];
    switch (ev.type) {
      case SIPNET_EV_IRRIGATION: {  // :484-506
        const double amount = ev.p[0];
        double soilAmt, evapAmt;
        if (ev.method == 0) {
          evapAmt = SIP_P(immedEvapFrac) * amount;
          soilAmt = amount - evapAmt;
        } else {
          evapAmt = 0.0;
          soilAmt = amount;
        }
        r.eventEvap += dv.byLen(evapAmt);
        r.eventSoilWater += dv.byLen(soilAmt);
        const double v[2] = {soilAmt, evapAmt};
        rec.add(mb, ev.type, 0, 2, v);
      } break;
      case SIPNET_EV_PLANTING: {  // :507-542
        const double leafC = ev.p[0], woodC = ev.p[1], fineC = ev.p[2], coarseC = ev.p[3];
        r.eventLeafC += dv.byLen(leafC);
        r.eventWoodC += dv.byLen(woodC);
        r.eventFineRootC += dv.byLen(fineC);
        r.eventCoarseRootC += dv.byLen(coarseC);
        const double inC = leafC + woodC + fineC + coarseC;
        double inN = 0.0;
        r.eventInputC += dv.byLen(inC);
        if (fl.on(F_NITROGEN)) {
          inN = dv.byLeafCN(leafC) + dv.byWoodCN(woodC) + dv.byFineCN(fineC) + dv.byWoodCN(coarseC);
          r.eventInputN += dv.byLen(inN);
        }
        const double v[6] = {leafC, woodC, fineC, coarseC, inC, inN};
        rec.add(mb, ev.type, 0, 6, v);
      } break;
      case SIPNET_EV_HARVEST: {  // :543-635
        const double fRA = ev.p[0], fRB = ev.p[1], fTA = ev.p[2], fTB = ev.p[3];
        const double woodC = mb.wood + mb.delta;
        const double above = woodC + mb.leaf;
        const double below = mb.fine + mb.coarse;
        const double total = above + below;
        if (total > kTiny) {
          const double removed = fRA * above + fRB * below;
          const double moved = fTA * above + fTB * below;
          harvRemoved += nm.div(removed, total);
          harvTransferred += nm.div(moved, total);
        }
        double litterAdd = fTA * (mb.leaf + woodC);
        double soilAdd = fTB * (mb.fine + mb.coarse);
        const double dLeaf = -mb.leaf * (fRA + fTA);
        const double dWood = -woodC * (fRA + fTA);
        const double dFine = -mb.fine * (fRB + fTB);
        const double dCoarse = -mb.coarse * (fRB + fTB);
        if (!fl.on(F_LITTER_POOL)) {
          soilAdd += litterAdd;
          litterAdd = 0.0;
        }
        r.eventLitterC += dv.byLen(litterAdd);
        r.eventSoilC += dv.byLen(soilAdd);
        r.eventLeafC += dv.byLen(dLeaf);
        r.eventWoodC += dv.byLen(dWood);
        r.eventFineRootC += dv.byLen(dFine);
        r.eventCoarseRootC += dv.byLen(dCoarse);
        double litterNAdd = 0.0, soilNAdd = 0.0;
        if (fl.on(F_NITROGEN)) {
          const double nAbove = (dv.byLeafCN(mb.leaf)) + (dv.byWoodCN(mb.wood));
          const double nBelow = (dv.byFineCN(mb.fine)) + (dv.byWoodCN(mb.coarse));
          litterNAdd = fTA * nAbove;
          soilNAdd = fTB * nBelow;
          r.eventSoilOrgN += dv.byLen(soilNAdd);
          r.eventLitterN += dv.byLen(litterNAdd);
        }
        const double outC = ((woodC + mb.leaf) * fRA + (mb.fine + mb.coarse) * fRB);
        double outN = 0.0;
        r.eventOutputC += dv.byLen(outC);
        if (fl.on(F_NITROGEN)) {
          outN = (dv.byWoodCN(mb.wood) + dv.byLeafCN(mb.leaf)) * fRA +
                 (dv.byFineCN(mb.fine) + dv.byWoodCN(mb.coarse)) * fRB;
          r.eventOutputN += dv.byLen(outN);
        }
        const double v[10] = {soilAdd, litterAdd, dLeaf, dWood, dFine, dCoarse, soilNAdd, litterNAdd, outC, outN};
        rec.add(mb, ev.type, 0, 10, v);
      } break;
      case SIPNET_EV_TILLAGE: {  // :636-646
        mb.dTill += ev.p[0];
        const double v[1] = {ev.p[0]};
        rec.add(mb, ev.type, 0, 1, v);
      } break;
      case SIPNET_EV_FERTILIZATION: {  // :647-685
        const double orgC = ev.p[1];
        double orgN = 0.0, minN = 0.0;
        if (fl.on(F_NITROGEN)) {
          orgN = ev.p[0];
          minN = ev.p[2];
        }
        if (fl.on(F_LITTER_POOL)) {
          r.eventLitterC += dv.byLen(orgC);
        } else {
          r.eventSoilC += dv.byLen(orgC);
        }
        if (fl.on(F_NITROGEN)) {
          r.eventLitterN += dv.byLen(orgN);
          r.eventMinN += dv.byLen(minN);
        }
        r.eventInputC += dv.byLen(orgC);
        if (fl.on(F_NITROGEN)) r.eventInputN += dv.byLen(orgN + minN);
        const double v[6] = {fl.on(F_LITTER_POOL) ? orgC : 0.0, fl.on(F_LITTER_POOL) ? 0.0 : orgC, minN, orgN, orgC,
                             (orgN + minN)};
        rec.add(mb, ev.type, 0, 6, v);
      } break;
      case SIPNET_EV_LEAFON: {  // :686-705
        double flux = dv.byLen(SIP_P(leafGrowth));
        limit_leaf_on(fl, dv, mb, len, flux, DEBUG ? &ext.cnt[SIPNET_GPU_CNT_LEAFON_LIMITED] : nullptr);
        r.eventLeafOnCreation += flux;
        const double src = mb.wood + mb.coarse;
        if (src > kTiny) r.eventLeafOnCreationFromWood += nm.div(flux * mb.wood, src);
      } break;
      case SIPNET_EV_LEAFOFF: {  // :706-728
        const double leafOff = mb.leaf * SIP_P(fracLeafFall);
        r.eventLeafOffLitter += dv.byLen(leafOff);
        double litterNAdd = 0.0, resorb = 0.0;
        if (fl.on(F_NITROGEN)) {
          const double leafN = dv.byLeafCN(leafOff);
          resorb = leafN * SIP_P(leafNResorptionFrac);
          litterNAdd = leafN - resorb;
          r.eventLeafOffNResorption += dv.byLen(resorb);
          r.eventLitterN += dv.byLen(litterNAdd);
        }
        const double v[3] = {leafOff, resorb, litterNAdd};
        rec.add(mb, ev.type, 1, 3, v);
      } break;
      default:
        break;  // PLANTDEATH is ignored (events.c:729-734); unknown types are rejected at init
    }
  }
  }

  // ---------------- calculateFluxes, sipnet.c:1256-1336 -------------------------------
  const double whc = SIP_P(soilWHC);
  const double lai = nm.divs(mb.leaf, SIP_P(leafCSpWt), SIP_K(kSeedLeafCSpWt));  // :1274
  const double meanNpp = nm.divs(mb.ringSum, kMeanNppDays, kc.seed5);          // getMeanTrackerMean, runmean.c:118
  const double woodTot = mb.wood + mb.delta;                                   // getTotalWoodC

  double dLight;
  if ((lai > 0) & (c.par > 0)) {  // calcLightEff, :517-570 (Simpson, 6 layers, coefficients 1,4,2,4,2,4,2 then -last)
    const double att = SIP_P(attenuation), hsp = SIP_P(halfSatPar), seedHsp = SIP_K(kSeedHalfSatPar);
    double eff[7];
    // the seven layers are independent: each stage is issued for all layers before the next one
#pragma unroll
    for (int layer = 6; layer >= 0; --layer) {
      const double cumLai = lai * ((double)layer / 6);
      // optimistic policy, top of the canopy: cumLai = lai * 0 = +0 and exp(-att * +0) = exp(+-0) = 1 for finite
      // att and lai -- and layer 6's guard (|att * lai| < 512) has just flagged the member if either is not
      if (NM::kFast && layer == 0) eff[layer] = 1.0;
      else eff[layer] = nm.exp(-1.0 * att * cumLai);
    }
#pragma unroll
    for (int layer = 0; layer <= 6; ++layer) {
      const double inten = c.par * eff[layer];
      eff[layer] = nm.divs(-1.0 * inten, hsp, seedHsp);
    }
#pragma unroll
    for (int layer = 0; layer <= 6; ++layer) eff[layer] = (1 - nm.powc(2.0, kc.log2Hi, kc.log2Lo, eff[layer]));
    double cum = 0.0;
#pragma unroll
    for (int layer = 0; layer <= 6; ++layer) {
      const int coeff = (layer == 0) ? 1 : 2 * (1 + layer % 2);
      cum += coeff * eff[layer];
    }
    cum -= eff[6];
    dLight = nm.divs(cum, 3.0 * 6, kc.seed18);
  } else {
    dLight = 0;
  }
  // State-independent Q10 / VPD factors: six independent exp-class evaluations.  They sit AFTER the canopy integral:
  // computed before it they only lengthen the live ranges across the step's most register-hungry stretch; here they
  // fill the issue slots of the serial tail that follows (measured: 88.1 -> 85.9 ms).
  const double q10Fol = nm.powc(SIP_P(vegRespQ10), SIP_K(kLogVegQ10), SIP_K(kLogVegQ10 + 1),
                                nm.divs(c.tair - SIP_P(psnTOpt), 10.0, kc.seed10));      // sipnet.c:1056
  const double q10Wood = nm.powc(SIP_P(vegRespQ10), SIP_K(kLogVegQ10), SIP_K(kLogVegQ10 + 1), c.tair10);  // :1067
  const double tsoil10 = c.tsoil10;
  const double q10Coarse = nm.powc(SIP_P(coarseRootQ10), SIP_K(kLogCoarseQ10), SIP_K(kLogCoarseQ10 + 1), tsoil10);  // :1076
  const double q10Fine = nm.powc(SIP_P(fineRootQ10), SIP_K(kLogFineQ10), SIP_K(kLogFineQ10 + 1), tsoil10);
  const double tempEffect = nm.powc(SIP_P(soilRespQ10), SIP_K(kLogSoilQ10), SIP_K(kLogSoilQ10 + 1), tsoil10);  // depeffects.c:72-75
  const double vpdPow = nm.powc(c.vpd, c.logVpdHi, c.logVpdLo, SIP_P(dVpdExp));                // :626

  // potPsn, :590-641
  const double respPerGram = SIP_K(kRespPerGram);
  const double grossAMax = SIP_K(kGrossAMax);
  // kPsnTRangeSqSlot holds pow((psnTMax - psnTMin) / 2.0, 2), evaluated once per member by the setup kernel
  double dTemp = nm.divs((SIP_P(psnTMax) - c.tair) * (c.tair - SIP_P(psnTMin)), SIP_K(kPsnTRangeSqSlot),
                         SIP_K(kSeedPsnTRangeSq));
  dTemp = max0(dTemp);
  double dVpd = 1.0 - SIP_P(dVpdSlope) * vpdPow;
  dVpd = max0(dVpd);
  const double conv = SIP_K(kConvBase) * lai * 86400.0;
  const double potPsn = grossAMax * dTemp * dVpd * dLight * conv;
  const double baseFolResp = respPerGram * conv;

  // moisture, :656-699
  double dWater;
  if (potPsn < kTiny) {
    r.transpiration = 0.0;
    dWater = 1;
  } else {
    const double wue = nm.div(SIP_P(wueConst), c.vpd);
    const double potTrans = nm.div(potPsn, wue) * 1000.0 * (44.0 / 12.0) * (1.0 / 10000.0);
    double removable = fmin(mb.water, whc) * SIP_P(waterRemoveFrac);
    if (c.tsoil < SIP_P(frozenSoilThreshold)) removable *= SIP_P(frozenSoilEff);
    r.transpiration = fmin(removable, potTrans);
    dWater = nm.div(r.transpiration, potTrans);
  }

  // calcPrecip, :848-882
  if (c.tair <= 0) {
    r.snowFall = c.precipRate;
    r.rain = 0;
  } else {
    r.snowFall = 0;
    r.rain = c.precipRate;
  }
  r.immedEvap = r.rain * SIP_P(immedEvapFrac);
  if (fl.on(F_LEAF_WATER)) {
    const double maxPool = lai * SIP_P(leafPoolDepth);
    if (r.immedEvap > maxPool) r.immedEvap = maxPool;
  }
  const double netRain = r.rain - r.immedEvap;  // :1281

  // snowPack, :888-946
  {
    if (mb.snow <= 0) {
      r.snowMelt = 0;
      r.sublimation = 0;
    } else {
      const double rd = nm.div(SIP_P(rdConst), c.wspd);
      r.sublimation = nm.div(c.sublK, rd);
      double left = mb.snow + (r.snowFall * len);
      if (r.sublimation < 0) r.sublimation = 0;
      if (left - (r.sublimation * len) < 0) {
        r.sublimation = dv.byLen(left);
        left = 0;
      } else {
        left -= (r.sublimation * len);
      }
      if (c.tair <= 0) {
        r.snowMelt = 0;
      } else {
        r.snowMelt = SIP_P(snowMelt) * c.tair;
        if (left - (r.snowMelt * len) < 0) r.snowMelt = dv.byLen(left);
      }
    }
  }

  // calcSoilWaterFluxes, :963-1031 (sums written with nc_*: same operations as `a + b * len`, never contracted)
  const double waterFrac = clip01(dv.byWhcW(mb.water));  // getClippedWaterFrac, depeffects.c:11
  {
    double netIn = nc_add(netRain, r.snowMelt);
    r.fastFlow = nc_mul(netIn, SIP_P(fastFlowFrac));
    netIn = nc_sub(netIn, r.fastFlow);
    double left = nc_sub(nc_add(mb.water, nc_mul(netIn, len)), nc_mul(r.transpiration, len));
    if (mb.snow > 0) {
      r.evaporation = 0;
    } else {
      const double rd = nm.divw(SIP_P(rdConst), c.wspd);
      const double rsoil = nm.exp(nc_sub(SIP_P(rSoilConst1), nc_mul(SIP_P(rSoilConst2), waterFrac)));
      r.evaporation = nm.divw(c.evapK, nc_add(rd, rsoil));
      if (r.evaporation < 0) r.evaporation = 0;
      if (nc_sub(left, nc_mul(r.evaporation, len)) < kTiny) {
        r.evaporation = dv.byLenW(nc_sub(left, kTiny));
        left = 0;
      } else {
        left = nc_sub(left, nc_mul(r.evaporation, len));
      }
    }
    if (!fl.on(F_FLOODING)) {  // the quotient on every step, kept where there is an excess: a select, no branch
      const bool over = left > whc;
      const double excess = nc_sub(left, whc);
      const double q = dv.byLenW(over ? excess : 0.0);
      r.drainage = over ? q : 0.0;
    } else
    if (left > whc) {
      const double excess = nc_sub(left, whc);
      if (fl.on(F_FLOODING)) {
        r.drainage = fmin(nc_mul(excess, SIP_P(waterDrainFrac)), dv.byLenW(excess));
      } else {
        r.drainage = dv.byLenW(excess);
      }
    } else {
      r.drainage = 0;
    }
  }

  RingHead ringHead = ring_head(mb, rg);  // consumed by ring_push at the end of the step
  r.photosynthesis = potPsn * dWater;  // getGpp, :1034

  // vegResp / vegResp2, :1051-1103
  {
    double fol = baseFolResp * q10Fol;
    if (c.tsoil < SIP_P(frozenSoilThreshold)) fol *= SIP_P(frozenSoilFolREff);
    const double woodR = SIP_P(baseVegResp) * woodTot * q10Wood;
    if (fl.on(F_GROWTH_RESP)) {
      double growth = SIP_P(growthRespFrac) * meanNpp;
      if (growth < 0) growth = 0;
      r.rVeg = fol + woodR + growth;
    } else {
      r.rVeg = fol + woodR;
    }
  }

  // calcWoodAndLeafFluxes, :756-782
  r.woodLitter += woodTot * SIP_P(woodTurnoverRate);
  r.leafLitter += mb.leaf * SIP_P(leafTurnoverRate);
  r.leafCreation += meanNpp * SIP_P(leafAllocation);
  r.woodCreation += meanNpp * SIP_P(woodAllocation);

  // calcLeafOnOffFluxes, :800-842
  {
    if (c.year > mb.phenLastYear) {
      mb.didGrowth = 0;
      mb.didFall = 0;
      mb.phenLastYear = c.year;
    }
    if (!(mb.didGrowth & mb.didFall)) {  // both done for the rest of the year: one test instead of two
    if (!mb.didGrowth) {
      bool past;  // pastLeafGrowth, :705-729
      if (fl.on(F_GDD)) {
        double g = c.gdd;
        if (c.year == mb.trkLastYear) g += mb.gdd;
        past = g >= SIP_P(gddLeafOn);
      } else if (fl.on(F_SOIL_PHENOL)) {
        past = c.tsoil >= SIP_P(soilTempLeafOn);
      } else if (SIP_P(leafOnDay) > 0) {
        past = c.dayFrac >= SIP_P(leafOnDay);  // (double)day + time / 24.0, from the host
      } else {
        past = false;
      }
      if (past) {
        double on = dv.byLen(SIP_P(leafGrowth));
        limit_leaf_on(fl, dv, mb, len, on, DEBUG ? &ext.cnt[SIPNET_GPU_CNT_LEAFON_LIMITED] : nullptr);
        r.leafOnCreation += on;
        const double src = mb.wood + mb.coarse;
        if (src > kTiny) r.leafOnCreationFromWood += nm.div(on * mb.wood, src);
        mb.didGrowth = 1;
      }
    }
    if (!mb.didFall) {
      const bool past = (SIP_P(leafOffDay) > 0) & (c.dayFrac >= SIP_P(leafOffDay));  // pastLeafFall, :733-742
      if (past) {
        const double off = dv.byLen(mb.leaf * SIP_P(fracLeafFall));
        r.leafLitter += off;
        mb.didFall = 1;
        if (off > kTiny && fl.on(F_EVENTS)) {
          const double v[1] = {off * len};
          rec.add(mb, SIPNET_EV_LEAFOFF, 0, 1, v);
        }
      }
    }
    }
  }

  // shared dependency terms (depeffects.c); each is a pure function of (tsoil, soilWater, params)
  double anaerobicIdx = 0.0;  // calcAnaerobicIndex :15-22
  if (fl.on(F_ANAEROBIC) || fl.on(F_NITROGEN)) {
    anaerobicIdx = clip01(nm.divs(waterFrac - SIP_P(fAnoxia), SIP_K(kOneMinusFa), SIP_K(kSeedOneMinusFa)));
  }
  double moistEffect;  // calcRespMoistEffect :24-63
  if (!fl.on(F_WATER_HRESP) || c.tsoil < 0) {
    moistEffect = 1.0;
  } else if (!fl.on(F_ANAEROBIC)) {
    moistEffect = nm.pow(waterFrac, SIP_P(soilRespMoistEffect));
  } else {
    const double dAer = clip01(nm.divs(waterFrac, SIP_P(fAnoxia), SIP_K(kSeedFAnoxia)));
    moistEffect = (1 - anaerobicIdx) * dAer + SIP_P(anaerobicDecompRate) * anaerobicIdx;
  }
  const double tillEffect = 1 + mb.dTill;  // calcTillageEffect :77

  // C:N ratios of the litter and soil pools (calcRatio); shared by calcCNEffect and calcNPoolFluxes
  double litterCN = 0.0, soilCN = 0.0;
  if (fl.on(F_NITROGEN)) {
    litterCN = ratio(nm, mb.litter, mb.litN);
    soilCN = ratio(nm, mb.soil, mb.orgN);
  }

  // calcLitterFluxes, :1150-1171
  if (fl.on(F_LITTER_POOL)) {
    double cn = 1.0;  // calcCNEffect, depeffects.c:79-88
    if (fl.on(F_NITROGEN)) cn = nm.div(SIP_P(kCN), SIP_P(kCN) + litterCN);
    const double breakdown = mb.litter * SIP_P(litterBreakdownRate) * tempEffect * moistEffect * tillEffect * cn;
    r.rLitter = breakdown * SIP_P(fracLitterRespired);
    r.litterToSoil = breakdown * SIP_K(kOneMinusFracLitResp);
  }

  // calcRootFluxes, :1176-1196
  r.coarseRootLoss += SIP_P(coarseRootTurnoverRate) * mb.coarse;
  r.fineRootLoss += SIP_P(fineRootTurnoverRate) * mb.fine;
  r.coarseRootCreation += SIP_P(coarseRootAllocation) * meanNpp;
  r.fineRootCreation += SIP_P(fineRootAllocation) * meanNpp;
  r.rCoarseRoot = SIP_P(baseCoarseRootResp) * mb.coarse * q10Coarse;
  r.rFineRoot = SIP_P(baseFineRootResp) * mb.fine * q10Fine;

  // calcSoilRespiration, :1132-1148
  {
    double cn = 1.0;
    if (fl.on(F_NITROGEN)) cn = nm.div(SIP_P(kCN), SIP_P(kCN) + soilCN);
    r.rSoil = mb.soil * SIP_P(baseSoilResp) * moistEffect * tempEffect * tillEffect * cn;
  }

  // calcMethaneFlux, :1201-1214
  if (fl.on(F_ANAEROBIC)) {
    // calcMethaneMoistEffect: pow(A, e).  A is exactly +0 whenever the soil is below the anoxia threshold, and
    // pow(+0, e) = +0 for every finite e > 0 (e_pow.c zero branch): skip the evaluation in that (common) case.
    const double te = SIP_P(anaerobicTransExp);
    double mm;
    if ((__double_as_longlong(anaerobicIdx) == 0ll) & (te > 0.0) & (te < 1e300)) {
      mm = 0.0;
    } else {
      mm = nm.pow(anaerobicIdx, te);
    }
    r.soilMethane = SIP_P(soilMethaneRate) * mb.soil * tempEffect * mm;
    if (fl.on(F_LITTER_POOL)) r.litterMethane = SIP_P(litterMethaneRate) * mb.litter * tempEffect * mm;
  }

  // checkNegativeCreation, limitations.c:146-182
  {
    const double turnover = mb.leaf * SIP_P(leafTurnoverRate);
    const double leafDef = dv.byLen(mb.leaf) + r.leafCreation - turnover;
    if (leafDef < 0) {
      r.woodCreation += leafDef;
      r.leafCreation -= leafDef;
    }
    const double fineDef = dv.byLen(mb.fine) + r.fineRootCreation - r.fineRootLoss;
    const double coarseDef = dv.byLen(mb.coarse) + r.coarseRootCreation - r.coarseRootLoss;
    if ((fineDef < 0.0) != (coarseDef < 0.0)) {
      if (fineDef < 0.0) {
        r.coarseRootCreation += fineDef;
        r.fineRootCreation -= fineDef;
      }
      if (coarseDef < 0.0) {
        r.fineRootCreation += coarseDef;
        r.coarseRootCreation -= coarseDef;
      }
    }
  }

  double nDemand = 0.0;  // calcPlantNDemandFlux() of the final creation fluxes
  if (fl.on(F_NITROGEN)) {
    // calcNResorptionFluxes, nitrogen.c:170-196
    if (r.woodCreation + r.leafCreation + r.fineRootCreation + r.coarseRootCreation < 0.0) {
      r.reductionNResorption -= (dv.byLeafCN(r.leafCreation) + dv.byWoodCN(r.woodCreation) +
                                 dv.byWoodCN(r.coarseRootCreation) + dv.byFineCN(r.fineRootCreation));
    }
    r.leafOffNResorption += dv.byLeafCN(SIP_P(leafNResorptionFrac) * r.leafLitter);
    // calcNVolatilizationFlux, nitrogen.c:15-25 (+ calcVolatilizationMoistEffect, depeffects.c:90-96)
    {
      const double dw = 0.05 + 3.8 * anaerobicIdx * (1 - anaerobicIdx);
      r.nVolatilization = SIP_P(nVolatilizationFrac) * mb.minN * tempEffect * dw;
    }
    // calcNLeachingFlux, nitrogen.c:30-40
    {
      const double dOverWhc = dv.byWhc(r.drainage);
      const double phi = (dOverWhc < 1) ? dOverWhc : 1;
      r.nLeaching = mb.minN * phi * SIP_P(nLeachingFrac);
    }
    // calcNPoolFluxes, nitrogen.c:45-83
    {
      const double litterMin = nm.div(r.rLitter, litterCN);
      const double soilMin = nm.div(r.rSoil, soilCN);
      const double l2sN = nm.div(r.litterToSoil, litterCN);
      const double inputs = l2sN + dv.byFineCN(r.fineRootLoss) + dv.byWoodCN(r.coarseRootLoss);
      const double sat =
          fl.on(F_CSAT) ? clip01(nm.divs(mb.soil, SIP_P(soilCSaturation), SIP_K(kSeedCSat))) : 0.0;
      r.nOrgLitter = dv.byLeafCN(r.leafLitter) - r.leafOffNResorption + dv.byWoodCN(r.woodLitter) - litterMin -
                     l2sN + (inputs * sat);
      r.nOrgSoil = inputs * (1 - sat) - soilMin;
      r.nMin = litterMin + soilMin;
    }
    nDemand = n_fix_and_uptake(fl, dv, mb, r, len);  // nitrogen.c:156-168

    // checkMineralNLimitation, limitations.c:119-130, then checkNitrogenLimitation, limitations.c:69-114
    {
      const double pool = mb.minN + (r.nMin + r.eventMinN) * len;
      const double loss = (r.nLeaching + r.nVolatilization) * len;
      const bool lossLimited = (loss > kTiny) & (loss > pool);
      // both tests behind one: the second one's inputs only change when the first one acts, so evaluated up front it
      // is the real test whenever the first is false -- and when the first is true the block is entered anyway
      const double uptakeDemand0 = r.nUptake * len;
      const bool uptakeMaybe = (uptakeDemand0 > kTiny) & (uptakeDemand0 > mb.minN + n_non_uptake(r) * len);
      if (lossLimited | uptakeMaybe)
      {
        if (lossLimited) {
          const double red = nm.div(pool, loss);
          r.nLeaching *= red;
          r.nVolatilization *= red;
          mb.status |= SIPNET_GPU_ST_MINN_LIMITED;
          if (DEBUG) ++ext.cnt[SIPNET_GPU_CNT_MINN_LIMITED];
        }
        const double uptakeDemand = r.nUptake * len;
        const double nonUptake = n_non_uptake(r) * len;
        const double avail = mb.minN + nonUptake;
        if (uptakeDemand > kTiny && uptakeDemand > avail) {
          const double unclaimed = n_unclaimed_storage(dv, mb, r, len);
          const double demand = n_demand(fl, dv, r) * len;
          const double uptakeFrac = 1 - n_fix_frac(dv, mb);
          const double red = nm.div(nm.div(avail, uptakeFrac) + unclaimed, demand);
          mb.status |= SIPNET_GPU_ST_N_LIMITED;  // the reference's logInfo, limitations.c:98-102
          if (DEBUG) ++ext.cnt[SIPNET_GPU_CNT_N_LIMITED];
          r.woodCreation *= red;
          r.leafCreation *= red;
          r.fineRootCreation *= red;
          r.coarseRootCreation *= red;
          nDemand = n_fix_and_uptake(fl, dv, mb, r, len);
        }
      }
    }
  }

  // writeLeafOnEventIfNeeded, sipnet.c:1230-1247
  if (fl.on(F_EVENTS) && ((r.leafOnCreation > kTiny) | (r.eventLeafOnCreation > kTiny))) {
    if (r.leafOnCreation > kTiny) {
      const double v[2] = {r.leafOnCreation * len, r.leafOnCreationFromWood * len};
      rec.add(mb, SIPNET_EV_LEAFON, 0, 2, v);
    }
    if (r.eventLeafOnCreation > kTiny) {
      const double v[2] = {r.eventLeafOnCreation * len, r.eventLeafOnCreationFromWood * len};
      rec.add(mb, SIPNET_EV_LEAFON, 1, 2, v);
    }
  }

  // ---------------- updatePoolsAndBalance, sipnet.c:1769-1806 ---------------------------
  // the mass-balance tracker (balance.c) is diagnostic: only the validation dump evaluates it
  double balPreC = 0.0, balPreN = 0.0, balPostC = 0.0, balPostN = 0.0;
  if (DEBUG) mass_totals(fl, nm, prm, mb, balPreC, balPreN);  // updateBalanceTrackerPreUpdate, balance.c:35-38
  // updatePoolsForEvents, events.c:744-790
  mb.wood += r.eventWoodC * len;
  mb.leaf += r.eventLeafC * len;
  mb.soil += r.eventSoilC * len;
  if (fl.on(F_LITTER_POOL)) mb.litter += r.eventLitterC * len;
  mb.wood -= r.eventLeafOnCreationFromWood * len;
  {
    const double evFromRoot = r.eventLeafOnCreation - r.eventLeafOnCreationFromWood;
    mb.coarse -= evFromRoot * len;
  }
  mb.leaf += (r.eventLeafOnCreation - r.eventLeafOffLitter) * len;
  if (fl.on(F_LITTER_POOL)) {
    mb.litter += r.eventLeafOffLitter * len;
  } else {
    mb.soil += r.eventLeafOffLitter * len;
  }
  mb.coarse += r.eventCoarseRootC * len;
  mb.fine += r.eventFineRootC * len;
  mb.water = nc_add(mb.water, nc_mul(r.eventSoilWater, len));
  if (fl.on(F_NITROGEN)) {
    mb.minN += r.eventMinN * len;
    mb.orgN += r.eventSoilOrgN * len;
    mb.litN += r.eventLitterN * len;
    const double onN = n_leafon_from_c(dv, r.eventLeafOnCreation);
    mb.storN += (r.eventLeafOffNResorption - onN) * len;
  }

  // updateMainPools, sipnet.c:1579-1626
  {
    const double ra = r.rVeg + r.rFineRoot + r.rCoarseRoot;
    const double alloc = r.leafCreation + r.woodCreation + r.fineRootCreation + r.coarseRootCreation;
    mb.delta += ((r.photosynthesis - ra) - alloc) * len;
    mb.wood += (r.woodCreation - r.woodLitter - r.leafOnCreationFromWood) * len;
    mb.leaf += (r.leafCreation + r.leafOnCreation - r.leafLitter) * len;
    mb.water = nc_add(mb.water, nc_mul(r.rain + r.snowMelt - r.immedEvap - r.fastFlow - r.evaporation - r.transpiration - r.drainage, len));
    mb.snow += (r.snowFall - r.snowMelt - r.sublimation) * len;
  }

  // updatePoolsForSoil, sipnet.c:1634-1680
  if (fl.on(F_LITTER_POOL)) {
    const double inputs = r.coarseRootLoss + r.fineRootLoss + r.litterToSoil;
    const double sat = fl.on(F_CSAT) ? clip01(nm.divs(mb.soil, SIP_P(soilCSaturation), SIP_K(kSeedCSat))) : 0.0;
    mb.litter +=
        (r.woodLitter + r.leafLitter + (inputs * sat) - r.litterToSoil - r.rLitter - r.litterMethane) * len;
    mb.soil += (inputs * (1 - sat) - r.rSoil - r.soilMethane) * len;
  } else {
    mb.soil += (r.coarseRootLoss + r.fineRootLoss + r.woodLitter + r.leafLitter - r.rSoil - r.soilMethane) * len;
  }
  {
    const double fromRoot = r.leafOnCreation - r.leafOnCreationFromWood;
    mb.coarse += (r.coarseRootCreation - r.coarseRootLoss - fromRoot) * len;
    mb.fine += (r.fineRootCreation - r.fineRootLoss) * len;
  }

  // updateNitrogenPools, nitrogen.c:210-239
  if (fl.on(F_NITROGEN)) {
    const double demand = nDemand;  // the creation fluxes have not changed since n_fix_and_uptake() evaluated it
    const double fromStorage = demand - r.nUptake - r.nFixation;
    const double onN = n_leafon_from_c(dv, r.leafOnCreation);
    mb.storN += (r.leafOffNResorption + r.reductionNResorption - fromStorage - onN) * len;
    const double nonUptake = n_non_uptake(r);
    mb.minN += (nonUptake - r.nUptake) * len;
    mb.orgN += r.nOrgSoil * len;
    mb.litN += r.nOrgLitter * len;
  }

  if (DEBUG) mass_totals(fl, nm, prm, mb, balPostC, balPostN);  // updateBalanceTrackerPostUpdate, balance.c:40-43

  // checkForMortality, sipnet.c:1688-1767
  const bool hasBio = has_biomass(mb);
  // updateEventTrackers, events.c:811-822: the tillage modifier is last read by the fluxes above, so its decay can sit
  // here, behind the same "anything unusual on this step?" test as the mortality check
  const bool tilled = mb.dTill > 0;
  if ((alive == hasBio) & !tilled) {
  } else {
  if (tilled) {
    mb.dTill *= c.tillDecay;
    if (mb.dTill < 0.01) mb.dTill = 0.0;
  }
  if (alive == hasBio) {  // the usual step: nothing changes (one test)
  } else if (!alive) {
    alive = true;
  } else {
    alive = false;
    mb.status |= SIPNET_GPU_ST_DIED;
    const double totWood = mb.wood + mb.delta;
    const double totRoot = mb.fine + mb.coarse;
    mb.soil += totRoot;
    if (fl.on(F_LITTER_POOL)) {
      mb.litter += mb.wood + mb.leaf + mb.delta;
    } else {
      mb.soil += mb.wood + mb.leaf + mb.delta;
    }
    if (fl.on(F_NITROGEN)) {
      mb.orgN += dv.byFineCN(mb.fine) + dv.byWoodCN(mb.coarse);
      mb.litN += dv.byWoodCN(mb.wood) + dv.byLeafCN(mb.leaf) + mb.storN;
    }
    mb.wood = 0.0;
    mb.leaf = 0.0;
    mb.coarse = 0.0;
    mb.fine = 0.0;
    mb.delta = 0.0;
    if (fl.on(F_NITROGEN)) mb.storN = 0.0;
    ring_reset(mb, rg, 0.0, ringHead);
    if (fl.on(F_EVENTS)) {
      const double v[4] = {harvRemoved, harvTransferred, totWood, totRoot};
      rec.add(mb, SIPNET_EV_PLANTDEATH, 0, 4, v);
    }
  }

  }
  // ensureNonNegativeStocks, sipnet.c:1368-1397
  {
    bool clamped = false;
    const auto clamp = [&](double &v, double floorv) {
      bool one = false;
      if (floorv == 0.0) clamp_stock(v, one);
      else clamp_stock(v, floorv, one);
      clamped = clamped | one;
      if (DEBUG) ext.cnt[SIPNET_GPU_CNT_CLAMPED] += one ? 1u : 0u;  // one warning per clamped stock
    };
    clamp(mb.wood, 0);
    clamp(mb.leaf, 0);
    if (fl.on(F_LITTER_POOL)) clamp(mb.litter, 0);
    clamp(mb.soil, 0);
    clamp(mb.coarse, 0);
    clamp(mb.fine, 0);
    clamp(mb.water, 0);
    clamp(mb.snow, kTiny);
    clamp(mb.minN, 0);
    clamp(mb.orgN, 0);
    clamp(mb.litN, 0);
    clamp(mb.storN, 0);
    mb.status |= clamped ? SIPNET_GPU_ST_CLAMPED : 0u;
  }

  double balDeltaC = 0.0, balDeltaN = 0.0;
  if (DEBUG) {  // updateBalanceTrackerPostClamp + checkBalance, balance.c:45-148
    double finalC, finalN;
    mass_totals(fl, nm, prm, mb, finalC, finalN);
    double clampedC = finalC - balPostC;
    if (clampedC < kEps) clampedC = 0;
    double clampedN = finalN - balPostN;
    if (clampedN < kEps) clampedN = 0;
    double inputsC = r.photosynthesis + r.eventInputC;
    double outputsC = r.rVeg + r.rFineRoot + r.rCoarseRoot + r.rSoil + r.soilMethane + r.eventOutputC;
    if (fl.on(F_LITTER_POOL)) outputsC += r.rLitter + r.litterMethane;
    inputsC *= len;
    outputsC *= len;
    double inputsN = 0.0, outputsN = 0.0;
    if (fl.on(F_NITROGEN)) {
      inputsN = r.nFixation + r.eventInputN;
      outputsN = r.nLeaching + r.nVolatilization + r.eventOutputN;
      inputsN *= len;
      outputsN *= len;
    }
    inputsC += clampedC;
    if (fl.on(F_NITROGEN)) inputsN += clampedN;
    const double poolCDelta = finalC - balPreC;
    const double systemCDelta = inputsC - outputsC;
    balDeltaC = poolCDelta - systemCDelta;
    const double poolNDelta = finalN - balPreN;
    const double systemNDelta = outputsN - inputsN;
    balDeltaN = poolNDelta + systemNDelta;
    if (fabs(balDeltaC) < kEps) balDeltaC = 0.0;
    if (fabs(balDeltaN) < kEps) balDeltaN = 0.0;
    if (fabs(balDeltaC) > 0.0 || fabs(balDeltaN) > 0.0) mb.status |= SIPNET_GPU_ST_BALANCE;  // the reference warns
  }

  // ---------------- updateTrackers, sipnet.c:1420-1496 ------------------------------------
  StepTrack t;
  if (c.year != mb.trkLastYear) {
    if (DEBUG) ext.yGpp = ext.yRtot = ext.yRa = ext.yRh = ext.yNpp = ext.yNee = 0.0;
    mb.gdd = 0.0;
    mb.trkLastYear = c.year;
  }
  t.gpp = r.photosynthesis * len;
  t.rh = (r.rLitter + r.rSoil) * len;
  t.rAboveground = (r.rVeg) * len;
  t.rRoot = (r.rCoarseRoot + r.rFineRoot) * len;
  t.rSoil = t.rRoot + t.rh;
  t.ra = t.rRoot + t.rAboveground;
  t.rtot = t.ra + t.rh;
  t.npp = t.gpp - t.ra;
  t.nee = -1.0 * (t.npp - t.rh);
  if (DEBUG) {
    ext.yGpp += t.gpp;
    ext.yRa += t.ra;
    ext.yRh += t.rh;
    ext.yRtot += t.rtot;
    ext.yNpp += t.npp;
    ext.yNee += t.nee;
    ext.tGpp += t.gpp;
    ext.tRa += t.ra;
    ext.tRh += t.rh;
    ext.tRtot += t.rtot;
    ext.tNpp += t.npp;
  }
  mb.totNee += t.nee;
  t.woodCreation = r.woodCreation * len;
  t.methane = (r.soilMethane + r.litterMethane) * len;
  t.evapotranspiration = (r.transpiration + r.immedEvap + r.evaporation + r.sublimation + r.eventEvap) * len;
  mb.wetFrac = nm.divs(oldSoilWater + mb.water, SIP_K(kTwoWhc), SIP_K(kSeedTwoWhc));
  if (DEBUG) ext.yLitter += r.leafLitter + r.eventLeafOffLitter;
  if (DEBUG) {
    ext.harvRemoved = harvRemoved;
    ext.harvTransferred = harvTransferred;
  }
  if (fl.on(F_GDD)) {
    mb.gdd += c.gdd;
  } else {
    mb.gdd = 0.0;
  }
  t.meanNPP = nm.divs(mb.ringSum, kMeanNppDays, kc.seed5);  // read before this step's insert (sipnet.c:1486 precedes :1852)
  if (fl.on(F_NITROGEN)) {
    t.n2o = r.nVolatilization * len;
    t.nLeaching = r.nLeaching * len;
    t.nFixation = r.nFixation * len;
    t.nUptake = r.nUptake * len;
  } else {
    t.n2o = t.nLeaching = t.nFixation = t.nUptake = 0.0;
  }

  // ---------------- outputs (outputState, sipnet.c:455-472) ---------------------------------
  // column -> value; with a compile-time column (all 32 kept) the switch folds away, with a run-time column (only
  // the summary columns are kept) it is one uniform jump per stored value instead of 32 tests per step
  const auto column = [&](int col) -> double {
    switch (col) {
      case SIPNET_O_plantWoodC: return mb.wood + mb.delta;
      case SIPNET_O_plantLeafC: return mb.leaf;
      case SIPNET_O_woodCreation: return t.woodCreation;
      case SIPNET_O_soilC: return mb.soil;
      case SIPNET_O_coarseRootC: return mb.coarse;
      case SIPNET_O_fineRootC: return mb.fine;
      case SIPNET_O_litterC: return mb.litter;
      case SIPNET_O_soilWater: return mb.water;
      case SIPNET_O_soilWetnessFrac: return mb.wetFrac;
      case SIPNET_O_snow: return mb.snow;
      case SIPNET_O_npp: return t.npp;
      case SIPNET_O_nee: return t.nee;
      case SIPNET_O_cumNEE: return mb.totNee;
      case SIPNET_O_gpp: return t.gpp;
      case SIPNET_O_rAboveground: return t.rAboveground;
      case SIPNET_O_rSoil: return t.rSoil;
      case SIPNET_O_rRoot: return t.rRoot;
      case SIPNET_O_ra: return t.ra;
      case SIPNET_O_rh: return t.rh;
      case SIPNET_O_rtot: return t.rtot;
      case SIPNET_O_evapotranspiration: return t.evapotranspiration;
      case SIPNET_O_fluxestranspiration: return r.transpiration;
      case SIPNET_O_minN: return mb.minN;
      case SIPNET_O_soilOrgN: return mb.orgN;
      case SIPNET_O_litterN: return mb.litN;
      case SIPNET_O_plantStorageN: return mb.storN;
      case SIPNET_O_n2o: return t.n2o;
      case SIPNET_O_nLeaching: return t.nLeaching;
      case SIPNET_O_nFixation: return t.nFixation;
      case SIPNET_O_nUptake: return t.nUptake;
      case SIPNET_O_ch4: return t.methane;
      default: return mb.delta;  // SIPNET_O_nppStorage
    }
  };
  emit.outputs(column);
  emit.nee(t.nee);

  // ---------------- updateMeanTrackers, sipnet.c:1546-1570 -----------------------------------
  {
    const double npp = r.photosynthesis - r.rVeg - r.rCoarseRoot - r.rFineRoot;
    if (!ring_push_usual(mb, rg, npp, len, ringHead, alive) && alive) ring_push(mb, rg, npp, len, ringHead);
  }

  if (DEBUG) {  // debug-log field order, debug_log.c:51-170
    int k = 0;
    const double ev[13] = {mb.wood, mb.leaf, mb.soil, mb.water, mb.litter, mb.snow, mb.coarse,
                           mb.fine, mb.minN, mb.orgN, mb.litN,  mb.storN,  mb.delta};
#pragma unroll
    for (int i = 0; i < 13; ++i) emit.dbg(k++, ev[i]);
    const double rv[56] = {r.photosynthesis, r.leafLitter, r.woodLitter, r.rVeg, r.rSoil, r.rain, r.transpiration,
                           r.drainage, r.litterToSoil, r.rLitter, r.snowFall, r.snowMelt, r.sublimation,
                           r.immedEvap, r.fastFlow, r.evaporation, r.fineRootLoss, r.coarseRootLoss,
                           r.fineRootCreation, r.coarseRootCreation, r.rCoarseRoot, r.rFineRoot, r.leafCreation,
                           r.woodCreation, r.leafOnCreation, r.leafOnCreationFromWood, r.nVolatilization,
                           r.nLeaching, r.nOrgSoil, r.nOrgLitter, r.nMin, r.nFixation, r.nUptake,
                           r.leafOffNResorption, r.reductionNResorption, r.eventLeafC, r.eventWoodC,
                           r.eventFineRootC, r.eventCoarseRootC, r.eventEvap, r.eventSoilWater, r.eventSoilC,
                           r.eventLitterC, r.eventMinN, r.eventSoilOrgN, r.eventLitterN, r.eventInputC,
                           r.eventOutputC, r.eventInputN, r.eventOutputN, r.eventLeafOnCreation,
                           r.eventLeafOnCreationFromWood, r.eventLeafOffLitter, r.eventLeafOffNResorption,
                           r.soilMethane, r.litterMethane};
#pragma unroll
    for (int i = 0; i < 56; ++i) emit.dbg(k++, rv[i]);
    const double tv[33] = {t.gpp, t.rtot, t.ra, t.rh, t.rRoot, t.rSoil, t.rAboveground, t.npp, t.nee,
                           t.woodCreation, mb.gdd, t.evapotranspiration, mb.wetFrac, ext.yGpp, ext.yRtot,
                           ext.yRa, ext.yRh, ext.yNpp, ext.yNee, ext.yLitter, ext.tGpp, ext.tRtot, ext.tRa,
                           ext.tRh, ext.tNpp, mb.totNee, (double)mb.trkLastYear, t.methane, t.n2o, t.nLeaching,
                           t.nFixation, t.nUptake, t.meanNPP};
#pragma unroll
    for (int i = 0; i < 33; ++i) emit.dbg(k++, tv[i]);
    emit.dbg(k++, (double)mb.didGrowth);
    emit.dbg(k++, (double)mb.didFall);
    emit.dbg(k++, (double)mb.phenLastYear);
    emit.dbg(k++, alive ? 1.0 : 0.0);
    emit.dbg(k++, balDeltaC);  // rows SIPNET_GPU_NDEBUG ..: SIPNET_GPU_GATHER_BALANCE
    emit.dbg(k++, balDeltaN);
  }

  // ---------------- updateEventTrackers, events.c:811-822 -------------------------------------

}

}  // namespace sip
