// sip_types.cuh -- device-side data layout of the batched SIPNET integrator.
//
// HBM layout (all FP64, structure-of-arrays, indexed by ensemble member m):
//   params  [80][ld]      derived parameters (post setupModel(), reference sipnet.c:1858-1916)
//   state   [NSTATE][ld]  carried state between run() segments (restart.c:148-308 field list)
//   ring_v  [cap][ld]     5-day mean-NPP ring values   (runmean.c)
//   ring_w  [cap][ld]     5-day mean-NPP ring weights
//   out     [ncols][n][ld] per-step output columns of the current run range (outputState order)
// Per site (shared by all its members, staged to shared memory per step chunk):
//   ClimRec [T_s]         one 112-byte record per climate step (ClimateVars, state.h:12-49,
//                         plus the pre-bound event range of that step)
//   EventDev[nev]         events.in rows in file order
#pragma once
#include <cstdint>

#include "../../include/sipnet_gpu.h"

namespace sip {

constexpr double kTiny = 0.000001;  // reference common/util.h:14
constexpr double kEps = 1e-8;       // reference balance.h:6
constexpr double kMeanNppDays = 5.0;  // MEAN_NPP_DAYS, sipnet.c:39
constexpr int kRingMax = 250;         // MEAN_NPP_MAX_ENTRIES, sipnet.c:40
// internal status bit (never visible after a run): the optimistic kernel met an input outside its guards in the
// CURRENT segment; the replay kernel re-runs the member from the segment's start state and clears it
constexpr uint32_t kStNeedsReplay = 0x80000000u;

// Device parameter rows: the 80 of struct Parameters plus derived per-member constants.
constexpr int kPsnTRangeSqSlot = SIPNET_GPU_NPARAMS;  // pow((psnTMax - psnTMin) / 2.0, 2), sipnet.c:622
// log_inline() of the member-constant pow() bases, as hi + lo (sip_libm.cuh pow_log); NaN = base is not
// "regular" and the full pow() runs instead.  Reusing them is bit-identical to calling pow(base, y).
constexpr int kLogVegQ10 = SIPNET_GPU_NPARAMS + 1;     // vegRespQ10, sipnet.c:1056,1067
constexpr int kLogCoarseQ10 = SIPNET_GPU_NPARAMS + 3;  // coarseRootQ10, sipnet.c:1076
constexpr int kLogFineQ10 = SIPNET_GPU_NPARAMS + 5;    // fineRootQ10, sipnet.c:1076
constexpr int kLogSoilQ10 = SIPNET_GPU_NPARAMS + 7;    // soilRespQ10, depeffects.c:74
// division seeds (sip_num.cuh FastNum::seed) of the member-constant divisors
constexpr int kSeedLeafCSpWt = SIPNET_GPU_NPARAMS + 9;
constexpr int kSeedPsnTRangeSq = SIPNET_GPU_NPARAMS + 10;
constexpr int kSeedHalfSatPar = SIPNET_GPU_NPARAMS + 11;
constexpr int kSeedWhc = SIPNET_GPU_NPARAMS + 12;
constexpr int kSeedTwoWhc = SIPNET_GPU_NPARAMS + 13;
constexpr int kSeedLeafCN = SIPNET_GPU_NPARAMS + 14;
constexpr int kSeedWoodCN = SIPNET_GPU_NPARAMS + 15;
constexpr int kSeedFineRootCN = SIPNET_GPU_NPARAMS + 16;
constexpr int kSeedFAnoxia = SIPNET_GPU_NPARAMS + 17;
constexpr int kSeedOneMinusFa = SIPNET_GPU_NPARAMS + 18;
constexpr int kSeedCSat = SIPNET_GPU_NPARAMS + 19;
// member-constant sub-expressions of potPsn() / depeffects (same operations, evaluated once)
constexpr int kRespPerGram = SIPNET_GPU_NPARAMS + 20;  // baseFolRespFrac * aMax, sipnet.c:614
constexpr int kGrossAMax = SIPNET_GPU_NPARAMS + 21;    // aMax * aMaxFrac + respPerGram, sipnet.c:617
constexpr int kConvBase = SIPNET_GPU_NPARAMS + 22;     // 12 * (1/1e9) * (leafCSpWt / cFracLeaf), sipnet.c:632-633
constexpr int kOneMinusFa = SIPNET_GPU_NPARAMS + 23;   // 1 - fAnoxia, depeffects.c:21
constexpr int kTwoWhc = SIPNET_GPU_NPARAMS + 24;       // 2.0 * soilWHC, sipnet.c:1474
constexpr int kOneMinusFracLitResp = SIPNET_GPU_NPARAMS + 25;  // 1.0 - fracLitterRespired, sipnet.c:1165
constexpr int kNParamDev = SIPNET_GPU_NPARAMS + 26;

// Rows the time loop never reads (initial conditions and factors already folded into derived rows):
// they stay in HBM only.  The shared-memory tile holds the remaining kNTileRows rows, which is what lets
// two 128-member blocks share an SM (2 x (90 x 128 x 8 B + 15 KB) < 227 KB).
// tests/test_abi.py::test_tile_skips_only_unused_rows keeps this list honest.
#define SIP_TILE_SKIP_LIST                                                                                         \
  SIPNET_P_plantWoodInit, SIPNET_P_laiInit, SIPNET_P_soilInit, SIPNET_P_soilWFracInit, SIPNET_P_aMax,                \
      SIPNET_P_aMaxFrac, SIPNET_P_baseFolRespFrac, SIPNET_P_cFracLeaf, SIPNET_P_litterInit, SIPNET_P_snowInit,       \
      SIPNET_P_fineRootFrac, SIPNET_P_coarseRootFrac, SIPNET_P_minNInit, SIPNET_P_soilOrgNInit,                      \
      SIPNET_P_litterOrgNInit, SIPNET_P_plantStorageNInit
constexpr int kTileSkip[] = {SIP_TILE_SKIP_LIST};
constexpr int kNTileSkip = (int)(sizeof(kTileSkip) / sizeof(kTileSkip[0]));
constexpr int kNTileRows = kNParamDev - kNTileSkip;
// device row k -> tile row, or -1 when the row is not staged (the list is ascending)
__host__ __device__ constexpr int tile_slot(int k) {
  constexpr int skip[] = {SIP_TILE_SKIP_LIST};
  int below = 0;
  for (int i = 0; i < kNTileSkip; ++i) {
    if (skip[i] == k) return -1;
    if (skip[i] < k) ++below;
  }
  return k - below;
}

// flag bits (runtime mask / compile-time specialisation)
enum : uint32_t {
  F_EVENTS = 1u << 0,
  F_GDD = 1u << 1,
  F_GROWTH_RESP = 1u << 2,
  F_LEAF_WATER = 1u << 3,
  F_LITTER_POOL = 1u << 4,
  F_SNOW = 1u << 5,
  F_SOIL_PHENOL = 1u << 6,
  F_WATER_HRESP = 1u << 7,
  F_NITROGEN = 1u << 8,
  F_ANAEROBIC = 1u << 9,
  F_FLOODING = 1u << 10,
  F_CSAT = 1u << 11,
};

// One climate step as staged to shared memory: 20 doubles + 4 ints = 176 bytes
// (multiple of 16 so a chunk is a legal cp.async.bulk size).
struct alignas(16) ClimRec {
  // dayFrac = (double)day + time / 24.0: the step's position in the year as pastLeafGrowth / pastLeafFall form it
  // (sipnet.c:720,739), evaluated on the host with the same two IEEE operations; `time` itself is not needed here
  double dayFrac, length, tair, tsoil, par, precip, vpd, vpdSoil, vPress, wspd, gdd;
  // exp(-length * (1 / 30.0)) evaluated on the host with the host libm: the
  // tillage decay factor of updateEventTrackers() (events.c:816) depends on the
  // step length only, so it is hoisted out of the member loop.
  double tillDecay;
  // log_inline(vpd) as hi + lo (sip_libm.cuh pow_log) for pow(vpd, dVpdExp), sipnet.c:626; NaN = not regular
  double logVpdHi, logVpdLo;
  // 1 / length when length is a power of two (then x / length == x * invLen exactly), else 0
  double invLenPow2;
  // site-level sub-expressions, evaluated on the host with the same IEEE operations:
  double tair10;      // tair / 10.0            (vegResp, sipnet.c:1067)
  double tsoil10;     // tsoil / 10 (== / 10.0) (calcRootResp sipnet.c:1076, calcTempEffect depeffects.c:74)
  double precipRate;  // precip / length        (calcPrecip, sipnet.c:851,858)
  double sublK;       // CONVERSION_S * (E_STAR_SNOW - vPress)   (snowPack, sipnet.c:910)
  double evapK;       // CONVERSION * vpdSoil                    (calcSoilWaterFluxes, sipnet.c:1000)
  int32_t year, day;
  int32_t evBegin, evEnd;  // events of this step: [evBegin, evEnd) in the site's EventDev array
};
static_assert(sizeof(ClimRec) == 176 && sizeof(ClimRec) % 16 == 0, "ClimRec must be a multiple of 16 bytes");

struct alignas(16) EventDev {
  double p[4];
  int32_t type, method, pad0, pad1;
};
static_assert(sizeof(EventDev) == 48, "EventDev must be 48 bytes");

struct SiteDev {
  const ClimRec *clim;
  const EventDev *events;
  const double *neeObs;  // may be null
  int64_t nsteps;
  int32_t member0, memberCount;
};

// One block of consecutive members.  They belong to `site`; with 128-member blocks the tail of a site may share its
// block with the head of the next site that has members (`site1`): lanes [0, count0) are `site`, lanes
// [count0, count) are `site1`.  A site of 100 members would otherwise leave 28 of every 128 lanes idle.
struct BlockDesc {
  int32_t site, member0, count, site1;
  int32_t count0, pad0, pad1, pad2;
};

// launch-lifetime constants of the step: log_inline(2.0) and the division seeds (sip_num.cuh FastNum::seed) of the
// literal divisors.  Evaluated once per handle on the device (consts_kernel) and handed to every launch in the
// kernel's parameter space, where they are constant-bank operands instead of registers.
struct StepConsts {
  double log2Hi, log2Lo;
  double seed10, seed5, seed18, seed24;  // 10.0 (Q10 exponents), MEAN_NPP_DAYS, 3.0 * NUM_LAYERS, 24.0 (hours)
};

struct RunArgs {
  StepConsts kc;
  int64_t ld;
  int64_t nmembers;
  const double *params;
  double *state;
  double *ringV;
  double *ringW;
  uint32_t *status;
  uint32_t *counters;  // [SIPNET_GPU_NCOUNTERS][ld] or null (DEBUG instantiation only)
  // segment-start copies used when a member is replayed by the general kernel
  const double *stateBackup;
  const double *ringVBackup;
  const double *ringWBackup;
  const uint32_t *statusBackup;
  const double *loglikBackup;
  const double *loglikNBackup;
  const int32_t *recCountBackup;
  const BlockDesc *blocks;
  const SiteDev *sites;
  int64_t stepBegin, stepEnd;
  // outputs (null => not requested)
  double *out;       // [ncols][outSteps][ld]
  int64_t outSteps;  // steps in this run range
  double *dbg;       // [NDEBUG][outSteps][ld]
  double *loglik;    // [ld] accumulators
  double *loglikN;   // [ld]
  sipnet_gpu_event_record *recs;  // [nmembers][maxRecs]
  int32_t *recCount;              // [nmembers]
  int32_t maxRecs;
  int32_t ringCap;
  uint32_t flags;          // runtime flag mask (generic kernel)
  double invSigma;         // 1 / nee_sigma
  double logNorm;          // -log(sigma) - 0.5*log(2*pi)
  int8_t colSlot[SIPNET_GPU_NOUT];  // output column -> slot in `out`, or -1
  int8_t slotCol[SIPNET_GPU_NOUT];  // slot -> output column (first nOutCols entries)
  int32_t nOutCols;
  // the usual summary columns are stored without going through the column switch: byte offset of the NEE / GPP slot
  // from the step's first column (-1 = not kept); the other kept columns follow as (column, byte offset) pairs
  int64_t neeOff, gppOff;
  int32_t onlyNeeGpp;  // exactly NEE and GPP are kept: the step stores them behind ONE launch-uniform test
  int32_t nSlowCols;
  int8_t slowCol[SIPNET_GPU_NOUT];
  int64_t slowOff[SIPNET_GPU_NOUT];
  // dynamic scheduling of (block descriptor, sub-range of steps) work items over a persistent grid; a null
  // workCounter means one CTA per block descriptor over the whole range (sip_kernels.cu: run_kernel)
  unsigned long long *workCounter;  // next work item
  unsigned int *progress;           // [nblocks] sub-ranges completed per block descriptor
  int32_t nblocks, itemSteps;
  int32_t mixedBlocks;  // some block descriptor holds members of two sites (the MIX kernel variants)
};

}  // namespace sip
