// sip_run.cuh -- K1, the fused step kernel (run_item / run_kernel) and its launch templates.
//
// Included by the translation units that instantiate it: sip_run_exact.cu (validation, debug dump, replay) and
// sip_run_fast_*.cu (the optimistic kernel, one flag policy x block size per file so that `make -j` builds them
// side by side; a single file took minutes).  sip_kernels.cu holds the setup kernels and the dispatcher.
// Compiled with -fmad=false: the model arithmetic keeps the reference's operation
// order, and exp/pow/division are explicit operation sequences (sip_libm.cuh,
// sip_num.cuh), so every kernel here is bit-identical to the reference binary.
// Two numerics policies of the SAME arithmetic:
//   FastNum  (production)  branch-free main paths + guard flag, ~2.3x fewer instructions
//   ExactNum (validation / debug dump / replay of members the fast kernel flagged)
//
// K1 layout: one thread = one ensemble member; a block holds members of ONE
// site, so forcing and the event schedule are block-uniform.  Per block:
//   * the members' parameter rows are copied once into a shared-memory tile
//     [kNParamDev][BLOCK] (conflict-free column access),
//   * the site's ClimRec stream is staged chunk by chunk into a 2-deep shared
//     memory ring with cp.async.bulk (TMA 1-D) completing on an mbarrier, so the
//     copy of chunk i+1 overlaps the arithmetic of chunk i,
//   * pools / trackers stay in registers for the whole run range and go back to
//     the SoA state rows once at the end,
//   * every requested output column is written with one coalesced streaming
//     store per step (consecutive lanes = consecutive members = 256 B per warp).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "sip_step.cuh"

#ifndef SIP_MIN_BLOCKS_128
#define SIP_MIN_BLOCKS_128 1
#endif

namespace sip {
namespace k1 {

constexpr int kChunkSteps = 32;  // steps staged per TMA chunk (default): 32 * 176 B = 5632 B

// Tuning policy of one instantiation: block size, resident blocks per SM asked of the compiler, steps per staged
// forcing chunk.
// 128-member blocks may hold members of TWO sites (BlockDesc): two forcing streams are staged, each in chunks of half
// the length, and a lane reads its own site's record.
// MIX instantiations exist for 128-member blocks only and are launched only when some block of the handle is mixed:
// the single-site variant keeps its longer chunks and block-uniform forcing reads (1 % faster on the C4 share).
template <int BLOCK, bool MIX = false, int MIN_BLOCKS = (BLOCK == 128 ? SIP_MIN_BLOCKS_128 : 1),
          int CHUNK = (MIX ? kChunkSteps / 2 : kChunkSteps)>
struct Tune {
  static_assert(!MIX || BLOCK == 128, "mixed blocks are a 128-member feature");
  static constexpr int kBlock = BLOCK, kMinBlocks = MIN_BLOCKS, kChunk = CHUNK;
  static constexpr bool kMix = MIX;
  static constexpr int kStreams = kMix ? 2 : 1;
};

// ---- mbarrier / bulk-copy PTX wrappers (sm_90+; SASS: SYNCS / UBLKCP) ------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- per-step output sink -----------------------------------------------------------
// FULL = all 32 outputState() columns are kept, slot == column: the store address is one running
// pointer bumped by a column stride (the emit calls come in column order).  Otherwise only the
// columns with a slot (summary outputs) are stored.
template <bool FULL>
struct Emitter {
  const RunArgs *a;
  double *outp;  // a->out + m (null => no column output)
  double *dbgp;  // a->dbg + m
  int64_t tLocal;
  const double *obs;  // site's observations or null
  int64_t tSite;
  double ll, lln;
  char *cur;
  int64_t colStride;  // BYTES between consecutive columns = outSteps * ld * 8
  __device__ __forceinline__ void begin(int64_t tl, int64_t ts) {
    tLocal = tl;
    tSite = ts;
    if (FULL || a->out != nullptr) cur = reinterpret_cast<char *>(outp + tl * a->ld);  // (launch-uniform tests)
  }
  // the following step of the same chunk: one row further (in FULL mode outputs() has walked cur through the columns)
  __device__ __forceinline__ void next() {
    ++tLocal;
    ++tSite;
    cur += a->ld * (int64_t)sizeof(double) - (FULL ? SIPNET_GPU_NOUT * colStride : 0);
  }
  // value(col) is the step's column -> value switch (sip_step.cuh)
  template <class F>
  __device__ __forceinline__ void outputs(const F &value) {
    if (FULL) {
#pragma unroll
      for (int c = 0; c < SIPNET_GPU_NOUT; ++c) {
        __stcs(reinterpret_cast<double *>(cur), value(c));
        cur += colStride;
      }
    } else if (a->onlyNeeGpp) {  // the usual summary request: one launch-uniform test, two stores
      __stcs(reinterpret_cast<double *>(cur + a->neeOff), value(SIPNET_O_nee));
      __stcs(reinterpret_cast<double *>(cur + a->gppOff), value(SIPNET_O_gpp));
    } else if (a->out != nullptr) {
      // NEE and GPP, the usual summary columns: a launch-uniform test and a store each (the column is a compile-time
      // constant, so value() folds to the variable); any other kept column goes through the column switch
      if (a->neeOff >= 0) __stcs(reinterpret_cast<double *>(cur + a->neeOff), value(SIPNET_O_nee));
      if (a->gppOff >= 0) __stcs(reinterpret_cast<double *>(cur + a->gppOff), value(SIPNET_O_gpp));
      for (int s = 0; s < a->nSlowCols; ++s)
        __stcs(reinterpret_cast<double *>(cur + a->slowOff[s]), value((int)a->slowCol[s]));
    }
  }
  __device__ __forceinline__ void dbg(int k, double v) const {
    if (dbgp != nullptr) __stcs(dbgp + ((int64_t)k * a->outSteps + tLocal) * a->ld, v);
  }
  __device__ __forceinline__ void nee(double v) {
    if (obs != nullptr) {
      const double o = __ldg(obs + tSite);
      if (o == o) {  // NaN = no observation
        const double d = (v - o) * a->invSigma;
        ll += -0.5 * d * d + a->logNorm;
        lln += 1.0;
      }
    }
  }
};

template <bool COHERENT>
__device__ __forceinline__ void load_member(const RunArgs &a, const double *state, const uint32_t *status, int64_t m,
                                            Member &mb, MemberExt &ext, bool debug) {
  // COHERENT: carried state may have been written by another SM earlier in this launch (dynamic scheduling)
  const double *s = state + m;
  const int64_t ld = a.ld;
  mb.wood = carried_load<COHERENT>(&s[SIPNET_S_plantWoodC * ld]);
  mb.leaf = carried_load<COHERENT>(&s[SIPNET_S_plantLeafC * ld]);
  mb.soil = carried_load<COHERENT>(&s[SIPNET_S_soilC * ld]);
  mb.water = carried_load<COHERENT>(&s[SIPNET_S_soilWater * ld]);
  mb.litter = carried_load<COHERENT>(&s[SIPNET_S_litterC * ld]);
  mb.snow = carried_load<COHERENT>(&s[SIPNET_S_snow * ld]);
  mb.coarse = carried_load<COHERENT>(&s[SIPNET_S_coarseRootC * ld]);
  mb.fine = carried_load<COHERENT>(&s[SIPNET_S_fineRootC * ld]);
  mb.minN = carried_load<COHERENT>(&s[SIPNET_S_minN * ld]);
  mb.orgN = carried_load<COHERENT>(&s[SIPNET_S_soilOrgN * ld]);
  mb.litN = carried_load<COHERENT>(&s[SIPNET_S_litterN * ld]);
  mb.storN = carried_load<COHERENT>(&s[SIPNET_S_plantStorageN * ld]);
  mb.delta = carried_load<COHERENT>(&s[SIPNET_S_plantCAccountingDelta * ld]);
  mb.gdd = carried_load<COHERENT>(&s[SIPNET_S_gdd * ld]);
  mb.wetFrac = carried_load<COHERENT>(&s[SIPNET_S_soilWetnessFrac * ld]);
  mb.totNee = carried_load<COHERENT>(&s[SIPNET_S_totNee * ld]);
  mb.dTill = carried_load<COHERENT>(&s[SIPNET_S_dTillMod * ld]);
  mb.ringSum = carried_load<COHERENT>(&s[SIPNET_S_meanSum * ld]);
  mb.ringStart = (int)carried_load<COHERENT>(&s[SIPNET_S_meanStart * ld]);
  mb.ringLast = (int)carried_load<COHERENT>(&s[SIPNET_S_meanLast * ld]);
  mb.trkLastYear = (int)carried_load<COHERENT>(&s[SIPNET_S_trackersLastYear * ld]);
  mb.phenLastYear = (int)carried_load<COHERENT>(&s[SIPNET_S_phenLastYear * ld]);
  mb.didGrowth = (int)carried_load<COHERENT>(&s[SIPNET_S_didLeafGrowth * ld]);
  mb.didFall = (int)carried_load<COHERENT>(&s[SIPNET_S_didLeafFall * ld]);
  mb.status = carried_load<COHERENT>(&status[m]);
  if (debug) {
    ext.yGpp = carried_load<COHERENT>(&s[SIPNET_S_yearlyGpp * ld]);
    ext.yRtot = carried_load<COHERENT>(&s[SIPNET_S_yearlyRtot * ld]);
    ext.yRa = carried_load<COHERENT>(&s[SIPNET_S_yearlyRa * ld]);
    ext.yRh = carried_load<COHERENT>(&s[SIPNET_S_yearlyRh * ld]);
    ext.yNpp = carried_load<COHERENT>(&s[SIPNET_S_yearlyNpp * ld]);
    ext.yNee = carried_load<COHERENT>(&s[SIPNET_S_yearlyNee * ld]);
    ext.yLitter = carried_load<COHERENT>(&s[SIPNET_S_yearlyLitter * ld]);
    ext.tGpp = carried_load<COHERENT>(&s[SIPNET_S_totGpp * ld]);
    ext.tRtot = carried_load<COHERENT>(&s[SIPNET_S_totRtot * ld]);
    ext.tRa = carried_load<COHERENT>(&s[SIPNET_S_totRa * ld]);
    ext.tRh = carried_load<COHERENT>(&s[SIPNET_S_totRh * ld]);
    ext.tNpp = carried_load<COHERENT>(&s[SIPNET_S_totNpp * ld]);
    ext.harvRemoved = carried_load<COHERENT>(&s[SIPNET_S_harvestFracRemoved * ld]);
    ext.harvTransferred = carried_load<COHERENT>(&s[SIPNET_S_harvestFracTransferred * ld]);
  }
}

__device__ __forceinline__ void store_member(const RunArgs &a, int64_t m, const Member &mb, const MemberExt &ext,
                                             bool debug) {
  double *s = a.state + m;
  const int64_t ld = a.ld;
  s[SIPNET_S_plantWoodC * ld] = mb.wood;
  s[SIPNET_S_plantLeafC * ld] = mb.leaf;
  s[SIPNET_S_soilC * ld] = mb.soil;
  s[SIPNET_S_soilWater * ld] = mb.water;
  s[SIPNET_S_litterC * ld] = mb.litter;
  s[SIPNET_S_snow * ld] = mb.snow;
  s[SIPNET_S_coarseRootC * ld] = mb.coarse;
  s[SIPNET_S_fineRootC * ld] = mb.fine;
  s[SIPNET_S_minN * ld] = mb.minN;
  s[SIPNET_S_soilOrgN * ld] = mb.orgN;
  s[SIPNET_S_litterN * ld] = mb.litN;
  s[SIPNET_S_plantStorageN * ld] = mb.storN;
  s[SIPNET_S_plantCAccountingDelta * ld] = mb.delta;
  s[SIPNET_S_gdd * ld] = mb.gdd;
  s[SIPNET_S_soilWetnessFrac * ld] = mb.wetFrac;
  s[SIPNET_S_totNee * ld] = mb.totNee;
  s[SIPNET_S_dTillMod * ld] = mb.dTill;
  s[SIPNET_S_meanSum * ld] = mb.ringSum;
  s[SIPNET_S_meanStart * ld] = (double)mb.ringStart;
  s[SIPNET_S_meanLast * ld] = (double)mb.ringLast;
  s[SIPNET_S_trackersLastYear * ld] = (double)mb.trkLastYear;
  s[SIPNET_S_phenLastYear * ld] = (double)mb.phenLastYear;
  s[SIPNET_S_didLeafGrowth * ld] = (double)mb.didGrowth;
  s[SIPNET_S_didLeafFall * ld] = (double)mb.didFall;
  uint32_t st = mb.status;
  if (!(isfinite(mb.wood) && isfinite(mb.leaf) && isfinite(mb.soil) && isfinite(mb.water) && isfinite(mb.litter) &&
        isfinite(mb.snow) && isfinite(mb.coarse) && isfinite(mb.fine) && isfinite(mb.minN) && isfinite(mb.orgN) &&
        isfinite(mb.litN) && isfinite(mb.storN) && isfinite(mb.delta))) {
    st |= SIPNET_GPU_ST_NONFINITE;
  }
  a.status[m] = st;
  if (debug) {
    s[SIPNET_S_yearlyGpp * ld] = ext.yGpp;
    s[SIPNET_S_yearlyRtot * ld] = ext.yRtot;
    s[SIPNET_S_yearlyRa * ld] = ext.yRa;
    s[SIPNET_S_yearlyRh * ld] = ext.yRh;
    s[SIPNET_S_yearlyNpp * ld] = ext.yNpp;
    s[SIPNET_S_yearlyNee * ld] = ext.yNee;
    s[SIPNET_S_yearlyLitter * ld] = ext.yLitter;
    s[SIPNET_S_totGpp * ld] = ext.tGpp;
    s[SIPNET_S_totRtot * ld] = ext.tRtot;
    s[SIPNET_S_totRa * ld] = ext.tRa;
    s[SIPNET_S_totRh * ld] = ext.tRh;
    s[SIPNET_S_totNpp * ld] = ext.tNpp;
    s[SIPNET_S_harvestFracRemoved * ld] = ext.harvRemoved;
    s[SIPNET_S_harvestFracTransferred * ld] = ext.harvTransferred;
  }
}

// ---- K1: fused [events -> fluxes -> pools -> trackers -> mean tracker] over a step range ----
// REPLAY = true: only members the optimistic kernel flagged (SIPNET_GPU_ST_REPLAY set during this
// segment) are integrated, starting again from the segment's start state (RunArgs::*Backup).
constexpr int kLibmTabWords = 2 * 128 + 4 * 128;  // exp table + pow-log table, 6 KB

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One work item: block descriptor `blk` (up to BLOCK members of one site) over steps [itemBegin, itemEnd).
// `sc` counts the forcing chunks this CTA has staged so far (chunk sc uses buffer sc & 1, mbarrier parity (sc >> 1) & 1).
template <class FL, bool DEBUG, class NM, int BLOCK, bool REPLAY, bool FULL, bool DYN, class TN>
__device__ __forceinline__ void run_item(const RunArgs &a, const FL &fl, int64_t blk, int64_t itemBegin, int64_t itemEnd,
                                         int &sc, double *tile, ClimRec *climBuf, uint64_t *libmTab, uint64_t *bars) {
  constexpr int kChunkSteps = TN::kChunk;
  const int tid = threadIdx.x;
  const BlockDesc bd = a.blocks[blk];
  const SiteDev site = a.sites[bd.site];
  // second site of a mixed block (block-uniform); `second` = this lane belongs to it
  const bool mixed = TN::kMix && bd.count0 < bd.count;
  const SiteDev siteB = mixed ? a.sites[bd.site1] : site;
  const bool second = TN::kMix && tid >= bd.count0;
  const int64_t m = (int64_t)bd.member0 + tid;
  bool active = tid < bd.count;
  if (REPLAY) {
    // kStNeedsReplay is per segment: set by the optimistic kernel of THIS segment only (statusBackup never holds it),
    // so a member that leaves the guards in several segments is replayed in every one of them
    active = active && ((a.status[m] & kStNeedsReplay) != 0);
    if (!__syncthreads_or(active ? 1 : 0)) return;  // nothing to replay in this block (the normal case)
  }

  const int64_t t0 = itemBegin;
  const int64_t t1A = itemEnd < site.nsteps ? itemEnd : site.nsteps;
  const int64_t t1B = itemEnd < siteB.nsteps ? itemEnd : siteB.nsteps;
  const int64_t t1 = t1A > t1B ? t1A : t1B;          // the block's range
  const int64_t myT1 = second ? t1B : t1A;           // this lane's site may end earlier
  const EventDev *myEvents = second ? siteB.events : site.events;
  const double *myObs = second ? siteB.neeObs : site.neeObs;

  // parameter tile: coalesced global reads, column-per-thread shared layout (each thread reads only its column)
  for (int k = 0; k < kNParamDev; ++k) {
    const int slot = tile_slot(k);
    if (slot >= 0) tile[slot * BLOCK + tid] = active ? a.params[(int64_t)k * a.ld + m] : 1.0;
  }
  const DirectTile prm{tile + tid, BLOCK};

  auto issue = [&](int64_t cs, int serial) {  // stage steps [cs, cs + kChunkSteps) of the block's site(s) as chunk `serial`
    const int buf = serial & 1;
    const int64_t nA = t1A - cs < kChunkSteps ? t1A - cs : kChunkSteps;
    const int64_t nB = mixed ? (t1B - cs < kChunkSteps ? t1B - cs : kChunkSteps) : 0;
    const uint32_t bytesA = nA > 0 ? (uint32_t)(nA * sizeof(ClimRec)) : 0u;
    const uint32_t bytesB = nB > 0 ? (uint32_t)(nB * sizeof(ClimRec)) : 0u;
    mbar_expect_tx(&bars[buf], bytesA + bytesB);
    if (bytesA) bulk_g2s(climBuf + buf * kChunkSteps, site.clim + cs, bytesA, &bars[buf]);
    if (bytesB) bulk_g2s(climBuf + (2 + buf) * kChunkSteps, siteB.clim + cs, bytesB, &bars[buf]);
  };
  if (tid == 0 && t1 > t0) issue(t0, sc);

  Member mb;
  MemberExt ext = {};
  if (active) {
    if (REPLAY) {  // restore the member's segment-start state: ring columns, accumulators, status
      for (int sl = 0; sl < a.ringCap; ++sl) {
        a.ringV[(int64_t)sl * a.ld + m] = a.ringVBackup[(int64_t)sl * a.ld + m];
        a.ringW[(int64_t)sl * a.ld + m] = a.ringWBackup[(int64_t)sl * a.ld + m];
      }
      if (a.loglik != nullptr) {
        a.loglik[m] = a.loglikBackup[m];
        a.loglikN[m] = a.loglikNBackup[m];
      }
      if (a.recCount != nullptr) a.recCount[m] = a.recCountBackup[m];
    }
    load_member<DYN>(a, REPLAY ? a.stateBackup : a.state, REPLAY ? a.statusBackup : a.status, m, mb, ext, DEBUG);
    if (REPLAY) mb.status = (mb.status & ~kStNeedsReplay) | SIPNET_GPU_ST_REPLAY;  // sticky, informational
    if (DEBUG && a.counters != nullptr)
      for (int k = 0; k < SIPNET_GPU_NCOUNTERS; ++k) ext.cnt[k] = carried_load<DYN>(&a.counters[(int64_t)k * a.ld + m]);
    if (mb.status & SIPNET_GPU_ST_BAD_ALLOCATION) {  // reference would have exited (sipnet.c:1117-1122)
      active = false;
      // the member is not integrated: its outputs of this range are NaN (summaries skip non-finite members)
      const double nanv = __longlong_as_double(0x7ff8000000000000ll);
      const int64_t o0 = t0 - a.stepBegin, o1 = myT1 - a.stepBegin;
      if (a.out != nullptr)
        for (int c = 0; c < SIPNET_GPU_NOUT; ++c)
          if (a.colSlot[c] >= 0)
            for (int64_t t = o0; t < o1; ++t) a.out[((int64_t)a.colSlot[c] * a.outSteps + t) * a.ld + m] = nanv;
      if (a.dbg != nullptr)
        for (int k = 0; k < SIPNET_GPU_NDEBUG + SIPNET_GPU_NBALANCE; ++k)
          for (int64_t t = o0; t < o1; ++t) a.dbg[((int64_t)k * a.outSteps + t) * a.ld + m] = nanv;
    }
  }
  NM nm;
  if constexpr (NM::kFast) {
    nm.expTab = libmTab;
    nm.powlogTab = libmTab + 2 * 128;
    // the member-constant divisors must be ordinary numbers (sip_num.cuh divisor_check)
    nm.divisor_check(SIP_P(leafCSpWt));
    nm.divisor_check(SIP_K(kPsnTRangeSqSlot));
    nm.divisor_check(SIP_P(halfSatPar));
    nm.divisor_check(SIP_P(soilWHC));
    nm.divisor_check(SIP_K(kTwoWhc));
    nm.divisor_check(SIP_P(leafCN));
    nm.divisor_check(SIP_P(woodCN));
    nm.divisor_check(SIP_P(fineRootCN));
    nm.divisor_check(SIP_P(fAnoxia));
    nm.divisor_check(SIP_K(kOneMinusFa));
    if (fl.on(F_CSAT)) nm.divisor_check(SIP_P(soilCSaturation));
  }
  const StepConsts &kc = a.kc;
  const RingRefT<DYN> rg{a.ringV + m, a.ringW + m, (unsigned)a.ld, a.ringCap};
  RecSinkT<DYN> rec{nullptr, nullptr, a.maxRecs, 0};
  if (a.recCount != nullptr && active) {
    rec.count = a.recCount + m;
    rec.recs = a.recs != nullptr ? a.recs + m * (int64_t)a.maxRecs : nullptr;
  }
  Emitter<FULL> emit{&a, a.out != nullptr ? a.out + m : nullptr, a.dbg != nullptr ? a.dbg + m : nullptr, 0,
                     myObs, 0, 0.0, 0.0, nullptr, a.outSteps * a.ld * (int64_t)sizeof(double)};
  if (active && a.loglik != nullptr) {  // continue the member's running sums (same addition order as one long run)
    emit.ll = carried_load<DYN>(&a.loglik[m]);
    emit.lln = carried_load<DYN>(&a.loglikN[m]);
  }

  for (int64_t cs = t0; cs < t1; cs += kChunkSteps, ++sc) {
    if (tid == 0 && cs + kChunkSteps < t1) issue(cs + kChunkSteps, sc + 1);  // that buffer was released by the barrier below
    mbar_wait(&bars[sc & 1], (uint32_t)((sc >> 1) & 1));
    const ClimRec *cbuf = climBuf + ((second ? 2 : 0) + (sc & 1)) * kChunkSteps;  // this lane's site's records
    const int n = (int)((myT1 - cs) < kChunkSteps ? (myT1 - cs) : kChunkSteps);   // <= 0 once its site has ended
    if (active) {
      emit.begin(cs - a.stepBegin, cs);
      rec.step = (int32_t)cs;
      for (int i = 0; i < n; ++i) {
        step<FL, DEBUG>(fl, nm, prm, cbuf[i], myEvents, mb, ext, rg, rec, emit, kc);
        emit.next();  // (one past the chunk's last row after the last step: never dereferenced)
        ++rec.step;
      }
    }
    __syncthreads();  // everyone is done reading this buffer before it is refilled
  }
  if (active) {
    if (NM::kFast && nm.bad) mb.status |= kStNeedsReplay;  // outside the optimistic guards: general kernel re-runs it
    store_member(a, m, mb, ext, DEBUG);
    if (DEBUG && a.counters != nullptr)
      for (int k = 0; k < SIPNET_GPU_NCOUNTERS; ++k) a.counters[(int64_t)k * a.ld + m] = ext.cnt[k];
    if (a.loglik != nullptr && myObs != nullptr) {  // running sums continue across segments in step order
      a.loglik[m] = emit.ll;
      a.loglikN[m] = emit.lln;
    }
  }
}

// DYN = false: CTA b integrates block descriptor b over the whole step range of the launch.
// DYN = true (persistent grid, one CTA per resident slot): work items are (block descriptor, sub-range of
// itemSteps steps), handed out in sub-range-major order by an atomic counter, so a member count that fills a
// fractional number of waves no longer leaves SMs idle.  Item (b, s) needs (b, s-1); items are claimed in order,
// so the predecessor was claimed earlier by a running CTA that waits for nothing later -- no deadlock.
template <class TN>
__host__ __device__ constexpr size_t fixed_smem_bytes() {  // forcing ring + libm tables + mbarriers
  return (size_t)TN::kStreams * 2 * TN::kChunk * sizeof(ClimRec) + kLibmTabWords * sizeof(uint64_t) + 2 * sizeof(uint64_t);
}

template <class FL, bool DEBUG, class NM, int BLOCK, bool REPLAY, bool FULL, bool DYN, class TN = Tune<BLOCK>>
__global__ void __launch_bounds__(BLOCK, TN::kMinBlocks) run_kernel(const __grid_constant__ RunArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [forcing ring 2 x kChunk][libm tables][mbarriers][parameter tile]
  ClimRec *climBuf = reinterpret_cast<ClimRec *>(smem_raw);                     // [kStreams][2][kChunk]
  uint64_t *libmTab = reinterpret_cast<uint64_t *>(climBuf + TN::kStreams * 2 * TN::kChunk);  // [kLibmTabWords]
  uint64_t *bars = libmTab + kLibmTabWords;                                     // [2]
  double *tile = reinterpret_cast<double *>(bars + 2);                          // direct: [kNTileRows][BLOCK]

  const int tid = threadIdx.x;
  const FL fl(a.flags);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (NM::kFast) {  // libm tables -> shared memory (table lookups become LDS)
    for (int i = tid; i < 2 * 128; i += BLOCK) libmTab[i] = libm::d_exp_tab[i];
    for (int i = tid; i < 4 * 128; i += BLOCK) libmTab[2 * 128 + i] = libm::d_powlog_tab[i];
  }
  __syncthreads();

  int sc = 0;
  if constexpr (!DYN) {
    run_item<FL, DEBUG, NM, BLOCK, REPLAY, FULL, false, TN>(a, fl, blockIdx.x, a.stepBegin, a.stepEnd, sc, tile, climBuf, libmTab, bars);
  } else {
    __shared__ long long sItem;
    const int64_t nsub = (a.stepEnd - a.stepBegin + a.itemSteps - 1) / a.itemSteps;
    const int64_t nItems = (int64_t)a.nblocks * nsub;
    for (;;) {
      if (tid == 0) sItem = (long long)atomicAdd(a.workCounter, 1ull);
      __syncthreads();
      const int64_t w = sItem;
      if (w >= nItems) break;
      const int64_t sub = w / a.nblocks;
      const int64_t blk = w - sub * a.nblocks;
      if (sub > 0) {
        if (tid == 0)
          while (ld_acquire_u32(&a.progress[blk]) < (unsigned)sub) __nanosleep(256);
        __syncthreads();
        __threadfence();  // acquire side for every thread: the predecessor's state/ring/status stores are visible
      }
      const int64_t itemBegin = a.stepBegin + sub * (int64_t)a.itemSteps;
      const int64_t itemEnd = itemBegin + a.itemSteps < a.stepEnd ? itemBegin + a.itemSteps : a.stepEnd;
      run_item<FL, DEBUG, NM, BLOCK, REPLAY, FULL, true, TN>(a, fl, blk, itemBegin, itemEnd, sc, tile, climBuf, libmTab, bars);
      __syncthreads();  // every member's state is stored ...
      if (tid == 0) {   // ... before the sub-range is published (release)
        const int64_t w1 = sItem;
        const int64_t s1 = w1 / a.nblocks;
        __threadfence();
        st_release_u32(&a.progress[w1 - s1 * a.nblocks], (unsigned)(s1 + 1));
      }
    }
  }
}

// ---- launch templates ----------------------------------------------------------------------------
constexpr uint32_t kMaskDefault = F_EVENTS | F_GDD | F_SNOW | F_WATER_HRESP;                       // context.c:35-46
constexpr uint32_t kMaskCropN = kMaskDefault | F_LITTER_POOL | F_ANAEROBIC | F_NITROGEN;          // russell_2 / C2-C5

constexpr int kItemSteps = 256;  // steps per dynamically scheduled work item (8 forcing chunks)
constexpr int kDynamicMaxWaves = 1 << 20;  // dynamic scheduling below this many waves of block descriptors (= always)

// A/B switch for measurements: SIPNET_GPU_DYNAMIC_MAX_WAVES overrides the wave threshold of dynamic scheduling
inline int dynamic_max_waves() {
  static const int v = [] {
    const char *e = getenv("SIPNET_GPU_DYNAMIC_MAX_WAVES");
    const int n = e ? atoi(e) : 0;
    return n > 0 ? n : kDynamicMaxWaves;
  }();
  return v;
}

template <class FL, bool DEBUG, class NM, int BLOCK, bool REPLAY, bool FULL, bool MIX = false>
static cudaError_t launch_one(const RunArgs &a, int nblocks, cudaStream_t stream) {
  using TN = Tune<BLOCK, MIX>;
  const size_t smem = fixed_smem_bytes<TN>() + sizeof(double) * kNTileRows * BLOCK;
  auto kern = run_kernel<FL, DEBUG, NM, BLOCK, REPLAY, FULL, false, TN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  RunArgs args = a;
  args.nblocks = nblocks;
  args.itemSteps = kItemSteps;
  int grid = nblocks;
  if constexpr (!REPLAY && !DEBUG && NM::kFast) {
    if (a.workCounter != nullptr) {
      // More block descriptors than resident CTAs: whole waves would quantise the run time (1.4 waves cost 2,
      // 3.46 cost ~3.6), so a persistent grid pulls (block, sub-range) items instead.  Measured: 32 768 members
      // 57.6 -> 42.3 ms, 131 072 members 122.6 -> 108.3 ms, 262 144 members (6.9 waves) 219.4 -> 216.8 ms.
      auto dyn = run_kernel<FL, DEBUG, NM, BLOCK, REPLAY, FULL, true, TN>;
      int dev = 0, sms = 0, perSm = 0;
      if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
      if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, dyn, BLOCK, smem)) != cudaSuccess) return e;
      const int resident = sms * perSm;
      if (resident > 0 && nblocks > resident && nblocks < dynamic_max_waves() * resident) {
        grid = resident;
        dyn<<<grid, BLOCK, smem, stream>>>(args);
        return cudaGetLastError();
      }
    }
  }
  args.workCounter = nullptr;
  kern<<<grid, BLOCK, smem, stream>>>(args);
  return cudaGetLastError();
}

template <class FL, int BLOCK, class NM = FastNum>
static cudaError_t launch_fast(const RunArgs &a, int nblocks, bool full, cudaStream_t stream) {
  if constexpr (BLOCK == 128) {
    if (a.mixedBlocks)  // some block holds members of two sites
      return full ? launch_one<FL, false, NM, BLOCK, false, true, true>(a, nblocks, stream)
                  : launch_one<FL, false, NM, BLOCK, false, false, true>(a, nblocks, stream);
  }
  return full ? launch_one<FL, false, NM, BLOCK, false, true>(a, nblocks, stream)
              : launch_one<FL, false, NM, BLOCK, false, false>(a, nblocks, stream);
}

}  // namespace k1
}  // namespace sip
