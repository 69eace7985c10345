// sip_reduce.cu -- on-device ensemble summaries (K3) and the FP64 roofline probe.
//
// Summaries work on the column buffer K1 wrote for the current run range:
//   cols[t * ld + m]   (one output column, n steps, ld-padded member rows)
// and reduce every (site, step) row over that site's members.  The reference
// has no ensemble code at all (SURVEY 8c "unpinned"): the definitions here are
//   mean      = sum(x) / N                      (finite members only)
//   variance  = sum((x - mean)^2) / N           (population variance, two pass)
//   quantile  = numpy's default "linear" rule:  pos = p (N-1), x[lo] + (x[hi]-x[lo]) (pos-lo)
// and tests/ check them against a trivially-written numpy loop over the
// oracle's per-step output.  All reductions use a fixed order (strided
// per-thread partials, then a fixed shared-memory tree), so results are
// bit-reproducible run to run.
#include <cuda_runtime.h>

#include <cstdint>

#include "sip_libm.cuh"
#include "sip_num.cuh"
#include "sip_types.cuh"

namespace sip {

// ---- shared helpers -----------------------------------------------------------------------------
constexpr int kSelThreads = 512;           // block size of the per-row kernel
constexpr int kSelBins = 2048;             // 11 key bits per histogram level
constexpr int kSelMaxQ = 4;                // quantiles per launch (the launcher loops over groups)
constexpr int kSelMaxRanks = 2 * kSelMaxQ; // each quantile needs the order statistics lo and lo+1
constexpr int kCandCap = 8192;             // candidate keys finished in shared memory (64 KB, shares the histograms' storage)
constexpr int kSelDynBytes = kSelMaxRanks * kSelBins * 4;
static_assert(kSelDynBytes == kCandCap * 8, "histograms and candidate buffer share one allocation");
constexpr int kWarpRowMax = 256;           // rows up to this long take the warp-per-row moments kernel

// fixed-order block sum (strided per-thread partials were accumulated by the caller)
__device__ __forceinline__ double block_sum(double v, double *sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = kSelThreads / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// fixed-order block min / max (inputs may be +-inf, never NaN)
__device__ __forceinline__ double block_minmax(double v, double *sh, bool wantMax) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = kSelThreads / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] = wantMax ? fmax(sh[tid], sh[tid + s]) : fmin(sh[tid], sh[tid + s]);
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// order-preserving map double -> uint64 (ascending)
__device__ __forceinline__ uint64_t key_of(double x) {
  const uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double val_of(uint64_t k) {
  const uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

// histogram levels over the 64-bit key: 11,11,11,11,11,9 bits from the top
__device__ __forceinline__ int level_shift(int level) { return level < 5 ? 53 - 11 * level : 0; }
__device__ __forceinline__ int level_bins(int level) { return level < 5 ? kSelBins : 512; }

// Histogram update for all 32 lanes (code < 0: nothing to count).  Ensemble values crowd into a few bins of the
// leading levels, where plain shared-memory atomics would serialise 32-fold: the lanes that agree with lane 0
// are counted by one ballot and added once, the others add individually.
__device__ __forceinline__ void hist_add(unsigned int *hist, int code) {
  const int lead = __shfl_sync(0xffffffffu, code, 0);
  const unsigned same = __ballot_sync(0xffffffffu, code == lead);
  if ((threadIdx.x & 31) == 0) {
    if (lead >= 0) atomicAdd(&hist[lead], (unsigned)__popc(same));
  } else if (code >= 0 && code != lead) {
    atomicAdd(&hist[code], 1u);
  }
}

constexpr int kSelUnroll = 4;  // independent row loads in flight per thread

// kernel parameters passed by value, so a caller needs no device scratch (and no synchronising allocation)
struct RowSpan {
  int member0, memberCount;
};
struct QuantileProbs {
  double p[kSelMaxQ];
};

struct SelState {
  uint64_t prefix[kSelMaxRanks];       // resolved leading key bits of each wanted order statistic
  unsigned long long k[kSelMaxRanks];  // its rank among the keys that share the prefix
  unsigned int pop[kSelMaxRanks];      // how many keys share the prefix
  int slotOf[kSelMaxRanks];            // ranks with equal prefixes share a histogram slot
  uint64_t slotTop[kSelMaxRanks];      // prefix >> shift of the last resolved level
  unsigned int slotTop32[kSelMaxRanks]; // the same while it still fits the key's high word (levels 0 and 1)
  unsigned int slotBase[kSelMaxRanks + 1];
  unsigned int slotFill[kSelMaxRanks];
  double frac[kSelMaxQ];
  int nslots;
};

// warp w resolves rank w on the histogram of its slot: find the bin holding the k-th key
__device__ __forceinline__ void resolve_level(SelState &S, const unsigned int *hist, int nr, int level) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < nr) {
    const unsigned int *h = hist + S.slotOf[warp] * kSelBins;
    const int per = level_bins(level) / 32;
    unsigned long long mine = 0;
    for (int b = 0; b < per; ++b) mine += h[lane * per + b];
    unsigned long long incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const unsigned long long k = S.k[warp];
    const unsigned owner = __ballot_sync(0xffffffffu, incl > k);
    const int who = __ffs(owner) - 1;  // first lane whose inclusive count exceeds k (exists: k < population)
    if (lane == who) {
      unsigned long long acc = incl - mine;
      int bin = lane * per;
      for (;; ++bin) {
        const unsigned long long c = h[bin];
        if (acc + c > k) break;
        acc += c;
      }
      S.prefix[warp] |= (uint64_t)bin << level_shift(level);
      S.k[warp] = k - acc;
      S.pop[warp] = h[bin];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // ranks ascend, so equal prefixes are adjacent
    int ns = 0;
    unsigned int base = 0;
    for (int r = 0; r < nr; ++r) {
      if (r > 0 && S.prefix[r] == S.prefix[r - 1]) {
        S.slotOf[r] = ns - 1;
        continue;
      }
      S.slotOf[r] = ns;
      S.slotTop[ns] = S.prefix[r] >> level_shift(level);
      S.slotTop32[ns] = (unsigned int)S.slotTop[ns];
      S.slotBase[ns] = base;
      base += S.pop[r];
      ++ns;
    }
    S.slotBase[ns] = base;
    S.nslots = ns;
  }
  __syncthreads();
}

// The first two levels (22 key bits) live in the high word of the double: the passes that only need those
// (level-0/1 histograms, compaction after level 1 -- i.e. all three passes of a typical row) classify an element
// with 32-bit integer work.  HI32 = false is the general 64-bit path for deeper levels.
__device__ __forceinline__ unsigned int key_hi_of(double x) {
  const unsigned int hi = (unsigned int)__double2hiint(x);
  return hi ^ ((unsigned int)((int)hi >> 31) | 0x80000000u);
}
__device__ __forceinline__ bool finite_hi(double x) { return (((unsigned int)__double2hiint(x) >> 20) & 0x7ffu) != 0x7ffu; }

template <bool HI32>
__device__ __forceinline__ int slot_of(const SelState &S, double x, int level, int nslots) {
  if (level == 0) return 0;
  int slot = -1;
  if (HI32) {
    const unsigned int top = key_hi_of(x) >> (level_shift(level - 1) - 32);
    for (int s = 0; s < nslots; ++s)
      if (S.slotTop32[s] == top) slot = s;
  } else {
    const uint64_t top = key_of(x) >> level_shift(level - 1);
    for (int s = 0; s < nslots; ++s)
      if (S.slotTop[s] == top) slot = s;
  }
  return slot;
}
template <bool HI32>
__device__ __forceinline__ int bin_of(double x, int shift, int bmask) {
  return HI32 ? (int)((key_hi_of(x) >> (shift - 32)) & (unsigned int)bmask) : (int)((key_of(x) >> shift) & (uint64_t)bmask);
}

// one histogram pass at `level` over the row (optionally accumulating the squared deviations on the way)
template <bool HI32>
__device__ __forceinline__ void hist_pass(const double *row, int count, const SelState &S, unsigned int *hist, int level,
                                          int nslots, bool needSS, double mu, double &q2) {
  const int tid = threadIdx.x;
  const int shift = level_shift(level), bmask = level_bins(level) - 1;
  for (int i0 = 0; i0 < count; i0 += kSelUnroll * kSelThreads) {
    double x[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int i = i0 + u * kSelThreads + tid;
      x[u] = i < count ? row[i] : nan("");
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      int code = -1;
      if (finite_hi(x[u])) {
        if (needSS) {
          const double d = x[u] - mu;
          q2 += d * d;
        }
        const int slot = slot_of<HI32>(S, x[u], level, nslots);
        if (slot >= 0) code = slot * kSelBins + bin_of<HI32>(x[u], shift, bmask);
      }
      hist_add(hist, code);
    }
  }
}

// keys sharing a wanted prefix (`level` levels resolved) -> shared memory, grouped by slot
template <bool HI32>
__device__ __forceinline__ void compact_pass(const double *row, int count, SelState &S, uint64_t *cand, int level,
                                             int nslots, bool needSS, double mu, double &q2) {
  const int tid = threadIdx.x, lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < count; i0 += kSelUnroll * kSelThreads) {
    double x[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int i = i0 + u * kSelThreads + tid;
      x[u] = i < count ? row[i] : nan("");
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      int slot = -1;
      if (finite_hi(x[u])) {
        if (needSS) {
          const double d = x[u] - mu;
          q2 += d * d;
        }
        slot = slot_of<HI32>(S, x[u], level, nslots);
      }
      const unsigned any = __ballot_sync(0xffffffffu, slot >= 0);
      if (any) {  // (rare on long rows) lanes of the first candidate's slot reserve their places with one atomic
        const int lead = __shfl_sync(0xffffffffu, slot, __ffs(any) - 1);
        const unsigned same = __ballot_sync(0xffffffffu, slot == lead);
        const int leader = __ffs(same) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&S.slotFill[lead], (unsigned)__popc(same));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (slot == lead) {
          cand[S.slotBase[slot] + base + __popc(same & ((1u << lane) - 1u))] = key_of(x[u]);
        } else if (slot >= 0) {
          cand[S.slotBase[slot] + atomicAdd(&S.slotFill[slot], 1u)] = key_of(x[u]);
        }
      }
    }
  }
}

// ---- per-row summary: moments and exact quantiles in (typically) three reads of the row ----------
// grid = (nsteps, nsites); block = kSelThreads; dynamic shared = kSelDynBytes.
//   mean/var[site * momStride + t]              (either both or neither)
//   quant[site * qStride + q * nsteps + t]      (nq <= kSelMaxQ)
// Pass 1: finite count + sum (+ level-0 histogram when the row is longer than kCandCap).
// Pass 2: sum of squared deviations (+ level-1 histogram).  Further histogram passes only while the
// wanted prefixes still hold more than kCandCap keys (heavily duplicated data).  Last pass: the few keys
// sharing a wanted prefix are compacted to shared memory, sorted (bitonic) and indexed.
__global__ void __launch_bounds__(kSelThreads, 2)
    row_summary_kernel(const double *cols, int64_t ld, int64_t nsteps, const SiteDev *sites, const RowSpan oneRow,
                       const QuantileProbs probs, int nq,
                       double *mean, double *var, int64_t momStride, double *quant, int64_t qStride) {
  extern __shared__ __align__(16) unsigned char dyn[];
  unsigned int *hist = reinterpret_cast<unsigned int *>(dyn);
  uint64_t *cand = reinterpret_cast<uint64_t *>(dyn);
  __shared__ double red[kSelThreads];
  __shared__ SelState S;
  const int tid = threadIdx.x;
  const int64_t t = blockIdx.x;
  const int site = blockIdx.y;
  // the members of the row: a site of the handle, or (sites == nullptr) the one span passed by value
  const int member0 = sites != nullptr ? sites[site].member0 : oneRow.member0;
  const int count = sites != nullptr ? sites[site].memberCount : oneRow.memberCount;
  const double *row = cols + t * ld + member0;
  const int nr = 2 * nq;
  const bool big = nq > 0 && count > kCandCap;

  // ---- pass 1
  if (big) {
    for (int i = tid; i < kSelBins; i += kSelThreads) hist[i] = 0;
    __syncthreads();
  }
  double s = 0.0, c = 0.0;
  double vmin = __longlong_as_double(0x7ff0000000000000ll), vmax = -vmin;
  for (int i0 = 0; i0 < count; i0 += kSelUnroll * kSelThreads) {
    double x[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int i = i0 + u * kSelThreads + tid;
      x[u] = i < count ? row[i] : nan("");
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {  // per thread the elements are still accumulated in index order
      const bool fin = finite_hi(x[u]);
      if (fin) {
        s += x[u];
        c += 1.0;
        if (big) {
          vmin = fmin(vmin, x[u]);
          vmax = fmax(vmax, x[u]);
        }
      }
      if (big) hist_add(hist, fin ? (int)(key_hi_of(x[u]) >> 21) : -1);
    }
  }
  const double total = block_sum(s, red);
  const double n = block_sum(c, red);
  const double mu = n > 0 ? total / n : nan("");
  bool needSS = mean != nullptr && n > 0;
  double q2 = 0.0;
  // a constant row (GPP of a night step is exactly 0 for every member) would walk all six levels: settle it here
  bool constantRow = false;
  if (big && n > 0) {
    vmin = block_minmax(vmin, red, false);
    vmax = block_minmax(vmax, red, true);
    constantRow = vmin == vmax;
  }

  if (constantRow) {
    if (tid < nq) quant[(int64_t)site * qStride + (int64_t)tid * nsteps + t] = vmin;
  } else if (nq > 0 && n > 0) {
    if (tid < nq) {
      const double pos = probs.p[tid] * (n - 1.0);
      const double lo = floor(pos);
      const double fr = pos - lo;
      S.frac[tid] = fr;
      S.k[2 * tid] = (unsigned long long)lo;
      S.k[2 * tid + 1] = (unsigned long long)(fr > 0.0 ? lo + 1.0 : lo);
    }
    if (tid < nr) {
      S.prefix[tid] = 0;
      S.slotOf[tid] = 0;
    }
    if (tid == 0) {
      S.nslots = 1;
      S.slotBase[0] = 0;
      S.slotBase[1] = (unsigned int)n;
    }
    __syncthreads();
    int level = 0;
    if (big) {
      resolve_level(S, hist, nr, 0);
      level = 1;
      while (S.slotBase[S.nslots] > (unsigned)kCandCap && level < 6) {
        const int nslots = S.nslots;
        for (int i = tid; i < nslots * kSelBins; i += kSelThreads) hist[i] = 0;
        __syncthreads();
        if (level_shift(level) >= 32)
          hist_pass<true>(row, count, S, hist, level, nslots, needSS, mu, q2);
        else
          hist_pass<false>(row, count, S, hist, level, nslots, needSS, mu, q2);
        needSS = false;  // accumulated (reduced below)
        __syncthreads();
        resolve_level(S, hist, nr, level);
        ++level;
      }
    }
    if (level < 6) {
      // ---- compaction pass: keys sharing a wanted prefix -> shared memory
      const int nslots = S.nslots;
      const unsigned int ctot = S.slotBase[nslots];
      if (tid < nslots) S.slotFill[tid] = 0;
      __syncthreads();
      if (level == 0 || level_shift(level - 1) >= 32)
        compact_pass<true>(row, count, S, cand, level, nslots, needSS, mu, q2);
      else
        compact_pass<false>(row, count, S, cand, level, nslots, needSS, mu, q2);
      needSS = false;
      unsigned int n2 = 2;
      while (n2 < ctot) n2 <<= 1;
      __syncthreads();
      for (unsigned int i = ctot + tid; i < n2; i += kSelThreads) cand[i] = ~0ull;
      __syncthreads();
      // slots hold ascending key ranges in ascending order, so one sort of the whole buffer sorts each slot in place
      for (unsigned int kk = 2; kk <= n2; kk <<= 1) {
        for (unsigned int j = kk >> 1; j > 0; j >>= 1) {
          for (unsigned int i = tid; i < n2; i += kSelThreads) {
            const unsigned int ixj = i ^ j;
            if (ixj > i) {
              const uint64_t a = cand[i], b = cand[ixj];
              if ((a > b) == ((i & kk) == 0)) {
                cand[i] = b;
                cand[ixj] = a;
              }
            }
          }
          __syncthreads();
        }
      }
      if (tid < nr) S.prefix[tid] = cand[S.slotBase[S.slotOf[tid]] + (unsigned int)S.k[tid]];
      __syncthreads();
    }
    // S.prefix[r] is now the complete key of order statistic r
    if (tid < nq) {
      const double xlo = val_of(S.prefix[2 * tid]), xhi = val_of(S.prefix[2 * tid + 1]);
      const double fr = S.frac[tid];
      // numpy's _lerp: a + (b - a) t, evaluated from the b side when t >= 0.5
      quant[(int64_t)site * qStride + (int64_t)tid * nsteps + t] =
          (fr >= 0.5) ? xhi - (xhi - xlo) * (1.0 - fr) : xlo + (xhi - xlo) * fr;
    }
  } else if (nq > 0 && tid < nq) {
    quant[(int64_t)site * qStride + (int64_t)tid * nsteps + t] = nan("");
  }

  if (mean != nullptr) {
    if (needSS) {  // no later pass carried the squared deviations (moments only)
      for (int i = tid; i < count; i += kSelThreads) {
        const double x = row[i];
        if (isfinite(x)) {
          const double d = x - mu;
          q2 += d * d;
        }
      }
    }
    const double ss = block_sum(q2, red);
    if (tid == 0) {
      mean[(int64_t)site * momStride + t] = mu;
      var[(int64_t)site * momStride + t] = n > 0 ? ss / n : nan("");
    }
  }
}

// ---- moments of short rows: one warp per (site, step) row, values held in registers ---------------
// C3-shaped runs have 10^4 sites x 100 members: a block per row would idle most of its threads.
// grid.x covers sites in groups of (blockDim / 32) -- neighbouring warps read neighbouring members --, grid.y = steps
constexpr int kWarpRowsPerBlock = 8;
__global__ void __launch_bounds__(kWarpRowsPerBlock * 32)
    moments_warp_kernel(const double *cols, int64_t ld, int64_t nsteps, const SiteDev *sites, int64_t nsites, double *mean,
                        double *var, int64_t momStride) {
  const int lane = threadIdx.x & 31;
  const int64_t site = (int64_t)blockIdx.x * kWarpRowsPerBlock + (threadIdx.x >> 5);
  const int64_t t = blockIdx.y;
  if (site >= nsites) return;
  const SiteDev sd = sites[site];
  const double *row = cols + t * ld + sd.member0;
  constexpr int kPer = kWarpRowMax / 32;
  double x[kPer];
  double s = 0.0, c = 0.0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = j * 32 + lane;
    x[j] = i < sd.memberCount ? row[i] : nan("");
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j)
    if (isfinite(x[j])) {
      s += x[j];
      c += 1.0;
    }
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, d);
    c += __shfl_xor_sync(0xffffffffu, c, d);
  }
  const double mu = c > 0 ? s / c : nan("");
  double q = 0.0;
#pragma unroll
  for (int j = 0; j < kPer; ++j)
    if (isfinite(x[j])) {
      const double d = x[j] - mu;
      q += d * d;
    }
  for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  if (lane == 0) {
    mean[site * momStride + t] = mu;
    var[site * momStride + t] = c > 0 ? q / c : nan("");
  }
}

// One column's summaries for every (site, step) row.  mean/var may be null (quantiles only), nq may be 0.
// `sites` = the handle's device site table, or null: then every row is columns [0, maxMembers) (nsites must be 1).
// `probs` is a HOST array.
cudaError_t launch_row_summary(const double *cols, int64_t ld, int64_t nsteps, const SiteDev *sites, int64_t nsites,
                               int64_t maxMembers, const double *probs, int nq, double *mean, double *var,
                               int64_t momStride, double *quant, int64_t qStride, cudaStream_t stream) {
  {
    cudaError_t e = cudaFuncSetAttribute(row_summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelDynBytes);
    if (e != cudaSuccess) return e;
  }
  const RowSpan one{0, (int)maxMembers};
  if (sites != nullptr && nq == 0 && mean != nullptr && maxMembers <= kWarpRowMax) {
    // gridDim.y <= 65535: tile steps
    for (int64_t t0 = 0; t0 < nsteps; t0 += 65535) {
      const int64_t nt = (nsteps - t0) < 65535 ? (nsteps - t0) : 65535;
      dim3 grid((unsigned)((nsites + kWarpRowsPerBlock - 1) / kWarpRowsPerBlock), (unsigned)nt);
      moments_warp_kernel<<<grid, kWarpRowsPerBlock * 32, 0, stream>>>(cols + t0 * ld, ld, nsteps, sites, nsites,
                                                                       mean + t0, var + t0, momStride);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  for (int q0 = 0; q0 < (nq > 0 ? nq : 1); q0 += kSelMaxQ) {
    const int nqq = nq - q0 < kSelMaxQ ? nq - q0 : kSelMaxQ;
    const bool first = q0 == 0;
    QuantileProbs qp{};
    for (int i = 0; i < nqq && nq > 0; ++i) qp.p[i] = probs[q0 + i];
    for (int64_t s0 = 0; s0 < nsites; s0 += 65535) {  // gridDim.y <= 65535: tile sites
      const int ns = (int)((nsites - s0) < 65535 ? (nsites - s0) : 65535);
      dim3 grid((unsigned)nsteps, (unsigned)ns);
      row_summary_kernel<<<grid, kSelThreads, kSelDynBytes, stream>>>(
          cols, ld, nsteps, sites != nullptr ? sites + s0 : nullptr, one, qp, nq > 0 ? nqq : 0,
          first && mean ? mean + s0 * momStride : nullptr, first && var ? var + s0 * momStride : nullptr, momStride,
          quant ? quant + s0 * qStride + (int64_t)q0 * nsteps : nullptr, qStride);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// ---- FP64 issue-rate probe -------------------------------------------------------------------
// Eight independent register-resident DFMA chains per thread; 2 flops per DFMA.
__global__ void __launch_bounds__(256) fp64_probe_kernel(double *sink, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b);
    x1 = fma(x1, a, b);
    x2 = fma(x2, a, b);
    x3 = fma(x3, a, b);
    x4 = fma(x4, a, b);
    x5 = fma(x5, a, b);
    x6 = fma(x6, a, b);
    x7 = fma(x7, a, b);
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.6789) sink[0] = s;  // never true; keeps the chains alive
}

cudaError_t measure_fp64_peak(int device, double *tflops) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return e;
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
  double *sink = nullptr;
  if ((e = cudaMalloc(&sink, 8)) != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sm * 8, threads = 256, iters = 1 << 16;
  fp64_probe_kernel<<<blocks, threads>>>(sink, 1 << 12, 0.999999, 1e-9);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fp64_probe_kernel<<<blocks, threads>>>(sink, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return e;
}

// ---- validation hook: device exp / pow / division on arrays ---------------------------------
// ops 0, 1: the general restatement (libm::exp, libm::pow).  ops 2..5: the optimistic policy of the production
// kernel (FastNum: exp, pow with a varying base, pow with a cached log of the base, a / b); an input outside its
// guards yields kFlagged instead of a value -- "FastNum never has to be right outside its guards, only to notice".
constexpr unsigned long long kFlagged = 0x7ff8bad0bad0bad0ull;
__global__ void eval_libm_kernel(int op, const double *x, const double *y, double *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op < 2) {
    out[i] = op == 0 ? libm::exp(x[i]) : libm::pow(x[i], y[i]);
    return;
  }
  FastNum nm;
  nm.expTab = libm::d_exp_tab;
  nm.powlogTab = libm::d_powlog_tab;
  double v;
  if (op == 2) {
    v = nm.exp(x[i]);
  } else if (op == 3) {
    v = nm.pow(x[i], y[i]);
  } else if (op == 4) {
    const libm::LogHL l = libm::pow_log(x[i]);
    v = nm.powc(x[i], l.hi, l.lo, y[i]);
  } else {
    v = nm.div(x[i], y[i]);
  }
  out[i] = nm.bad ? __longlong_as_double((long long)kFlagged) : v;
}

cudaError_t eval_libm(int device, int op, const double *x, const double *y, double *out, int64_t n) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return e;
  double *dx = nullptr, *dy = nullptr, *dout = nullptr;
  const size_t bytes = (size_t)n * sizeof(double);
  if ((e = cudaMalloc(&dx, bytes)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&dy, bytes)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&dout, bytes)) != cudaSuccess) return e;
  cudaMemcpy(dx, x, bytes, cudaMemcpyHostToDevice);
  if (y != nullptr) cudaMemcpy(dy, y, bytes, cudaMemcpyHostToDevice);
  eval_libm_kernel<<<(unsigned)((n + 255) / 256), 256>>>(op, dx, dy, dout, n);
  e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dy);
  cudaFree(dout);
  return e;
}

}  // namespace sip
