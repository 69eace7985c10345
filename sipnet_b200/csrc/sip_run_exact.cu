// sip_run_exact.cu -- instantiations of K1 with the general numerics (ExactNum, runtime flags): validation run,
// debug dump, replay of members the optimistic kernel flagged.  See sip_run.cuh.
#include "sip_run.cuh"

namespace sip {
namespace k1 {

// mode: 0 = exact (validation), 2 = replay of flagged members; debug => the validation dump
template <int BLOCK>
static cudaError_t launch_exact_block(const RunArgs &a, int nblocks, bool debug, int mode, cudaStream_t stream) {
  constexpr bool MIX = BLOCK == 128;  // the general kernels always take the two-site path (speed is not their job)
  if (mode == 2) return launch_one<RuntimeFlags, false, ExactNum, BLOCK, true, false, MIX>(a, nblocks, stream);
  if (debug) return launch_one<RuntimeFlags, true, ExactNum, BLOCK, false, false, MIX>(a, nblocks, stream);
  return launch_one<RuntimeFlags, false, ExactNum, BLOCK, false, false, MIX>(a, nblocks, stream);
}

cudaError_t launch_exact(const RunArgs &a, int nblocks, int blockThreads, bool debug, int mode, cudaStream_t stream) {
  switch (blockThreads) {
    case 32: return launch_exact_block<32>(a, nblocks, debug, mode, stream);
    case 128: return launch_exact_block<128>(a, nblocks, debug, mode, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace k1
}  // namespace sip
