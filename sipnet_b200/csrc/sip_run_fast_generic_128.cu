// optimistic kernel, flag policy "generic", 128-member blocks (see sip_run_fast.inc)
#define SIP_FL RuntimeFlags
#define SIP_BLOCK 128
#define SIP_NAME launch_fast_generic_128
#define SIP_PACK launch_pack_generic
#include "sip_run_fast.inc"

namespace sip {
namespace k1 {
cudaError_t launch_pack_generic_3(const RunArgs &a, int nblocks, bool full, cudaStream_t stream);
cudaError_t launch_pack_generic_4(const RunArgs &a, int nblocks, bool full, cudaStream_t stream);
cudaError_t launch_pack_generic(const RunArgs &a, int nblocks, bool full, int occ, cudaStream_t stream) {
  return occ == 4 ? launch_pack_generic_4(a, nblocks, full, stream) : launch_pack_generic_3(a, nblocks, full, stream);
}
}  // namespace k1
}  // namespace sip
