// optimistic kernel, flag policy "cropn", 128-member blocks (see sip_run_fast.inc)
#define SIP_FL StaticFlags<kMaskCropN>
#define SIP_BLOCK 128
#define SIP_NAME launch_fast_cropn_128
#define SIP_PACK launch_pack_cropn
#include "sip_run_fast.inc"

namespace sip {
namespace k1 {
cudaError_t launch_pack_cropn_3(const RunArgs &a, int nblocks, bool full, cudaStream_t stream);
cudaError_t launch_pack_cropn_4(const RunArgs &a, int nblocks, bool full, cudaStream_t stream);
cudaError_t launch_pack_cropn(const RunArgs &a, int nblocks, bool full, int occ, cudaStream_t stream) {
  return occ == 4 ? launch_pack_cropn_4(a, nblocks, full, stream) : launch_pack_cropn_3(a, nblocks, full, stream);
}
}  // namespace k1
}  // namespace sip
