// optimistic kernel, flag policy "cropn", 32-member blocks (see sip_run_fast.inc)
#define SIP_FL StaticFlags<kMaskCropN>
#define SIP_BLOCK 32
#define SIP_NAME launch_fast_cropn_32
#include "sip_run_fast.inc"
