// optimistic kernel, flag policy "cropn", 128-member blocks (see sip_run_fast.inc)
#define SIP_FL StaticFlags<kMaskCropN>
#define SIP_BLOCK 128
#define SIP_NAME launch_fast_cropn_128
#include "sip_run_fast.inc"
