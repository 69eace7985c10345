// throughput variant, flag policy "cropn", 4 resident 128-member blocks per SM (see sip_run_pack.inc)
#define SIP_FL StaticFlags<kMaskCropN>
#define SIP_OCC 4
#define SIP_NAME launch_pack_cropn_4
#include "sip_run_pack.inc"
