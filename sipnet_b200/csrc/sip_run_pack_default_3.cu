// throughput variant, flag policy "default", 3 resident 128-member blocks per SM (see sip_run_pack.inc)
#define SIP_FL StaticFlags<kMaskDefault>
#define SIP_OCC 3
#define SIP_NAME launch_pack_default_3
#include "sip_run_pack.inc"
