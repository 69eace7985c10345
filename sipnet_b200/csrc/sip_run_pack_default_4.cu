// throughput variant, flag policy "default", 4 resident 128-member blocks per SM (see sip_run_pack.inc)
#define SIP_FL StaticFlags<kMaskDefault>
#define SIP_OCC 4
#define SIP_NAME launch_pack_default_4
#include "sip_run_pack.inc"
