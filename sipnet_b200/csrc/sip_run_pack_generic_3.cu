// throughput variant, flag policy "generic", 3 resident 128-member blocks per SM (see sip_run_pack.inc)
#define SIP_FL RuntimeFlags
#define SIP_OCC 3
#define SIP_NAME launch_pack_generic_3
#include "sip_run_pack.inc"
