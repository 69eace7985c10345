// throughput variant, flag policy "generic", 4 resident 128-member blocks per SM (see sip_run_pack.inc)
#define SIP_FL RuntimeFlags
#define SIP_OCC 4
#define SIP_NAME launch_pack_generic_4
#include "sip_run_pack.inc"
