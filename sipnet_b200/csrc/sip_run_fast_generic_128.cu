// optimistic kernel, flag policy "generic", 128-member blocks (see sip_run_fast.inc)
#define SIP_FL RuntimeFlags
#define SIP_BLOCK 128
#define SIP_NAME launch_fast_generic_128
#include "sip_run_fast.inc"
