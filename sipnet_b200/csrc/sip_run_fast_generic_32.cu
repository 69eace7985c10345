// optimistic kernel, flag policy "generic", 32-member blocks (see sip_run_fast.inc)
#define SIP_FL RuntimeFlags
#define SIP_BLOCK 32
#define SIP_NAME launch_fast_generic_32
#include "sip_run_fast.inc"
