// throughput policy, flag policy "generic", 32-member blocks (see sip_run_thr.inc); compiled with -fmad=true
#define SIP_FL RuntimeFlags
#define SIP_BLOCK 32
#define SIP_NAME launch_thr_generic_32
#include "sip_run_thr.inc"
