// throughput policy, flag policy "generic", 128-member blocks (see sip_run_thr.inc); compiled with -fmad=true
#define SIP_FL RuntimeFlags
#define SIP_BLOCK 128
#define SIP_NAME launch_thr_generic_128
#include "sip_run_thr.inc"
