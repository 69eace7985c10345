// throughput policy, flag policy "default", 128-member blocks (see sip_run_thr.inc); compiled with -fmad=true
#define SIP_FL StaticFlags<kMaskDefault>
#define SIP_BLOCK 128
#define SIP_NAME launch_thr_default_128
#include "sip_run_thr.inc"
