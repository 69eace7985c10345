// throughput policy, flag policy "cropn", 32-member blocks (see sip_run_thr.inc); compiled with -fmad=true
#define SIP_FL StaticFlags<kMaskCropN>
#define SIP_BLOCK 32
#define SIP_NAME launch_thr_cropn_32
#include "sip_run_thr.inc"
