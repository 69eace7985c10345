// optimistic kernel, flag policy "default", 128-member blocks (see sip_run_fast.inc)
#define SIP_FL StaticFlags<kMaskDefault>
#define SIP_BLOCK 128
#define SIP_NAME launch_fast_default_128
#include "sip_run_fast.inc"
