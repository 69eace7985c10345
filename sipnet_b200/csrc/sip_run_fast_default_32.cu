// optimistic kernel, flag policy "default", 32-member blocks (see sip_run_fast.inc)
#define SIP_FL StaticFlags<kMaskDefault>
#define SIP_BLOCK 32
#define SIP_NAME launch_fast_default_32
#include "sip_run_fast.inc"
