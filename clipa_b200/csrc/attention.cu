// placeholder replaced below in this round: attention kernels
#include "host_common.h"
extern "C" int clipa_attention_fwd(const void*, void*, float*, int32_t, int32_t, int32_t, int32_t, int32_t, void*) {
  clipa::set_error("attention_fwd: not built yet");
  return CLIPA_ERR_UNSUPPORTED;
}
extern "C" int clipa_attention_bwd(const void*, const void*, const void*, const float*, void*, int32_t, int32_t, int32_t, int32_t, int32_t, void*) {
  clipa::set_error("attention_bwd: not built yet");
  return CLIPA_ERR_UNSUPPORTED;
}
