// Fish S1-DAC decode path on B200 (placeholder while the DiT path is brought up; replaced below in this round).
#include "handle.h"

using namespace echo;

namespace echo {
int dac_set_weight(echo_handle*, const char* key, const void*, const int64_t*, int, int, cudaStream_t) {
  set_error("dac weights not supported yet ('%s')", key);
  return ECHO_ERR_STATE;
}
}  // namespace echo

extern "C" int echo_dac_configure(echo_handle*, const echo_dac_config*) { set_error("dac: not built yet"); return ECHO_ERR_STATE; }
extern "C" int echo_dac_finalize(echo_handle*, void*) { set_error("dac: not built yet"); return ECHO_ERR_STATE; }
extern "C" int echo_dac_decode(echo_handle*, const float*, const float*, const float*, float, int, int, float*, void*) {
  set_error("dac: not built yet"); return ECHO_ERR_STATE;
}
extern "C" int echo_dac_decode_zq(echo_handle*, const float*, int, int, float*, void*) {
  set_error("dac: not built yet"); return ECHO_ERR_STATE;
}
