// capi.cu -- version / error text of the C ABI (include/odwscl.h).
#include "common.cuh"

ODW_API int odwscl_version(void) { return ODWSCL_VERSION; }

ODW_API const char* odwscl_strerror(int code) {
  if (code == 0) return "ok";
  if (code == ODWSCL_EINVAL) return "odwscl: invalid argument";
  if (code == ODWSCL_ENOWS) return "odwscl: workspace too small";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "odwscl: unknown error";
}

static int g_sm_margin = 0;
int odw_sm_margin() { return g_sm_margin; }

ODW_API int odwscl_set_sm_margin(int sms) {
  if (sms < 0 || sms > ODW_NUM_SMS - 16) return ODWSCL_EINVAL;
  g_sm_margin = sms & ~1;                       // whole CTA pairs
  return 0;
}
