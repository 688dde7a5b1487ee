// Library identity + error strings for libmpcg_b200.so (see include/mpcg_b200.h).
#include "common.cuh"

extern "C" int mpcg_abi_version(void) { return MPCG_ABI_VERSION; }

extern "C" const char* mpcg_error_string(int code) {
  switch (code) {
    case MPCG_OK: return "ok";
    case MPCG_EINVAL: return "invalid argument (null pointer, negative or inconsistent size)";
    case MPCG_ERANGE: return "size outside the range this build supports";
    case MPCG_EUNSUPPORTED: return "option not supported";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown mpcg error";
}
