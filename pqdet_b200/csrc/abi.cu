// Version / error-string entry points of the C ABI (include/pqdet_b200.h).
#include "pq_common.cuh"

extern "C" int pqdet_version(void) { return PQDET_VERSION; }

extern "C" const char* pqdet_strerror(int code) {
  switch (code) {
    case PQDET_OK: return "ok";
    case PQDET_ERR_INVALID_ARG: return "invalid argument";
    case PQDET_ERR_CUDA: return "CUDA runtime error (launch or device selection failed)";
    case PQDET_ERR_UNSUPPORTED: return "unsupported configuration";
    case PQDET_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}
