// ABI version + error strings of libpropainter_b200.so (see include/propainter_b200.h).
#include "pp_common.cuh"
#include "../../include/propainter_b200.h"

extern "C" int pp_abi_version(void) { return PP_ABI_VERSION; }

extern "C" const char* pp_error_string(int code) {
  switch (code) {
    case PP_OK: return "ok";
    case PP_ERR_SHAPE: return "bad shape";
    case PP_ERR_DTYPE: return "unsupported dtype";
    case PP_ERR_WORKSPACE: return "workspace too small";
    case PP_ERR_LAUNCH: return "kernel launch failed";
    case PP_ERR_ALIGN: return "misaligned pointer or leading dimension";
    default: return "unknown error";
  }
}
