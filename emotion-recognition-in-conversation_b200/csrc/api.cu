// Library-level entry points of libercgraph.so.
#include "common.cuh"

namespace ercg {
std::atomic<unsigned long long> g_launches{0};
}

extern "C" const char* ercg_strerror(int code) {
  switch (code) {
    case ERCG_OK: return "ok";
    case ERCG_EINVAL: return "invalid argument";
    case ERCG_EALIGN: return "pointer or leading dimension not 16-byte aligned";
    case ERCG_ERANGE: return "size exceeds the packed int32/uint8 format";
    case ERCG_ECUDA: return "CUDA launch error";
    case ERCG_EWORKSPACE: return "workspace too small";
    case ERCG_P2P_ETIMEOUT: return "a peer did not arrive at a peer-memory collective";
    default: return "unknown error";
  }
}

extern "C" int ercg_version(void) { return 100; }

extern "C" unsigned long long ercg_launch_count(void) { return ercg::g_launches.load(std::memory_order_relaxed); }
