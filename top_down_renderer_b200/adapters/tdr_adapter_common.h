// tdr_adapter_common.h — shared by the adapter bodies: the process-wide device context (one, like the single ROS spinner
// thread that owns every call, SURVEY 8b "Threading") and the "log and carry on" error convention of the reference's
// void methods.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "tdr.h"

namespace tdr_adapter {
inline tdr_ctx* tdr() {
  static tdr_ctx* ctx = nullptr;
  if (!ctx) {
    const char* dev = getenv("TDR_DEVICE");
    if (tdr_create(&ctx, dev ? atoi(dev) : 0) != TDR_OK) {
      fprintf(stderr, "[XView] libtdr_b200: %s\n", tdr_last_error());
      ctx = nullptr;
    }
  }
  return ctx;
}
// One device context serves every adapter object of the process and is not re-entrant (SURVEY 8b "Threading": all public
// calls arrive on the ROS spinner thread — except ParticleFilter's GMM thread, which reads the particle set once a second).
// Every adapter member that reaches the C ABI holds this (recursive: members call each other).
inline std::recursive_mutex& device_mutex() { static std::recursive_mutex* m = new std::recursive_mutex; return *m; }   // never destroyed: the GMM thread outlives main()
struct Guard {
  std::lock_guard<std::recursive_mutex> g;
  Guard() : g(device_mutex()) {}
};
inline bool ok(int rc) {
  if (rc != TDR_OK) fprintf(stderr, "[XView] libtdr_b200: %s\n", tdr_last_error());
  return rc == TDR_OK;
}
}  // namespace tdr_adapter
