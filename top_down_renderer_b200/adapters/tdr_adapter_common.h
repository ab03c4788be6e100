// tdr_adapter_common.h — shared by the adapter bodies: the process-wide device context (one, like the single ROS spinner
// thread that owns every call, SURVEY 8b "Threading") and the "log and carry on" error convention of the reference's
// void methods.
#pragma once
#include <cstdio>
#include <cstdlib>

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
inline bool ok(int rc) {
  if (rc != TDR_OK) fprintf(stderr, "[XView] libtdr_b200: %s\n", tdr_last_error());
  return rc == TDR_OK;
}
}  // namespace tdr_adapter
