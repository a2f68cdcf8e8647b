// Temporary: entry points of the functional-map stages that are not implemented yet report
// DM_ERR_UNSUPPORTED (never a silent fallback).
#include "dm_internal.cuh"
using namespace dm;
extern "C" {
size_t dm_icp_workspace_bytes(int, int64_t, int64_t, int, int, int, int, int) { return 0; }
int dm_icp(const double*, int, int, int, const double*, int64_t, const int64_t*, int64_t, int, const double*, int64_t,
           const int64_t*, int64_t, int, int, double*, void*, int, void*, size_t, dm_stream_t) {
  DM_FAIL(DM_ERR_UNSUPPORTED, "dm_icp not implemented yet");
}
}
