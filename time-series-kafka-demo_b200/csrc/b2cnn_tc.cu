// b2cnn_tc.cu -- tcgen05 fast path (placeholder until the kernel lands: reports "unsupported"
// so every call takes the exact generic path; never a CPU fallback).
#include "b2cnn_tc.cuh"

namespace b2cnn {
static thread_local const char *g_tc_err = "";
const char *tc_error() { return g_tc_err; }
int tc_prepare(TcState &s, const Dims &, const ConvWeights &, const float *, const HeadWeights &, int splits, int, cudaStream_t) {
    s.splits = splits;
    s.ready = false;
    return 0;
}
void tc_release(TcState &s) { s.ready = false; }
bool tc_supported(const TcState &s, const Dims &, int, int64_t, int) { return s.ready; }
bool tc_can_emit_features(const TcState &) { return false; }
int64_t tc_workspace_bytes(const TcState &, const Dims &, int64_t) { return 0; }
int tc_forward(TcState &, const Dims &, const ConvWeights &, const HeadWeights &, const void *, int64_t, const float *, int64_t,
               int, float *, float *, float *, float *, void *, int, cudaStream_t, const char **err) {
    *err = "tensor-core path not built";
    return -1;
}
int tc_features(TcState &, const Dims &, const void *, int64_t, float *, int, cudaStream_t, const char **err) {
    *err = "tensor-core path not built";
    return -1;
}
}  // namespace b2cnn
