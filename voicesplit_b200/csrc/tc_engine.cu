// Tensor-core path orchestration (placeholder until the tcgen05 kernels land in this file).
#include "tc.cuh"

namespace vs {

int tc_create(vs_engine*) { return VS_OK; }
void tc_destroy(vs_engine*) {}
int tc_pack(vs_engine*, cudaStream_t) { return VS_OK; }
size_t tc_workspace_bytes(const vs_engine*, int, int, int) { return 1024; }
static int nyi() { set_error("tensor-core precision modes are not built yet"); return VS_ERR_UNSUPPORTED; }
int tc_forward(vs_engine*, const float*, const float*, float*, float*, int, int, int, void*, const TcLstmBuffers&, cudaStream_t) { return nyi(); }
int tc_conv_stack(vs_engine*, const float*, float*, int, int, int, void*, cudaStream_t) { return nyi(); }
int tc_debug_layer(vs_engine*, int, const float*, const float*, float*, int, int, int, cudaStream_t) { return nyi(); }
int tc_debug_lstm_head(vs_engine*, const float*, const float*, const float*, float*, int, int, int, const TcLstmBuffers&, cudaStream_t) { return nyi(); }

}  // namespace vs
