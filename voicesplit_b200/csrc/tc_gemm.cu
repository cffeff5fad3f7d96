// LSTM input projection, recurrence and FC head of the tensor-core precision modes.
// v1: cnn8 writes the fp32 LSTM input and the contractions after the conv stack still run on the
// fp32 CUDA-core kernels (fp32_kernels.cu); the tcgen05 GEMM replaces them next.
#include "tc.cuh"

namespace vs {

int tc_gemm_pack(vs_engine*, cudaStream_t) { return VS_OK; }
size_t tc_gemm_workspace_bytes(const vs_engine*, int, int, int) { return 1024; }

int tc_lstm_head(vs_engine* e, const elt16* plane_hi, const elt16* plane_lo, const float* conv_out32,
                 const float* emb, const float* x, float* mask, float* masked, int B, int T, int precision, float* xcat32,
                 float* fc1, void*, const TcLstmBuffers& lb, cudaStream_t st) {
    const int F = e->d.num_freq, H = e->d.lstm_dim, E = e->d.emb_dim, N1 = e->d.fc1_dim;
    const int M = B * T;
    const float* xin = conv_out32;
    if (!xin) {
        VS_LAUNCH(e, KID_POINT8, st, tc_launch_point8(e, plane_hi, plane_lo, tc_elt(precision), xcat32, nullptr, nullptr, 8 * F, B, T, st));
        xin = xcat32;
    }
    VS_LAUNCH(e, KID_EMB_BIAS, st, launch_gemm_fp32(emb, E, e->wih_e, E, e->b_lstm, nullptr, 1, lb.bias_u, 8 * H, B, 8 * H, E,
                                                    false, EPI_NONE, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_INPROJ, st, launch_gemm_fp32(xin, 8 * F, e->wih_x, 8 * F, nullptr, lb.bias_u, T, lb.gates, 8 * H, M, 8 * H, 8 * F,
                                                  false, EPI_NONE, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_LSTM_REC, st, launch_lstm_rec_fp32(e, lb.gates, lb.hout, lb.hx, lb.barrier, B, T, st));
    VS_LAUNCH(e, KID_FC1, st, launch_gemm_fp32(lb.hout, 2 * H, e->fc1_w, 2 * H, e->fc1_b, nullptr, 1, fc1, N1, M, N1, 2 * H,
                                               true, EPI_RELU, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_FC2, st, launch_gemm_fp32(fc1, N1, e->fc2_w, N1, e->fc2_b, nullptr, 1, mask, F, M, F, N1,
                                               false, EPI_SIGMOID_MASK, x, masked, st));
    return VS_OK;
}

int tc_debug_lstm_head(vs_engine* e, const float* conv_out, const float* emb, const float* x, float* mask, int B, int T,
                       int precision, const TcLstmBuffers& lb, cudaStream_t st) {
    float* fc1 = nullptr;
    VS_CUDA_TRY(cudaMalloc(&fc1, (size_t)B * T * e->d.fc1_dim * sizeof(float)));
    int rc = tc_lstm_head(e, nullptr, nullptr, conv_out, emb, x, mask, nullptr, B, T, precision, nullptr, fc1, nullptr, lb, st);
    cudaStreamSynchronize(st);
    cudaFree(fc1);
    return rc;
}

}  // namespace vs
