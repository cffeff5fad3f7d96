// C ABI of the engine (include/voicesplit_b200.h): parameter packing, workspace carving and the
// forward orchestration.  Everything here is host code plus a few tiny packing kernels.
#include "common.cuh"
#include "tc.cuh"

#include <mutex>
#include <new>
#include <vector>

namespace vs {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

// ---- per-kernel timing ---------------------------------------------------------------------------
struct Prof {
    std::vector<cudaEvent_t> ev;  // ev[0] = start, ev[i] = after launch i
    std::vector<int> ids;
    int n = 0;
};
static cudaEvent_t prof_event(Prof* p, int i) {
    while ((int)p->ev.size() <= i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        p->ev.push_back(e);
    }
    return p->ev[i];
}
void prof_begin(vs_engine* e, cudaStream_t st) {
    if (!e->profiling) return;
    if (!e->prof) e->prof = new Prof();
    Prof* p = (Prof*)e->prof;
    p->n = 0;
    p->ids.clear();
    cudaEventRecord(prof_event(p, 0), st);
}
void prof_after(vs_engine* e, int id, cudaStream_t st) {
    Prof* p = (Prof*)e->prof;
    if (!p) return;
    p->n++;
    p->ids.push_back(id);
    cudaEventRecord(prof_event(p, p->n), st);
}

// ---- packing kernels --------------------------------------------------------------------------
// conv weight [co][ci][kh][kw] -> [tap][ci][co]
__global__ void k_pack_conv(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = cout * cin * taps;
    if (i >= n) return;
    int co = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
    out[i] = w[((size_t)co * cin + ci) * taps + tap];
}
// eval BatchNorm folded behind the conv: y = acc*scale + shift, eps = 1e-5 (torch default)
__global__ void k_fold_bn(const float* bias, const float* g, const float* b, const float* m, const float* v,
                          float* scale, float* shift, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = g[i] / sqrtf(v[i] + 1e-5f);
    scale[i] = s;
    shift[i] = (bias[i] - m[i]) * s + b[i];
}
// W_ih [4H][8F+E] of one direction -> rows d*4H.. of wih_x [8H][8F] and wih_e [8H][E]
__global__ void k_split_wih(const float* __restrict__ w, float* __restrict__ wx, float* __restrict__ we,
                            int rows, int KX, int E) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)rows * (KX + E);
    if (i >= n) return;
    int c = (int)(i % (KX + E));
    long long r = i / (KX + E);
    if (c < KX) wx[r * KX + c] = w[i];
    else we[r * E + (c - KX)] = w[i];
}
__global__ void k_add2(const float* a, const float* b, float* o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}

static int free_params(vs_engine* e) {
    for (int l = 0; l < 8; ++l) {
        cudaFree(e->conv_w32[l]); cudaFree(e->conv_scale[l]); cudaFree(e->conv_shift[l]);
        e->conv_w32[l] = e->conv_scale[l] = e->conv_shift[l] = nullptr;
    }
    cudaFree(e->wih_x); cudaFree(e->wih_e); cudaFree(e->b_lstm); cudaFree(e->whh);
    cudaFree(e->fc1_w); cudaFree(e->fc1_b); cudaFree(e->fc2_w); cudaFree(e->fc2_b);
    e->wih_x = e->wih_e = e->b_lstm = e->whh = e->fc1_w = e->fc1_b = e->fc2_w = e->fc2_b = nullptr;
    return VS_OK;
}

// ---- workspace --------------------------------------------------------------------------------
struct Workspace {
    float *planeA, *planeB, *xcat, *gates, *bias_u, *hout, *hx, *fc1;
    unsigned int* barrier;
    void* tc;  // tensor-core path scratch
    size_t total;
};

static Workspace carve(const vs_engine* e, int B, int T, int precision, void* base) {
    const int F = e->d.num_freq, H = e->d.lstm_dim, Fp = padded_freq(F);
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
    Workspace w{};
    const size_t rows = (size_t)B * T;
    if (precision == VS_PREC_FP32) {
        w.planeA = (float*)take(rows * Fp * 64 * sizeof(float));
        w.planeB = (float*)take(rows * Fp * 64 * sizeof(float));
        w.xcat = (float*)take(rows * 8 * F * sizeof(float));
        w.fc1 = (float*)take(rows * e->d.fc1_dim * sizeof(float));
        w.tc = nullptr;
    } else {
        w.tc = take(tc_workspace_bytes(e, B, T, precision));
    }
    w.gates = (float*)take(rows * 8 * H * sizeof(float));
    w.bias_u = (float*)take((size_t)B * 8 * H * sizeof(float));
    w.hout = (float*)take(rows * 2 * H * sizeof(float));
    {
        size_t a = lstm_rec_scratch_bytes(e, B), b = tc_lstm_scratch_bytes(e, B);
        w.hx = (float*)take(a > b ? a : b);
    }
    w.barrier = (unsigned int*)take(256);
    w.total = off;
    return w;
}

static int check_common(const vs_engine* e, int B, int T, int precision) {
    if (!e) { set_error("null engine"); return VS_ERR_INVALID; }
    if (!e->loaded) { set_error("parameters not loaded: call vs_engine_load_params first"); return VS_ERR_STATE; }
    if (B < 1 || T < 1) { set_error("B and T must be >= 1"); return VS_ERR_INVALID; }
    if (precision < VS_PREC_FP32 || precision > VS_PREC_FP16_F8C) {
        set_error("unknown precision"); return VS_ERR_INVALID;
    }
    // the kernels' tile schedulers keep tile and row indices in 32 bits
    if ((long long)B * T * padded_freq(e->d.num_freq) >= (1LL << 31)) { set_error("batch too large: B*T*Fp must be < 2^31"); return VS_ERR_INVALID; }
    if (precision != VS_PREC_FP32 && (e->d.lstm_dim % 4) != 0) {
        set_error("tensor-core precisions need lstm_dim % 4 == 0 (16-byte operand rows); use VS_PREC_FP32"); return VS_ERR_INVALID;
    }
    return VS_OK;
}


// conv stack in fp32: x -> xcat [B*T][8F]
static int conv_stack_fp32(vs_engine* e, const float* x, const Workspace& w, float* xcat, int B, int T, cudaStream_t st) {
    VS_LAUNCH(e, KID_FRONT, st, launch_front_fp32(e, x, w.planeA, B, T, st));
    float *src = w.planeA, *dst = w.planeB;
    for (int l = 1; l <= 6; ++l) {
        VS_LAUNCH(e, KID_CONV1 + l - 1, st, launch_conv_fp32(e, l, src, dst, B, T, st));
        float* t = src; src = dst; dst = t;
    }
    VS_LAUNCH(e, KID_POINT8, st, launch_point8_fp32(e, src, xcat, B, T, st));
    return VS_OK;
}

// BiLSTM + head in fp32 on a given xcat
static int lstm_head_fp32(vs_engine* e, const float* xcat, const float* emb, const float* x, const Workspace& w,
                          float* fc1buf, float* mask, float* masked, int B, int T, cudaStream_t st) {
    const int F = e->d.num_freq, H = e->d.lstm_dim, E = e->d.emb_dim, N1 = e->d.fc1_dim;
    const int M = B * T;
    // d-vector folded into a per-utterance gate bias: bias_u = W_ih[:, 8F:] emb + b_ih + b_hh
    VS_LAUNCH(e, KID_EMB_BIAS, st, launch_gemm_fp32(emb, E, e->wih_e, E, e->b_lstm, nullptr, 1, w.bias_u, 8 * H, B, 8 * H, E,
                                      false, EPI_NONE, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_INPROJ, st, launch_gemm_fp32(xcat, 8 * F, e->wih_x, 8 * F, nullptr, w.bias_u, T, w.gates, 8 * H, M, 8 * H, 8 * F,
                                      false, EPI_NONE, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_LSTM_REC, st, launch_lstm_rec_fp32(e, w.gates, w.hout, w.hx, w.barrier, B, T, st));
    VS_LAUNCH(e, KID_FC1, st, launch_gemm_fp32(w.hout, 2 * H, e->fc1_w, 2 * H, e->fc1_b, nullptr, 1, fc1buf, N1, M, N1, 2 * H,
                                      true, EPI_RELU, nullptr, nullptr, st));
    VS_LAUNCH(e, KID_FC2, st, launch_gemm_fp32(fc1buf, N1, e->fc2_w, N1, e->fc2_b, nullptr, 1, mask, F, M, F, N1,
                                      false, EPI_SIGMOID_MASK, x, masked, st));
    return VS_OK;
}

}  // namespace vs

using namespace vs;

extern "C" {

static void pipe_destroy(vs_engine* e);  // pipelined host entry, defined below

int vs_abi_version(void) { return VS_ABI_VERSION; }
const char* vs_last_error(void) { return g_err.c_str(); }

int vs_engine_create(const vs_dims* dims, vs_engine** out) {
    if (!dims || !out) { set_error("null argument"); return VS_ERR_INVALID; }
    if (dims->num_freq < 1 || dims->emb_dim < 1 || dims->lstm_dim < 1 || dims->fc1_dim < 1) {
        set_error("dimensions must be positive"); return VS_ERR_INVALID;
    }
    if (dims->fc2_dim != dims->num_freq) {
        set_error("fc2_dim must equal num_freq (the mask is applied bin by bin)"); return VS_ERR_INVALID;
    }
    if (dims->activation != VS_ACT_MISH && dims->activation != VS_ACT_RELU) {
        set_error("unknown activation"); return VS_ERR_INVALID;
    }
    int dev = 0;
    VS_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    VS_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error(std::string("voicesplit_b200 needs an sm_100a (B200) device, found sm_") + std::to_string(prop.major) +
                  std::to_string(prop.minor));
        return VS_ERR_UNSUPPORTED;
    }
    if (2 * ((dims->lstm_dim + 7) / 8) > prop.multiProcessorCount || (size_t)dims->lstm_dim * 96 * 4 > 220 * 1024) {
        set_error("lstm_dim too large for the persistent recurrent kernel (max 584 on 148 SMs)"); return VS_ERR_INVALID;
    }
    vs_engine* e = new (std::nothrow) vs_engine();
    if (!e) { set_error("out of host memory"); return VS_ERR_STATE; }
    e->d = *dims;
    e->device = dev;
    e->num_sms = prop.multiProcessorCount;
    int rc = tc_create(e);
    if (rc != VS_OK) { delete e; return rc; }
    *out = e;
    return VS_OK;
}

int vs_engine_destroy(vs_engine* e) {
    if (!e) return VS_OK;
    pipe_destroy(e);
    audio_free(e);
    loss_free(e);
    encoder_free(e);
    free_params(e);
    train_free(e);
    tc_destroy(e);
    if (e->prof) {
        for (cudaEvent_t ev : ((Prof*)e->prof)->ev) cudaEventDestroy(ev);
        delete (Prof*)e->prof;
    }
    cudaFree(e->stage);
    delete e;
    return VS_OK;
}

int vs_engine_load_params(vs_engine* e, const vs_params* p, void* stream) {
    if (!e || !p) { set_error("null argument"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int F = e->d.num_freq, H = e->d.lstm_dim, E = e->d.emb_dim, N1 = e->d.fc1_dim;
    if (!e->conv_w32[0]) {
        for (int l = 0; l < 8; ++l) {
            const ConvGeom g = kConv[l];
            VS_CUDA_TRY(cudaMalloc(&e->conv_w32[l], sizeof(float) * g.cout * g.cin * g.kh * g.kw));
            VS_CUDA_TRY(cudaMalloc(&e->conv_scale[l], sizeof(float) * 64));
            VS_CUDA_TRY(cudaMalloc(&e->conv_shift[l], sizeof(float) * 64));
        }
        VS_CUDA_TRY(cudaMalloc(&e->wih_x, sizeof(float) * 8 * H * 8 * F));
        VS_CUDA_TRY(cudaMalloc(&e->wih_e, sizeof(float) * 8 * H * E));
        VS_CUDA_TRY(cudaMalloc(&e->b_lstm, sizeof(float) * 8 * H));
        VS_CUDA_TRY(cudaMalloc(&e->whh, sizeof(float) * 8 * H * H));
        VS_CUDA_TRY(cudaMalloc(&e->fc1_w, sizeof(float) * N1 * 2 * H));
        VS_CUDA_TRY(cudaMalloc(&e->fc1_b, sizeof(float) * N1));
        VS_CUDA_TRY(cudaMalloc(&e->fc2_w, sizeof(float) * F * N1));
        VS_CUDA_TRY(cudaMalloc(&e->fc2_b, sizeof(float) * F));
    }
    for (int l = 0; l < 8; ++l) {
        const ConvGeom g = kConv[l];
        int n = g.cout * g.cin * g.kh * g.kw;
        k_pack_conv<<<(n + 255) / 256, 256, 0, st>>>(p->conv_w[l], e->conv_w32[l], g.cout, g.cin, g.kh * g.kw);
        k_fold_bn<<<1, 64, 0, st>>>(p->conv_b[l], p->bn_gamma[l], p->bn_beta[l], p->bn_mean[l], p->bn_var[l],
                                    e->conv_scale[l], e->conv_shift[l], g.cout);
    }
    for (int d = 0; d < 2; ++d) {
        long long n = (long long)4 * H * (8 * F + E);
        k_split_wih<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p->w_ih[d], e->wih_x + (size_t)d * 4 * H * 8 * F,
                                                                 e->wih_e + (size_t)d * 4 * H * E, 4 * H, 8 * F, E);
        k_add2<<<(4 * H + 255) / 256, 256, 0, st>>>(p->b_ih[d], p->b_hh[d], e->b_lstm + d * 4 * H, 4 * H);
        VS_CUDA_TRY(cudaMemcpyAsync(e->whh + (size_t)d * 4 * H * H, p->w_hh[d], sizeof(float) * 4 * H * H, cudaMemcpyDeviceToDevice, st));
    }
    VS_CUDA_TRY(cudaMemcpyAsync(e->fc1_w, p->fc1_w, sizeof(float) * N1 * 2 * H, cudaMemcpyDeviceToDevice, st));
    VS_CUDA_TRY(cudaMemcpyAsync(e->fc1_b, p->fc1_b, sizeof(float) * N1, cudaMemcpyDeviceToDevice, st));
    VS_CUDA_TRY(cudaMemcpyAsync(e->fc2_w, p->fc2_w, sizeof(float) * F * N1, cudaMemcpyDeviceToDevice, st));
    VS_CUDA_TRY(cudaMemcpyAsync(e->fc2_b, p->fc2_b, sizeof(float) * F, cudaMemcpyDeviceToDevice, st));
    VS_CUDA_TRY(cudaGetLastError());
    int rc = train_pack(e, p, st);      // first: tc_pack also tiles the data-gradient weights it produces
    if (rc != VS_OK) return rc;
    rc = tc_pack(e, st);
    if (rc != VS_OK) return rc;
    e->loaded = true;
    return VS_OK;
}

size_t vs_workspace_bytes(const vs_engine* e, int32_t B, int32_t T, int32_t precision) {
    if (!e || B < 1 || T < 1) return 0;
    return carve(e, B, T, precision, nullptr).total;
}

int vs_forward(vs_engine* e, const float* x, const float* emb, float* mask, float* masked, int32_t B, int32_t T,
               int32_t precision, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    if (!x || !emb || !mask || !workspace) { set_error("null buffer"); return VS_ERR_INVALID; }
    Workspace w = carve(e, B, T, precision, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    e->launches = 0;
    prof_begin(e, st);
    if (precision == VS_PREC_FP32) {
        rc = conv_stack_fp32(e, x, w, w.xcat, B, T, st);
        if (rc != VS_OK) return rc;
        return lstm_head_fp32(e, w.xcat, emb, x, w, w.fc1, mask, masked, B, T, st);
    }
    TcLstmBuffers lb{w.gates, w.bias_u, w.hout, w.hx, w.barrier};
    return tc_forward(e, x, emb, mask, masked, B, T, precision, w.tc, lb, st);
}

int vs_conv_stack(vs_engine* e, const float* x, float* conv_out, int32_t B, int32_t T, int32_t precision,
                  void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    if (!x || !conv_out || !workspace) { set_error("null buffer"); return VS_ERR_INVALID; }
    Workspace w = carve(e, B, T, precision, workspace);
    if (workspace_bytes < w.total) { set_error("workspace too small"); return VS_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    e->launches = 0;
    prof_begin(e, st);
    if (precision == VS_PREC_FP32) return conv_stack_fp32(e, x, w, conv_out, B, T, st);
    return tc_conv_stack(e, x, conv_out, B, T, precision, w.tc, st);
}

int vs_forward_host(vs_engine* e, const float* x_host, const float* emb_host, float* mask_host, float* masked_host,
                    int32_t B, int32_t T, int32_t precision, void* stream) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    if (!x_host || !emb_host || !mask_host) { set_error("null buffer"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nx = (size_t)B * T * e->d.num_freq * sizeof(float), ne = (size_t)B * e->d.emb_dim * sizeof(float);
    const size_t wsb = vs_workspace_bytes(e, B, T, precision);
    const size_t need = align_up(nx, 1024) * 3 + align_up(ne, 1024) + wsb;
    if (need > e->stage_bytes) {
        VS_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(e->stage);
        e->stage = nullptr; e->stage_bytes = 0;
        VS_CUDA_TRY(cudaMalloc(&e->stage, need));
        e->stage_bytes = need;
    }
    char* p = (char*)e->stage;
    float* dx = (float*)p; p += align_up(nx, 1024);
    float* dmask = (float*)p; p += align_up(nx, 1024);
    float* dmasked = (float*)p; p += align_up(nx, 1024);
    float* demb = (float*)p; p += align_up(ne, 1024);
    VS_CUDA_TRY(cudaMemcpyAsync(dx, x_host, nx, cudaMemcpyHostToDevice, st));
    VS_CUDA_TRY(cudaMemcpyAsync(demb, emb_host, ne, cudaMemcpyHostToDevice, st));
    rc = vs_forward(e, dx, demb, dmask, masked_host ? dmasked : nullptr, B, T, precision, p, wsb, stream);
    if (rc != VS_OK) return rc;
    VS_CUDA_TRY(cudaMemcpyAsync(mask_host, dmask, nx, cudaMemcpyDeviceToHost, st));
    if (masked_host) VS_CUDA_TRY(cudaMemcpyAsync(masked_host, dmasked, nx, cudaMemcpyDeviceToHost, st));
    VS_CUDA_TRY(cudaStreamSynchronize(st));
    return VS_OK;
}

// ---- pipelined host entry --------------------------------------------------------------------
struct HostSlot {
    void* io = nullptr;       // device staging: x, emb, mask, masked
    size_t io_bytes = 0;
    cudaEvent_t h2d = nullptr, fwd = nullptr, done = nullptr;
    bool busy = false;
};
struct HostPipe {
    cudaStream_t copy_in = nullptr, compute = nullptr, copy_out = nullptr;
    void* ws = nullptr;       // one workspace: forwards of the two slots are serialised on `compute`
    size_t ws_bytes = 0;
    HostSlot slot[2];
};

static int pipe_get(vs_engine* e, HostPipe** out) {
    if (!e->pipe) {
        HostPipe* p = new (std::nothrow) HostPipe();
        if (!p) { set_error("out of host memory"); return VS_ERR_STATE; }
        VS_CUDA_TRY(cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking));
        VS_CUDA_TRY(cudaStreamCreateWithFlags(&p->compute, cudaStreamNonBlocking));
        VS_CUDA_TRY(cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking));
        for (HostSlot& s : p->slot) {
            VS_CUDA_TRY(cudaEventCreateWithFlags(&s.h2d, cudaEventDisableTiming));
            VS_CUDA_TRY(cudaEventCreateWithFlags(&s.fwd, cudaEventDisableTiming));
            VS_CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        e->pipe = p;
    }
    *out = (HostPipe*)e->pipe;
    return VS_OK;
}

static void pipe_destroy(vs_engine* e) {
    HostPipe* p = (HostPipe*)e->pipe;
    if (!p) return;
    cudaStreamSynchronize(p->copy_in); cudaStreamSynchronize(p->compute); cudaStreamSynchronize(p->copy_out);
    for (HostSlot& s : p->slot) { cudaFree(s.io); cudaEventDestroy(s.h2d); cudaEventDestroy(s.fwd); cudaEventDestroy(s.done); }
    cudaFree(p->ws);
    cudaStreamDestroy(p->copy_in); cudaStreamDestroy(p->compute); cudaStreamDestroy(p->copy_out);
    delete p;
    e->pipe = nullptr;
}

int vs_forward_host_submit(vs_engine* e, int32_t slot, const float* x_host, const float* emb_host, float* mask_host,
                           float* masked_host, int32_t B, int32_t T, int32_t precision) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    if (slot < 0 || slot > 1) { set_error("slot must be 0 or 1"); return VS_ERR_INVALID; }
    if (!x_host || !emb_host || !mask_host) { set_error("null buffer"); return VS_ERR_INVALID; }
    HostPipe* p = nullptr;
    rc = pipe_get(e, &p);
    if (rc != VS_OK) return rc;
    HostSlot& s = p->slot[slot];
    if (s.busy) { set_error("slot still in flight: call vs_forward_host_wait first"); return VS_ERR_STATE; }
    const size_t nx = (size_t)B * T * e->d.num_freq * sizeof(float), ne = (size_t)B * e->d.emb_dim * sizeof(float);
    const size_t io_need = align_up(nx, 1024) * 3 + align_up(ne, 1024);
    const size_t wsb = vs_workspace_bytes(e, B, T, precision);
    if (io_need > s.io_bytes) {     // grow-only; steady-state calls of one shape never allocate
        cudaFree(s.io); s.io = nullptr; s.io_bytes = 0;
        VS_CUDA_TRY(cudaMalloc(&s.io, io_need));
        s.io_bytes = io_need;
    }
    if (wsb > p->ws_bytes) {
        VS_CUDA_TRY(cudaStreamSynchronize(p->compute));
        cudaFree(p->ws); p->ws = nullptr; p->ws_bytes = 0;
        VS_CUDA_TRY(cudaMalloc(&p->ws, wsb));
        p->ws_bytes = wsb;
    }
    char* q = (char*)s.io;
    float* dx = (float*)q; q += align_up(nx, 1024);
    float* dmask = (float*)q; q += align_up(nx, 1024);
    float* dmasked = (float*)q; q += align_up(nx, 1024);
    float* demb = (float*)q;
    VS_CUDA_TRY(cudaMemcpyAsync(dx, x_host, nx, cudaMemcpyHostToDevice, p->copy_in));
    VS_CUDA_TRY(cudaMemcpyAsync(demb, emb_host, ne, cudaMemcpyHostToDevice, p->copy_in));
    VS_CUDA_TRY(cudaEventRecord(s.h2d, p->copy_in));
    VS_CUDA_TRY(cudaStreamWaitEvent(p->compute, s.h2d, 0));
    rc = vs_forward(e, dx, demb, dmask, masked_host ? dmasked : nullptr, B, T, precision, p->ws, p->ws_bytes, p->compute);
    if (rc != VS_OK) return rc;
    VS_CUDA_TRY(cudaEventRecord(s.fwd, p->compute));
    VS_CUDA_TRY(cudaStreamWaitEvent(p->copy_out, s.fwd, 0));
    VS_CUDA_TRY(cudaMemcpyAsync(mask_host, dmask, nx, cudaMemcpyDeviceToHost, p->copy_out));
    if (masked_host) VS_CUDA_TRY(cudaMemcpyAsync(masked_host, dmasked, nx, cudaMemcpyDeviceToHost, p->copy_out));
    VS_CUDA_TRY(cudaEventRecord(s.done, p->copy_out));
    s.busy = true;
    return VS_OK;
}

int vs_forward_host_reserve(vs_engine* e, int32_t B, int32_t T, int32_t precision) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    const size_t nx = (size_t)B * T * e->d.num_freq * sizeof(float), ne = (size_t)B * e->d.emb_dim * sizeof(float);
    const size_t io_need = align_up(nx, 1024) * 3 + align_up(ne, 1024);
    const size_t wsb = vs_workspace_bytes(e, B, T, precision);
    if (io_need + wsb > e->stage_bytes) {
        VS_CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(e->stage);
        e->stage = nullptr; e->stage_bytes = 0;
        VS_CUDA_TRY(cudaMalloc(&e->stage, io_need + wsb));
        e->stage_bytes = io_need + wsb;
    }
    HostPipe* p = nullptr;
    rc = pipe_get(e, &p);
    if (rc != VS_OK) return rc;
    for (HostSlot& s : p->slot) {
        if (s.busy) { set_error("a slot is in flight: wait for it before reserving"); return VS_ERR_STATE; }
        if (io_need > s.io_bytes) {
            cudaFree(s.io); s.io = nullptr; s.io_bytes = 0;
            VS_CUDA_TRY(cudaMalloc(&s.io, io_need));
            s.io_bytes = io_need;
        }
    }
    if (wsb > p->ws_bytes) {
        VS_CUDA_TRY(cudaStreamSynchronize(p->compute));
        cudaFree(p->ws); p->ws = nullptr; p->ws_bytes = 0;
        VS_CUDA_TRY(cudaMalloc(&p->ws, wsb));
        p->ws_bytes = wsb;
    }
    return VS_OK;
}

int vs_forward_host_wait(vs_engine* e, int32_t slot) {
    if (!e || !e->pipe || slot < 0 || slot > 1) { set_error("nothing submitted on this slot"); return VS_ERR_STATE; }
    HostSlot& s = ((HostPipe*)e->pipe)->slot[slot];
    if (!s.busy) { set_error("nothing submitted on this slot"); return VS_ERR_STATE; }
    VS_CUDA_TRY(cudaEventSynchronize(s.done));
    s.busy = false;
    return VS_OK;
}

int vs_debug_conv_layer(vs_engine* e, int32_t layer, const float* in_nchw, float* out_nchw, int32_t B, int32_t T,
                        int32_t precision, void* stream) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    if (layer < 0 || layer > 6) { set_error("layer must be 0..6"); return VS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int F = e->d.num_freq, Fp = padded_freq(F);
    const size_t pb = (size_t)B * T * Fp * 64 * sizeof(float);
    float *pin = nullptr, *pout = nullptr;
    VS_CUDA_TRY(cudaMalloc(&pin, pb));
    VS_CUDA_TRY(cudaMalloc(&pout, pb));
    int ret = VS_OK;
    cudaError_t ce = cudaSuccess;
    if (layer == 0) {
        // input is the spectrogram itself: [B][1][T][F] == [B][T][F]
        if (precision == VS_PREC_FP32) ce = launch_front_fp32(e, in_nchw, pout, B, T, st);
        else ret = tc_debug_layer(e, 0, in_nchw, nullptr, pout, B, T, precision, st);
    } else {
        ce = launch_nchw_to_plane(in_nchw, pin, B, 64, T, F, st);
        if (ce == cudaSuccess) {
            if (precision == VS_PREC_FP32) ce = launch_conv_fp32(e, layer, pin, pout, B, T, st);
            else ret = tc_debug_layer(e, layer, nullptr, pin, pout, B, T, precision, st);
        }
    }
    if (ce == cudaSuccess && ret == VS_OK) ce = launch_plane_to_nchw(pout, out_nchw, B, 64, T, F, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(pin); cudaFree(pout);
    if (ce != cudaSuccess) { set_error(std::string("vs_debug_conv_layer: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    return ret;
}

int vs_debug_lstm_head(vs_engine* e, const float* conv_out, const float* emb, const float* x, float* lstm_out,
                       float* mask, int32_t B, int32_t T, int32_t precision, void* stream) {
    int rc = check_common(e, B, T, precision);
    if (rc != VS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t wsb = vs_workspace_bytes(e, B, T, VS_PREC_FP32);
    void* ws = nullptr;
    VS_CUDA_TRY(cudaMalloc(&ws, wsb));
    Workspace w = carve(e, B, T, VS_PREC_FP32, ws);
    int ret;
    if (precision == VS_PREC_FP32) {
        ret = lstm_head_fp32(e, conv_out, emb, x, w, w.fc1, mask, nullptr, B, T, st);
    } else {
        TcLstmBuffers lb{w.gates, w.bias_u, w.hout, w.hx, w.barrier};
        ret = tc_debug_lstm_head(e, conv_out, emb, x, mask, B, T, precision, lb, st);
    }
    cudaError_t ce = cudaSuccess;
    if (ret == VS_OK && lstm_out)
        ce = cudaMemcpyAsync(lstm_out, w.hout, (size_t)B * T * 2 * e->d.lstm_dim * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(ws);
    if (ce != cudaSuccess) { set_error(std::string("vs_debug_lstm_head: ") + cudaGetErrorString(ce)); return VS_ERR_CUDA; }
    return ret;
}

int vs_debug_lstm_timing(vs_engine* e, int64_t* cycles8) {
    if (!e || !cycles8) { set_error("null argument"); return VS_ERR_INVALID; }
    return tc_lstm_read_timing(tc_lstm_slot(e), (long long*)cycles8);
}

int vs_last_launch_count(const vs_engine* e) { return e ? e->launches : 0; }

int vs_engine_set_profiling(vs_engine* e, int32_t enabled) {
    if (!e) { set_error("null engine"); return VS_ERR_INVALID; }
    e->profiling = enabled != 0;
    return VS_OK;
}

int vs_profile_read(vs_engine* e, int32_t max_entries, int32_t* kernel_ids, float* milliseconds) {
    if (!e || !e->prof) return 0;
    Prof* p = (Prof*)e->prof;
    if (p->n == 0) return 0;
    if (cudaEventSynchronize(p->ev[p->n]) != cudaSuccess) return 0;
    int n = p->n < max_entries ? p->n : max_entries;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p->ev[i], p->ev[i + 1]);
        kernel_ids[i] = p->ids[i];
        milliseconds[i] = ms;
    }
    return n;
}

}  // extern "C"
