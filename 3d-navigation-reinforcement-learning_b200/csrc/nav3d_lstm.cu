// nav3d_lstm.cu — one-layer LSTM over a whole sequence minibatch with episode-start resets, forward and backward, for the
// PPO update of the trainer.  The recurrent GEMMs are plain library calls (cuBLAS, optionally TF32
// tensor-op math); what is hand-written is everything between them: ONE fused pointwise kernel per timestep in each
// direction that applies the bias, the gate non-linearities, the cell update, the episode-start mask of the NEXT step
// (forward) or the mask and the four gate derivatives (backward), in place in the gate buffer.
//
// Why not cuDNN (what torch.nn.LSTM calls): (1) it has no notion of an episode start inside a sequence, so a rollout with
// resets has to be cut into pieces on the host; here the mask is an operand.  (2) its backward runs a bulk pass over all
// gate gradients (GENERIC_elementWise_bp2, 3.6 ms for 128 x 2048 x 1024 floats on B200) that costs as much as all the
// per-step kernels together; keeping gate gradients in the layout the weight-gradient GEMM reads removes it.
//
// Replaces, in the reference's third-party stack, the `nn.LSTM` calls inside sb3-contrib's
// RecurrentActorCriticPolicy._process_sequence during RecurrentPPO.train() (driven by train/Grid_Train.py:228).
// Weight layout is torch's: W_ih [4H, F], W_hh [4H, H], gate order i, f, g, o.
#include "../../include/nav3d.h"

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>
#include <string>

namespace nav3d { int fail_with(int code, const std::string &msg); }

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// h_in[0] = h0 * keep_0
__global__ void __launch_bounds__(256) lstm_mask_h0_kernel(const float *__restrict__ h0, const uint8_t *__restrict__ starts0,
                                                           int B, int H, float *__restrict__ h_in0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H;
    h_in0[idx] = starts0[b] ? 0.f : h0[idx];
}

// One timestep of the forward pass, one thread per (env b, unit j).  gates_t holds W_ih x_t + W_hh h_in_t (pre-activation,
// without bias) on entry and the activated gates i, f, g, o on exit.
__global__ void __launch_bounds__(256) lstm_fwd_step_kernel(float *__restrict__ gates_t, const float *__restrict__ bias_ih,
                                                            const float *__restrict__ bias_hh,
                                                            const float *__restrict__ c_prev,      // c_{t-1} or c0, unmasked
                                                            const uint8_t *__restrict__ starts_t,  // [B]
                                                            const uint8_t *__restrict__ starts_next,   // [B] or NULL at t = S-1
                                                            int B, int H, float *__restrict__ h_t, float *__restrict__ c_t,
                                                            float *__restrict__ h_in_next) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");     // launched with programmatic stream serialisation (see pdl_launch)
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    float *g4 = gates_t + (size_t)b * 4 * H;
    const float pi = g4[j] + bias_ih[j] + bias_hh[j];
    const float pf = g4[H + j] + bias_ih[H + j] + bias_hh[H + j];
    const float pg = g4[2 * H + j] + bias_ih[2 * H + j] + bias_hh[2 * H + j];
    const float po = g4[3 * H + j] + bias_ih[3 * H + j] + bias_hh[3 * H + j];
    const float i = sigmoidf_(pi), f = sigmoidf_(pf), g = tanhf(pg), o = sigmoidf_(po);
    const float c_in = starts_t[b] ? 0.f : c_prev[idx];
    const float c = f * c_in + i * g;
    const float h = o * tanhf(c);
    g4[j] = i; g4[H + j] = f; g4[2 * H + j] = g; g4[3 * H + j] = o;
    c_t[idx] = c;
    h_t[idx] = h;
    if (h_in_next) h_in_next[idx] = starts_next[b] ? 0.f : h;
}

// One timestep of the backward pass.  gates_t: activated gates on entry, gradients w.r.t. the pre-activations on exit.
// dh_rec = (dgates_{t+1} W_hh), unmasked; dc_rec = dL/dc_t carried from step t+1 (already masked there); both may be NULL
// at t = S-1.  dc_rec is updated in place to dL/dc_{t-1}.
__global__ void __launch_bounds__(256) lstm_bwd_step_kernel(float *__restrict__ gates_t, const float *__restrict__ dh_out_t,
                                                            const float *__restrict__ dh_rec, float *__restrict__ dc_rec,
                                                            const float *__restrict__ c_t, const float *__restrict__ c_prev,
                                                            const uint8_t *__restrict__ starts_t,
                                                            const uint8_t *__restrict__ starts_next, int first, int B, int H) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    float *g4 = gates_t + (size_t)b * 4 * H;
    const float i = g4[j], f = g4[H + j], g = g4[2 * H + j], o = g4[3 * H + j];
    float dh = dh_out_t ? dh_out_t[idx] : 0.f;
    float dc_in = 0.f;
    if (!first) {
        if (!starts_next[b]) dh += dh_rec[idx];       // h_in_{t+1} = keep_{t+1} * h_t
        dc_in = dc_rec[idx];
    }
    const float keep = starts_t[b] ? 0.f : 1.f;
    const float c_in = keep * c_prev[idx];
    const float tc = tanhf(c_t[idx]);
    const float d_o = dh * tc;
    const float dc = dh * o * (1.f - tc * tc) + dc_in;
    g4[j] = dc * g * i * (1.f - i);
    g4[H + j] = dc * c_in * f * (1.f - f);
    g4[2 * H + j] = dc * i * (1.f - g * g);
    g4[3 * H + j] = d_o * o * (1.f - o);
    dc_rec[idx] = dc * f * keep;
}

// db[c] = sum over rows of dG[row][c].  grid.x tiles the 4H columns (one thread per column: coalesced row reads),
// grid.y splits the rows; partial sums meet in db with one atomicAdd per thread (db zeroed beforehand).
__global__ void __launch_bounds__(256) column_sum_kernel(const float *__restrict__ dg, long long rows, int cols,
                                                         float *__restrict__ db) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const long long per = (rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    long long r = r0;
    for (; r + 3 < r1; r += 4) {
        acc0 += dg[r * cols + c]; acc1 += dg[(r + 1) * cols + c];
        acc2 += dg[(r + 2) * cols + c]; acc3 += dg[(r + 3) * cols + c];
    }
    for (; r < r1; r++) acc0 += dg[r * cols + c];
    atomicAdd(db + c, (acc0 + acc1) + (acc2 + acc3));
}

std::mutex g_handle_mu;
std::map<std::pair<int, void *>, cublasHandle_t> g_handles;

// One handle per (device, stream): the actor's and the critic's recurrences run concurrently on two streams, and a cuBLAS
// handle's workspace must not be shared by GEMMs in flight on different streams.
int get_handle(void *stream, cublasHandle_t *out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nav3d::fail_with(NAV3D_ERR_CUDA, "nav3d_lstm: no current CUDA device");
    std::lock_guard<std::mutex> lk(g_handle_mu);
    auto key = std::make_pair(dev, stream);
    auto it = g_handles.find(key);
    if (it == g_handles.end()) {
        cublasHandle_t h;
        if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) return nav3d::fail_with(NAV3D_ERR_CUDA, "nav3d_lstm: cublasCreate failed");
        if (cublasSetStream(h, (cudaStream_t)stream) != CUBLAS_STATUS_SUCCESS)
            return nav3d::fail_with(NAV3D_ERR_CUDA, "nav3d_lstm: cublasSetStream failed");
        // An explicit workspace: inside a CUDA-graph capture cuBLAS may not allocate, and a captured GEMM must find the
        // same workspace at every replay.  (Held for the life of the process, like the handle.)
        void *ws = nullptr;
        constexpr size_t kWorkspace = 32u << 20;
        if (cudaMalloc(&ws, kWorkspace) != cudaSuccess || cublasSetWorkspace(h, ws, kWorkspace) != CUBLAS_STATUS_SUCCESS)
            return nav3d::fail_with(NAV3D_ERR_CUDA, "nav3d_lstm: cuBLAS workspace allocation failed");
        it = g_handles.emplace(key, h).first;
    }
    *out = it->second;
    return NAV3D_OK;
}

#define CUBLAS_TRY(expr)                                                                                   \
    do {                                                                                                   \
        cublasStatus_t _s = (expr);                                                                        \
        if (_s != CUBLAS_STATUS_SUCCESS)                                                                   \
            return nav3d::fail_with(NAV3D_ERR_CUDA, std::string(#expr) + ": cuBLAS status " + std::to_string((int)_s)); \
    } while (0)

unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

// The per-timestep kernels sit between cuBLAS GEMMs in a chain of ~10 us links: launched with programmatic stream
// serialisation their blocks become resident while the GEMM before them drains and wait in griddepcontrol.wait.
template <typename... KArgs, typename... Args>
cudaError_t pdl_launch(void (*kernel)(KArgs...), unsigned grid, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace

extern "C" {

int nav3d_lstm_prepare(void *stream) {
    cublasHandle_t hnd;
    return get_handle(stream, &hnd);
}

int nav3d_lstm_forward(const float *x, const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh,
                       const float *h0, const float *c0, const uint8_t *starts, int32_t S, int32_t B, int32_t F, int32_t H,
                       int32_t tf32, float *gates, float *h_in, float *h_all, float *c_all, void *stream) {
    if (!x || !w_ih || !w_hh || !b_ih || !b_hh || !h0 || !c0 || !starts || !gates || !h_in || !h_all || !c_all)
        return nav3d::fail_with(NAV3D_ERR_INVALID, "nav3d_lstm_forward: NULL buffer");
    if (S < 1 || B < 1 || F < 1 || H < 1) return nav3d::fail_with(NAV3D_ERR_INVALID, "nav3d_lstm_forward: bad sizes");
    cublasHandle_t hnd;
    if (int rc = get_handle(stream, &hnd)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUBLAS_TRY(cublasSetMathMode(hnd, tf32 ? CUBLAS_TF32_TENSOR_OP_MATH : CUBLAS_PEDANTIC_MATH));
    const float one = 1.f, zero = 0.f;
    const long long SB = (long long)S * B;
    const int G4 = 4 * H;
    // gates[S*B, 4H] = x[S*B, F] @ W_ih^T        (row-major; one GEMM for the whole sequence)
    CUBLAS_TRY(cublasSgemm(hnd, CUBLAS_OP_T, CUBLAS_OP_N, G4, (int)SB, F, &one, w_ih, F, x, F, &zero, gates, G4));
    const long long BH = (long long)B * H;
    lstm_mask_h0_kernel<<<blocks_for(BH), 256, 0, s>>>(h0, starts, B, H, h_in);
    for (int t = 0; t < S; t++) {
        float *g_t = gates + (size_t)t * B * G4;
        // gates_t += h_in_t[B, H] @ W_hh^T
        CUBLAS_TRY(cublasSgemm(hnd, CUBLAS_OP_T, CUBLAS_OP_N, G4, B, H, &one, w_hh, H, h_in + (size_t)t * BH, H, &one, g_t, G4));
        const float *c_prev = t == 0 ? c0 : c_all + (size_t)(t - 1) * BH;
        const bool last = t == S - 1;
        pdl_launch(lstm_fwd_step_kernel, blocks_for(BH), s, g_t, b_ih, b_hh, c_prev, starts + (size_t)t * B,
                   last ? nullptr : starts + (size_t)(t + 1) * B, B, H, h_all + (size_t)t * BH, c_all + (size_t)t * BH,
                   last ? nullptr : h_in + (size_t)(t + 1) * BH);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return nav3d::fail_with(NAV3D_ERR_CUDA, std::string("nav3d_lstm_forward: ") + cudaGetErrorString(err));
    return NAV3D_OK;
}

int nav3d_lstm_backward(const float *x, const float *w_hh, const float *c0, const uint8_t *starts, const float *h_in,
                        const float *c_all, float *gates, const float *dh_all, int32_t S, int32_t B, int32_t F, int32_t H,
                        int32_t tf32, float *dw_ih, float *dw_hh, float *db, float *scratch, void *stream) {
    if (!x || !w_hh || !c0 || !starts || !h_in || !c_all || !gates || !dh_all || !dw_ih || !dw_hh || !db || !scratch)
        return nav3d::fail_with(NAV3D_ERR_INVALID, "nav3d_lstm_backward: NULL buffer");
    if (S < 1 || B < 1 || F < 1 || H < 1) return nav3d::fail_with(NAV3D_ERR_INVALID, "nav3d_lstm_backward: bad sizes");
    cublasHandle_t hnd;
    if (int rc = get_handle(stream, &hnd)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUBLAS_TRY(cublasSetMathMode(hnd, tf32 ? CUBLAS_TF32_TENSOR_OP_MATH : CUBLAS_PEDANTIC_MATH));
    const float one = 1.f, zero = 0.f;
    const long long SB = (long long)S * B, BH = (long long)B * H;
    const int G4 = 4 * H;
    float *dh_rec = scratch, *dc_rec = scratch + BH;                               // scratch: 2*B*H floats (+ S*B unused)
    for (int t = S - 1; t >= 0; t--) {
        float *g_t = gates + (size_t)t * B * G4;
        const float *c_prev = t == 0 ? c0 : c_all + (size_t)(t - 1) * BH;
        const bool first = t == S - 1;
        pdl_launch(lstm_bwd_step_kernel, blocks_for(BH), s, g_t, dh_all + (size_t)t * BH, dh_rec, dc_rec,
                   c_all + (size_t)t * BH, c_prev, starts + (size_t)t * B, first ? nullptr : starts + (size_t)(t + 1) * B,
                   first ? 1 : 0, B, H);
        // dh_rec[B, H] = dgates_t[B, 4H] @ W_hh[4H, H]
        if (t > 0) CUBLAS_TRY(cublasSgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_N, H, B, G4, &one, w_hh, H, g_t, G4, &zero, dh_rec, H));
    }
    // weight gradients over the whole sequence: dW_hh[4H, H] = dG^T @ h_in, dW_ih[4H, F] = dG^T @ x, db = column sums of dG
    CUBLAS_TRY(cublasSgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_T, H, G4, (int)SB, &one, h_in, H, gates, G4, &zero, dw_hh, H));
    CUBLAS_TRY(cublasSgemm(hnd, CUBLAS_OP_N, CUBLAS_OP_T, F, G4, (int)SB, &one, x, F, gates, G4, &zero, dw_ih, F));
    if (cudaMemsetAsync(db, 0, sizeof(float) * G4, s) != cudaSuccess)
        return nav3d::fail_with(NAV3D_ERR_CUDA, "nav3d_lstm_backward: cudaMemsetAsync failed");
    {
        const unsigned gx = (unsigned)((G4 + 255) / 256);
        const unsigned gy = (unsigned)(SB < 592 ? SB : 592);       // 148 SMs x 4 row slabs per column tile
        column_sum_kernel<<<dim3(gx, gy), 256, 0, s>>>(gates, SB, G4, db);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return nav3d::fail_with(NAV3D_ERR_CUDA, std::string("nav3d_lstm_backward: ") + cudaGetErrorString(err));
    return NAV3D_OK;
}

}  // extern "C"
