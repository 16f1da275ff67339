// sr.cu -- SR autoencoder inference (encoder_10 + decoder_400) behind the C ABI of include/srcfd.h.
// Architecture: sr-ae-conv.ipynb cell lines 162-169 (encoder) and 277-287 (decoder); call sites
// PyCFD_ML_accelerated.py:831-858.  NHWC float32 like Keras; kernel layouts are converted once at
// upload so that the output-channel index is the contiguous (coalesced) one.
//
// This file holds the fp32 CUDA-core path: every layer is one kernel, one thread per output element
// (output channel fastest => weight reads coalesced, activation reads broadcast).
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/srcfd.h"
#include "sr_tc.cuh"

namespace {

thread_local std::string g_sr_err;
int sr_fail(int code, const std::string& m) { g_sr_err = m; return code; }
#define SRCK(call)                                                                                     \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return sr_fail(SRCFD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }

// out[b, n] = act(in[b, :] . W[:, n] + bias[n]);  W is (K, N) row-major (Keras Dense layout)
template <typename TOUT>
__device__ __forceinline__ void st_act(TOUT* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_act<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename TOUT>
__global__ void k_dense(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                        TOUT* __restrict__ out, int B, int K, int N, int act) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int n = (int)(t % N), b = (int)(t / N);
    const float* x = in + (long long)b * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(x[k], W[(long long)k * N + n], acc);
    acc += bias[n];
    st_act<TOUT>(out + t, act ? swishf(acc) : acc);
}

// Same layer for small K (the decoder's Dense 50 -> 36 864): a thread keeps one output column for BT samples, so a weight
// is fetched once per BT samples instead of once per sample (the one-output-per-thread form re-reads the 7.4 MB matrix
// from L2 for every sample: 76 us per 128 samples, L2-bound).  Per output the k-ascending fma order is unchanged.
template <typename TOUT, int BT>
__global__ void __launch_bounds__(256) k_dense_bt(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                                   TOUT* __restrict__ out, int B, int K, int N, int act) {
    __shared__ float xs[BT * 128];                             // K <= 128
    const int b0 = blockIdx.y * BT, nb = min(BT, B - b0);
    for (int c = threadIdx.x; c < BT * K; c += blockDim.x) { const int j = c / K; xs[c] = j < nb ? in[(long long)(b0 + j) * K + (c - j * K)] : 0.f; }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float w = W[(long long)k * N + n];
#pragma unroll
        for (int j = 0; j < BT; ++j) acc[j] = fmaf(xs[j * K + k], w, acc[j]);
    }
    const float bn = bias[n];
#pragma unroll
    for (int j = 0; j < BT; ++j)
        if (j < nb) { const float v = acc[j] + bn; st_act<TOUT>(out + (long long)(b0 + j) * N + n, act ? swishf(v) : v); }
}

// Conv2D, cross-correlation, zero padding (pt, pl) at the top/left; W (kh, kw, Cin, Cout)
template <typename TIN>
__device__ __forceinline__ float ld_act(const TIN* p) { return (float)*p; }

template <typename TIN>
__global__ void k_conv2d(const TIN* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                         float* __restrict__ out, int B, int H, int Wd, int Cin, int OH, int OW, int Cout, int kh, int kw,
                         int stride, int pt, int pl, int act) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * OH * OW * Cout) return;
    const int co = (int)(t % Cout);
    long long p = t / Cout;
    const int ox = (int)(p % OW); p /= OW;
    const int oy = (int)(p % OH);
    const int b = (int)(p / OH);
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
        const int iy = oy * stride + ky - pt;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < kw; ++kx) {
            const int ix = ox * stride + kx - pl;
            if (ix < 0 || ix >= Wd) continue;
            const TIN* x = in + (((long long)b * H + iy) * Wd + ix) * Cin;
            const float* w = W + ((long long)(ky * kw + kx) * Cin) * Cout + co;
            for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ld_act<TIN>(x + ci), w[(long long)ci * Cout], acc);
        }
    }
    acc += bias[co];
    out[t] = act ? swishf(acc) : acc;
}

// Final layer of decoder_400: Conv2D(1, 3x3, 'same', linear) on (B, H, W, 8) -> (B, H, W, 1).  HBM-bound (2.56 MB of bf16
// in, 0.64 MB out per sample): a CTA stages an (18 x 66) pixel tile with its halo in shared memory (one 16-byte or
// 32-byte vector per pixel), every thread produces 4 vertically adjacent outputs so each staged row is read once per
// three taps, weights (72 floats) sit in registers.  W layout (3, 3, 8, 1).
constexpr int FC_TX = 64, FC_TY = 16;
struct FinalConvW { float w[72]; float b; };      // passed by value: lives in the constant bank, no registers
template <typename TIN>
__global__ void __launch_bounds__(256, 4) k_conv3x3_c8_final(const TIN* __restrict__ in, const FinalConvW fw, float* __restrict__ out,
                                                             int B, int H, int Wd) {
    __shared__ float tile[(FC_TY + 2) * (FC_TX + 2) * 8];       // staged as fp32 (19 KB... 38 KB): one conversion per input value
    const int b = blockIdx.z, y0 = blockIdx.y * FC_TY, x0 = blockIdx.x * FC_TX;
    const TIN* src = in + (long long)b * H * Wd * 8;
    // all vector loads of the tile are in flight before the first conversion/store (one memory round trip per CTA)
    constexpr int NPIX = (FC_TY + 2) * (FC_TX + 2), PER = (NPIX + 255) / 256;
    using VecT = typename std::conditional<sizeof(TIN) == 2, uint4, float4>::type;     // 8 bf16, or 4 of the 8 floats
    VecT va[PER], vb[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int p = threadIdx.x + q * 256;
        const int ty = p / (FC_TX + 2), tx = p - ty * (FC_TX + 2);
        const int y = y0 + ty - 1, x = x0 + tx - 1;
        va[q] = VecT{}; vb[q] = VecT{};
        if (p < NPIX && y >= 0 && y < H && x >= 0 && x < Wd) {
            const VecT* qv = reinterpret_cast<const VecT*>(src + ((long long)y * Wd + x) * 8);
            va[q] = __ldg(qv);
            if constexpr (sizeof(TIN) != 2) vb[q] = __ldg(qv + 1);
        }
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int p = threadIdx.x + q * 256;
        if (p >= NPIX) continue;
        float v[8];
        if constexpr (sizeof(TIN) == 2) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&va[q]);
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float2 f = __bfloat1622float2(h2[c]); v[2 * c] = f.x; v[2 * c + 1] = f.y; }
        } else {
            const float4 a = *reinterpret_cast<const float4*>(&va[q]), c4 = *reinterpret_cast<const float4*>(&vb[q]);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
        }
        float4* d = reinterpret_cast<float4*>(tile + (size_t)p * 8);
        d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    const float bs = fw.b;
    __syncthreads();
    const int tx = threadIdx.x & (FC_TX - 1), tq = threadIdx.x / FC_TX;        // 4 row groups of 4 rows
    float acc[4] = {bs, bs, bs, bs};
#pragma unroll
    for (int r = 0; r < 6; ++r) {                                               // staged rows 4*tq + r feed outputs r-2 .. r
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float4* q = reinterpret_cast<const float4*>(tile + ((size_t)(4 * tq + r) * (FC_TX + 2) + tx + kx) * 8);
            const float4 a = q[0], c4 = q[1];
            const float xv[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int ky = r - o;
                if (ky >= 0 && ky < 3) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[o] = fmaf(xv[c], fw.w[(ky * 3 + kx) * 8 + c], acc[o]);
                }
            }
        }
    }
    const int x = x0 + tx;
    if (x < Wd) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int y = y0 + 4 * tq + o;
            if (y < H) out[((long long)b * H + y) * Wd + x] = acc[o];
        }
    }
}

// Conv2DTranspose 'valid' in gather form; Wt (kh, kw, Cin, Cout) (= Keras (kh, kw, Cout, Cin) transposed at upload)
__global__ void k_conv2d_transpose(const float* __restrict__ in, const float* __restrict__ Wt, const float* __restrict__ bias,
                                   float* __restrict__ out, int B, int H, int Wd, int Cin, int OH, int OW, int Cout, int k,
                                   int stride, int act) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * OH * OW * Cout) return;
    const int co = (int)(t % Cout);
    long long p = t / Cout;
    const int X = (int)(p % OW); p /= OW;
    const int Y = (int)(p % OH);
    const int b = (int)(p / OH);
    float acc = 0.f;
    for (int ky = Y % stride; ky < k; ky += stride) {
        const int y = (Y - ky) / stride;
        if (y < 0 || y >= H) continue;
        for (int kx = X % stride; kx < k; kx += stride) {
            const int x = (X - kx) / stride;
            if (x < 0 || x >= Wd) continue;
            const float* xin = in + (((long long)b * H + y) * Wd + x) * Cin;
            const float* w = Wt + ((long long)(ky * k + kx) * Cin) * Cout + co;
            for (int ci = 0; ci < Cin; ++ci) acc = fmaf(xin[ci], w[(long long)ci * Cout], acc);
        }
    }
    acc += bias[co];
    out[t] = act ? swishf(acc) : acc;
}

// ---- pre / post of ml_super_resolution (LDC.py:841-876, BFS.py:1084-1134) on the device ----------------------------------
// stats per field: {mean_lr, std_lr, mean_hr, std_hr}.  One block per field: optional blend of the training statistics
// with the field's own mean / std (BFS.py:1090-1100, np.mean / np.std over the 100 coarse values), then
// standardize_with_stats (LDC.py:665-668: (x - mean) / std, std == 0 -> 1e-8) in float32.
__global__ void k_sr_pre(const float* __restrict__ x, const double* __restrict__ stats, int adaptive, double blend,
                         float* __restrict__ xs) {
    __shared__ double red[128];
    const int b = blockIdx.x, t = threadIdx.x;
    const float v = t < 100 ? x[(size_t)b * 100 + t] : 0.f;
    double mean = stats[4 * b + 0], sd = stats[4 * b + 1];
    if (adaptive) {
        red[t] = t < 100 ? (double)v : 0.0;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
        const double m = red[0] / 100.0;
        __syncthreads();
        const double d = t < 100 ? (double)v - m : 0.0;
        red[t] = d * d;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
        const double in_sd = sqrt(red[0] / 100.0);
        mean = (1.0 - blend) * mean + blend * (double)(float)m;
        sd = (1.0 - blend) * sd + blend * fmax((double)(float)in_sd, 1e-8);
    }
    if (sd == 0.0) sd = 1e-8;
    if (t < 100) xs[(size_t)b * 100 + t] = (v - (float)mean) / (float)sd;
}
// inverse_standardize (LDC.py:671-673: pred * std_hr + mean_hr) and the NaN/Inf guard (LDC.py:869-876), in place.
__global__ void k_sr_post(float* __restrict__ y, const double* __restrict__ stats, long long per, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const float v = y[i] * (float)stats[4 * b + 3] + (float)stats[4 * b + 2];
        y[i] = (isnan(v) || isinf(v)) ? 0.f : v;
    }
}

struct Layer { float* W = nullptr; float* b = nullptr; __nv_bfloat16* Wbf = nullptr; __nv_bfloat16* Wlo = nullptr;   // Wlo = bf16(W - Wbf): split-operand path
               __nv_bfloat16* Wimg = nullptr; };   // 3x3 ConvT only: both halves as the shared-memory image of k_convT3x3_l1_tc3's stages

}  // namespace

struct srcfd_sr {
    int dev = 0;
    cudaStream_t stream = nullptr;
    Layer enc[4], dec[7];
    bool has_enc = false, has_dec = false;
    float* act[8] = {nullptr};      // activation buffers for one chunk
    float *zin = nullptr, *xin = nullptr;
    int chunk = 0;
    cudaEvent_t ea = nullptr, eb = nullptr;
    int64_t launches = 0;
    int precision = 3;                 // 3 (default): split-operand tcgen05 (fp32-grade accuracy); 0: fp32 CUDA cores; 1: bf16 tcgen05 (fastest, bf16 activations)
    __nv_bfloat16* actbf[6] = {nullptr};   // bf16 activations of layers 1..5 (index = layer) for one chunk
    int* tc_err = nullptr;
    FinalConvW fcw;                    // host copy of the last layer's 72 weights + bias (kernel argument)
    // environment knobs, read once at creation
    int tc_persist = 16;               // SRCFD_TC_PERSIST: CTAs per SM of the persistent ConvT launch (batch 1024: 8.37 ms with one
                                       // tile per CTA, 8.01 ms persistent, 20.6 ms with ONE persistent CTA per SM)
    int tail_fused = 1;                // SRCFD_TAIL_FUSED=0: split-operand path with the last ConvT and the final conv as two launches
    int tail_fused_default = 1;
    int tc3_wide256 = 1;               // SRCFD_TC3_WIDE256=0: the 128->64 and 64->32 split-operand layers with 128 threads per CTA
    int l1_one_cta = 1, l1_one_cta_default = 1;   // SRCFD_L1_ONE_CTA=0: split-operand 3x3 ConvT as per-tap tiles (k_convT2x2_tc3) instead of k_convT3x3_l1_tc3
    int tc3_l1_ctas = 16;              // SRCFD_TC3_L1_CTAS: CTAs per tap of the split-operand 3x3 ConvT (9 taps x 16 = 144 CTAs on 148 SMs)
    int final_tc = 1;                  // SRCFD_FINAL_TC=0: final conv on the CUDA-core tile kernel
    int final_tc_rows = 8;             // SRCFD_FINAL_TC_ROWS: 4 | 8 | 16 output rows per CTA of the tensor-core final conv
    double* stats_dev = nullptr;       // per-field {mean_lr, std_lr, mean_hr, std_hr} of srcfd_sr_super_resolve
    int stats_cap = 0;
};

namespace {

const int DEC_ACT_ELEMS[7] = {12 * 12 * 256, 25 * 25 * 128, 50 * 50 * 64, 100 * 100 * 32, 200 * 200 * 16, 400 * 400 * 8, 400 * 400};

int upload(float** dst, const float* src, size_t n, cudaStream_t s) {
    if (!*dst) SRCK(cudaMalloc(dst, n * sizeof(float)));
    SRCK(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice, s));
    return SRCFD_OK;
}
// Keras Conv2DTranspose kernel (kh,kw,Cout,Cin) -> (kh,kw,Cin,Cout)
std::vector<float> transpose_last2(const float* k, int taps, int Cout, int Cin) {
    std::vector<float> o((size_t)taps * Cin * Cout);
    for (int t = 0; t < taps; ++t)
        for (int co = 0; co < Cout; ++co)
            for (int ci = 0; ci < Cin; ++ci) o[((size_t)t * Cin + ci) * Cout + co] = k[((size_t)t * Cout + co) * Cin + ci];
    return o;
}
int ensure_chunk(srcfd_sr* h, int chunk) {
    if (h->chunk >= chunk) return SRCFD_OK;
    for (int i = 0; i < 8; ++i) { cudaFree(h->act[i]); h->act[i] = nullptr; }
    cudaFree(h->zin); cudaFree(h->xin); h->zin = h->xin = nullptr;
    for (int i = 0; i < 7; ++i) SRCK(cudaMalloc(&h->act[i], (size_t)chunk * DEC_ACT_ELEMS[i] * sizeof(float)));
    SRCK(cudaMalloc(&h->act[7], (size_t)chunk * 3200 * sizeof(float)));   // encoder scratch (5*5*128)
    SRCK(cudaMalloc(&h->zin, (size_t)chunk * 192 * sizeof(float)));   // (chunk,128) dense scratch + (chunk,50) latents
    SRCK(cudaMalloc(&h->xin, (size_t)chunk * 1600 * sizeof(float)));
    for (int i = 0; i < 6; ++i) { cudaFree(h->actbf[i]); SRCK(cudaMalloc(&h->actbf[i], (size_t)chunk * DEC_ACT_ELEMS[i] * sizeof(__nv_bfloat16))); }
    if (!h->tc_err) { SRCK(cudaMalloc(&h->tc_err, sizeof(int))); SRCK(cudaMemset(h->tc_err, 0, sizeof(int))); }
    h->chunk = chunk;
    return SRCFD_OK;
}
inline int nblk(long long n) { return (int)((n + 255) / 256); }

// encoder on `B` samples: x_dev (B,10,10,1) -> z_dev (B,50)
int run_encoder(srcfd_sr* h, const float* x_dev, int B, float* z_dev) {
    float* a0 = h->xin;          // (B,5,5,64)
    float* a1 = h->act[7];       // (B,5,5,128)
    float* a2 = h->zin;          // (B,128)
    k_conv2d<float><<<nblk((long long)B * 25 * 64), 256, 0, h->stream>>>(x_dev, h->enc[0].W, h->enc[0].b, a0, B, 10, 10, 1, 5, 5, 64, 3, 3, 2, 0, 0, 1);
    k_conv2d<float><<<nblk((long long)B * 25 * 128), 256, 0, h->stream>>>(a0, h->enc[1].W, h->enc[1].b, a1, B, 5, 5, 64, 5, 5, 128, 3, 3, 1, 1, 1, 1);
    k_dense<float><<<nblk((long long)B * 128), 256, 0, h->stream>>>(a1, h->enc[2].W, h->enc[2].b, a2, B, 3200, 128, 1);
    k_dense<float><<<nblk((long long)B * 50), 256, 0, h->stream>>>(a2, h->enc[3].W, h->enc[3].b, z_dev, B, 128, 50, 0);
    h->launches += 4;
    SRCK(cudaGetLastError());
    return SRCFD_OK;
}
template <int KD, int ND>
int launch_convT_tc(srcfd_sr* h, const __nv_bfloat16* in, const __nv_bfloat16* Wbf, const float* bias, __nv_bfloat16* out,
                    int B, int H) {
    const long long M = (long long)B * H * H;
    const size_t smem = srtc::convT_tc_smem<KD, ND>();
    static bool attr_done[64] = {false};                 // per device: the attribute belongs to the (device, kernel) pair
    if (!attr_done[h->dev & 63]) { SRCK(cudaFuncSetAttribute(srtc::k_convT2x2_tc<KD, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[h->dev & 63] = true; }
    // The kernel walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... with the weights staged once per CTA.
    // SRCFD_TC_PERSIST = CTAs per SM of the persistent launch (capped by the 512 TMEM columns per SM); 0 = one tile per CTA.
    static int sms_dev[64] = {0};
    if (!sms_dev[h->dev & 63]) SRCK(cudaDeviceGetAttribute(&sms_dev[h->dev & 63], cudaDevAttrMultiProcessorCount, h->dev));
    const long long ntiles = (M + 127) / 128;
    const int per_sm = std::min(h->tc_persist, 512 / (ND < 32 ? 32 : ND));
    const unsigned grid = per_sm > 0 ? (unsigned)std::min<long long>(ntiles, (long long)per_sm * sms_dev[h->dev & 63]) : (unsigned)ntiles;
    srtc::k_convT2x2_tc<KD, ND><<<grid, 128, smem, h->stream>>>(in, Wbf, bias, out, M, H, H, h->tc_err);
    h->launches += 1;
    SRCK(cudaGetLastError());
    return SRCFD_OK;
}
// split-operand (bf16 x 3) ConvT layer: fp32 activation in and out
template <int KD, int ND, int NTv = 128>
int launch_convT_tc3(srcfd_sr* h, const float* in, const Layer& L, float* out, int B, int H) {
    const long long M = (long long)B * H * H;
    const size_t smem = srtc::convT_tc3_smem<KD, ND, 1>();
    constexpr int NT = NTv;
    static bool attr_done[64] = {false};
    if (!attr_done[h->dev & 63]) { SRCK(cudaFuncSetAttribute(srtc::k_convT2x2_tc3<KD, ND, 0, 1, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[h->dev & 63] = true; }
    static int sms_dev[64] = {0};
    if (!sms_dev[h->dev & 63]) SRCK(cudaDeviceGetAttribute(&sms_dev[h->dev & 63], cudaDevAttrMultiProcessorCount, h->dev));
    const long long ntiles = (M + 127) / 128;
    const int per_sm = std::max(1, std::min(std::min(h->tc_persist, 512 / (ND < 32 ? 32 : ND)), (int)((size_t)220 * 1024 / smem)));
    const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)per_sm * sms_dev[h->dev & 63]);
    srtc::k_convT2x2_tc3<KD, ND, 0, 1, NT><<<grid, NT, smem, h->stream>>>(in, L.Wbf, L.Wlo, L.b, out, M, H, H, h->tc_err);
    h->launches += 1;
    SRCK(cudaGetLastError());
    return SRCFD_OK;
}

// tensor-core ConvT layer l (1..4): input H = 25 * 2^(l-1), Cin = 128 >> (l-1)
int run_convT_tc(srcfd_sr* h, int l, const __nv_bfloat16* in, __nv_bfloat16* out, int B) {
    switch (l) {
        case 1: return launch_convT_tc<128, 256>(h, in, h->dec[2].Wbf, h->dec[2].b, out, B, 25);
        case 2: return launch_convT_tc<64, 128>(h, in, h->dec[3].Wbf, h->dec[3].b, out, B, 50);
        case 3: return launch_convT_tc<32, 64>(h, in, h->dec[4].Wbf, h->dec[4].b, out, B, 100);
        case 4: return launch_convT_tc<16, 32>(h, in, h->dec[5].Wbf, h->dec[5].b, out, B, 200);
    }
    return sr_fail(SRCFD_ERR_ARG, "bad tensor-core layer");
}

// decoder on `B` samples: z_dev (B,50) -> out_dev (B,400,400,1)
int run_decoder(srcfd_sr* h, const float* z_dev, int B, float* out_dev) {
    {
        const dim3 gd(36864 / 256, (B + 7) / 8);
        if (h->precision == 0 || h->precision == 3) k_dense_bt<float, 8><<<gd, 256, 0, h->stream>>>(z_dev, h->dec[0].W, h->dec[0].b, h->act[0], B, 50, 36864, 1);
        else k_dense_bt<__nv_bfloat16, 8><<<gd, 256, 0, h->stream>>>(z_dev, h->dec[0].W, h->dec[0].b, h->actbf[0], B, 50, 36864, 1);
    }
    const int hw[6] = {12, 25, 50, 100, 200, 400}, ch[6] = {256, 128, 64, 32, 16, 8};
    h->launches += 1;
    if (h->precision == 0) {
        for (int l = 0; l < 5; ++l) {
            const int k = (l == 0) ? 3 : 2;
            k_conv2d_transpose<<<nblk((long long)B * hw[l + 1] * hw[l + 1] * ch[l + 1]), 256, 0, h->stream>>>(
                h->act[l], h->dec[l + 1].W, h->dec[l + 1].b, h->act[l + 1], B, hw[l], hw[l], ch[l], hw[l + 1], hw[l + 1], ch[l + 1], k, 2, 1);
        }
        k_conv3x3_c8_final<float><<<dim3((400 + FC_TX - 1) / FC_TX, (400 + FC_TY - 1) / FC_TY, B), 256, 0, h->stream>>>(h->act[5], h->fcw, out_dev, B, 400, 400);
        h->launches += 6;
    } else if (h->precision == 3) {
        // split-operand tensor-core path: fp32 activations end to end, every ConvT on tcgen05 as three bf16 MMAs per K-step
        {
            const long long M = (long long)B * 144;
            if (h->l1_one_cta) {
                // one CTA per 128-row tile: A staged once, the nine taps' weights streamed by the TMA engine
                const size_t smem = srtc::l1_smem();
                static bool attr1_done[64] = {false};
                if (!attr1_done[h->dev & 63]) { SRCK(cudaFuncSetAttribute(srtc::k_convT3x3_l1_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr1_done[h->dev & 63] = true; }
                srtc::k_convT3x3_l1_tc3<<<(unsigned)((M + 127) / 128), 128, smem, h->stream>>>(h->act[0], h->dec[1].Wimg, h->act[5], M, h->tc_err);
            } else {
            const size_t smem = srtc::convT_tc3_smem<256, 128, 2>();
            static bool attr_done[64] = {false};
            if (!attr_done[h->dev & 63]) { SRCK(cudaFuncSetAttribute(srtc::k_convT2x2_tc3<256, 128, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[h->dev & 63] = true; }
            // CTAs per tap: each stages its tap's 128 KB of split weights once and then walks its share of the 128-row tiles
            const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((M + 127) / 128, h->tc3_l1_ctas));
            srtc::k_convT2x2_tc3<256, 128, 1, 2><<<dim3(gx, 9), 128, smem, h->stream>>>(
                h->act[0], h->dec[1].Wbf, h->dec[1].Wlo, h->dec[1].b, nullptr, M, 12, 12, h->tc_err, h->act[5], 1152);
            }
            srtc::k_col2im_3x3s2_f32<<<nblk((long long)B * 25 * 25 * 128), 256, 0, h->stream>>>(h->act[5], h->dec[1].b, h->act[1], B);
            h->launches += 2;
        }
        if (h->tc3_wide256) {      // 256 threads: a second warpgroup for staging and for half of each accumulator row's columns
            if (int rc = launch_convT_tc3<128, 256, 256>(h, h->act[1], h->dec[2], h->act[2], B, 25)) return rc;
            if (int rc = launch_convT_tc3<64, 128, 256>(h, h->act[2], h->dec[3], h->act[3], B, 50)) return rc;
        } else {
            if (int rc = launch_convT_tc3<128, 256>(h, h->act[1], h->dec[2], h->act[2], B, 25)) return rc;
            if (int rc = launch_convT_tc3<64, 128>(h, h->act[2], h->dec[3], h->act[3], B, 50)) return rc;
        }
        if (int rc = launch_convT_tc3<32, 64>(h, h->act[3], h->dec[4], h->act[4], B, 100)) return rc;
        if (h->tail_fused) {
            // last ConvT + final conv in one kernel: the (B, 400, 400, 8) activation between them stays in shared memory
            srtc::FinalW fw;
            memcpy(fw.w, h->fcw.w, sizeof(fw.w)); fw.b = h->fcw.b;
            const size_t smem = srtc::tail_fused_smem();
            static bool attr_done[64] = {false};
            if (!attr_done[h->dev & 63]) { SRCK(cudaFuncSetAttribute(srtc::k_tail_fused_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[h->dev & 63] = true; }
            const unsigned np = (200 + srtc::TF_OWN - 1) / srtc::TF_OWN;
            srtc::k_tail_fused_tc3<<<dim3(np, np, B), 128, smem, h->stream>>>(h->act[4], h->dec[5].Wbf, h->dec[5].Wlo, h->dec[5].b, fw, out_dev, 200, 200, h->tc_err);
            h->launches += 1;
            SRCK(cudaGetLastError());
        } else {
            if (int rc = launch_convT_tc3<16, 32>(h, h->act[4], h->dec[5], h->act[5], B, 200)) return rc;
            k_conv3x3_c8_final<float><<<dim3((400 + FC_TX - 1) / FC_TX, (400 + FC_TY - 1) / FC_TY, B), 256, 0, h->stream>>>(h->act[5], h->fcw, out_dev, B, 400, 400);
            h->launches += 1;
        }
    } else {
        // ConvT1 (3x3, stride 2: overlapping taps): tensor-core GEMM per tap into Y (act[5] reused as fp32 scratch,
        // B*144 x 1152), then col2im + bias + swish -> bf16 activation
        {
            const long long M = (long long)B * 144;
            const size_t smem = srtc::convT_tc_smem<256, 128>();
            static bool attr_done[64] = {false};
            if (!attr_done[h->dev & 63]) { SRCK(cudaFuncSetAttribute(srtc::k_convT2x2_tc<256, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[h->dev & 63] = true; }
            srtc::k_convT2x2_tc<256, 128, 1><<<dim3((unsigned)((M + 127) / 128), 9), 128, smem, h->stream>>>(
                h->actbf[0], h->dec[1].Wbf, h->dec[1].b, nullptr, M, 12, 12, h->tc_err, h->act[5], 1152);
            srtc::k_col2im_3x3s2<<<nblk((long long)B * 25 * 25 * 128), 256, 0, h->stream>>>(h->act[5], h->dec[1].b, h->actbf[1], B);
        }
        h->launches += 2;
        for (int l = 1; l <= 4; ++l)
            if (int rc = run_convT_tc(h, l, h->actbf[l], h->actbf[l + 1], B)) return rc;
        if (h->final_tc) {
            srtc::FinalW fw;
            memcpy(fw.w, h->fcw.w, sizeof(fw.w)); fw.b = h->fcw.b;
            const int R = h->final_tc_rows;
            const dim3 grid((400 + 127) / 128, (400 + R - 1) / R, B);
            if (R == 4) srtc::k_conv3x3_c8_final_tc<4><<<grid, 128, srtc::ft_smem<4>(), h->stream>>>(h->actbf[5], fw, out_dev, 400, 400, h->tc_err);
            else if (R == 16) srtc::k_conv3x3_c8_final_tc<16><<<grid, 128, srtc::ft_smem<16>(), h->stream>>>(h->actbf[5], fw, out_dev, 400, 400, h->tc_err);
            else srtc::k_conv3x3_c8_final_tc<8><<<grid, 128, srtc::ft_smem<8>(), h->stream>>>(h->actbf[5], fw, out_dev, 400, 400, h->tc_err);
        } else {
            k_conv3x3_c8_final<__nv_bfloat16><<<dim3((400 + FC_TX - 1) / FC_TX, (400 + FC_TY - 1) / FC_TY, B), 256, 0, h->stream>>>(h->actbf[5], h->fcw, out_dev, B, 400, 400);
        }
        h->launches += 1;
    }
    SRCK(cudaGetLastError());
    return SRCFD_OK;
}

}  // namespace

extern "C" {

const char* srcfd_sr_last_error(void) { return g_sr_err.c_str(); }

int srcfd_sr_create(int device, srcfd_sr** out) {
    if (!out) return sr_fail(SRCFD_ERR_ARG, "null out");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return sr_fail(SRCFD_ERR_CUDA, "no CUDA device: libsrcfd has no CPU fallback");
    if (device < 0 || device >= ndev) return sr_fail(SRCFD_ERR_ARG, "bad device ordinal");
    srcfd_sr* h = new srcfd_sr();
    h->dev = device;
    SRCK(cudaSetDevice(device));
    SRCK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    SRCK(cudaEventCreate(&h->ea)); SRCK(cudaEventCreate(&h->eb));
    if (const char* e = getenv("SRCFD_TC_PERSIST")) h->tc_persist = atoi(e);
    if (const char* e = getenv("SRCFD_FINAL_TC")) h->final_tc = atoi(e);
    if (const char* e = getenv("SRCFD_TAIL_FUSED")) h->tail_fused = h->tail_fused_default = atoi(e);
    if (const char* e = getenv("SRCFD_TC3_L1_CTAS")) h->tc3_l1_ctas = std::max(1, atoi(e));
    if (const char* e = getenv("SRCFD_TC3_WIDE256")) h->tc3_wide256 = atoi(e);
    if (const char* e = getenv("SRCFD_L1_ONE_CTA")) h->l1_one_cta = h->l1_one_cta_default = atoi(e);
    if (const char* e = getenv("SRCFD_FINAL_TC_ROWS")) h->final_tc_rows = atoi(e);
    *out = h;
    return SRCFD_OK;
}

int srcfd_sr_destroy(srcfd_sr* h) {
    if (!h) return SRCFD_OK;
    cudaSetDevice(h->dev);
    cudaStreamSynchronize(h->stream);
    for (auto& l : h->enc) { cudaFree(l.W); cudaFree(l.b); }
    for (auto& l : h->dec) { cudaFree(l.W); cudaFree(l.b); cudaFree(l.Wbf); cudaFree(l.Wlo); cudaFree(l.Wimg); }
    for (int i = 0; i < 6; ++i) cudaFree(h->actbf[i]);
    cudaFree(h->tc_err); cudaFree(h->stats_dev);
    for (int i = 0; i < 8; ++i) cudaFree(h->act[i]);
    cudaFree(h->zin); cudaFree(h->xin);
    cudaEventDestroy(h->ea); cudaEventDestroy(h->eb);
    cudaStreamDestroy(h->stream);
    delete h;
    return SRCFD_OK;
}

// kernels/biases in Keras layouts: conv2d (3,3,1,64), conv2d_1 (3,3,64,128), dense (3200,128), latent_vector (128,50)
int srcfd_sr_set_encoder(srcfd_sr* h, const float* const kernels[4], const float* const biases[4]) {
    if (!h || !kernels || !biases) return sr_fail(SRCFD_ERR_ARG, "null argument");
    SRCK(cudaSetDevice(h->dev));
    const size_t kn[4] = {3 * 3 * 1 * 64, 3 * 3 * 64 * 128, 3200 * 128, 128 * 50}, bn[4] = {64, 128, 128, 50};
    for (int i = 0; i < 4; ++i) {
        if (int rc = upload(&h->enc[i].W, kernels[i], kn[i], h->stream)) return rc;
        if (int rc = upload(&h->enc[i].b, biases[i], bn[i], h->stream)) return rc;
    }
    SRCK(cudaStreamSynchronize(h->stream));
    h->has_enc = true;
    return SRCFD_OK;
}

// kernels/biases in Keras layouts: dense (50,36864); conv2d_transpose (3,3,128,256), _1 (2,2,64,128), _2 (2,2,32,64),
// _3 (2,2,16,32), _4 (2,2,8,16); output_image_400 (3,3,8,1)
int srcfd_sr_set_decoder(srcfd_sr* h, const float* const kernels[7], const float* const biases[7]) {
    if (!h || !kernels || !biases) return sr_fail(SRCFD_ERR_ARG, "null argument");
    SRCK(cudaSetDevice(h->dev));
    const int cin[5] = {256, 128, 64, 32, 16}, cout[5] = {128, 64, 32, 16, 8}, taps[5] = {9, 4, 4, 4, 4};
    if (int rc = upload(&h->dec[0].W, kernels[0], (size_t)50 * 36864, h->stream)) return rc;
    if (int rc = upload(&h->dec[0].b, biases[0], 36864, h->stream)) return rc;
    for (int l = 0; l < 5; ++l) {
        std::vector<float> wt = transpose_last2(kernels[l + 1], taps[l], cout[l], cin[l]);
        if (int rc = upload(&h->dec[l + 1].W, wt.data(), wt.size(), h->stream)) return rc;
        SRCK(cudaStreamSynchronize(h->stream));      // wt is a temporary
        if (int rc = upload(&h->dec[l + 1].b, biases[l + 1], cout[l], h->stream)) return rc;
        {                 // tensor-core operand: the Keras layout (ky,kx,co | ci) is already the K-major (N, K) matrix
            const size_t n = (size_t)taps[l] * cout[l] * cin[l];
            std::vector<__nv_bfloat16> wb(n), wl(n);
            for (size_t i = 0; i < n; ++i) {
                wb[i] = __float2bfloat16(kernels[l + 1][i]);
                wl[i] = __float2bfloat16(kernels[l + 1][i] - __bfloat162float(wb[i]));      // the part the first bf16 dropped
            }
            if (!h->dec[l + 1].Wbf) SRCK(cudaMalloc(&h->dec[l + 1].Wbf, n * sizeof(__nv_bfloat16)));
            if (!h->dec[l + 1].Wlo) SRCK(cudaMalloc(&h->dec[l + 1].Wlo, n * sizeof(__nv_bfloat16)));
            SRCK(cudaMemcpy(h->dec[l + 1].Wbf, wb.data(), n * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
            SRCK(cudaMemcpy(h->dec[l + 1].Wlo, wl.data(), n * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
            if (l == 0) {     // the 3x3 layer's weights once more, laid out as the stages the TMA engine copies (sr_tc.cuh)
                std::vector<__nv_bfloat16> img(2 * n);
                for (int tap = 0; tap < 9; ++tap)
                    for (int nn = 0; nn < 128; ++nn)
                        for (int k = 0; k < 256; ++k) {
                            const size_t src = ((size_t)tap * 128 + nn) * 256 + k;
                            img[srtc::l1_img_index(tap, nn, k, 0)] = wb[src];
                            img[srtc::l1_img_index(tap, nn, k, 1)] = wl[src];
                        }
                if (!h->dec[1].Wimg) SRCK(cudaMalloc(&h->dec[1].Wimg, img.size() * sizeof(__nv_bfloat16)));
                SRCK(cudaMemcpy(h->dec[1].Wimg, img.data(), img.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
            }
        }
    }
    if (int rc = upload(&h->dec[6].W, kernels[6], 3 * 3 * 8 * 1, h->stream)) return rc;
    if (int rc = upload(&h->dec[6].b, biases[6], 1, h->stream)) return rc;
    memcpy(h->fcw.w, kernels[6], 72 * sizeof(float)); h->fcw.b = biases[6][0];
    SRCK(cudaStreamSynchronize(h->stream));
    h->has_dec = true;
    return SRCFD_OK;
}

static int sr_run(srcfd_sr* h, const float* x, const float* z, int B, float* zout, float* out) {
    SRCK(cudaSetDevice(h->dev));
    const int CH = std::min(B, 128);     // samples per pass: large enough that the early (small) layers fill the GPU
    if (int rc = ensure_chunk(h, CH)) return rc;
    float* zdev = h->zin + (size_t)h->chunk * 128;   // latents (CH,50) live behind the (chunk,128) dense scratch
    for (int b0 = 0; b0 < B; b0 += CH) {
        const int nb = std::min(CH, B - b0);
        if (x) {
            float* xdev = h->act[6];              // reuse the output buffer as input staging (consumed before it is written)
            SRCK(cudaMemcpyAsync(xdev, x + (size_t)b0 * 100, (size_t)nb * 100 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            if (int rc = run_encoder(h, xdev, nb, zdev)) return rc;
        } else {
            SRCK(cudaMemcpyAsync(zdev, z + (size_t)b0 * 50, (size_t)nb * 50 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        }
        if (zout) SRCK(cudaMemcpyAsync(zout + (size_t)b0 * 50, zdev, (size_t)nb * 50 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (out) {
            if (int rc = run_decoder(h, zdev, nb, h->act[6])) return rc;
            SRCK(cudaMemcpyAsync(out + (size_t)b0 * 160000, h->act[6], (size_t)nb * 160000 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        }
        SRCK(cudaStreamSynchronize(h->stream));
    }
    return SRCFD_OK;
}

// encoder_10.predict: x (B,10,10,1) host -> z (B,50) host
int srcfd_sr_encode(srcfd_sr* h, const float* x, int B, float* z) {
    if (!h || !x || !z || B < 1) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    if (!h->has_enc) return sr_fail(SRCFD_ERR_ARG, "encoder weights not set");
    return sr_run(h, x, nullptr, B, z, nullptr);
}
// decoder_400.predict: z (B,50) host -> out (B,400,400,1) host
int srcfd_sr_decode(srcfd_sr* h, const float* z, int B, float* out) {
    if (!h || !z || !out || B < 1) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    if (!h->has_dec) return sr_fail(SRCFD_ERR_ARG, "decoder weights not set");
    return sr_run(h, nullptr, z, B, nullptr, out);
}
// SuperResolutionAE.call (PyCFD_ML_accelerated.py:686-689): x (B,10,10,1) host -> (B,400,400,1) host
int srcfd_sr_predict(srcfd_sr* h, const float* x, int B, float* out) {
    if (!h || !x || !out || B < 1) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    if (!h->has_enc || !h->has_dec) return sr_fail(SRCFD_ERR_ARG, "encoder/decoder weights not set");
    return sr_run(h, x, nullptr, B, nullptr, out);
}
// ml_super_resolution's per-field pipeline for B fields in ONE call, everything between the two host copies on the device:
// [adaptive statistics] -> standardize -> encoder_10 -> decoder_400 -> inverse standardize -> NaN/Inf guard.
int srcfd_sr_super_resolve(srcfd_sr* h, const float* x, int B, const double* stats, int adaptive, double blend, float* out) {
    if (!h || !x || !out || !stats || B < 1) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    if (!h->has_enc || !h->has_dec) return sr_fail(SRCFD_ERR_ARG, "encoder/decoder weights not set");
    SRCK(cudaSetDevice(h->dev));
    const int CH = std::min(B, 128);
    if (int rc = ensure_chunk(h, CH)) return rc;
    if (h->stats_cap < CH) {
        cudaFree(h->stats_dev); h->stats_dev = nullptr; h->stats_cap = 0;
        SRCK(cudaMalloc(&h->stats_dev, sizeof(double) * 4 * (size_t)CH));
        h->stats_cap = CH;
    }
    float* zdev = h->zin + (size_t)h->chunk * 128;
    for (int b0 = 0; b0 < B; b0 += CH) {
        const int nb = std::min(CH, B - b0);
        float* xraw = h->act[6];                     // staging: raw coarse fields, then (behind them) the standardized ones
        float* xstd = h->act[6] + (size_t)nb * 100;
        SRCK(cudaMemcpyAsync(xraw, x + (size_t)b0 * 100, (size_t)nb * 100 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        SRCK(cudaMemcpyAsync(h->stats_dev, stats + (size_t)b0 * 4, (size_t)nb * 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        k_sr_pre<<<nb, 128, 0, h->stream>>>(xraw, h->stats_dev, adaptive, blend, xstd);
        h->launches += 1;
        if (int rc = run_encoder(h, xstd, nb, zdev)) return rc;
        if (int rc = run_decoder(h, zdev, nb, h->act[6])) return rc;
        const long long n = (long long)nb * 160000;
        k_sr_post<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, h->stream>>>(h->act[6], h->stats_dev, 160000, n);
        h->launches += 1;
        SRCK(cudaGetLastError());
        SRCK(cudaMemcpyAsync(out + (size_t)b0 * 160000, h->act[6], (size_t)nb * 160000 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        SRCK(cudaStreamSynchronize(h->stream));
    }
    return SRCFD_OK;
}
// Throughput entry: latents and outputs resident in HBM (device pointers), whole batch, CUDA-event timed.
int srcfd_sr_decode_device(srcfd_sr* h, uint64_t z_dev, int B, uint64_t out_dev, double* ms) {
    if (!h || !z_dev || !out_dev || B < 1) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    if (!h->has_dec) return sr_fail(SRCFD_ERR_ARG, "decoder weights not set");
    SRCK(cudaSetDevice(h->dev));
    const int CH = std::min(B, 128);     // samples per pass: large enough that the early (small) layers fill the GPU
    if (int rc = ensure_chunk(h, CH)) return rc;
    SRCK(cudaEventRecord(h->ea, h->stream));
    for (int b0 = 0; b0 < B; b0 += CH) {
        const int nb = std::min(CH, B - b0);
        if (int rc = run_decoder(h, (const float*)(uintptr_t)z_dev + (size_t)b0 * 50, nb, (float*)(uintptr_t)out_dev + (size_t)b0 * 160000)) return rc;
    }
    SRCK(cudaEventRecord(h->eb, h->stream));
    SRCK(cudaEventSynchronize(h->eb));
    float f = 0.f;
    SRCK(cudaEventElapsedTime(&f, h->ea, h->eb));
    if (ms) *ms = f;
    return SRCFD_OK;
}
// 0 = fp32 CUDA cores (default, parity path); 1 = bf16 tcgen05 tensor cores for the four 2x2/stride-2 ConvT layers
int srcfd_sr_set_precision(srcfd_sr* h, int mode) {
    if (!h || mode < 0 || mode > 4) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    h->precision = mode == 0 ? 0 : mode >= 3 ? 3 : 1;
    if (mode == 3) { h->tail_fused = h->tail_fused_default; h->l1_one_cta = h->l1_one_cta_default; }
    if (mode == 4) { h->tail_fused = 0; h->l1_one_cta = 0; }   // the round-1 kernels for the 3x3 ConvT and the tail (same bits; parity tests)
    if (mode == 1) h->final_tc = 1;        // mode 2: bf16 layers with the final conv on the CUDA-core tile kernel (parity tests)
    if (mode == 2) h->final_tc = 0;
    return SRCFD_OK;
}
// 1 if a tensor-core kernel ever timed out waiting for its accumulator (diagnostic; the kernels never hang)
int srcfd_sr_tc_error(srcfd_sr* h, int* flag) {
    if (!h || !flag) return sr_fail(SRCFD_ERR_ARG, "null argument");
    *flag = 0;
    if (h->tc_err) { SRCK(cudaSetDevice(h->dev)); SRCK(cudaMemcpy(flag, h->tc_err, sizeof(int), cudaMemcpyDeviceToHost)); }
    return SRCFD_OK;
}
// One tensor-core ConvT layer in isolation (layer 1..4 = conv2d_transpose_1.._4): in (B,H,H,Cin) fp32 host ->
// out (B,2H,2H,Cout) fp32 host, operands rounded to bf16 on the way in, bf16 result widened on the way out.
int srcfd_sr_debug_convT_tc(srcfd_sr* h, int layer, const float* in, int B, float* out) {
    if (!h || !in || !out || layer < 1 || layer > 4 || B < 1) return sr_fail(SRCFD_ERR_ARG, "bad argument");
    if (!h->has_dec) return sr_fail(SRCFD_ERR_ARG, "decoder weights not set");
    SRCK(cudaSetDevice(h->dev));
    if (int rc = ensure_chunk(h, std::max(B, 1))) return rc;
    const long long nin = (long long)B * DEC_ACT_ELEMS[layer], nout = (long long)B * DEC_ACT_ELEMS[layer + 1];
    SRCK(cudaMemcpyAsync(h->act[layer], in, nin * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    srtc::k_f32_to_bf16<<<1024, 256, 0, h->stream>>>(h->act[layer], h->actbf[layer], nin);
    if (int rc = run_convT_tc(h, layer, h->actbf[layer], h->actbf[layer + 1], B)) return rc;
    srtc::k_bf16_to_f32<<<1024, 256, 0, h->stream>>>(h->actbf[layer + 1], h->act[layer + 1], nout);
    SRCK(cudaMemcpyAsync(out, h->act[layer + 1], nout * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    SRCK(cudaStreamSynchronize(h->stream));
    return SRCFD_OK;
}
int srcfd_sr_launch_count(srcfd_sr* h, int64_t* n) {
    if (!h || !n) return sr_fail(SRCFD_ERR_ARG, "null argument");
    *n = h->launches;
    return SRCFD_OK;
}

}  // extern "C"
