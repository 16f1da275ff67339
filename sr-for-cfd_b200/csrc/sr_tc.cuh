// sr_tc.cuh -- tensor-core (tcgen05 / TMEM) path for the decoder's stride-2, 2x2 transposed convolutions.
//
// Conv2DTranspose(Cout, 2x2, stride 2, 'valid') has non-overlapping taps, so it IS a GEMM:
//     D[m, n] = sum_ci A[m, ci] * Wk[n, ci]      m = (b, y, x) input pixel,  n = (dy, dx, co)
//     out[b, 2y+dy, 2x+dx, co] = swish(D[m, n] + bias[co])
// A is the NHWC activation itself (M x K, K = Cin contiguous => "K-major", no im2col: implicit GEMM); Wk is the
// Keras kernel (kh, kw, Cout, Cin) read as (N = 4*Cout) x K, also K-major.  One CTA computes a 128 x N tile:
//   * 128 threads stage the A tile and the whole weight matrix in shared memory in the canonical K-major
//     SWIZZLE_NONE layout of the UMMA shared-memory descriptor: 8-row x 16-byte core matrices, 128 B each,
//     consecutive 8-row groups SBO = 128 B apart, consecutive 16-byte K-chunks LBO = (rows/8)*128 B apart;
//   * one thread issues K/16 tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128, N = 4*Cout) into a TMEM
//     accumulator allocated by warp 0, then tcgen05.commit -> mbarrier;
//   * each warp reads its 32 TMEM lanes with tcgen05.ld (32x32b.x16), adds the bias, applies swish and writes
//     the pixel-shuffled bf16 output (for a fixed row and dy the (dx, co) run is contiguous in NHWC).
// K = Cin <= 128 fits one stage, so there is no K pipeline inside a CTA; several CTAs per SM overlap instead.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace srtc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);              // start address  [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;     // leading byte offset [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (sm_100)
    return d;                                              // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor for kind::f16: bf16 x bf16 -> f32, both operands K-major (InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// swish(x) = x * sigmoid(x) with sigmoid(x) = 0.5 + 0.5 * tanh(x / 2): one MUFU (tanh.approx.f32, relative error 2^-11) and
// three FMA-pipe operations instead of exp + full-precision division (2 MUFU + ~10).  The result is rounded to bf16
// (2^-8) right after, so the approximation is invisible at the output precision of this path; the fp32 path keeps expf.
__device__ __forceinline__ float swishf(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return x * fmaf(0.5f, t, 0.5f);
}

// A: (M, KD) bf16 row-major.  Wk: (ND, KD) bf16 row-major (ND = 4*Cout, n = (dy*2+dx)*Cout + co).  bias: (Cout) fp32.
// out: (B, 2H, 2W, Cout) bf16.  M = B*H*W.  err: set to 1 if the MMA completion wait times out (never hang).
// EPI = 0: the fused bias + swish + pixel-shuffle epilogue described above (out = bf16 NHWC activation).
// EPI = 1: plain GEMM tile, fp32 D stored to Y[m * ldy + blockIdx.y * ND + n]; Wk is offset by blockIdx.y * ND rows.
//          Used for the 3x3/stride-2 ConvT (overlapping taps): Y holds the 9 per-tap products, k_col2im_3x3s2 sums them.
template <int KD, int ND, int EPI = 0>
__global__ void __launch_bounds__(128) k_convT2x2_tc(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Wk,
                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                      long long M, int H, int Wd, int* err, float* __restrict__ Y = nullptr,
                                                      int ldy = 0) {
    constexpr int MT = 128, COUT = ND / 4;
    if (EPI == 1) Wk += (size_t)blockIdx.y * ND * KD;
    constexpr uint32_t LBO_A = (MT / 8) * 128, LBO_B = (ND / 8) * 128, SBO = 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;                              // MT x KD bf16, canonical layout
    unsigned char* sB = smem_raw + (size_t)MT * KD * 2;        // ND x KD bf16
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[COUT];               // read as float4 broadcasts by the epilogue
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (EPI == 0) for (int c = tid; c < COUT; c += 128) s_bias[c] = bias[c];

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(ND < 32 ? 32 : ND) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- the weights are staged ONCE per CTA; the CTA then walks its tiles (persistent over blockIdx.x):
    // one 16-byte chunk = 8 bf16 of one row
    constexpr int KC = KD / 8;
    for (int c = tid; c < ND * KC; c += 128) {
        const int n = c / KC, kc = c - n * KC;
        const uint4 v = *reinterpret_cast<const uint4*>(Wk + (size_t)n * KD + kc * 8);
        *reinterpret_cast<uint4*>(sB + (size_t)kc * LBO_B + (n >> 3) * SBO + (n & 7) * 16) = v;
    }
    const long long ntiles = (M + MT - 1) / MT;
    uint32_t phase = 0;
    const int OW = 2 * Wd;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, phase ^= 1u) {
        const long long m0 = tile * MT;
        // ---- stage A (zero rows past M)
        for (int c = tid; c < MT * KC; c += 128) {
            const int kc = c / MT, r = c - kc * MT;               // row fastest: consecutive lanes store consecutive 16-byte core-matrix rows (no bank conflicts)
            uint4 v = make_uint4(0, 0, 0, 0);
            if (m0 + r < M) v = *reinterpret_cast<const uint4*>(A + (m0 + r) * KD + kc * 8);
            *reinterpret_cast<uint4*>(sA + (size_t)kc * LBO_A + (r >> 3) * SBO + (r & 7) * 16) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                               // also: every warp has drained the previous tile's accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem = tmem_base_s;

        if (tid == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(MT, ND);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int k = 0; k < KD / 16; ++k) {                       // one MMA consumes K = 16 = two 16-byte chunks
                const uint64_t ad = make_smem_desc(a0 + (uint32_t)k * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bd = make_smem_desc(b0 + (uint32_t)k * 2 * LBO_B, LBO_B, SBO);
                umma_bf16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // ---- wait for the accumulator (bounded: never hang the GPU); the barrier's phase flips once per tile
        {
            uint32_t done = 0;
            for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
            }
            if (!done) { if (lane == 0) atomicExch(err, 1); break; }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: thread = accumulator row (TMEM lane) 32*warp + lane
        const long long m = m0 + warp * 32 + lane;
        const bool valid = m < M;
        long long pix = valid ? m : 0;
        const int x = (int)(pix % Wd); pix /= Wd;
        const int y = (int)(pix % H);
        const long long b = pix / H;
#pragma unroll 1
        for (int c0 = 0; c0 < ND; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (EPI == 1) {
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(Y + m * ldy + (long long)blockIdx.y * ND + c0);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
                }
            } else if (valid) {
                // columns c0..c0+15 : n = tap*COUT + co (COUT is a multiple of 8, so 8-column groups stay inside one tap)
#pragma unroll
                for (int g = 0; g < 16; g += 8) {
                    const int n = c0 + g, tap = n / COUT, co = n - tap * COUT;
                    const int dy = tap >> 1, dx = tap & 1;
                    const float4 b0 = *reinterpret_cast<const float4*>(s_bias + co), b1 = *reinterpret_cast<const float4*>(s_bias + co + 4);
                    __nv_bfloat162 o[4];                         // cvt.rn.bf16x2.f32: two values per conversion
                    o[0] = __floats2bfloat162_rn(swishf(__uint_as_float(v[g + 0]) + b0.x), swishf(__uint_as_float(v[g + 1]) + b0.y));
                    o[1] = __floats2bfloat162_rn(swishf(__uint_as_float(v[g + 2]) + b0.z), swishf(__uint_as_float(v[g + 3]) + b0.w));
                    o[2] = __floats2bfloat162_rn(swishf(__uint_as_float(v[g + 4]) + b1.x), swishf(__uint_as_float(v[g + 5]) + b1.y));
                    o[3] = __floats2bfloat162_rn(swishf(__uint_as_float(v[g + 6]) + b1.z), swishf(__uint_as_float(v[g + 7]) + b1.w));
                    __nv_bfloat16* dst = out + ((((long long)b * (2 * H) + (2 * y + dy)) * OW + (2 * x + dx)) * COUT + co);
                    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // pairs with the fence after the next tile's barrier
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(ND < 32 ? 32 : ND) : "memory");
}

template <int KD, int ND>
constexpr size_t convT_tc_smem() { return (size_t)(128 + ND) * KD * 2; }

// ---- split-operand variant (round 2): fp32 accuracy from the bf16 tensor cores ---------------------------------------------
// a = a_hi + a_lo, w = w_hi + w_lo with every part a bf16: a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (the dropped a_lo*w_lo and
// the parts' own rounding are ~2^-16 of the product); three MMAs per K-step accumulate in the same fp32 TMEM accumulator.
// The activations stay fp32 in HBM (same traffic as the fp32 CUDA-core path): they are split while they are staged
// (hi = bf16(a), lo = bf16(a - hi)); the weights are split once at upload; the epilogue is fp32 (expf swish, fp32 store).
// KPART > 1 stages the A tile in K-parts (the 3x3 layer: K = 256 would need 256 KB for both operands' two halves); the
// weights stay resident.  Same tile / descriptor / epilogue structure as k_convT2x2_tc.
// expf + approximate division (MUFU.EX2, MUFU.RCP: ~2 ulp each) -- three orders of magnitude inside the 1e-4 bar of this path
// (round 2: ex2.approx.ftz / rcp.approx.ftz spelled out -- __expf / __fdividef wrap the same two MUFU operations in denormal and
// range fix-ups, 13.7 instructions per value in the fused tail's epilogue against 5 here; flushing matters only for |x| > 87,
// where the result is 0 or x to well inside the bar)
__device__ __forceinline__ float swish_exact(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    __nv_bfloat162 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        l[e] = __floats2bfloat162_rn(v[2 * e] - __low2float(h[e]), v[2 * e + 1] - __high2float(h[e]));
    }
    hi = *reinterpret_cast<const uint4*>(h); lo = *reinterpret_cast<const uint4*>(l);
}
// NT = 128 or 256 threads: with 256 the second warpgroup shares the staging and takes the upper half of every accumulator row's
// columns in the epilogue (a warp reads the TMEM lane quarter warp % 4), for the wide layers that fit one or three CTAs per SM.
template <int KD, int ND, int EPI, int KPART, int NT = 128>
__global__ void __launch_bounds__(NT) k_convT2x2_tc3(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Whi,
                                                       const __nv_bfloat16* __restrict__ Wlo, const float* __restrict__ bias,
                                                       float* __restrict__ out, long long M, int H, int Wd, int* err,
                                                       float* __restrict__ Y = nullptr, int ldy = 0) {
    constexpr int MT = 128, COUT = ND / 4, KP = KD / KPART, KCP = KP / 8, KC = KD / 8;
    if (EPI == 1) { Whi += (size_t)blockIdx.y * ND * KD; Wlo += (size_t)blockIdx.y * ND * KD; }
    constexpr uint32_t LBO_A = (MT / 8) * 128, LBO_B = (ND / 8) * 128, SBO = 128;
    constexpr uint32_t ABYTES = MT * KP * 2, BBYTES = ND * KD * 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sAh = smem_raw;
    unsigned char* sAl = smem_raw + ABYTES;
    unsigned char* sBh = smem_raw + 2 * (size_t)ABYTES;
    unsigned char* sBl = sBh + BBYTES;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[COUT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wq = warp & 3, whalf = warp >> 2;             // TMEM lane quarter / column half of this warp
    if (EPI == 0) for (int c = tid; c < COUT; c += NT) s_bias[c] = bias[c];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(ND < 32 ? 32 : ND) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = tid; c < ND * KC; c += NT) {                    // both halves of the weights, once per CTA
        const int n = c / KC, kc = c - n * KC;
        const size_t so = (size_t)kc * LBO_B + (n >> 3) * SBO + (n & 7) * 16;
        *reinterpret_cast<uint4*>(sBh + so) = *reinterpret_cast<const uint4*>(Whi + (size_t)n * KD + kc * 8);
        *reinterpret_cast<uint4*>(sBl + so) = *reinterpret_cast<const uint4*>(Wlo + (size_t)n * KD + kc * 8);
    }
    const long long ntiles = (M + MT - 1) / MT;
    uint32_t phase = 0;
    const int OW = 2 * Wd;
    bool dead = false;
    for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x) {
        const long long m0 = tile * MT;
#pragma unroll 1
        for (int part = 0; part < KPART && !dead; ++part, phase ^= 1u) {
            // ---- stage this K-part of the A tile, split into its two bf16 halves (zero rows past M)
            for (int c = tid; c < MT * KCP; c += NT) {
                const int kc = c / MT, r = c - kc * MT;           // row fastest: conflict-free shared-memory stores (kc fastest: all lanes on one bank)
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (m0 + r < M) {
                    const float4* src = reinterpret_cast<const float4*>(A + (m0 + r) * KD + part * KP + kc * 8);
                    const float4 a = src[0], b = src[1];
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
                uint4 hi, lo;
                split8(v, hi, lo);
                const size_t so = (size_t)kc * LBO_A + (r >> 3) * SBO + (r & 7) * 16;
                *reinterpret_cast<uint4*>(sAh + so) = hi;
                *reinterpret_cast<uint4*>(sAl + so) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();                                     // also: every warp has drained the previous tile's accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem = tmem_base_s;
            if (tid == 0) {
                constexpr uint32_t idesc = make_idesc_bf16(MT, ND);
                const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), bh = smem_u32(sBh), bl = smem_u32(sBl);
#pragma unroll
                for (int k = 0; k < KP / 16; ++k) {
                    const uint32_t ao = (uint32_t)k * 2 * LBO_A, bo = (uint32_t)(part * (KP / 16) + k) * 2 * LBO_B;
                    const uint64_t adh = make_smem_desc(ah + ao, LBO_A, SBO), adl = make_smem_desc(al + ao, LBO_A, SBO);
                    const uint64_t bdh = make_smem_desc(bh + bo, LBO_B, SBO), bdl = make_smem_desc(bl + bo, LBO_B, SBO);
                    umma_bf16(tmem, adh, bdh, idesc, (part > 0 || k > 0) ? 1u : 0u);
                    umma_bf16(tmem, adh, bdl, idesc, 1u);
                    umma_bf16(tmem, adl, bdh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
            {                                                    // bounded wait: never hang the GPU
                uint32_t done = 0;
                for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
                }
                if (!done) { if (lane == 0) atomicExch(err, 1); dead = true; }
            }
            dead = __syncthreads_or(dead ? 1 : 0) != 0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (dead) break;
        // ---- epilogue: thread = accumulator row (TMEM lane) 32*warp + lane
        const long long m = m0 + wq * 32 + lane;
        const bool valid = m < M;
        long long pix = valid ? m : 0;
        const int x = (int)(pix % Wd); pix /= Wd;
        const int y = (int)(pix % H);
        const long long b = pix / H;
        const uint32_t tmem = tmem_base_s;
#pragma unroll 1
        for (int c0 = whalf * (ND / (NT / 128)); c0 < (whalf + 1) * (ND / (NT / 128)); c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (EPI == 1) {
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(Y + m * ldy + (long long)blockIdx.y * ND + c0);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
                }
            } else if (valid) {
#pragma unroll
                for (int g = 0; g < 16; g += 8) {
                    const int n = c0 + g, tap = n / COUT, co = n - tap * COUT;
                    const int dy = tap >> 1, dx = tap & 1;
                    const float4 b0 = *reinterpret_cast<const float4*>(s_bias + co), b1 = *reinterpret_cast<const float4*>(s_bias + co + 4);
                    float4* dst = reinterpret_cast<float4*>(out + ((((long long)b * (2 * H) + (2 * y + dy)) * OW + (2 * x + dx)) * COUT + co));
                    dst[0] = make_float4(swish_exact(__uint_as_float(v[g + 0]) + b0.x), swish_exact(__uint_as_float(v[g + 1]) + b0.y),
                                         swish_exact(__uint_as_float(v[g + 2]) + b0.z), swish_exact(__uint_as_float(v[g + 3]) + b0.w));
                    dst[1] = make_float4(swish_exact(__uint_as_float(v[g + 4]) + b1.x), swish_exact(__uint_as_float(v[g + 5]) + b1.y),
                                         swish_exact(__uint_as_float(v[g + 6]) + b1.z), swish_exact(__uint_as_float(v[g + 7]) + b1.w));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(ND < 32 ? 32 : ND) : "memory");
}
template <int KD, int ND, int KPART>
constexpr size_t convT_tc3_smem() { return (size_t)2 * 128 * (KD / KPART) * 2 + (size_t)2 * ND * KD * 2; }

// col2im of the 3x3 / stride-2 layer with an fp32 activation out (split-operand path)
__global__ void k_col2im_3x3s2_f32(const float* __restrict__ Y, const float* __restrict__ bias, float* __restrict__ out, int B) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * 25 * 25 * 128) return;
    const int co = (int)(t & 127);
    long long p = t >> 7;
    const int X = (int)(p % 25); p /= 25;
    const int Yo = (int)(p % 25);
    const long long b = p / 25;
    float acc = bias[co];
    for (int ky = Yo & 1; ky < 3; ky += 2) {
        const int y = (Yo - ky) >> 1;
        if (y < 0 || y >= 12) continue;
        for (int kx = X & 1; kx < 3; kx += 2) {
            const int x = (X - kx) >> 1;
            if (x < 0 || x >= 12) continue;
            acc += Y[((b * 12 + y) * 12 + x) * 1152 + (ky * 3 + kx) * 128 + co];
        }
    }
    out[t] = swish_exact(acc);
}

// Conv2DTranspose(128, 3x3, stride 2, valid) from the per-tap products Y (B*144, 9*128): sum the <= 4 taps that hit
// each output pixel, add bias, swish, write the bf16 NHWC activation (B, 25, 25, 128).
__global__ void k_col2im_3x3s2(const float* __restrict__ Y, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * 25 * 25 * 128) return;
    const int co = (int)(t & 127);
    long long p = t >> 7;
    const int X = (int)(p % 25); p /= 25;
    const int Yo = (int)(p % 25);
    const long long b = p / 25;
    float acc = bias[co];
    for (int ky = Yo & 1; ky < 3; ky += 2) {
        const int y = (Yo - ky) >> 1;
        if (y < 0 || y >= 12) continue;
        for (int kx = X & 1; kx < 3; kx += 2) {
            const int x = (X - kx) >> 1;
            if (x < 0 || x >= 12) continue;
            acc += Y[((b * 12 + y) * 12 + x) * 1152 + (ky * 3 + kx) * 128 + co];
        }
    }
    out[t] = __float2bfloat16(swishf(acc));
}

// fp32 -> bf16 elementwise
__global__ void k_f32_to_bf16(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        out[t] = __float2bfloat16(in[t]);
}
__global__ void k_bf16_to_f32(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        out[t] = __bfloat162float(in[t]);
}

// ---- final Conv2D(1, 3x3, 'same', linear) over the 8-channel bf16 activation, on the tensor cores ------------------
// With Cin = 8 one pixel is exactly one 16-byte K-chunk, so the staged input rows (pixel-contiguous, 16 B per pixel)
// ARE the canonical K-major SWIZZLE_NONE operand: rows r = 128 consecutive pixels (8-row groups SBO = 128 B apart),
// and the operand of tap (dy, dx) for output row yy is the same buffer at byte offset (yy + dy) * ROWB + dx * 16 --
// an implicit GEMM with no im2col at all.  One tcgen05.mma consumes K = 16 = two chunks LBO apart: LBO = ROWB pairs
// the taps (dy, dx) and (dy + 1, dx); the odd third row is paired with a zero-weight chunk (a zeroed extra row).
// N = 16 (the smallest N for M = 128) with only column 0 carrying weights; one accumulator (16 TMEM columns) per
// output row, R rows per CTA: 6 MMAs per row.  Epilogue: thread = pixel reads column 0 of each accumulator, adds the
// bias and stores fp32 (128 B per warp and row).
struct FinalW { float w[72]; float b; };
constexpr int FT_PX = 130;                               // staged pixels per row (128 + halo)
constexpr int FT_ROWB = FT_PX * 16;                      // bytes per staged row
template <int FT_R> constexpr size_t ft_smem() { return (size_t)(FT_R + 3) * FT_ROWB + 6 * 512; }
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {     // volatile: issued where written, so a batch of them is in flight together
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// FT_R = output rows per CTA (16 * FT_R TMEM columns: a power of two >= 32, i.e. 2, 4, 8 or 16 rows)
template <int FT_R>
__global__ void __launch_bounds__(128) k_conv3x3_c8_final_tc(const __nv_bfloat16* __restrict__ in, const FinalW fw, float* __restrict__ out,
                                                             int H, int Wd, int* err) {
    constexpr int FT_ROWS = FT_R + 3;                        // halo above / below + one zero row for the padded K-chunk
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;
    unsigned char* sB = smem_raw + (size_t)FT_ROWS * FT_ROWB;          // 6 blocks of 16 x 16 bf16 (512 B each)
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, y0 = blockIdx.y * FT_R, x0 = blockIdx.x * 128;
    constexpr int NCOL = 16 * FT_R;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(NCOL) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- stage the input rows (zeros outside the image and in the extra row): all loads in flight, then the stores
    const __nv_bfloat16* src = in + (long long)b * H * Wd * 8;
    constexpr int NCH = FT_ROWS * FT_PX, PER = (NCH + 127) / 128;
    uint4 v[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int c = tid + q * 128;
        const int sy = c / FT_PX, px = c - sy * FT_PX;
        const int y = y0 - 1 + sy, x = x0 - 1 + px;
        v[q] = make_uint4(0, 0, 0, 0);
        if (c < NCH && sy < FT_R + 2 && y >= 0 && y < H && x >= 0 && x < Wd)
            v[q] = ldg_nc_v4(src + ((long long)y * Wd + x) * 8);
    }
    __syncthreads();            // scheduling fence for ptxas: every load above is issued before the first store below
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int c = tid + q * 128;
        if (c < NCH) *reinterpret_cast<uint4*>(sA + (size_t)c * 16) = v[q];
    }
    // ---- weights: block (dx, half): chunk 0 = tap (2*half, dx), chunk 1 = tap (2*half + 1, dx) or zeros; row n = 0 only
    for (int c = tid; c < 6 * 512 / 16; c += 128) *reinterpret_cast<uint4*>(sB + (size_t)c * 16) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (tid < 6 * 2) {
        const int blk = tid >> 1, kc = tid & 1, dx = blk >> 1, half = blk & 1, dy = 2 * half + kc;
        if (dy < 3) {
            __nv_bfloat16 o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16(fw.w[(dy * 3 + dx) * 8 + e]);
            *reinterpret_cast<uint4*>(sB + (size_t)blk * 512 + kc * 256) = *reinterpret_cast<const uint4*>(o);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 16);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
        for (int yy = 0; yy < FT_R; ++yy) {
#pragma unroll
            for (int blk = 0; blk < 6; ++blk) {
                const int dx = blk >> 1, half = blk & 1;
                const uint64_t ad = make_smem_desc(a0 + (uint32_t)(yy + 2 * half) * FT_ROWB + (uint32_t)dx * 16, FT_ROWB, 128);
                const uint64_t bd = make_smem_desc(b0 + (uint32_t)blk * 512, 256, 128);
                umma_bf16(tmem + 16u * yy, ad, bd, idesc, blk > 0 ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        }
        if (!done && lane == 0) atomicExch(err, 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc[FT_R];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int yy = 0; yy < FT_R; ++yy)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(acc[yy]) : "r"(taddr + 16u * yy) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int x = x0 + tid;
    if (x < Wd) {
#pragma unroll
        for (int yy = 0; yy < FT_R; ++yy)
            if (y0 + yy < H) out[((long long)b * H + y0 + yy) * Wd + x] = __uint_as_float(acc[yy]) + fw.b;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NCOL) : "memory");
}

// ---- fused tail of the split-operand decoder (round 2): ConvT(16 -> 8, 2x2, stride 2) + swish + Conv(8 -> 1, 3x3, 'same') ----
// (sr-ae-conv.ipynb cell 277-287: the last two layers of decoder_400.)  Unfused, the (B, 400, 400, 8) fp32 activation between
// them is written and read back once: 1.3 GB per 128 samples, more than every other tensor of the decoder together.  Here a
// CTA takes a patch of 16 x 16 pixels of the (B, 200, 200, 16) input -- two 128-row tiles of the same implicit GEMM as
// k_convT2x2_tc3<16, 32> (three bf16 MMAs per tile into a 32-column fp32 TMEM accumulator each) --, its epilogue (bias, expf
// swish) writes the 32 x 32 x 8 activation tile to SHARED memory (zeros where the input pixel lies outside the image: the
// final conv's 'same' padding), and the final conv runs from there on the CUDA cores: the middle 28 x 28 output pixels, the
// ring of one input pixel around the owned 14 x 14 being the conv's halo (recomputed by the neighbouring patches:
// (16/14)^2 = 1.31x the ConvT work and input reads).  Same operands, MMA order, swish and FMA order as the two kernels it
// replaces, so the output is the same bits.
constexpr int TF_P = 16, TF_OWN = TF_P - 2, TF_AT = 2 * TF_P;
constexpr size_t TF_ACT_BYTES = (size_t)TF_AT * TF_AT * 8 * 4, TF_ATILE = 128 * 16 * 2, TF_B_BYTES = 32 * 16 * 2;
constexpr size_t tail_fused_smem() { return TF_ACT_BYTES + 2 * TF_B_BYTES; }   // (the A operand lives inside the activation tile's space)
// 16-byte chunk of channel half h (0: c0-3, 1: c4-7) of pixel (row, col) in the activation tile.  The epilogue's lanes write
// pixel PAIRS 64 B apart and the conv's lanes read consecutive pixels; the XOR keeps both free of bank conflicts.
__device__ __forceinline__ int tf_chunk(int row, int col, int h) {
    const int pp = col >> 1, q = ((col & 1) << 1) | h;
    return (row * TF_AT + 2 * pp) * 2 + (q ^ ((pp >> 1) & 3));
}

__global__ void __launch_bounds__(128) k_tail_fused_tc3(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Whi,
                                                         const __nv_bfloat16* __restrict__ Wlo, const float* __restrict__ bias,
                                                         const FinalW fw, float* __restrict__ out, int H, int Wd, int* err) {
    constexpr int KD = 16, ND = 32, MT = 128, KC = KD / 8;
    constexpr uint32_t LBO_A = (MT / 8) * 128, LBO_B = (ND / 8) * 128, SBO = 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* act = reinterpret_cast<float*>(smem_raw);          // [32][32][8] fp32
    unsigned char* sAh = smem_raw;                            // two tiles of 128 rows, canonical K-major layout: dead once the MMAs
    unsigned char* sAl = sAh + 2 * TF_ATILE;                  // have completed, i.e. before the first activation is written over them
    unsigned char* sBh = smem_raw + TF_ACT_BYTES;
    unsigned char* sBl = sBh + TF_B_BYTES;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, y0 = (int)blockIdx.y * TF_OWN - 1, x0 = (int)blockIdx.x * TF_OWN - 1;   // input pixel of patch (0, 0)
    if (tid < 8) s_bias[tid] = bias[tid];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- the A patch: 256 rows x 2 chunks of 8 floats, every load in flight before the first split
    const float* Ab = A + (long long)b * H * Wd * KD;
    float4 va[4], vb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = tid + q * 128, r = c >> 1, kc = c & 1;
        const int Y = y0 + (r >> 4), X = x0 + (r & 15);
        va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q];
        if (Y >= 0 && Y < H && X >= 0 && X < Wd) {
            const float4* src = reinterpret_cast<const float4*>(Ab + ((long long)Y * Wd + X) * KD + kc * 8);
            va[q] = __ldg(src); vb[q] = __ldg(src + 1);
        }
    }
    for (int c = tid; c < ND * KC; c += 128) {                // both halves of the (32 x 16) weights
        const int n = c / KC, kc = c - n * KC;
        const size_t so = (size_t)kc * LBO_B + (n >> 3) * SBO + (n & 7) * 16;
        *reinterpret_cast<uint4*>(sBh + so) = *reinterpret_cast<const uint4*>(Whi + (size_t)n * KD + kc * 8);
        *reinterpret_cast<uint4*>(sBl + so) = *reinterpret_cast<const uint4*>(Wlo + (size_t)n * KD + kc * 8);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = tid + q * 128, r = c >> 1, kc = c & 1, t = r >> 7, rr = r & 127;
        const float v[8] = {va[q].x, va[q].y, va[q].z, va[q].w, vb[q].x, vb[q].y, vb[q].z, vb[q].w};
        uint4 hi, lo;
        split8(v, hi, lo);
        const size_t so = (size_t)t * TF_ATILE + (size_t)kc * LBO_A + (rr >> 3) * SBO + (rr & 7) * 16;
        *reinterpret_cast<uint4*>(sAh + so) = hi;
        *reinterpret_cast<uint4*>(sAl + so) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(MT, ND);
        const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), bh = smem_u32(sBh), bl = smem_u32(sBl);
        const uint64_t bdh = make_smem_desc(bh, LBO_B, SBO), bdl = make_smem_desc(bl, LBO_B, SBO);
#pragma unroll
        for (int t = 0; t < 2; ++t) {                          // K = 16: one K-step per tile
            const uint64_t adh = make_smem_desc(ah + (uint32_t)(t * TF_ATILE), LBO_A, SBO), adl = make_smem_desc(al + (uint32_t)(t * TF_ATILE), LBO_A, SBO);
            umma_bf16(tmem + (uint32_t)(t * ND), adh, bdh, idesc, 0u);
            umma_bf16(tmem + (uint32_t)(t * ND), adh, bdl, idesc, 1u);
            umma_bf16(tmem + (uint32_t)(t * ND), adl, bdh, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    bool dead = false;
    {                                                          // bounded wait: never hang the GPU
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        }
        if (!done) { if (lane == 0) atomicExch(err, 1); dead = true; }
    }
    dead = __syncthreads_or(dead ? 1 : 0) != 0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!dead) {
        // ---- ConvT epilogue: thread = accumulator row of both tiles; bias + swish into the shared activation tile
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = t * MT + warp * 32 + lane, py = r >> 4, px = r & 15;
            const int Y = y0 + py, X = x0 + px;
            const bool inside = Y >= 0 && Y < H && X >= 0 && X < Wd;
#pragma unroll
            for (int c0 = 0; c0 < ND; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * ND + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 16; g += 8) {
                    const int tap = (c0 + g) >> 3, dy = tap >> 1, dx = tap & 1;
                    const float4 b0 = *reinterpret_cast<const float4*>(s_bias), b1 = *reinterpret_cast<const float4*>(s_bias + 4);
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                    if (inside) {
                        o0 = make_float4(swish_exact(__uint_as_float(v[g + 0]) + b0.x), swish_exact(__uint_as_float(v[g + 1]) + b0.y),
                                         swish_exact(__uint_as_float(v[g + 2]) + b0.z), swish_exact(__uint_as_float(v[g + 3]) + b0.w));
                        o1 = make_float4(swish_exact(__uint_as_float(v[g + 4]) + b1.x), swish_exact(__uint_as_float(v[g + 5]) + b1.y),
                                         swish_exact(__uint_as_float(v[g + 6]) + b1.z), swish_exact(__uint_as_float(v[g + 7]) + b1.w));
                    }
                    float4* at = reinterpret_cast<float4*>(act);
                    at[tf_chunk(2 * py + dy, 2 * px + dx, 0)] = o0;
                    at[tf_chunk(2 * py + dy, 2 * px + dx, 1)] = o1;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
    if (dead) return;
    // ---- final conv on the middle 28 x 28 pixels of the tile: thread = column, 7 rows each (k_conv3x3_c8_final's FMA order)
    const int tx = lane, tq = warp;
    if (tx < 2 * TF_OWN) {
        const float bs = fw.b;
        float acc[7] = {bs, bs, bs, bs, bs, bs, bs};
#pragma unroll
        for (int r = 0; r < 9; ++r) {                          // tile rows 7*tq + 1 + r feed outputs r-2 .. r
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4* at = reinterpret_cast<const float4*>(act);
                const float4 a = at[tf_chunk(7 * tq + 1 + r, tx + 1 + kx, 0)], c4 = at[tf_chunk(7 * tq + 1 + r, tx + 1 + kx, 1)];
                const float xv[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int o = 0; o < 7; ++o) {
                    const int ky = r - o;
                    if (ky >= 0 && ky < 3) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[o] = fmaf(xv[c], fw.w[(ky * 3 + kx) * 8 + c], acc[o]);
                    }
                }
            }
        }
        const int OH = 2 * H, OW = 2 * Wd;
        const int gx = 2 * (int)blockIdx.x * TF_OWN + tx;
        if (gx < OW) {
#pragma unroll
            for (int o = 0; o < 7; ++o) {
                const int gy = 2 * (int)blockIdx.y * TF_OWN + 7 * tq + o;
                if (gy < OH) out[((long long)b * OH + gy) * OW + gx] = acc[o];
            }
        }
    }
}

// ---- the 3x3 / stride-2 ConvT of the split-operand decoder as ONE CTA per 128-row tile (round 2) -----------------------
// (sr-ae-conv.ipynb cell 277-287, conv2d_transpose 256 -> 128.)  Overlapping taps: nine per-tap GEMMs Y[m, tap, :] = A[m, :] W[tap]^T
// (K = 256, N = 128), summed by k_col2im_3x3s2_f32.  As nine launches' worth of k_convT2x2_tc3 tiles every tap re-staged (and
// re-split) the same 128 KB A tile and its own 128 KB of weights.  Here the A tile is staged and split ONCE (2 x 64 KB, resident)
// and the weights STREAM through a two-deep ring of 32 KB stages (one quarter of K for both bf16 halves) filled by the TMA
// engine: the host lays the split weights out once, at upload, as the exact shared-memory image of every stage (canonical
// K-major core matrices), so a stage is one cp.async.bulk with an mbarrier transaction count.  One thread issues copies and
// MMAs: wait full[s] -> 12 tcgen05.mma (4 K-steps x 3 split products) -> tcgen05.commit -> empty[s]; the copy of stage i+1
// goes out as soon as stage i-1's MMAs have released its buffer, so copies run under the MMAs.  After a tap's fourth stage a
// second commit signals the accumulator; all four warps drain it (tcgen05.ld) to Y while the next tap's first stages land.
constexpr int L1_K = 256, L1_N = 128, L1_TAPS = 9, L1_Q = 4, L1_KQ = L1_K / L1_Q;          // 64 K per stage
constexpr uint32_t L1_HALF_BYTES = L1_N * L1_KQ * 2, L1_STAGE_BYTES = 2 * L1_HALF_BYTES;   // 16 KB per bf16 half, 32 KB per stage
constexpr size_t L1_A_BYTES = (size_t)128 * L1_K * 2;                                      // 64 KB per bf16 half of the A tile
constexpr size_t l1_smem() { return 2 * L1_A_BYTES + 2 * (size_t)L1_STAGE_BYTES; }          // 192 KB
// element index of weight (tap, n, k), bf16 half `half`, in the staged image
__host__ __device__ inline size_t l1_img_index(int tap, int n, int k, int half) {
    const int q = k / L1_KQ, kc = (k % L1_KQ) / 8, kk = k % 8;
    return ((size_t)((tap * L1_Q + q) * 2 + half)) * (L1_HALF_BYTES / 2) + (size_t)kc * (L1_N / 8) * 64 + (size_t)(n >> 3) * 64 + (size_t)(n & 7) * 8 + kk;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 20) && !done; ++spin)     // <= ~1 us per try: a lost arrival costs about a second, never a hang
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000u) : "memory");
    return done != 0;
}

__global__ void __launch_bounds__(128) k_convT3x3_l1_tc3(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Wimg,
                                                          float* __restrict__ Y, long long M, int* err) {
    constexpr uint32_t LBO_A = (128 / 8) * 128, LBO_B = (L1_N / 8) * 128, SBO = 128;
    constexpr int NST = L1_TAPS * L1_Q;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sAh = smem_raw;
    unsigned char* sAl = smem_raw + L1_A_BYTES;
    unsigned char* sW = smem_raw + 2 * L1_A_BYTES;            // two stages
    __shared__ __align__(8) unsigned long long bars[5];       // full[2], empty[2], acc
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * 128;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]), accb = smem_u32(&bars[4]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first two weight stages are on their way while the A tile is staged
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8u * i), "r"(L1_STAGE_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sW + (size_t)i * L1_STAGE_BYTES)), "l"(Wimg + (size_t)i * (L1_STAGE_BYTES / 2)), "r"(L1_STAGE_BYTES), "r"(full0 + 8u * i) : "memory");
        }
    }
    // ---- the A tile: 128 rows x 32 chunks of 8 floats, split into its bf16 halves (row fastest: conflict-free stores)
#pragma unroll 1
    for (int it = 0; it < 32; it += 4) {
        float4 va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = tid + (it + u) * 128, kc = c >> 7, r = c & 127;
            va[u] = make_float4(0.f, 0.f, 0.f, 0.f); vb[u] = va[u];
            if (m0 + r < M) {
                const float4* src = reinterpret_cast<const float4*>(A + (m0 + r) * L1_K + kc * 8);
                va[u] = __ldg(src); vb[u] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = tid + (it + u) * 128, kc = c >> 7, r = c & 127;
            const float v[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
            uint4 hi, lo;
            split8(v, hi, lo);
            const size_t so = (size_t)kc * LBO_A + (r >> 3) * SBO + (r & 7) * 16;
            *reinterpret_cast<uint4*>(sAh + so) = hi;
            *reinterpret_cast<uint4*>(sAl + so) = lo;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const long long m = m0 + warp * 32 + lane;
    const bool valid = m < M;
    bool dead = false, dead0 = false;
#pragma unroll 1
    for (int tap = 0; tap < L1_TAPS; ++tap) {
        if (tid == 0 && !dead0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, L1_N);
            const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl);
#pragma unroll 1
            for (int q = 0; q < L1_Q; ++q) {
                const int i = tap * L1_Q + q, sidx = i & 1;
                if (!mbar_wait(full0 + 8u * sidx, (uint32_t)(i >> 1) & 1u)) { dead0 = true; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t wh = smem_u32(sW + (size_t)sidx * L1_STAGE_BYTES), wl = wh + L1_HALF_BYTES;
#pragma unroll
                for (int k = 0; k < L1_KQ / 16; ++k) {
                    const uint32_t ao = (uint32_t)(q * (L1_KQ / 8) + 2 * k) * LBO_A, bo = (uint32_t)(2 * k) * LBO_B;
                    const uint64_t adh = make_smem_desc(ah + ao, LBO_A, SBO), adl = make_smem_desc(al + ao, LBO_A, SBO);
                    const uint64_t bdh = make_smem_desc(wh + bo, LBO_B, SBO), bdl = make_smem_desc(wl + bo, LBO_B, SBO);
                    umma_bf16(tmem, adh, bdh, idesc, (q > 0 || k > 0) ? 1u : 0u);
                    umma_bf16(tmem, adh, bdl, idesc, 1u);
                    umma_bf16(tmem, adl, bdh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8u * sidx) : "memory");
                if (q == L1_Q - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(accb) : "memory");
                // refill: stage i+1 goes into the buffer stage i-1 used, once that stage's MMAs have completed
                if (i >= 1 && i + 1 < NST) {
                    const int pidx = (i - 1) & 1;
                    if (!mbar_wait(empty0 + 8u * pidx, (uint32_t)((i - 1) >> 1) & 1u)) { dead0 = true; break; }
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8u * pidx), "r"(L1_STAGE_BYTES) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(sW + (size_t)pidx * L1_STAGE_BYTES)), "l"(Wimg + (size_t)(i + 1) * (L1_STAGE_BYTES / 2)), "r"(L1_STAGE_BYTES), "r"(full0 + 8u * pidx) : "memory");
                }
            }
        }
        if (!mbar_wait(accb, (uint32_t)tap & 1u)) { if (lane == 0) atomicExch(err, 1); dead = true; }
        dead = __syncthreads_or(dead ? 1 : 0) != 0;
        if (dead) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < L1_N; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(Y + m * (L1_TAPS * L1_N) + (long long)tap * L1_N + c0);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                      // the accumulator is drained: the next tap may overwrite it
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

}  // namespace srtc
