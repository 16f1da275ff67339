// Host-only check of the weight image k_convT3x3_l1_tc3 streams with cp.async.bulk (sr_tc.cuh): l1_img_index must place weight
// (tap, n, k) of each bf16 half where the kernel's UMMA descriptors look for it -- stage (tap, K quarter) of 32 KB = [hi 16 KB | lo 16 KB],
// each half in the canonical K-major SWIZZLE_NONE layout (8-row x 16-byte core matrices, SBO = 128 B, LBO = (128 / 8) * 128 B) --
// and must be a bijection onto the image.  Built and run by tests/test_sr_layout_cpu.py (no GPU needed).
#include <cstdio>
#include <vector>
#include "../../sr-for-cfd_b200/csrc/sr_tc.cuh"

int main() {
    using namespace srtc;
    const size_t total = (size_t)L1_TAPS * L1_Q * 2 * (L1_HALF_BYTES / 2);
    std::vector<unsigned char> seen(total, 0);
    const unsigned LBO_B = (L1_N / 8) * 128, SBO = 128;
    for (int tap = 0; tap < L1_TAPS; ++tap)
        for (int n = 0; n < L1_N; ++n)
            for (int k = 0; k < L1_K; ++k)
                for (int half = 0; half < 2; ++half) {
                    const size_t idx = l1_img_index(tap, n, k, half);
                    if (idx >= total) { std::printf("index out of range\n"); return 1; }
                    if (seen[idx]++) { std::printf("collision at %zu\n", idx); return 2; }
                    const int q = k / L1_KQ, kl = k % L1_KQ;
                    const size_t stage = (size_t)(tap * L1_Q + q) * L1_STAGE_BYTES;                 // bytes: what the kernel copies per stage
                    const size_t want = stage + (size_t)half * L1_HALF_BYTES + (size_t)(kl / 8) * LBO_B + (size_t)(n >> 3) * SBO + (size_t)(n & 7) * 16 + (size_t)(kl & 7) * 2;
                    if (idx * 2 != want) { std::printf("tap %d n %d k %d half %d: byte %zu, canonical %zu\n", tap, n, k, half, idx * 2, want); return 3; }
                }
    for (size_t i = 0; i < total; ++i) if (!seen[i]) { std::printf("hole at %zu\n", i); return 4; }
    if (l1_smem() != 2 * L1_A_BYTES + 2 * (size_t)L1_STAGE_BYTES || l1_smem() > 227 * 1024) { std::printf("smem\n"); return 5; }
    if (tail_fused_smem() != TF_ACT_BYTES + 2 * TF_B_BYTES || 4 * TF_ATILE > TF_ACT_BYTES) { std::printf("tail smem / alias\n"); return 6; }
    std::printf("ok %zu elements\n", total);
    return 0;
}
