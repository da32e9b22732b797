// Hardware probe (test-only ABI): does a UMMA shared-memory descriptor whose start address is shifted by whole
// 128-byte rows inside a 128B-swizzled tile address the shifted rows correctly, and does it need base_offset?
//   mode 0: K-major A tile [144 rows][64 k] -> D[128,64] = A[shift : shift+128, :] * B^T        (conv fwd/dgrad tap shift)
//   mode 1: MN-major B tile [144 k-rows][64 n], A MN-major [128 k-rows][64 m... M=128 via 2 blocks]
//           -> D[128,64] = sum_k A[k, m] * B[k + shift, n]                                       (conv wgrad tap shift)
#pragma once
#include "gemm_tc.cuh"

namespace emb {

__device__ __forceinline__ uint64_t umma_desc_bo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_offset) {
    return umma_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)(base_offset & 7u) << 49);
}

__global__ void __launch_bounds__(128, 1)
umma_shift_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int mode, int shift,
                        int use_bo, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                    // up to 2 x 18 KB
    uint8_t* sb = smem + 40960;            // 18 KB
    uint64_t* bar = (uint64_t*)(smem + 65536);
    uint64_t* done = bar + 1;
    uint32_t* slot = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        if (mode == 0) {
            mbar_expect_tx(bar, 144 * 128 + 64 * 128);
            tma_load_3d(sa, &map_a, bar, 0, 0, 0);          // box {64, 144, 1}
            tma_load_3d(sb, &map_b, bar, 0, 0, 0);          // box {64, 64, 1}
        } else {
            mbar_expect_tx(bar, 2 * 128 * 128 + 144 * 128);
            tma_load_3d(sa, &map_a, bar, 0, 0, 0);          // box {64 m, 128 k}
            tma_load_3d(sa + 16384, &map_a, bar, 64, 0, 0);
            tma_load_3d(sb, &map_b, bar, 0, 0, 0);          // box {64 n, 144 k}
        }
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
        if (mode == 0) {
            const uint32_t idesc = make_idesc_dev(0, 0, 64);
            for (int s = 0; s < 4; ++s) {
                uint64_t da = umma_desc_bo(a0 + shift * 128 + s * 32, 16, 1024, use_bo ? shift : 0);
                uint64_t db = umma_desc(b0 + s * 32, 16, 1024);
                tc_mma_f16(tmem, da, db, idesc, s ? 1u : 0u);
            }
        } else {
            const uint32_t idesc = make_idesc_dev(1, 1, 64);
            for (int s = 0; s < 8; ++s) {
                uint64_t da = umma_desc(a0 + s * 2048, 16384, 1024);
                uint64_t db = umma_desc_bo(b0 + shift * 128 + s * 2048, 18432, 1024, use_bo ? shift : 0);
                tc_mma_f16(tmem, da, db, idesc, s ? 1u : 0u);
            }
        }
        tc_commit(done);
        if (use_bo >= 2) {
            // timing: reps x 8 MMAs back to back, N = use_bo * 32 columns (garbage operands beyond the staged tiles are fine)
            mbar_wait(done, 0);
            const int ncols = (use_bo >> 1) * 32;
            const uint32_t idesc = make_idesc_dev(mode, mode, ncols);
            const uint32_t kstep = mode ? 2048 : 32;
            long long t0 = clock64();
            for (int rep = 0; rep < 200; ++rep)
                for (int s2 = 0; s2 < 8; ++s2) {
                    uint64_t da = umma_desc(a0 + (mode == 0 ? shift * 128 : 0) + (s2 & 3) * kstep, mode ? 16384 : 16, 1024);
                    uint64_t db = umma_desc(b0 + (mode == 1 ? shift * 128 : 0) + (s2 & 3) * kstep, mode ? 18432 : 16, 1024);
                    tc_mma_f16(tmem, da, db, idesc, 1u);
                }
            tc_commit(done);
            mbar_wait(done, 1);
            long long t1 = clock64();
            out[128 * 64] = (float)(t1 - t0) / 1600.f;
        }
    }
    if (use_bo < 2) mbar_wait(done, 0);
    else { __syncthreads(); }
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

}  // namespace emb
