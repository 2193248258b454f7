// Hardware probe (kept for documentation / regression): how a tcgen05 K-major SWIZZLE_128B A-operand descriptor
// behaves when its start address is NOT 1024-byte aligned (shifted by whole 128-byte rows) and when the 8-row group
// pitch (SBO) is not 1024 bytes.  This is what a halo-reuse 3x3 convolution needs: one TMA load of a (th+2)x(tw+2)
// pixel tile, nine MMAs reading shifted windows of it.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "../../include/b200sd.h"  // C-ABI conventions only
#include "igemm.cuh"
#include "ptx.cuh"

namespace b2 {

struct ProbeParams {
    CUtensorMap tmA, tmB;
    float* D;
    int rows_a;       // rows loaded (<= 256)
    int shift_rows;   // descriptor start = base + shift_rows*128
    int pitch_rows;   // 8-row groups are pitch_rows rows apart (SBO = pitch_rows*128)
    int base_mode;    // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
    long long* timing; // optional: cycles for 1024 back-to-back MMAs
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                      // rows_a x 128 B
    uint8_t* sB = smem + 256 * 128;          // 64 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
    uint64_t* mma_bar = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(mma_bar + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(mma_bar, 1);
        mbar_init(mma_bar + 2, 1);
        mbar_init(mma_bar + 3, 1);
        mbar_init(mma_bar + 4, 1);
        *(reinterpret_cast<uint32_t*>(mma_bar + 8) + 1) = 0;
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)p.rows_a * 128 + 64 * 128);
        tma_load_2d(sA, &p.tmA, bar, 0, 0);
        tma_load_2d(sB, &p.tmB, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)p.shift_rows * 128;
        uint64_t da = 0;
        da |= (uint64_t)((a_addr & 0x3ffff) >> 4);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((p.pitch_rows * 128) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        if (p.base_mode & 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
        da |= (uint64_t)2 << 61;
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(sB));
        const uint32_t idesc = make_idesc_f16(128, 64);
        for (int k = 0; k < 4; ++k) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
        umma_commit(mma_bar);
        if (p.timing && !(p.base_mode & (128 | 256))) {
            // (A) divergent single-thread issue: loop runs under `if (threadIdx.x == 0)`
            mbar_wait(mma_bar, 0);
            const long long c0 = clock64();
            const unsigned long long g0 = globaltimer_ns();
            for (int it = 0; it < 256; ++it) {
                if (p.base_mode & 32) mbar_wait(mma_bar, 0);
                if (p.base_mode & 16) tc_fence_after();
                for (int k = 0; k < 4; ++k) umma_f16(tmem + ((p.base_mode & 64) ? (it & 1) * 64 : 0), da + 2 * k, db + 2 * k, idesc, 1);
                if ((p.base_mode & 4) && (it & 1)) umma_commit(mma_bar + 2);
            }
            umma_commit(bar);
            mbar_wait(bar, 1);
            if (blockIdx.x == 0) {
                p.timing[0] = clock64() - c0;
                p.timing[1] = (long long)(globaltimer_ns() - g0);
            }
            *reinterpret_cast<volatile uint32_t*>(slot + 1) = 1;  // release the spinners
        }
    }
    if (p.timing && (p.base_mode & 128) && !(p.base_mode & 256) && warp == 0) {
        // (B) warp-uniform issue: every lane of warp 0 runs the loop, one elected lane issues -> operands can live in
        // uniform registers
        mbar_wait(mma_bar, 0);
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)p.shift_rows * 128;
        uint64_t da = 0;
        da |= (uint64_t)((a_addr & 0x3ffff) >> 4);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((p.pitch_rows * 128) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        da |= (uint64_t)2 << 61;
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(sB));
        const uint32_t idesc = make_idesc_f16(128, 64);
        const long long c0 = clock64();
        const unsigned long long g0 = globaltimer_ns();
        for (int it = 0; it < 256; ++it) {
            if (p.base_mode & 32) mbar_wait(mma_bar, 0);
            if (p.base_mode & 16) tc_fence_after();
            if (elect_one()) {
                for (int k = 0; k < 4; ++k) umma_f16(tmem + ((p.base_mode & 64) ? (it & 1) * 64 : 0), da + 2 * k, db + 2 * k, idesc, 1);
                if ((p.base_mode & 4) && (it & 1)) umma_commit(mma_bar + 2);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 1);
        if (blockIdx.x == 0 && lane == 0) {
            p.timing[0] = clock64() - c0;
            p.timing[1] = (long long)(globaltimer_ns() - g0);
        }
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(slot + 1) = 1;
    }
    if (p.timing && (p.base_mode & 256) && warp == 0) {
        // (C) warp-uniform issue with the stage loop unrolled by 4: all descriptors are loop-invariant per unrolled slot
        mbar_wait(mma_bar, 0);
        const uint32_t idesc = make_idesc_f16(128, 64);
        uint64_t da[4], db[4];
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            da[st] = make_kmajor_sw128_desc(smem_u32(sA) + st * 8192);   // four "stages" inside the loaded tile
            db[st] = make_kmajor_sw128_desc(smem_u32(sB));
        }
        const long long c0 = clock64();
        const unsigned long long g0 = globaltimer_ns();
        for (int it = 0; it < 64; ++it) {
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                if (p.base_mode & 32) mbar_wait(mma_bar, 0);
                if (p.base_mode & 16) tc_fence_after();
                if (elect_one()) {
                    umma_f16(tmem, da[st], db[st], idesc, 1);
                    umma_f16(tmem, da[st] + 2, db[st] + 2, idesc, 1);
                    umma_f16(tmem, da[st] + 4, db[st] + 4, idesc, 1);
                    umma_f16(tmem, da[st] + 6, db[st] + 6, idesc, 1);
                    if (p.base_mode & 4) umma_commit(mma_bar + 2);
                }
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 1);
        if (blockIdx.x == 0 && lane == 0) {
            p.timing[0] = clock64() - c0;
            p.timing[1] = (long long)(globaltimer_ns() - g0);
        }
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(slot + 1) = 1;
    }
    if (p.timing && (p.base_mode & 2) && threadIdx.x >= 32) {
        // like the epilogue warps of the real kernels: wait on an mbarrier that completes only at the very end
        while (*reinterpret_cast<volatile uint32_t*>(slot + 1) == 0) mbar_try_wait(mma_bar + 3, 0);
    }
    if (p.timing && (p.base_mode & 8) && threadIdx.x == 32) {
        // concurrent TMA traffic into shared memory (reload A over and over into its own buffer)
        for (int it = 0; it < 24; ++it) {
            mbar_expect_tx(mma_bar + 4, (uint32_t)p.rows_a * 128);
            tma_load_2d(sA, &p.tmA, mma_bar + 4, 0, 0);
            mbar_wait(mma_bar + 4, it & 1);
        }
    }
    if (!(p.timing && (threadIdx.x == 0 || ((p.base_mode & (128 | 256)) && warp == 0)))) mbar_wait(mma_bar, 0);
    __syncthreads();
    tc_fence_after();
    const int r = warp * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) p.D[r * 64 + c + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace b2

extern "C" int b2sd_probe_umma_rowshift(const void* A, int rows_a, const void* B, void* D, int shift_rows, int pitch_rows,
                                        int base_mode, void* stream, long long* timing) {
    using namespace b2;
    if (igemm_init()) return -1;
    ProbeParams p{};
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return -1;
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
    cuuint64_t da[2] = {64, (cuuint64_t)rows_a}, db[2] = {64, 64};
    cuuint64_t st[1] = {128};
    cuuint32_t ba[2] = {64, (cuuint32_t)rows_a}, bb[2] = {64, 64}, es[2] = {1, 1};
    if (enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(A), da, st, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(B), db, st, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        fprintf(stderr, "probe: tensor map encode failed\n");
        return -1;
    }
    p.D = static_cast<float*>(D);
    p.rows_a = rows_a; p.shift_rows = shift_rows; p.pitch_rows = pitch_rows; p.base_mode = base_mode; p.timing = timing;
    const size_t smem = 256 * 128 + 64 * 128 + 256;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = getenv("B2_PROBE_GRID") ? atoi(getenv("B2_PROBE_GRID")) : 1;
    probe_kernel<<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
