// Persistent halo-tile 3x3 convolution with resident weights (see tconv.cuh).
#include "tconv.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

#include "epilogue.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b2 {

constexpr uint32_t TC_WTILE = TC_C * TC_C * 2;                               // one tap's [64 x 64] fp16 weight tile
constexpr uint32_t TC_HALO_BYTES = (TC_TW + 2) * (TC_TH + 2) * TC_C * 2;     // 10 x 18 pixels x 128 B

__global__ void __launch_bounds__(TC_THREADS, 1) tconv_kernel(const __grid_constant__ TconvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint8_t* sW = smem;                       // nine weight tiles, tap-major, each the canonical K-major SWIZZLE_128B tile
    uint8_t* sA = smem + 9 * TC_WTILE;        // halo ring
    uint64_t* w_full = reinterpret_cast<uint64_t*>(sA + (size_t)p.nbuf * p.abuf_bytes);
    uint64_t* a_full = w_full + 1;
    uint64_t* a_empty = a_full + TC_MAX_ABUF;
    uint64_t* tmem_full_bar = a_empty + TC_MAX_ABUF;   // [2] accumulator ready   (MMA -> epilogue)
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] accumulator drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        mbar_init(w_full, 1);
        for (int s = 0; s < p.nbuf; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * TC_C);   // two fp32 accumulators of 64 columns
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the weights are constants of the stream: request them before the programmatic-dependency wait
    if (warp == 0 && lane == 0) {
        mbar_expect_tx(w_full, 9 * TC_WTILE);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + tap * TC_WTILE, &p.tmB, w_full, tap * TC_C, 0);
    }
    pdl_launch_dependents();
    pdl_wait();

    const int tiles_per_img = p.tiles_w * p.tiles_h;
    if (warp == 0) {
        if (lane == 0) {
            // ===== halo producer: one (TH+2) x (TW+2) pixel tile per output tile =====
            int slot = 0;
            uint32_t phase = 0;
            for (int mt = blockIdx.x; mt < p.num_tiles; mt += gridDim.x) {
                const int tiw = mt % p.tiles_w, tih = (mt / p.tiles_w) % p.tiles_h, n0 = mt / tiles_per_img;
                mbar_wait(&a_empty[slot], phase ^ 1);
                mbar_expect_tx(&a_full[slot], TC_HALO_BYTES);
                tma_load_4d(sA + (size_t)slot * p.abuf_bytes, &p.tmA, &a_full[slot], 0, tiw * TC_TW - 1, tih * TC_TH - 1, n0);
                if (++slot == p.nbuf) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop, one elected lane issues (see igemm_kernel) =====
        const uint32_t idesc = make_idesc_f16(IG_BM, TC_C);
        const uint32_t sa_base = smem_u32(sA);
        const uint64_t db0 = make_kmajor_sw128_desc(smem_u32(sW));
        constexpr int pitch = TC_TW + 2;   // pixels per halo row: the 8-row core groups of the A operand are `pitch` pixels apart
        mbar_wait(w_full, 0);
        tc_fence_after();
        int slot = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int mt = blockIdx.x; mt < p.num_tiles; mt += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
            mbar_wait(&a_full[slot], phase);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)buf * TC_C;
            // A descriptor of tap (0,0): rows r = 8*hi + wi -> halo pixel hi*pitch + wi (+ tap shift).  The swizzle follows the
            // absolute shared-memory address bits, so the 128-byte-granular tap shifts need no base offset (tools/probe).
            uint64_t da0 = 0;
            da0 |= (uint64_t)(((sa_base + (uint32_t)slot * p.abuf_bytes) & 0x3ffff) >> 4);
            da0 |= (uint64_t)1 << 16;
            da0 |= (uint64_t)((pitch * 128) >> 4) << 32;
            da0 |= (uint64_t)1 << 46;
            da0 |= (uint64_t)2 << 61;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint64_t da = da0 + (uint64_t)(((tap / 3) * pitch + (tap % 3)) * 8);   // 128 B per pixel = 8 units of 16 B
                const uint64_t db = db0 + (uint64_t)(tap * (TC_WTILE >> 4));
                if (elect_one()) {
                    umma_f16(tacc, da, db, idesc, tap > 0 ? 1u : 0u);
                    umma_f16(tacc, da + 2, db + 2, idesc, 1u);
                    umma_f16(tacc, da + 4, db + 4, idesc, 1u);
                    umma_f16(tacc, da + 6, db + 6, idesc, 1u);
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(&a_empty[slot]);          // the halo buffer may be refilled when these MMAs retire
                umma_commit(&tmem_full_bar[buf]);
            }
            __syncwarp();
            if (++slot == p.nbuf) {
                slot = 0;
                phase ^= 1;
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias / residual / ReLU -> fp16 NHWC =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int hi = r >> 3, wl = r & 7;
        const IgEpilogue& e = p.epi;
        int it = 0;
        for (int mt = blockIdx.x; mt < p.num_tiles; mt += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t par = (it >> 1) & 1;
            const int tiw = mt % p.tiles_w, tih = (mt / p.tiles_w) % p.tiles_h, n0 = mt / tiles_per_img;
            const int h = tih * TC_TH + hi, w = tiw * TC_TW + wl;
            const bool row_ok = (h < p.Ho) && (w < p.Wo);
            const long orow = ((long)n0 * p.Ho + h) * p.Wo + w;
            const uint32_t taddr = tmem_base + (uint32_t)buf * TC_C + ((uint32_t)(q * 32) << 16);
            epi_row_fast(e, taddr, TC_C, 0, n0, orow, row_ok, &tmem_full_bar[buf], par);
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * TC_C);
}

// ------------------------------------------------------------------------------------------ host side
bool tconv_eligible(const IgemmDesc& d) {
    const IgEpilogue& e = d.epi;
    return d.nseg == 1 && d.ntap[0] == 9 && d.stride <= 1 && !d.swap && d.src[0].C == TC_C && e.n_valid == TC_C &&
           d.w_rows >= TC_C && d.w_ld == 9 * TC_C && d.src[0].H == d.Ho && d.src[0].W == d.Wo && d.src[0].N == d.Nb &&
           !(e.flags & (IG_GEGLU | IG_SPLITK)) && (e.ldc & 7) == 0 && (!e.res || (e.ldr & 7) == 0) && (e.colbias_bstride & 3) == 0 &&
           (d.src[0].ld & 7) == 0 && !(reinterpret_cast<uintptr_t>(d.src[0].ptr) & 15) && !(reinterpret_cast<uintptr_t>(d.w) & 15) &&
           !(reinterpret_cast<uintptr_t>(e.out) & 15) && !(reinterpret_cast<uintptr_t>(e.res) & 15) &&
           !(reinterpret_cast<uintptr_t>(e.colbias) & 15);
}

int igemm_encode_act_map(CUtensorMap* m, const ActView& a, int box_c, int box_w, int box_h, int box_n, int estride);
int igemm_encode_w_map(CUtensorMap* m, const __half* w, int rows, int ld, int box_rows);

int tconv_plan(const IgemmDesc& d, TconvPlan* plan) {
    *plan = TconvPlan{};
    if (!tconv_eligible(d)) {
        b2_set_error("tconv: needs a stride-1 3x3 convolution with 64 input and 64 output channels and a vectorisable epilogue");
        return -1;
    }
    TconvParams& p = plan->p;
    p.tiles_w = (d.Wo + TC_TW - 1) / TC_TW;
    p.tiles_h = (d.Ho + TC_TH - 1) / TC_TH;
    p.num_tiles = p.tiles_w * p.tiles_h * d.Nb;
    p.Wo = d.Wo; p.Ho = d.Ho; p.Nb = d.Nb;
    static const char* nb_env = getenv("B2_TCONV_NBUF");
    p.nbuf = nb_env ? atoi(nb_env) : 4;
    if (p.nbuf < 2) p.nbuf = 2;
    if (p.nbuf > TC_MAX_ABUF) p.nbuf = TC_MAX_ABUF;
    p.abuf_bytes = (TC_HALO_BYTES + 1023u) & ~1023u;
    p.epi = d.epi;
    if (igemm_encode_act_map(&p.tmA, d.src[0], TC_C, TC_TW + 2, TC_TH + 2, 1, 1)) return -1;
    if (igemm_encode_w_map(&p.tmB, d.w, d.w_rows, d.w_ld, TC_C)) return -1;
    int sms = 148;
    {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.multiProcessorCount > 0)
            sms = prop.multiProcessorCount;
    }
    plan->grid = dim3(p.num_tiles < sms ? p.num_tiles : sms, 1, 1);
    plan->smem = 9 * (size_t)TC_WTILE + (size_t)p.nbuf * p.abuf_bytes + 1024 /*align slack*/ + 512 /*barriers*/;
    plan->rows_total = (long)d.Nb * d.Ho * d.Wo;
    return 0;
}

int tconv_init() {
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(tconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            b2_set_error("cudaFuncSetAttribute(tconv): %s", cudaGetErrorString(e));
            return -1;
        }
        done = true;
    }
    return 0;
}

int tconv_launch(const TconvPlan& plan, cudaStream_t stream) {
    if (tconv_init()) return -1;
    cudaError_t e = launch_k(tconv_kernel, plan.grid, dim3(TC_THREADS), plan.smem, stream, 1, plan.p);
    if (e != cudaSuccess) {
        b2_set_error("tconv launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace b2
