// tcgen05 implicit-GEMM kernel + host-side plan builder. See igemm.cuh for the design.
#include "igemm.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------
// error string shared by the whole library (C-ABI: b2sd_last_error)
static thread_local char g_err[1024] = "";
const char* b2_last_error() { return g_err; }
void b2_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------
// epilogue math (shared by the main kernel and the split-K finalize kernel)
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void store_half16(__half* dst, const float* v, int nv, bool vec_ok) {
    if (nv == 16 && vec_ok) {
        uint4 u[2];
        __half2* h = reinterpret_cast<__half2*>(u);
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        reinterpret_cast<uint4*>(dst)[0] = u[0];
        reinterpret_cast<uint4*>(dst)[1] = u[1];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) dst[i] = __float2half_rn(v[i]);
    }
}

// 16 accumulator columns [col0, col0+16) of output row `orow` (batch item b); the accumulators are
// acc[OFF .. OFF+16) of a register array (compile-time indices only: nothing may spill to local memory).
template <int OFF, int N, typename T>
__device__ __forceinline__ void epi_store16(const IgEpilogue& e, const T (&acc)[N], int b, long orow, int col0) {
    int nv = e.n_valid - col0;
    if (nv <= 0) return;
    if (nv > 16) nv = 16;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if constexpr (sizeof(T) == 4 && !__is_same(T, float)) v[i] = __uint_as_float(acc[OFF + i]);
        else v[i] = acc[OFF + i];
    }
    if (e.colbias) {
        const float* bp = e.colbias + (long)b * e.colbias_bstride + col0;
        if (nv == 16 && (e.colbias_bstride & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 t = reinterpret_cast<const float4*>(bp)[i];
                v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nv) v[i] += bp[i];
        }
    }
    if (e.acc_scale != 1.0f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= e.acc_scale;
    }
    if (e.res) {
        const __half* rp = e.res + orow * e.ldr + col0;
        if (nv == 16 && (e.ldr & 7) == 0) {
            uint4 u[2];
            u[0] = reinterpret_cast<const uint4*>(rp)[0];
            u[1] = reinterpret_cast<const uint4*>(rp)[1];
            const __half2* h = reinterpret_cast<const __half2*>(u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 f = __half22float2(h[i]);
                v[2 * i] += e.res_scale * f.x;
                v[2 * i + 1] += e.res_scale * f.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nv) v[i] += e.res_scale * __half2float(rp[i]);
        }
    }
    if (e.flags & IG_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    store_half16(e.out + orow * e.ldc + col0, v, nv, (e.ldc & 7) == 0);
}

// GEGLU: val/gate are 16 accumulator columns each; packed-column index of val[0] is pcol0 (bias
// uses packed indexing), output column index is ocol0.
__device__ __forceinline__ void epi_store16_geglu(const IgEpilogue& e, const uint32_t (&val)[16],
                                                  const uint32_t (&gate)[16], long orow, int pcol_val,
                                                  int pcol_gate, int ocol0) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(val[i]), g = __uint_as_float(gate[i]);
        if (e.colbias) {
            a += e.colbias[pcol_val + i];
            g += e.colbias[pcol_gate + i];
        }
        v[i] = a * gelu_erf(g);
    }
    store_half16(e.out + orow * e.ldc + ocol0, v, 16, (e.ldc & 7) == 0);
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IG_THREADS) igemm_kernel(const __grid_constant__ IgemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t stage_bytes = IG_BM * IG_BK * 2 + (uint32_t)p.BN * IG_BK * 2;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * stage_bytes);
    uint64_t* empty_bar = full_bar + IG_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + IG_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    // tile coordinates
    const int bx = blockIdx.x;
    const int tiw = bx % p.tiles_w;
    const int tih = (bx / p.tiles_w) % p.tiles_h;
    const int tin = bx / (p.tiles_w * p.tiles_h);
    const int w0 = tiw * p.tw, h0 = tih * p.th, n0 = tin * p.tn;
    const int ntile = blockIdx.y;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.tmA[s]);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int seg = 0, base = 0;
            while (seg < p.nseg - 1 && kb_begin >= base + p.seg_ntap[seg] * p.seg_cblocks[seg]) {
                base += p.seg_ntap[seg] * p.seg_cblocks[seg];
                ++seg;
            }
            int tap = (kb_begin - base) / p.seg_cblocks[seg];
            int cb = (kb_begin - base) % p.seg_cblocks[seg];
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                uint8_t* sb = sa + IG_BM * IG_BK * 2;
                mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
                int dy = 0, dx = 0;
                if (p.seg_ntap[seg] == 9) {
                    dy = tap / 3 - 1;
                    dx = tap % 3 - 1;
                }
                tma_load_4d(sa, &p.tmA[seg], &full_bar[stage], p.seg_c0[seg] + cb * IG_BK,
                            w0 * p.stride + dx, h0 * p.stride + dy, n0);
                tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * IG_BK, ntile * p.BN);
                if (++cb == p.seg_cblocks[seg]) {
                    cb = 0;
                    if (++tap == p.seg_ntap[seg]) {
                        tap = 0;
                        ++seg;
                    }
                }
                if (++stage == p.num_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_idesc_f16(IG_BM, p.BN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t sb = sa + IG_BM * IG_BK * 2;
                const uint64_t da = make_kmajor_sw128_desc(sa);
                const uint64_t db = make_kmajor_sw128_desc(sb);
#pragma unroll
                for (int k = 0; k < IG_BK / 16; ++k) {
                    // +32 B per UMMA_K inside the 128 B swizzle row => +2 in the (addr>>4) field
                    umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                             (kb > kb_begin || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                if (++stage == p.num_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        const int wi = r % p.tw;
        const int hi = (r / p.tw) % p.th;
        const int ni = r / (p.tw * p.th);
        const int n = n0 + ni, h = h0 + hi, w = w0 + wi;
        const bool row_ok = (ni < p.tn) && (n < p.Nb) && (h < p.Ho) && (w < p.Wo);
        const long orow = ((long)n * p.Ho + h) * p.Wo + w;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const IgEpilogue& e = p.epi;
        if (e.flags & IG_SPLITK) {
            // Split-K inside a thread-block cluster (one CTA per K slice, cluster dims (1,1,splits)): every CTA
            // parks its fp32 partial tile in its own shared memory, laid out [4-column group][row] so that both
            // these stores and the peers' DSMEM reads are conflict-free; the reduction happens after the cluster
            // barrier below.
            float4* stg = reinterpret_cast<float4*>(smem);
            for (int c = 0; c < p.BN; c += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    stg[((c >> 2) + i) * IG_BM + r] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                  __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
        } else if (e.flags & IG_GEGLU) {
            const int half_n = p.BN / 2;
            for (int c = 0; c < half_n; c += 16) {
                uint32_t a[16], g[16];
                tmem_ld16(taddr + c, a);
                tmem_ld16(taddr + half_n + c, g);
                tmem_ld_wait();
                if (row_ok)
                    epi_store16_geglu(e, a, g, orow, ntile * p.BN + c, ntile * p.BN + half_n + c,
                                      ntile * half_n + c);
            }
        } else {
            int c = 0;
            for (; c + 32 <= p.BN; c += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c, v);
                tmem_ld_wait();
                if (row_ok) {
                    epi_store16<0>(e, v, n, orow, ntile * p.BN + c);
                    epi_store16<16>(e, v, n, orow, ntile * p.BN + c + 16);
                }
            }
            if (c < p.BN) {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
                if (row_ok) epi_store16<0>(e, v, n, orow, ntile * p.BN + c);
            }
        }
    }
    if (p.epi.flags & IG_SPLITK) {
        // ---- cluster-wide deterministic reduction over the K slices through distributed shared memory ----
        const int splits = (int)gridDim.z;
        cluster_sync_all();  // all partial tiles are in place (release/acquire over the cluster)
        if (warp >= 2) {
            const int rank = (int)cluster_ctarank();
            const int rows_per = IG_BM / splits;          // splits in {2,4,8}
            const int t = threadIdx.x - 64;               // 0..127
            const int chunks = p.BN >> 4;
            const uint32_t stg_local = smem_u32(smem);
            for (int item = t; item < rows_per * chunks; item += 128) {
                const int rl = item % rows_per;
                const int cc = item / rows_per;
                const int r = rank * rows_per + rl;       // row of the tile this CTA finalises
                const int wi = r % p.tw, hi = (r / p.tw) % p.th, ni = r / (p.tw * p.th);
                const int n = n0 + ni, h = h0 + hi, w = w0 + wi;
                const bool ok = (ni < p.tn) && (n < p.Nb) && (h < p.Ho) && (w < p.Wo);
                float acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.f;
                for (int sidx = 0; sidx < splits; ++sidx) {   // fixed order => bit-reproducible
                    const uint32_t peer = dsmem_map(stg_local, (uint32_t)sidx);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = dsmem_ld_f4(peer + (uint32_t)(((cc * 4 + i) * IG_BM + r) * 16));
                        acc[4 * i] += v.x; acc[4 * i + 1] += v.y; acc[4 * i + 2] += v.z; acc[4 * i + 3] += v.w;
                    }
                }
                if (ok) epi_store16<0>(p.epi, acc, n, ((long)n * p.Ho + h) * p.Wo + w, ntile * p.BN + cc * 16);
            }
        }
        cluster_sync_all();  // nobody may exit while a peer still reads its shared memory
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
// host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
        b2_set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                     cudaGetErrorString(err));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    return fn;
}

static int encode_act_map(CUtensorMap* m, const ActView& a, int box_c, int box_w, int box_h, int box_n,
                          int estride) {
    auto enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {(cuuint64_t)a.ld * 2, (cuuint64_t)a.W * a.ld * 2,
                             (cuuint64_t)a.H * a.W * a.ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(a.ptr), dims, strides,
                     box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        b2_set_error("cuTensorMapEncodeTiled(act) failed: %d (ptr %p dims %d,%d,%d,%d ld %d box "
                     "%d,%d,%d,%d es %d)",
                     (int)r, a.ptr, a.C, a.W, a.H, a.N, a.ld, box_c, box_w, box_h, box_n, estride);
        return -1;
    }
    return 0;
}

static int encode_w_map(CUtensorMap* m, const __half* w, int rows, int ld, int box_rows) {
    auto enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {IG_BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        b2_set_error("cuTensorMapEncodeTiled(weights) failed: %d (rows %d ld %d box %d)", (int)r, rows,
                     ld, box_rows);
        return -1;
    }
    return 0;
}

size_t igemm_partial_floats(int splits, long rows_total, int n_valid) {
    // worst case n_pad = n_tiles * BN < n_valid + 256
    long n_pad = ((n_valid + 255) / 256 + 1) * 256;
    return (size_t)splits * rows_total * n_pad;
}

int igemm_plan(const IgemmDesc& d, IgemmPlan* plan) {
    memset(plan, 0, sizeof(*plan));
    IgemmParams& p = plan->p;
    if (d.nseg < 1 || d.nseg > IG_MAX_SRC) {
        b2_set_error("igemm: bad nseg %d", d.nseg);
        return -1;
    }
    const bool geglu = (d.epi.flags & IG_GEGLU) != 0;
    const int n_gemm = geglu ? d.epi.n_valid * 2 : d.epi.n_valid;  // GEMM N (packed weight rows used)
    // ---- N tile
    int BN = d.BN;
    if (BN <= 0) {
        if (n_gemm <= 256 && n_gemm % 16 == 0) BN = n_gemm;
        else if (n_gemm < 16) BN = 16;
        else if (n_gemm % 128 == 0) BN = 128;
        else if (n_gemm % 160 == 0) BN = 160;
        else if (n_gemm % 64 == 0) BN = 64;
        else BN = 128;
    }
    if (BN % 16 != 0 || BN < 16 || BN > 256 || (geglu && (BN % 32 != 0 || n_gemm % BN != 0))) {
        b2_set_error("igemm: unsupported BN %d (n %d)", BN, n_gemm);
        return -1;
    }
    p.BN = BN;
    const int n_tiles = (n_gemm + BN - 1) / BN;
    if (d.w_rows < n_gemm) {
        // TMA zero-fills rows beyond w_rows; allowed (padded N) but flag obviously wrong descs
        if (d.w_rows <= 0) {
            b2_set_error("igemm: w_rows %d", d.w_rows);
            return -1;
        }
    }
    // ---- spatial tile
    int tw, th, tn;
    if (d.Ho == 1 && d.Nb == 1) {
        tw = IG_BM; th = 1; tn = 1;
    } else {
        if (d.Wo <= 16) tw = d.Wo;
        else if (d.Wo % 16 == 0) tw = 16;
        else if (d.Wo % 8 == 0) tw = 8;
        else tw = 16;
        th = IG_BM / tw;
        if (th > d.Ho) th = d.Ho;
        tn = IG_BM / (tw * th);
        if (tn > d.Nb) tn = d.Nb;
        if (tn < 1) tn = 1;
    }
    p.tw = tw; p.th = th; p.tn = tn;
    p.tiles_w = (d.Wo + tw - 1) / tw;
    p.tiles_h = (d.Ho + th - 1) / th;
    p.tiles_n = (d.Nb + tn - 1) / tn;
    p.Wo = d.Wo; p.Ho = d.Ho; p.Nb = d.Nb;
    p.stride = d.stride < 1 ? 1 : d.stride;
    if (p.stride > 2) {
        b2_set_error("igemm: stride %d", p.stride);
        return -1;
    }
    // ---- K segments + TMA maps
    p.nseg = d.nseg;
    int total_kb = 0;
    for (int s = 0; s < d.nseg; ++s) {
        const ActView& a = d.src[s];
        if (a.C % IG_BK != 0 || (a.ld % 8) != 0 || (reinterpret_cast<uintptr_t>(a.ptr) & 15)) {
            b2_set_error("igemm: source %d: C=%d must be a multiple of 64, ld=%d multiple of 8, ptr "
                         "16B aligned",
                         s, a.C, a.ld);
            return -1;
        }
        if (d.ntap[s] != 1 && d.ntap[s] != 9) {
            b2_set_error("igemm: ntap %d", d.ntap[s]);
            return -1;
        }
        p.seg_ntap[s] = d.ntap[s];
        p.seg_cblocks[s] = a.C / IG_BK;
        p.seg_c0[s] = 0;
        total_kb += d.ntap[s] * p.seg_cblocks[s];
        if (encode_act_map(&p.tmA[s], a, IG_BK, tw * p.stride, th * p.stride, tn, p.stride)) return -1;
    }
    p.total_kb = total_kb;
    if (d.w_ld < total_kb * IG_BK || (d.w_ld % 8) != 0 || (reinterpret_cast<uintptr_t>(d.w) & 15)) {
        b2_set_error("igemm: weight ld %d < K %d or misaligned", d.w_ld, total_kb * IG_BK);
        return -1;
    }
    if (encode_w_map(&p.tmB, d.w, d.w_rows, d.w_ld, BN)) return -1;
    p.a_bytes = (uint32_t)(tw * th * tn) * IG_BK * 2;
    p.b_bytes = (uint32_t)BN * IG_BK * 2;
    // ---- split-K
    int splits = d.splits < 1 ? 1 : d.splits;
    if (splits > total_kb) splits = total_kb;
    if (splits >= 8) splits = 8;        // portable cluster size; power of two so rows divide evenly
    else if (splits >= 4) splits = 4;
    else if (splits >= 2) splits = 2;
    p.kb_per_split = (total_kb + splits - 1) / splits;
    while (splits > 1 && (total_kb + p.kb_per_split - 1) / p.kb_per_split != splits) {  // keep every slice non-empty
        splits >>= 1;
        p.kb_per_split = (total_kb + splits - 1) / splits;
    }
    plan->splits = splits;
    plan->rows_total = (long)d.Nb * d.Ho * d.Wo;
    p.epi = d.epi;
    p.n_pad = n_tiles * BN;
    if (splits > 1) {
        if (geglu) {
            b2_set_error("igemm: split-K cannot be combined with GEGLU");
            return -1;
        }
        p.epi.flags |= IG_SPLITK;
    }
    // ---- pipeline depth / smem
    const size_t stage_bytes = (size_t)IG_BM * IG_BK * 2 + (size_t)BN * IG_BK * 2;
    int stages = (int)((100 * 1024) / stage_bytes);
    if (stages < 2) stages = 2;
    if (stages > IG_MAX_STAGES) stages = IG_MAX_STAGES;
    if (stages > p.kb_per_split) stages = p.kb_per_split < 2 ? 2 : p.kb_per_split;
    p.num_stages = stages;
    size_t pipe_bytes = stages * stage_bytes;
    if (splits > 1) {
        // the split-K staging tile [BN/4][128] float4 reuses the pipeline buffers
        const size_t stg = (size_t)BN * IG_BM * 4;
        if (stg > pipe_bytes) {
            // grow the ring rather than carving a second region (barriers live right after the ring)
            stages = (int)((stg + stage_bytes - 1) / stage_bytes);
            if (stages > IG_MAX_STAGES) {
                b2_set_error("igemm: split-K staging does not fit (BN %d)", BN);
                return -1;
            }
            p.num_stages = stages;
            pipe_bytes = stages * stage_bytes;
        }
    }
    plan->smem = pipe_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
    uint32_t cols = 32;
    while (cols < (uint32_t)BN) cols <<= 1;
    p.tmem_cols = cols;
    plan->grid = dim3(p.tiles_w * p.tiles_h * p.tiles_n, n_tiles, splits);
    return 0;
}

int igemm_init() {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) {
            b2_set_error("cudaFuncSetAttribute(igemm): %s", cudaGetErrorString(e));
            return -1;
        }
        if (!get_encode()) return -1;
        attr_set = true;
    }
    return 0;
}

int igemm_launch(const IgemmPlan& plan, cudaStream_t stream) {
    if (igemm_init()) return -1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = plan.grid;
    cfg.blockDim = dim3(IG_THREADS);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (plan.splits > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = (unsigned)plan.splits;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, igemm_kernel, plan.p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        b2_set_error("igemm launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace b2
