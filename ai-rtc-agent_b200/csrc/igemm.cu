// tcgen05 implicit-GEMM kernel + host-side plan builder. See igemm.cuh for the design.
#include "igemm.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "epilogue.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b2 {

// per-CTA timeline stamps (tools/timeline.py): compiled in only with -DB2_TIMELINE, they lengthen the MMA issue loop
#ifdef B2_TIMELINE
#define B2_TS(stmt) stmt
#else
#define B2_TS(stmt)
#endif

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

// Swapped orientation: an accumulator row (TMEM lane, thread) is an output CHANNEL, its columns are the pixels of the
// tile, so the NHWC store needs a transpose.  It goes through a small fp32 shared-memory tile T[pixel][128 channels]:
// thread r parks its channel's values for a chunk of <= 32 pixels (conflict-free 4-byte stores), then the 128 epilogue
// threads sweep T row-wise: one thread = 8 consecutive channels of one pixel (16-byte residual load, 16-byte store), a
// warp = two complete 256-byte pixel rows.  Bias, scale, residual and ReLU are applied in that coalesced sweep.
constexpr int SWAP_CH = 32;   // pixels per transposition chunk (T = 32 x 128 fp32 = 16 KB)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// What one thread needs from global memory for its (pixel, 8-channel) items of a chunk: fetched BEFORE the
// transposition barrier so the latency overlaps the TMEM/DSMEM reads that fill T.
struct SwapPre {
    uint4 res[SWAP_CH / 8];
    float4 b0[SWAP_CH / 8], b1[SWAP_CH / 8];
    long orow[SWAP_CH / 8];      // < 0: nothing to store
};

__device__ __forceinline__ void swap_prefetch(const IgemmParams& p, SwapPre& pre, int npix, int j0, int ntile, int n0, int h0,
                                              int w0, int t) {
    const IgEpilogue& e = p.epi;
    const int tw = 1 << p.tw_log2, th = 1 << p.th_log2;
#pragma unroll
    for (int k = 0; k < SWAP_CH / 8; ++k) {
        const int item = t + 128 * k;
        pre.orow[k] = -1;
        pre.res[k] = make_uint4(0, 0, 0, 0);
        pre.b0[k] = pre.b1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item >= npix * 16) continue;
        const int pl = item >> 4, q = item & 15;
        const int j = j0 + pl;
        const int wi = j & (tw - 1);
        const int hi = (j >> p.tw_log2) & (th - 1);
        const int ni = j >> (p.tw_log2 + p.th_log2);
        const int n = n0 + ni, h = h0 + hi, w = w0 + wi;
        const int cout0 = ntile * IG_BM + q * 8;
        if (ni >= p.tn || n >= p.Nb || h >= p.Ho || w >= p.Wo || cout0 >= e.n_valid) continue;
        const long orow = ((long)n * p.Ho + h) * p.Wo + w;
        pre.orow[k] = orow;
        if (e.colbias) {
            const float4* bp = reinterpret_cast<const float4*>(e.colbias + (long)n * e.colbias_bstride + cout0);
            pre.b0[k] = __ldg(bp);
            pre.b1[k] = __ldg(bp + 1);
        }
        if (e.res) pre.res[k] = __ldg(reinterpret_cast<const uint4*>(e.res + orow * e.ldr + cout0));
    }
}

__device__ __forceinline__ void swap_store_chunk(const IgemmParams& p, const SwapPre& pre, const float* T, int ntile, int t) {
    const IgEpilogue& e = p.epi;
#pragma unroll
    for (int k = 0; k < SWAP_CH / 8; ++k) {
        if (pre.orow[k] < 0) continue;
        const int item = t + 128 * k;
        const int pl = item >> 4, q = item & 15;
        const int cout0 = ntile * IG_BM + q * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(T + pl * IG_BM + q * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(T + pl * IG_BM + q * 8 + 4);
        const float4 b0 = pre.b0[k], b1 = pre.b1[k];
        float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= e.acc_scale;
        if (e.res) {
            const __half2* rh = reinterpret_cast<const __half2*>(&pre.res[k]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(rh[i]);
                v[2 * i] += e.res_scale * f.x;
                v[2 * i + 1] += e.res_scale * f.y;
            }
        }
        if (e.flags & IG_RELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(e.out + pre.orow[k] * e.ldc + cout0) = o;
    }
}

// Cluster split-K: sum one 16-column (4 x float4) strip of accumulator row `row` over the K slices held in the peers'
// shared memory.  All loads of up to four slices are in flight together (a dependent chain of DSMEM round trips was
// the dominant cost of the reduction); the summation order is fixed => bit-reproducible.
// (pair launches: cluster dims (2,1,splits), the K slice s of this CTA's M half lives in cluster rank 2*s + rank_add)
template <int SPL>
__device__ __forceinline__ void splitk_sum16(uint32_t stg_local, int cc, int row, float (&acc)[16], int rank_mul = 1, int rank_add = 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    constexpr int G = SPL < 4 ? SPL : 4;
#pragma unroll
    for (int s0 = 0; s0 < SPL; s0 += G) {
        float4 v[G][4];
#pragma unroll
        for (int s = 0; s < G; ++s) {
            const uint32_t peer = dsmem_map(stg_local, (uint32_t)((s0 + s) * rank_mul + rank_add));
#pragma unroll
            for (int i = 0; i < 4; ++i) v[s][i] = dsmem_ld_f4(peer + (uint32_t)(((cc * 4 + i) * IG_BM + row) * 16));
        }
#pragma unroll
        for (int s = 0; s < G; ++s)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[4 * i] += v[s][i].x; acc[4 * i + 1] += v[s][i].y; acc[4 * i + 2] += v[s][i].z; acc[4 * i + 3] += v[s][i].w;
            }
    }
}

// ------------------------------------------------------------------------------------------
// PAIR: the CTAs (2j, 2j+1) of grid.x form a CTA pair (cluster dims (2,1,splits)) that computes two neighbouring M tiles with
// ONE tcgen05.mma.cta_group::2 stream of M = 256: each CTA loads its own 128 pixel rows and half of the weight tile, the even
// CTA issues the MMAs for both, every CTA drains its own 128 accumulator rows.  Barrier plumbing across the pair: the odd
// CTA's (otherwise idle) MMA warp relays "my operands of this stage have landed" to the leader; ring slots are released in both
// CTAs by a multicast commit; both epilogues arrive on the leader's accumulator-drained barrier.
template <bool PAIR>
__device__ __forceinline__ void igemm_body(const IgemmParams& p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t stage_bytes = IG_BM * IG_BK * 2 + (uint32_t)(PAIR ? p.BN / 2 : p.BN) * IG_BK * 2;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * stage_bytes);
    uint64_t* empty_bar = full_bar + IG_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + IG_MAX_STAGES;   // [2] accumulator ready   (MMA -> epilogue)
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2] accumulator drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint64_t* peer_full = full_bar + 32;                   // [stages] (pair leader) the peer CTA's operands have landed
    const int cr = PAIR ? (int)(blockIdx.x & 1) : 0;       // rank inside the CTA pair (0 = leader)
    const int mt_first = PAIR ? (int)(blockIdx.x & ~1u) + cr : (int)blockIdx.x;   // first M tile of this CTA ...
    const int mt_step = (int)gridDim.x;                    // ... and the stride of a persistent launch (even for pairs)
    const int mt_guard = PAIR ? cr : 0;                    // pairs iterate together: the loop bound looks at the leader's tile

    // Persistent over M tiles: CTA x handles tiles x, x + gridDim.x, ... with two TMEM accumulators, so the epilogue of
    // tile i overlaps the mainloop of tile i+1 and the prologue (barriers, TMEM, descriptors) is paid once per CTA.
    const int num_mtiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int ntile = blockIdx.y;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
    [[maybe_unused]] unsigned long long* ts = p.dbg_ts ? p.dbg_ts + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
    B2_TS(if (ts && threadIdx.x == 0) ts[0] = globaltimer_ns();)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.tmA[s]);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            if (PAIR) mbar_init(&peer_full[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], PAIR ? 256 : 128);
        }
        fence_mbar_init();
    }
    if (PAIR) {   // the peer's barriers exist before anything arrives on them; both CTAs are resident (execution barrier only:
        cluster_arrive_relaxed();   // the release/acquire form costs a MEMBAR.ALL.GPU + L1 invalidate)
        cluster_wait();
    }
    if (warp == 1) {
        if (PAIR) {
            tmem_alloc_2cta(tmem_slot, p.tmem_cols);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(tmem_slot, p.tmem_cols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    pdl_launch_dependents();   // the next kernel may start its own prologue now
    pdl_wait();                // ... and everything below reads the previous kernel's output
    B2_TS(if (ts && threadIdx.x == 0) ts[1] = globaltimer_ns();)
    const uint32_t acc_stride = p.acc_bufs > 1 ? (uint32_t)p.BN : 0u;   // column offset of the second accumulator

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = mt_first; mt - mt_guard < num_mtiles; mt += mt_step) {
                const int w0 = (mt % p.tiles_w) * p.tw, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
                const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;   // (odd tile count: the last pair's second tile lies outside, TMA zero-fills)
                int seg = 0, base = 0;
                while (seg < p.nseg - 1 && kb_begin >= base + p.seg_ntap[seg] * p.seg_cblocks[seg]) {
                    base += p.seg_ntap[seg] * p.seg_cblocks[seg];
                    ++seg;
                }
                int tap = (kb_begin - base) / p.seg_cblocks[seg];
                int cb = (kb_begin - base) % p.seg_cblocks[seg];
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
#ifdef B2_BOUND_STUDY
                    if (p.dbg_mode == 1 && kb >= kb_begin + p.num_stages) {   // bound study: operands stay whatever they were
                        mbar_arrive(&full_bar[stage]);
                        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
#endif
                    mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + IG_BM * IG_BK * 2;
                    int dy = 0, dx = 0;
                    if (p.seg_ntap[seg] == 9) {
                        dy = tap / 3 - 1;
                        dx = tap % 3 - 1;
                    }
                    // normal: pixels -> A (M side), weights -> B.  swapped: weights (128 output channels) -> A, pixels -> B
                    tma_load_4d(p.swap ? sb : sa, &p.tmA[seg], &full_bar[stage], p.seg_c0[seg] + cb * IG_BK,
                                w0 * p.stride + dx, h0 * p.stride + dy, n0);
                    tma_load_2d(p.swap ? sa : sb, &p.tmB, &full_bar[stage], kb * IG_BK,
                                ntile * (p.swap ? IG_BM : p.BN) + (PAIR ? cr * (p.BN / 2) : 0));
                    B2_TS(if (ts && mt == (int)blockIdx.x && kb == kb_begin) ts[2] = globaltimer_ns();)
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
        }
    } else if (warp == 1 && PAIR && cr == 1) {
        // ===== pair follower: relay "operands of this stage have landed in MY shared memory" to the leader =====
        const uint32_t leader_peer_full = dsmem_map(smem_u32(peer_full), cluster_ctarank() & ~1u);
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = mt_first; mt - mt_guard < num_mtiles; mt += mt_step) {
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                if (elect_one()) mbar_arrive_remote(leader_peer_full + (uint32_t)stage * 8u);
                __syncwarp();
                if (++stage == p.num_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The whole warp walks the loop (warp-uniform control flow keeps descriptors in uniform registers, which is
        // what UTCHMMA consumes); one elected lane issues.  A divergent single-thread loop issues ~2.5x slower
        // (probe.cu / tools/probe_rowshift.py).
        const uint32_t idesc = make_idesc_f16(PAIR ? 2 * IG_BM : IG_BM, p.BN);
        const uint32_t smem_base = smem_u32(smem);
        [[maybe_unused]] const uint16_t pair_mask = (uint16_t)(3u << (cluster_ctarank() & ~1u));
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int mt = mt_first; mt < num_mtiles; mt += mt_step, ++it) {
            const int buf = it & 1;
            // epilogue has drained this accumulator (pairs: both CTAs' epilogues arrive here)
            mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)buf * acc_stride;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                if (PAIR) mbar_wait(&peer_full[stage], phase);
                tc_fence_after();
                B2_TS(if (ts && it == 0 && kb == kb_begin && lane == 0) ts[3] = globaltimer_ns();)
                const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                const uint64_t da = make_kmajor_sw128_desc(sa);
                const uint64_t db = make_kmajor_sw128_desc(sa + IG_BM * IG_BK * 2);
                const uint32_t acc0 = kb > kb_begin ? 1u : 0u;
#ifdef B2_BOUND_STUDY
                if (p.dbg_mode == 2) {
                    if (elect_one()) umma_commit(&empty_bar[stage]);
                } else
#endif
                if (elect_one()) {
                    // +32 B per UMMA_K inside the 128 B swizzle row => +2 in the (addr>>4) field
                    if (PAIR) {
                        umma_f16_2cta(tacc, da, db, idesc, acc0);
                        umma_f16_2cta(tacc, da + 2, db + 2, idesc, 1u);
                        umma_f16_2cta(tacc, da + 4, db + 4, idesc, 1u);
                        umma_f16_2cta(tacc, da + 6, db + 6, idesc, 1u);
                        umma_commit_2cta(&empty_bar[stage], pair_mask);  // frees the slot in BOTH CTAs
                    } else {
                        umma_f16(tacc, da, db, idesc, acc0);
                        umma_f16(tacc, da + 2, db + 2, idesc, 1u);
                        umma_f16(tacc, da + 4, db + 4, idesc, 1u);
                        umma_f16(tacc, da + 6, db + 6, idesc, 1u);
                        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    }
                }
                __syncwarp();
                if (++stage == p.num_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) {
                if (PAIR) umma_commit_2cta(&tmem_full_bar[buf], pair_mask);
                else umma_commit(&tmem_full_bar[buf]);
            }
            __syncwarp();
            B2_TS(if (ts && it == 0 && lane == 0) ts[4] = globaltimer_ns();)
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        const int wi = r % p.tw;
        const int hi = (r / p.tw) % p.th;
        const int ni = r / (p.tw * p.th);
        const IgEpilogue& e = p.epi;
        // LayerNorm-folded launches: stage this N tile's colsum / bias' in shared memory while the mainloop runs
        const bool ln_smem = e.colsum && !p.swap && !(e.flags & IG_SPLITK) && (e.ldc & 7) == 0 && (e.n_valid & 15) == 0;
        float* lnv = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512);
        if (ln_smem) {
            for (int i = threadIdx.x - 64; i < p.BN; i += 128) {
                const int gc = ntile * p.BN + i;
                const bool in = gc < ((e.flags & IG_GEGLU) ? 2 * e.n_valid : e.n_valid);
                lnv[i] = in ? e.colsum[gc] : 0.f;
                lnv[p.BN + i] = (in && e.colbias) ? e.colbias[gc] : 0.f;
            }
            epi_bar_sync();
        }
        [[maybe_unused]] const uint32_t leader_tmem_empty = PAIR ? dsmem_map(smem_u32(tmem_empty_bar), cluster_ctarank() & ~1u) : 0u;
        int it = 0;
        for (int mt = mt_first; mt - mt_guard < num_mtiles; mt += mt_step, ++it) {
            const int buf = it & 1;
            const uint32_t par = (it >> 1) & 1;
            const int w0 = (mt % p.tiles_w) * p.tw, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
            const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;
            const int n = n0 + ni, h = h0 + hi, w = w0 + wi;
            const bool row_ok = (ni < p.tn) && (n < p.Nb) && (h < p.Ho) && (w < p.Wo);
            const long orow = ((long)n * p.Ho + h) * p.Wo + w;
            const uint32_t taddr = tmem_base + (uint32_t)buf * acc_stride + ((uint32_t)(q * 32) << 16);
            if (p.swap) {
                mbar_wait(&tmem_full_bar[buf], par);
                tc_fence_after();
                if (e.flags & IG_SPLITK) {
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
                } else {
                    // the operand ring is idle (every MMA has retired): use its head as the transposition tile
                    float* T = reinterpret_cast<float*>(smem);
                    for (int c = 0; c < p.BN; c += SWAP_CH) {
                        SwapPre pre;
                        swap_prefetch(p, pre, SWAP_CH, c, ntile, n0, h0, w0, r);
                        uint32_t v[32];
                        tmem_ld32(taddr + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) T[i * IG_BM + r] = __uint_as_float(v[i]);
                        epi_bar_sync();
                        swap_store_chunk(p, pre, T, ntile, r);
                        epi_bar_sync();
                    }
                }
            } else if (ln_smem) {
                float mu, rstd;
                ln_row_stats(e, orow, row_ok, mu, rstd);     // global loads: in flight while the accumulator completes
                mbar_wait(&tmem_full_bar[buf], par);
                tc_fence_after();
                if (e.flags & IG_GEGLU) {
                    epi_row_geglu_ln(e, taddr, p.BN / 2, ntile * (p.BN / 2), orow, row_ok, lnv, p.BN, mu, rstd);
                } else {
                    int ncols = e.n_valid - ntile * p.BN;
                    if (ncols > p.BN) ncols = p.BN;
                    epi_row_ln(e, taddr, ncols, ntile * p.BN, orow, row_ok && ncols > 0, lnv, p.BN, mu, rstd);
                }
            } else if (epi_fast_ok(e)) {
                int ncols = e.n_valid - ntile * p.BN;
                if (ncols > p.BN) ncols = p.BN;
                epi_row_fast(e, taddr, ncols, ntile * p.BN, n, orow, row_ok && ncols > 0, &tmem_full_bar[buf], par);
            } else {
                mbar_wait(&tmem_full_bar[buf], par);
                tc_fence_after();
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
                    float mu, rstd;
                    ln_row_stats(e, orow, row_ok, mu, rstd);
                    for (int c = 0; c < half_n; c += 16) {
                        uint32_t a[16], g[16];
                        tmem_ld16(taddr + c, a);
                        tmem_ld16(taddr + half_n + c, g);
                        tmem_ld_wait();
                        if (row_ok)
                            epi_store16_geglu(e, a, g, orow, ntile * p.BN + c, ntile * p.BN + half_n + c, ntile * half_n + c, mu, rstd);
                    }
                } else {
                    float mu, rstd;
                    ln_row_stats(e, orow, row_ok, mu, rstd);
                    for (int c = 0; c + 32 <= p.BN; c += 32) {
                        uint32_t v[32];
                        tmem_ld32(taddr + c, v);
                        tmem_ld_wait();
                        if (row_ok) {
                            epi_store16<0>(e, v, n, orow, ntile * p.BN + c, mu, rstd);
                            epi_store16<16>(e, v, n, orow, ntile * p.BN + c + 16, mu, rstd);
                        }
                    }
                    if (p.BN & 31) {   // 16-column tail
                        const int c = p.BN & ~31;
                        uint32_t v[16];
                        tmem_ld16(taddr + c, v);
                        tmem_ld_wait();
                        if (row_ok) epi_store16<0>(e, v, n, orow, ntile * p.BN + c, mu, rstd);
                    }
                }
            }
            // accumulator `buf` is drained: the MMA warp may start the tile after next into it
            tc_fence_before();
            if (PAIR && cr == 1) mbar_arrive_remote(leader_tmem_empty + (uint32_t)buf * 8u);
            else mbar_arrive(&tmem_empty_bar[buf]);
            B2_TS(if (ts && it == 0 && threadIdx.x == 64) ts[5] = globaltimer_ns();)
        }
    }
    B2_TS(if (ts && threadIdx.x == 64) ts[6] = globaltimer_ns();)
    if (p.epi.flags & IG_SPLITK) {
        // ---- cluster-wide deterministic reduction over the K slices through distributed shared memory (one tile per CTA:
        // split-K launches are never persistent) ----
        const int mt = blockIdx.x;
        const int w0 = (mt % p.tiles_w) * p.tw, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;
        const int splits = (int)gridDim.z;
        const int rk_mul = PAIR ? 2 : 1, rk_add = cr;   // cluster rank of K slice s (same M half) = s * rk_mul + rk_add
        cluster_sync_all();  // all partial tiles are in place (release/acquire over the cluster)
        B2_TS(if (ts && threadIdx.x == 64) ts[6] = globaltimer_ns();)   // split launches: [5] staged, [6] cluster barrier passed, [7] reduced
        if (warp >= 2 && p.swap) {
            // swapped orientation: this CTA finalises the pixel columns [rank*cols_per, (rank+1)*cols_per) of the tile
            const int rank = (int)cluster_ctarank();
            const int cols_per = p.BN / splits;
            const int ch = cols_per < SWAP_CH ? cols_per : SWAP_CH;
            const int t = threadIdx.x - 64;               // 0..127 == accumulator row == output channel of the tile
            const uint32_t stg_local = smem_u32(smem);
            float* T = reinterpret_cast<float*>(smem + (size_t)p.BN * IG_BM * 4);   // right after the staging tile
            for (int c = rank * cols_per; c < (rank + 1) * cols_per; c += ch) {
                SwapPre pre;
                swap_prefetch(p, pre, ch, c, ntile, n0, h0, w0, t);
                for (int g = 0; g < (ch >> 2); g += 4) {     // 16 pixel columns per pass (ch is 8, 16 or 32)
                    float acc[16];
                    const int cc = ((c >> 2) + g) >> 2;      // 16-column strip index (c is a multiple of 16 when ch >= 16)
                    if (ch >= 16) {
                        switch (splits) {
                            case 2: splitk_sum16<2>(stg_local, cc, t, acc); break;
                            case 4: splitk_sum16<4>(stg_local, cc, t, acc); break;
                            default: splitk_sum16<8>(stg_local, cc, t, acc); break;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) T[(4 * g + i) * IG_BM + t] = acc[i];
                    } else {
                        // 8 pixel columns per CTA (BN 64 over 8 slices): two float4 groups
#pragma unroll
                        for (int gg = 0; gg < 2; ++gg) {
                            float4 v[8];
#pragma unroll
                            for (int sidx = 0; sidx < 8; ++sidx)
                                v[sidx] = dsmem_ld_f4(dsmem_map(stg_local, (uint32_t)sidx) + (uint32_t)((((c >> 2) + gg) * IG_BM + t) * 16));
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int sidx = 0; sidx < 8; ++sidx) { a.x += v[sidx].x; a.y += v[sidx].y; a.z += v[sidx].z; a.w += v[sidx].w; }
                            T[(4 * gg + 0) * IG_BM + t] = a.x;
                            T[(4 * gg + 1) * IG_BM + t] = a.y;
                            T[(4 * gg + 2) * IG_BM + t] = a.z;
                            T[(4 * gg + 3) * IG_BM + t] = a.w;
                        }
                    }
                }
                epi_bar_sync();
                swap_store_chunk(p, pre, T, ntile, t);
                epi_bar_sync();
            }
        } else if (warp >= 2) {
            const int rank = (int)cluster_ctarank() / rk_mul;
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
                switch (splits) {
                    case 2: splitk_sum16<2>(stg_local, cc, r, acc, rk_mul, rk_add); break;
                    case 4: splitk_sum16<4>(stg_local, cc, r, acc, rk_mul, rk_add); break;
                    default: splitk_sum16<8>(stg_local, cc, r, acc, rk_mul, rk_add); break;
                }
                if (ok) {
                    const long orow = ((long)n * p.Ho + h) * p.Wo + w;
                    float mu, rstd;
                    ln_row_stats(p.epi, orow, true, mu, rstd);
                    epi_store16<0>(p.epi, acc, n, orow, ntile * p.BN + cc * 16, mu, rstd);
                }
            }
        }
        B2_TS(if (ts && threadIdx.x == 64) ts[7] = globaltimer_ns();)
        cluster_sync_all();  // nobody may exit while a peer still reads its shared memory
    }
    tc_fence_before();
    if (PAIR) {   // both CTAs are done with the pair's tensor memory and with each other's barriers
        cluster_arrive_relaxed();
        cluster_wait();
    } else {
        __syncthreads();
    }
    if (warp == 1) {
        if (PAIR) tmem_dealloc_2cta(tmem_base, p.tmem_cols);
        else tmem_dealloc(tmem_base, p.tmem_cols);
    }
    B2_TS(if (ts && threadIdx.x == 32 && !(p.epi.flags & IG_SPLITK)) ts[7] = globaltimer_ns();)
}

__global__ void __launch_bounds__(IG_THREADS) igemm_kernel(const __grid_constant__ IgemmParams p) { igemm_body<false>(p); }
__global__ void __launch_bounds__(IG_THREADS) igemm_pair_kernel(const __grid_constant__ IgemmParams p) { igemm_body<true>(p); }

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

// Planning without a GPU (tests of the host logic, b2sd_igemm_plan_dry): the tensor maps are left zeroed.
static thread_local bool g_plan_dry = false;
void igemm_set_dry_run(bool on) { g_plan_dry = on; }

static int encode_act_map(CUtensorMap* m, const ActView& a, int box_c, int box_w, int box_h, int box_n,
                          int estride) {
    if (g_plan_dry) return 0;
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
    if (g_plan_dry) return 0;
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

// Swapped orientation plan: output channels on the M side (128 per CTA), a tile of BN pixels on the N side.
static int plan_swap(const IgemmDesc& d, IgemmPlan* plan) {
    IgemmParams& p = plan->p;
    if ((d.epi.flags & IG_GEGLU) || d.nseg < 1 || d.nseg > IG_MAX_SRC || d.epi.rowstat_out || d.epi.colsum || d.epi.out2) {
        b2_set_error("igemm(swap): unsupported (GEGLU / LayerNorm fold / row statistics / transposed V / nseg %d)", d.nseg);
        return -1;
    }
    int BN = d.BN;
    const long rows_total = (long)d.Nb * d.Ho * d.Wo;
    if (BN <= 0) BN = rows_total >= 256 ? 256 : (rows_total >= 128 ? 128 : 64);
    if (BN != 64 && BN != 128 && BN != 256) {
        b2_set_error("igemm(swap): BN %d must be 64, 128 or 256", BN);
        return -1;
    }
    p.swap = 1;
    { static const char* dm = getenv("B2_DBG_MODE"); p.dbg_mode = dm ? atoi(dm) : 0; }
    p.acc_bufs = 1;
    p.BN = BN;
    auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    int tw, th, tn;
    if (d.Ho == 1 && d.Nb == 1) { tw = BN; th = 1; tn = 1; }
    else if (d.Wo >= 16) { tw = 16; th = BN / 16; tn = 1; }
    else { tw = 8; th = 8; tn = BN / 64; }
    p.tw = tw; p.th = th; p.tn = tn;
    p.tw_log2 = lg(tw); p.th_log2 = lg(th);
    p.tiles_w = (d.Wo + tw - 1) / tw;
    p.tiles_h = (d.Ho + th - 1) / th;
    p.tiles_n = (d.Nb + tn - 1) / tn;
    p.Wo = d.Wo; p.Ho = d.Ho; p.Nb = d.Nb;
    p.stride = d.stride < 1 ? 1 : d.stride;
    p.nseg = d.nseg;
    int total_kb = 0;
    for (int s = 0; s < d.nseg; ++s) {
        const ActView& a = d.src[s];
        if (a.C % IG_BK != 0 || (a.ld % 8) != 0 || (reinterpret_cast<uintptr_t>(a.ptr) & 15) || (d.ntap[s] != 1 && d.ntap[s] != 9)) {
            b2_set_error("igemm(swap): bad source %d", s);
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
        b2_set_error("igemm(swap): weight ld %d < K %d or misaligned", d.w_ld, total_kb * IG_BK);
        return -1;
    }
    if (encode_w_map(&p.tmB, d.w, d.w_rows, d.w_ld, IG_BM)) return -1;
    p.a_bytes = IG_BM * IG_BK * 2;            // weight tile (lands in the A region)
    p.b_bytes = (uint32_t)BN * IG_BK * 2;     // pixel tile
    int splits = d.splits < 1 ? 1 : d.splits;
    if (splits > total_kb) splits = total_kb;
    if (splits >= 8) splits = 8; else if (splits >= 4) splits = 4; else if (splits >= 2) splits = 2;
    p.kb_per_split = (total_kb + splits - 1) / splits;
    while (splits > 1 && (total_kb + p.kb_per_split - 1) / p.kb_per_split != splits) {
        splits >>= 1;
        p.kb_per_split = (total_kb + splits - 1) / splits;
    }
    plan->splits = splits;
    plan->rows_total = rows_total;
    p.epi = d.epi;
    p.dbg_ts = d.dbg_ts;
    if (splits > 1) p.epi.flags |= IG_SPLITK;
    const int c_tiles = (d.epi.n_valid + IG_BM - 1) / IG_BM;
    p.n_pad = c_tiles * IG_BM;
    if ((d.epi.n_valid & 7) || (d.epi.ldc & 7) || (d.epi.res && (d.epi.ldr & 7)) || (d.epi.colbias && (d.epi.colbias_bstride & 3)) ||
        (reinterpret_cast<uintptr_t>(d.epi.out) & 15) || (reinterpret_cast<uintptr_t>(d.epi.res) & 15) ||
        (reinterpret_cast<uintptr_t>(d.epi.colbias) & 15)) {
        b2_set_error("igemm(swap): epilogue needs n_valid/ldc/ldr multiples of 8 and 16-byte aligned pointers");
        return -1;
    }
    const size_t stage_bytes = (size_t)IG_BM * IG_BK * 2 + (size_t)BN * IG_BK * 2;
    // swapped launches are small grids (<= ~1 CTA per SM): give the ring most of the shared memory
    static const char* sw_stage_env = getenv("B2_SWAP_STAGE_KB");
    int stages = (int)(((size_t)(sw_stage_env ? atoi(sw_stage_env) : (d.ring_kb > 0 ? d.ring_kb : 200)) * 1024) / stage_bytes);
    if (stages < 2) stages = 2;
    if (stages > IG_MAX_STAGES) stages = IG_MAX_STAGES;
    if (stages > p.kb_per_split) stages = p.kb_per_split < 2 ? 2 : p.kb_per_split;
    size_t pipe_bytes = stages * stage_bytes;
    {
        // epilogue scratch carved out of the ring: [split-K staging tile BN x 128 fp32] + transposition tile (<= 32 x 128 fp32)
        const int cols_per = BN / splits;
        const size_t need = (splits > 1 ? (size_t)BN * IG_BM * 4 : 0) + (size_t)(cols_per < SWAP_CH ? cols_per : SWAP_CH) * IG_BM * 4;
        if (need > pipe_bytes) {
            stages = (int)((need + stage_bytes - 1) / stage_bytes);
            if (stages > IG_MAX_STAGES) {
                b2_set_error("igemm(swap): split-K staging does not fit (BN %d)", BN);
                return -1;
            }
            pipe_bytes = stages * stage_bytes;
        }
    }
    p.num_stages = stages;
    plan->smem = pipe_bytes + 1024 + 512;
    uint32_t cols = 32;
    while (cols < (uint32_t)BN) cols <<= 1;
    p.tmem_cols = cols;
    plan->grid = dim3(p.tiles_w * p.tiles_h * p.tiles_n, c_tiles, splits);
    plan->mode = 0;
    return 0;
}

int igemm_plan(const IgemmDesc& d, IgemmPlan* plan) {
    *plan = IgemmPlan{};
    if (d.swap) return plan_swap(d, plan);
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
    const bool pair = d.pair != 0;
    if (pair && BN % 32 != 0) {
        b2_set_error("igemm(pair): BN %d must be a multiple of 32", BN);
        return -1;
    }
    p.BN = BN;
    { static const char* dm = getenv("B2_DBG_MODE"); p.dbg_mode = dm ? atoi(dm) : 0; }
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
    const int b_rows = pair ? BN / 2 : BN;   // weight rows one CTA stages per K-block
    if (encode_w_map(&p.tmB, d.w, d.w_rows, d.w_ld, b_rows)) return -1;
    p.a_bytes = (uint32_t)(tw * th * tn) * IG_BK * 2;
    p.b_bytes = (uint32_t)b_rows * IG_BK * 2;
    // ---- split-K
    int splits = d.splits < 1 ? 1 : d.splits;
    if (splits > total_kb) splits = total_kb;
    if (pair && splits > 4) splits = 4;   // cluster = 2 x splits CTAs, portable limit 8
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
    p.dbg_ts = d.dbg_ts;
    p.n_pad = n_tiles * BN;
    if (splits > 1) {
        if (geglu) {
            b2_set_error("igemm: split-K cannot be combined with GEGLU");
            return -1;
        }
        p.epi.flags |= IG_SPLITK;
    }
    // ---- pipeline depth / smem
    // one K-block (64 channels of one tap) per pipeline stage: packing several per stage was measured slower (shallower
    // prefetch, longer MMA issue code)
    const size_t stage_bytes = (size_t)IG_BM * IG_BK * 2 + (size_t)b_rows * IG_BK * 2;
    // TMA latency under load is ~1.3 us (tools/timeline.py): throughput per SM = bytes in flight / latency.  With at
    // most ~1 CTA per SM take the whole shared memory for the ring; with many CTAs keep two co-resident instead.
    static const char* pc_env = getenv("B2_PERSIST_CTAS");   // tuning: resident CTAs of a persistent launch (default 2 per SM)
    const int persist_ctas = pc_env ? atoi(pc_env) : 2 * 148;
    const long all_tiles = (long)p.tiles_w * p.tiles_h * p.tiles_n * n_tiles;
    static const bool no_persist_e = getenv("B2_NO_PERSIST") != nullptr;
    const bool will_persist = !no_persist_e && splits == 1 && all_tiles > 2 * 148 && 2 * BN <= 512;
    const long total_ctas = will_persist ? persist_ctas : all_tiles * splits;
    static const char* stage_env = getenv("B2_STAGE_KB");
    // <= 1 CTA per SM anyway: take (nearly) all the shared memory for the ring, the mainloop is TMA-latency bound
    const size_t ring_budget = stage_env ? (size_t)atoi(stage_env) * 1024
                                         : (d.ring_kb > 0 ? (size_t)d.ring_kb * 1024 : (size_t)((total_ctas <= 148 ? 200 : 100) * 1024));
    int stages = (int)(ring_budget / stage_bytes);
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
    plan->smem = pipe_bytes + 1024 /*align slack*/ + 512 /*barriers*/ + (d.epi.colsum ? 2 * 256 * sizeof(float) : 0) /*LN vectors*/;
    // persistent over M tiles when the launch would need more than one co-resident wave (2 CTAs per SM)
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    static const bool no_persist = getenv("B2_NO_PERSIST") != nullptr;
    int grid_x = m_tiles;
    p.acc_bufs = 1;
    if (!no_persist && splits == 1 && (long)m_tiles * n_tiles > 2 * 148 && 2 * BN <= 512) {
        grid_x = persist_ctas / n_tiles;
        if (grid_x < 1) grid_x = 1;
        if (grid_x > m_tiles) grid_x = m_tiles;
        p.acc_bufs = 2;
    }
    if (pair) {   // CTAs (2j, 2j+1) of x are a pair: an odd tile count leaves one masked tile; a persistent launch stays within
        // the co-resident wave (round down)
        grid_x = (p.acc_bufs == 2 && grid_x > 2) ? (grid_x & ~1) : ((grid_x + 1) & ~1);
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)(BN * p.acc_bufs)) cols <<= 1;
    p.tmem_cols = cols;
    plan->grid = dim3(grid_x, n_tiles, splits);
    plan->mode = pair ? 1 : 0;
    plan->pair = pair ? 1 : 0;
    return 0;
}

int igemm_encode_act_map(CUtensorMap* m, const ActView& a, int box_c, int box_w, int box_h, int box_n, int estride) {
    return encode_act_map(m, a, box_c, box_w, box_h, box_n, estride);
}
int igemm_encode_w_map(CUtensorMap* m, const __half* w, int rows, int ld, int box_rows) { return encode_w_map(m, w, rows, ld, box_rows); }

int igemm_init() {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) {
            b2_set_error("cudaFuncSetAttribute(igemm): %s", cudaGetErrorString(e));
            return -1;
        }
        e = cudaFuncSetAttribute(igemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            b2_set_error("cudaFuncSetAttribute(igemm pair): %s", cudaGetErrorString(e));
            return -1;
        }
        if (!get_encode()) return -1;
        attr_set = true;
    }
    return 0;
}

int igemm_launch(const IgemmPlan& plan, cudaStream_t stream) {
    if (igemm_init()) return -1;
    const int cz = plan.splits > 1 ? plan.splits : 1;
    cudaError_t e = plan.pair ? launch_kc(igemm_pair_kernel, plan.grid, dim3(IG_THREADS), plan.smem, stream, 2, cz, plan.p)
                              : launch_k(igemm_kernel, plan.grid, dim3(IG_THREADS), plan.smem, stream, cz, plan.p);
    if (e != cudaSuccess) {
        b2_set_error("igemm launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace b2
