// Epilogue helpers shared by the tcgen05 contraction kernels (igemm.cu, tconv.cu): accumulator row (TMEM lane = thread)
// -> bias / scale / residual / ReLU / GEGLU -> fp16 NHWC stores.
#pragma once
#include "igemm.cuh"
#include "ptx.cuh"

namespace b2 {

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

// LayerNorm statistics of row `orow` from the fixed-point sums its producer accumulated (see IgEpilogue)
__device__ __forceinline__ void ln_row_stats(const IgEpilogue& e, long orow, bool row_ok, float& mu, float& rstd) {
    mu = 0.f;
    rstd = 1.f;
    if (!e.rowstat_in || !row_ok) return;
    const ulonglong2 st = *reinterpret_cast<const ulonglong2*>(e.rowstat_in + 2 * orow);
    const float s1 = (float)((double)(long long)st.x * (1.0 / IG_STAT_SCALE));
    const float s2 = (float)((double)(long long)st.y * (1.0 / IG_STAT_SCALE));
    mu = s1 * e.ln_inv_c;
    const float var = fmaxf(s2 * e.ln_inv_c - mu * mu, 0.f);
    rstd = rsqrtf(var + e.ln_eps);
}
__device__ __forceinline__ void rowstat_add(const IgEpilogue& e, long orow, float s1, float s2) {
    atomicAdd(e.rowstat_out + 2 * orow, (unsigned long long)__float2ll_rn(s1 * IG_STAT_SCALE));
    atomicAdd(e.rowstat_out + 2 * orow + 1, (unsigned long long)__float2ll_rn(s2 * IG_STAT_SCALE));
}

// 16 accumulator columns [col0, col0+16) of output row `orow` (batch item b); the accumulators are
// acc[OFF .. OFF+16) of a register array (compile-time indices only: nothing may spill to local memory).
template <int OFF, int N, typename T>
__device__ __forceinline__ void epi_store16(const IgEpilogue& e, const T (&acc)[N], int b, long orow, int col0, float mu = 0.f,
                                            float rstd = 1.f) {
    int nv = e.n_valid - col0;
    if (nv <= 0) return;
    if (nv > 16) nv = 16;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if constexpr (sizeof(T) == 4 && !__is_same(T, float)) v[i] = __uint_as_float(acc[OFF + i]);
        else v[i] = acc[OFF + i];
    }
    if (e.colsum) {   // folded LayerNorm of the A rows
        const float* cs = e.colsum + col0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) v[i] = rstd * (v[i] - mu * cs[i]);
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
    if (e.out2 && col0 >= e.col2) {   // V block of the fused q/k/v projection: transposed store (32 lanes = 32 consecutive tokens)
        __half* tp = e.out2 + (long)(col0 - e.col2) * e.ld2 + orow;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) tp[(long)i * e.ld2] = __float2half_rn(v[i]);
        return;
    }
    if (e.rowstat_out) {   // statistics of the values as stored (fp16-rounded), like a LayerNorm reading them back
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) {
                const float x = __half2float(__float2half_rn(v[i]));
                s1 += x;
                s2 += x * x;
            }
        rowstat_add(e, orow, s1, s2);
    }
    store_half16(e.out + orow * e.ldc + col0, v, nv, (e.ldc & 7) == 0);
}

// GEGLU: val/gate are 16 accumulator columns each; packed-column index of val[0] is pcol0 (bias
// uses packed indexing), output column index is ocol0.
__device__ __forceinline__ void epi_store16_geglu(const IgEpilogue& e, const uint32_t (&val)[16],
                                                  const uint32_t (&gate)[16], long orow, int pcol_val,
                                                  int pcol_gate, int ocol0, float mu = 0.f, float rstd = 1.f) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(val[i]), g = __uint_as_float(gate[i]);
        if (e.colsum) {   // folded LayerNorm (norm3) of the A rows
            a = rstd * (a - mu * e.colsum[pcol_val + i]);
            g = rstd * (g - mu * e.colsum[pcol_gate + i]);
        }
        if (e.colbias) {
            a += e.colbias[pcol_val + i];
            g += e.colbias[pcol_gate + i];
        }
        v[i] = a * gelu_erf(g);
    }
    store_half16(e.out + orow * e.ldc + ocol0, v, 16, (e.ldc & 7) == 0);
}

// Fast epilogue of one output row (no split-K / GEGLU; n_valid % 16 == 0, vectorisable pitches): the residual row
// is prefetched 32 columns ahead -- the first chunk even before the accumulator is ready -- so its global-memory
// latency hides behind the mainloop instead of being paid once per 16-column chunk.
__device__ __forceinline__ void epi_row_fast(const IgEpilogue& e, uint32_t taddr, int ncols, int gcol0, int b, long orow,
                                             bool row_ok, uint64_t* wait_bar, uint32_t wait_parity = 0) {
    const bool has_res = e.res != nullptr && row_ok;
    const __half* rp = e.res ? e.res + orow * e.ldr + gcol0 : nullptr;
    const float* bp = e.colbias ? e.colbias + (long)b * e.colbias_bstride + gcol0 : nullptr;
    __half* op = e.out + orow * e.ldc + gcol0;
    uint4 rr[4];
    float4 bb[8];
    auto fetch = [&](int c, uint4 (&r4)[4], float4 (&b8)[8]) {   // residual + bias of columns [c, c+32)
#pragma unroll
        for (int i = 0; i < 4; ++i) r4[i] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) b8[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= ncols) return;
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (c + 8 * i < ncols) r4[i] = reinterpret_cast<const uint4*>(rp + c)[i];
        }
        if (bp) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c + 4 * i < ncols) b8[i] = reinterpret_cast<const float4*>(bp + c)[i];
        }
    };
    float st1 = 0.f, st2 = 0.f;
    fetch(0, rr, bb);
    if (wait_bar) {
        mbar_wait(wait_bar, wait_parity);
        tc_fence_after();
    }
    for (int c = 0; c < ncols; c += 32) {
        const int left = ncols - c;   // >= 16, multiple of 16
        uint32_t v[32];
        if (left >= 32) {
            tmem_ld32(taddr + c, v);
        } else {
            uint32_t lo[16];
            tmem_ld16(taddr + c, lo);
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = 0; }
        }
        uint4 rn[4];
        float4 bn[8];
        fetch(c + 32, rn, bn);   // next pass: latency hides behind this pass
        tmem_ld_wait();
        if (row_ok) {
            const float* bias = reinterpret_cast<const float*>(bb);
#pragma unroll
            for (int g = 0; g < 4; ++g) {   // 8 columns per 16-byte store
                if (8 * g < left) {
                    const __half2* rh = reinterpret_cast<const __half2*>(&rr[g]);
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = 8 * g + 2 * j;
                        float x0 = (__uint_as_float(v[i]) + bias[i]) * e.acc_scale;
                        float x1 = (__uint_as_float(v[i + 1]) + bias[i + 1]) * e.acc_scale;
                        if (e.res) {
                            const float2 f = __half22float2(rh[j]);
                            x0 += e.res_scale * f.x;
                            x1 += e.res_scale * f.y;
                        }
                        if (e.flags & IG_RELU) {
                            x0 = fmaxf(x0, 0.f);
                            x1 = fmaxf(x1, 0.f);
                        }
                        oh[j] = __floats2half2_rn(x0, x1);
                        if (e.rowstat_out) {   // LayerNorm statistics of the stored (fp16-rounded) row
                            const float2 f = __half22float2(oh[j]);
                            st1 += f.x + f.y;
                            st2 += f.x * f.x + f.y * f.y;
                        }
                    }
                    reinterpret_cast<uint4*>(op + c)[g] = o;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) rr[i] = rn[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = bn[i];
    }
    if (e.rowstat_out && row_ok && ncols > 0) rowstat_add(e, orow, st1, st2);
}

// LayerNorm-folded epilogue of one output row (non-split, see IgEpilogue::colsum): the tile's colsum / bias' vectors were
// staged in shared memory (`lnv`: [ncols_tile] colsum, then [ncols_tile] bias') while the mainloop ran, so the per-chunk loop
// has no global-memory loads on its critical path.  c_tile0 = first column of this call inside the staged tile.
__device__ __forceinline__ void epi_row_ln(const IgEpilogue& e, uint32_t taddr, int ncols, int gcol0, long orow, bool row_ok,
                                           const float* lnv, int tile_cols, float mu, float rstd) {
    const float* cs = lnv;
    const float* bb = lnv + tile_cols;
    const float nmr = -mu * rstd;
    for (int c = 0; c < ncols; c += 32) {
        const int left = ncols - c;
        uint32_t v[32];
        if (left >= 32) {
            tmem_ld32(taddr + c, v);
        } else {
            uint32_t lo[16];
            tmem_ld16(taddr + c, lo);
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = 0; }
        }
        tmem_ld_wait();
        if (!row_ok) continue;
        const bool transposed = e.out2 && gcol0 + c >= e.col2;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (8 * g >= left) break;
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = 8 * g + j;
                x[j] = fmaf(rstd, __uint_as_float(v[i]), fmaf(nmr, cs[c + i], bb[c + i]));   // rstd*acc - rstd*mu*colsum + bias'
            }
            if (transposed) {   // V block of the fused q/k/v projection -> V^T (32 lanes = 32 consecutive tokens)
                __half* tp = e.out2 + (long)(gcol0 + c + 8 * g - e.col2) * e.ld2 + orow;
#pragma unroll
                for (int j = 0; j < 8; ++j) tp[(long)j * e.ld2] = __float2half_rn(x[j]);
            } else {
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                *reinterpret_cast<uint4*>(e.out + orow * e.ldc + gcol0 + c + 8 * g) = o;
            }
        }
    }
}

// GEGLU with the LayerNorm (norm3) folded: value columns [0, half_n), gate columns [half_n, 2*half_n) of the tile
__device__ __forceinline__ void epi_row_geglu_ln(const IgEpilogue& e, uint32_t taddr, int half_n, int ocol0, long orow, bool row_ok,
                                                 const float* lnv, int tile_cols, float mu, float rstd) {
    const float* cs = lnv;
    const float* bb = lnv + tile_cols;
    const float nmr = -mu * rstd;
    for (int c = 0; c < half_n; c += 16) {
        uint32_t a[16], g[16];
        tmem_ld16(taddr + c, a);
        tmem_ld16(taddr + half_n + c, g);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float av = fmaf(rstd, __uint_as_float(a[i]), fmaf(nmr, cs[c + i], bb[c + i]));
            const float gv = fmaf(rstd, __uint_as_float(g[i]), fmaf(nmr, cs[half_n + c + i], bb[half_n + c + i]));
            v[i] = av * gelu_erf(gv);
        }
        store_half16(e.out + orow * e.ldc + ocol0 + c, v, 16, (e.ldc & 7) == 0);
    }
}

__device__ __forceinline__ bool epi_fast_ok(const IgEpilogue& e) {
    return !(e.flags & (IG_SPLITK | IG_GEGLU)) && (e.n_valid & 15) == 0 && (e.ldc & 7) == 0 && (!e.res || (e.ldr & 7) == 0) &&
           (e.colbias_bstride & 3) == 0 && !e.colsum && !e.out2;
}

}  // namespace b2
