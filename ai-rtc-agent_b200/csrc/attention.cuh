// Flash-style attention on tcgen05: S = Q K^T and O += P V as UMMA (accumulators in TMEM), online
// softmax in registers by 4 warps (one query row per thread == one TMEM lane), P staged through
// shared memory in the K-major SWIZZLE_128B layout.  Self-attention (seq 64..9216) and
// cross-attention against the cached prompt K/V (77 keys) use the same kernel.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace b2 {

struct AttnDesc {
    const __half* q;    // [nb*sq][ldq], head h occupies columns [h*dp, h*dp+dp)
    int ldq;
    const __half* k;    // [rows][ldk], same head layout; batch b starts at row b*k_bstride
    int ldk;
    long k_bstride;     // rows; 0 => K/V shared by every batch item (prompt cache)
    long k_rows;        // total rows addressable
    const __half* vt;   // V^T: [heads*dp][ldvt], kv index contiguous; batch b starts at column b*vt_bstride
    int ldvt;
    long vt_bstride;
    long vt_cols;       // total valid columns
    __half* out;        // [nb*sq][ldo], head h occupies columns [h*d_real, (h+1)*d_real)
    int ldo;
    int nb, heads, sq, skv;
    int d_real;         // true head dim (softmax scale = d_real^-0.5)
    int dp;             // padded head dim: 64, 128 or 192 (zero-padded columns)
};

struct AttnPlan {
    CUtensorMap tmq, tmk, tmv;
    AttnDesc d;
    dim3 grid;
    size_t smem;
};

int attn_plan(const AttnDesc& d, AttnPlan* plan);
int attn_launch(const AttnPlan& plan, cudaStream_t s);
int attn_init();

}  // namespace b2
