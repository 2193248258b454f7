/*
 * libb200sd -- C ABI of the B200-native per-frame img2img path.
 *
 * The reference (yondonfu/ai-rtc-agent) has no C/FFI boundary of its own: its drop-in boundary is
 * two Python classes (lib/pipeline.py:17-96 StreamDiffusionPipeline, lib/wrapper.py:34-407
 * StreamDiffusionWrapper) that reach the GPU through TensorRT engines built by the un-vendored
 * `streamdiffusion` package.  This header is what the Python shim in
 * ai-rtc-agent_b200/host/ binds with ctypes; every entry point names the
 * reference call it stands in for.
 *
 * Conventions: plain pointers and sizes only (no torch types); all device pointers are CUDA device
 * memory of the current device; `stream` is a cudaStream_t passed as void*; every function returns 0
 * on success and non-zero on failure, with the message available from b2sd_last_error().
 * No function synchronises the device unless its comment says so.
 */
#ifndef B200SD_H
#define B200SD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* b2sd_last_error(void);
int b2sd_version(void);

/* ------------------------------------------------------------------------------------------------
 * Operator level (the contractions inside the reference's unet.engine / vae_*.engine,
 * lib/wrapper.py:445-466): used by the parity tests and by the engine below.
 * ---------------------------------------------------------------------------------------------- */

/* NHWC fp16 view: element (n,h,w,c) at ptr[((n*H + h)*W + w)*ld + c] */
typedef struct {
    const void* ptr;
    int n, h, w, c;
    int ld;
} b2sd_act_view;

enum { B2SD_IG_RELU = 1, B2SD_IG_GEGLU = 2 };

/* conv3x3 / conv1x1 / Linear as one implicit GEMM:
 *   out[row][j] = acc_scale * (sum_seg sum_tap sum_c src[seg](row, tap, c) * w[j][k] + colbias[b][j])
 *                 + res_scale * res[row][j]        (then ReLU / GEGLU per flags)
 * K order of the packed weight rows = segments in sequence, each [tap][c]. */
typedef struct {
    b2sd_act_view src[3];
    int ntap[3];       /* 1 or 9 (3x3, pad 1) */
    int nseg;
    const void* w;     /* fp16 [w_rows][w_ld] */
    int w_rows, w_ld;
    int stride;        /* 1 or 2 */
    int nb, ho, wo;    /* output extents */
    int bn;            /* N tile, 0 = auto */
    int splits;        /* split-K factor, <=1 = off */
    void* partial;     /* fp32 workspace, b2sd_igemm_partial_floats() floats, when splits > 1 */
    void* out;         /* fp16 [nb*ho*wo][ldc] */
    int ldc;
    const float* colbias;
    int colbias_bstride;
    const void* res;   /* fp16, indexed like out with pitch ldr */
    int ldr;
    float acc_scale, res_scale;
    int flags;
    int n_valid;       /* output channels */
} b2sd_igemm_desc;

int b2sd_op_igemm(const b2sd_igemm_desc* d, void* stream);
uint64_t b2sd_igemm_partial_floats(int splits, int64_t rows_total, int n_valid);

/* Flash attention (self / cross) of BasicTransformerBlock.attn1 / attn2 (inside unet.engine).
 * q: [nb*sq][ldq], head h at columns [h*dp, (h+1)*dp); k likewise (batch b at row b*k_bstride, 0 = shared);
 * vt = V^T: [heads*dp][ldvt] with the key index contiguous (batch b at column b*vt_bstride);
 * out: [nb*sq][ldo], head h at columns [h*d_real, (h+1)*d_real). softmax scale = d_real^-0.5. */
typedef struct {
    const void* q; int ldq;
    const void* k; int ldk; int64_t k_bstride; int64_t k_rows;
    const void* vt; int ldvt; int64_t vt_bstride; int64_t vt_cols;
    void* out; int ldo;
    int nb, heads, sq, skv, d_real, dp;
} b2sd_attn_desc;
int b2sd_op_attention(const b2sd_attn_desc* d, void* stream);

/* GroupNorm(+SiLU) over the channel concatenation [xa | xb] (xb may be NULL), NHWC fp16. */
int b2sd_op_groupnorm(const void* xa, int ca, int lda, const void* xb, int cb, int ldb, const float* gamma,
                      const float* beta, void* y, int ldy, int nb, int hw, int groups, float eps, int silu,
                      void* stream);
int b2sd_op_layernorm(const void* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                      int64_t rows, int c, float eps, void* stream);
int b2sd_op_upsample2x(const void* x, void* y, int nb, int h, int w, int c, void* stream);
/* direct 3x3 conv for Cin in {3,4}; flags: 1 = input is u8 NHWC scaled by 1/255 (lib/pipeline.py:61),
 * 2 = tanh(x/3)*3 on the input (DecoderTiny), 4 = ReLU on the output */
int b2sd_op_smallconv(const void* x, const void* w_oihw, const float* bias, void* y, int ldy, int nb, int h,
                      int w, int cin, int cout, int in_h, int in_w, int flags, void* stream);
/* StreamDiffusion scheduler_step_batch + stream-batch buffer update (see elementwise.cuh) */
int b2sd_op_lcm_step(void* x, const void* eps, const void* noise, const float* coef, void* out_latent, int T,
                     int hw, int do_add_noise, void* stream);
/* decoder tail + lib/pipeline.py:72-74 on the fp16 grid -> u8 NCHW */
int b2sd_op_post_u8(const void* y_nhwc, int ldy, void* out_nchw_u8, int nb, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SD_H */
