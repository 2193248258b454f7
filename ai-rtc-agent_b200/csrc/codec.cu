// NV12 <-> RGB colour conversion kernels and the codec library probe (see codec.cuh).
#include "codec.cuh"

#include <dlfcn.h>

#include "igemm.cuh"   // b2_set_error

namespace b2 {

struct Csc {   // Y' = ky + (r*kr + g*kg + b*kb) * sy ; Cb/Cr = 128 + (B' - Y') / ... expressed as matrices both ways
    float kr, kb;       // luma coefficients (kg = 1 - kr - kb)
    float y_off, y_scale, c_scale;   // limited range: 16, 219/255, 224/255; full range: 0, 1, 1
};
__host__ __device__ inline Csc make_csc(int flags) {
    Csc c;
    if (flags & CSC_BT601) { c.kr = 0.299f; c.kb = 0.114f; } else { c.kr = 0.2126f; c.kb = 0.0722f; }
    if (flags & CSC_FULL_RANGE) { c.y_off = 0.f; c.y_scale = 1.f; c.c_scale = 1.f; }
    else { c.y_off = 16.f; c.y_scale = 219.f / 255.f; c.c_scale = 224.f / 255.f; }
    return c;
}
__device__ __forceinline__ uint8_t sat_u8(float v) { return (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f)); }

// one thread = one 2x2 pixel quad (one chroma sample): coalesced 2-byte Y / UV reads, 6-byte RGB writes per row
__global__ void nv12_to_rgb_u8_kernel(const uint8_t* __restrict__ yp, int y_pitch, const uint8_t* __restrict__ uvp, int uv_pitch,
                                      uint8_t* __restrict__ rgb, int h, int w, int flags) {
    const int qx = blockIdx.x * blockDim.x + threadIdx.x, qy = blockIdx.y * blockDim.y + threadIdx.y;
    if (2 * qx >= w || 2 * qy >= h) return;
    const Csc c = make_csc(flags);
    const float kg = 1.f - c.kr - c.kb;
    const float cb = ((float)uvp[(size_t)qy * uv_pitch + 2 * qx] - 128.f) / c.c_scale;
    const float cr = ((float)uvp[(size_t)qy * uv_pitch + 2 * qx + 1] - 128.f) / c.c_scale;
    const float r_add = 2.f * (1.f - c.kr) * cr, b_add = 2.f * (1.f - c.kb) * cb;
    const float g_add = -(c.kr * r_add + c.kb * b_add) / kg;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int x = 2 * qx + dx, y = 2 * qy + dy;
            if (x >= w || y >= h) continue;
            const float yy = ((float)yp[(size_t)y * y_pitch + x] - c.y_off) / c.y_scale;
            uint8_t* o = rgb + ((size_t)y * w + x) * 3;
            o[0] = sat_u8(yy + r_add);
            o[1] = sat_u8(yy + g_add);
            o[2] = sat_u8(yy + b_add);
        }
}

__global__ void rgb_u8_to_nv12_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ yp, int y_pitch,
                                      uint8_t* __restrict__ uvp, int uv_pitch, int h, int w, int flags) {
    const int qx = blockIdx.x * blockDim.x + threadIdx.x, qy = blockIdx.y * blockDim.y + threadIdx.y;
    if (2 * qx >= w || 2 * qy >= h) return;
    const Csc c = make_csc(flags);
    const float kg = 1.f - c.kr - c.kb;
    const size_t plane = (size_t)h * w;
    float cb = 0.f, cr = 0.f;
    int n = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int x = 2 * qx + dx, y = 2 * qy + dy;
            if (x >= w || y >= h) continue;
            const size_t i = (size_t)y * w + x;
            const float r = rgb[i], g = rgb[plane + i], b = rgb[2 * plane + i];
            const float yl = c.kr * r + kg * g + c.kb * b;
            yp[(size_t)y * y_pitch + x] = sat_u8(c.y_off + c.y_scale * yl);
            cb += (b - yl) / (2.f * (1.f - c.kb));
            cr += (r - yl) / (2.f * (1.f - c.kr));
            ++n;
        }
    uvp[(size_t)qy * uv_pitch + 2 * qx] = sat_u8(128.f + c.c_scale * cb / n);
    uvp[(size_t)qy * uv_pitch + 2 * qx + 1] = sat_u8(128.f + c.c_scale * cr / n);
}

int nv12_to_rgb_u8_launch(const uint8_t* y, int y_pitch, const uint8_t* uv, int uv_pitch, uint8_t* rgb_nhwc, int h, int w,
                          int flags, cudaStream_t s) {
    if (!y || !uv || !rgb_nhwc || h < 2 || w < 2 || y_pitch < w || uv_pitch < ((w + 1) & ~1)) {
        b2_set_error("nv12_to_rgb: bad arguments (h %d w %d pitches %d %d)", h, w, y_pitch, uv_pitch);
        return -1;
    }
    dim3 block(32, 8), grid(((w + 1) / 2 + 31) / 32, ((h + 1) / 2 + 7) / 8);
    nv12_to_rgb_u8_kernel<<<grid, block, 0, s>>>(y, y_pitch, uv, uv_pitch, rgb_nhwc, h, w, flags);
    if (cudaGetLastError() != cudaSuccess) { b2_set_error("nv12_to_rgb launch failed"); return -1; }
    return 0;
}

int rgb_u8_to_nv12_launch(const uint8_t* rgb_nchw, uint8_t* y, int y_pitch, uint8_t* uv, int uv_pitch, int h, int w, int flags,
                          cudaStream_t s) {
    if (!y || !uv || !rgb_nchw || h < 2 || w < 2 || y_pitch < w || uv_pitch < ((w + 1) & ~1)) {
        b2_set_error("rgb_to_nv12: bad arguments (h %d w %d pitches %d %d)", h, w, y_pitch, uv_pitch);
        return -1;
    }
    dim3 block(32, 8), grid(((w + 1) / 2 + 31) / 32, ((h + 1) / 2 + 7) / 8);
    rgb_u8_to_nv12_kernel<<<grid, block, 0, s>>>(rgb_nchw, y, y_pitch, uv, uv_pitch, h, w, flags);
    if (cudaGetLastError() != cudaSuccess) { b2_set_error("rgb_to_nv12 launch failed"); return -1; }
    return 0;
}

int codec_probe() {
    int mask = 0;
    const char* dec[] = {"libnvcuvid.so.1", "libnvcuvid.so"};
    const char* enc[] = {"libnvidia-encode.so.1", "libnvidia-encode.so"};
    for (const char* n : dec)
        if (void* hnd = dlopen(n, RTLD_LAZY | RTLD_LOCAL)) { mask |= 1; dlclose(hnd); break; }
    for (const char* n : enc)
        if (void* hnd = dlopen(n, RTLD_LAZY | RTLD_LOCAL)) { mask |= 2; dlclose(hnd); break; }
    return mask;
}

}  // namespace b2
