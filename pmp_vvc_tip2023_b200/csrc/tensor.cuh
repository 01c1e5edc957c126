// Activation tensor views shared by the conv engines and the element-wise kernels (internal).
//
// Formats
//   FMT_F32   : fp32 NCHW                                   (SIMT engine activations, net outputs)
//   FMT_U8    : uint8 NCHW                                  (pixel blocks)
//   FMT_SPLIT : split-precision planar-8 layout, 16-bit elements.   (TC engine activations)
//               Layout [N][C/16][4][H][W][8]: per group of 16 channels four planes of 8 channels each:
//               plane 0 = hi(c0..7), 1 = hi(c8..15), 2 = lo(c0..7), 3 = lo(c8..15), hi = rn16(a),
//               lo = rn16(a - hi); a ~= hi + lo to ~22 (fp16) / ~16 (bf16) mantissa bits.  C is padded to a
//               multiple of 16 (padding channels are zero).  8 channels x 16 bit = one 16-byte unit, pixels
//               of a row are contiguous: a K-major UMMA core matrix (8 pixels x 8 channels) is 128 contiguous
//               bytes and ONE TMA box (8, W+halo, rows, 4, 1) per K=16 step lands hi and lo operands of a
//               halo tile directly in the no-swizzle canonical layout (OOB zero fill == conv zero padding).
//   FMT_PAIR  : two fp32 planes per image addressed by separate base pointers + batch stride
//               (the [B,2,16,16] module outputs or the regrouped bt/dire [B,3,16,16] of inference_pre_QBD).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace pmp {

enum Fmt { FMT_F32 = 0, FMT_U8 = 1, FMT_SPLIT = 2, FMT_PAIR = 3 };

struct Act {
    void *p = nullptr;
    void *p2 = nullptr;     // FMT_PAIR: channel-1 plane
    int fmt = FMT_F32;
    int C = 0;              // logical channels
    int Cp = 0;             // stored channels (FMT_SPLIT: padded to 16; else == C)
    int H = 0, W = 0;
    long long bstride = 0;  // FMT_PAIR: elements between images
    int bf16 = 0;           // FMT_SPLIT element type: 0 fp16, 1 bf16
    size_t bytes = 0;       // allocation size (arena bookkeeping)
};

__host__ __device__ inline int pad16(int c) { return (c + 15) & ~15; }
// plane index of 8-channel chunk `ch` (hi: lo=0, lo: lo=1) inside an image of a FMT_SPLIT tensor
__host__ __device__ inline int split_plane(int ch, int lo) { return ((ch >> 1) << 2) + (lo << 1) + (ch & 1); }

// 16-bit split helpers -------------------------------------------------------------------------
__device__ __forceinline__ void split16(float a, bool bf, uint16_t &hi, uint16_t &lo)
{
    if (bf) {
        __nv_bfloat16 h = __float2bfloat16_rn(a);
        float r = a - __bfloat162float(h);
        __nv_bfloat16 l = __float2bfloat16_rn(r);
        hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(l);
    } else {
        a = fminf(fmaxf(a, -65504.f), 65504.f);          // fp16 range guard (trained nets peak at ~4e3)
        __half h = __float2half_rn(a);
        float r = a - __half2float(h);
        __half l = __float2half_rn(r);
        hi = __half_as_ushort(h); lo = __half_as_ushort(l);
    }
}
__device__ __forceinline__ float join16(uint16_t hi, uint16_t lo, bool bf)
{
    if (bf) return __bfloat162float(__ushort_as_bfloat16(hi)) + __bfloat162float(__ushort_as_bfloat16(lo));
    return __half2float(__ushort_as_half(hi)) + __half2float(__ushort_as_half(lo));
}

// 8 channels (one chunk) of one pixel -----------------------------------------------------------
__device__ __forceinline__ void store_chunk_split(const Act &t, int n, int chunk, int y, int x, const float v[8])
{
    const int npl = t.Cp >> 2;
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; e++) split16(v[e], t.bf16, hi[e], lo[e]);
    uint4 H, L;
    H.x = hi[0] | ((uint32_t)hi[1] << 16); H.y = hi[2] | ((uint32_t)hi[3] << 16);
    H.z = hi[4] | ((uint32_t)hi[5] << 16); H.w = hi[6] | ((uint32_t)hi[7] << 16);
    L.x = lo[0] | ((uint32_t)lo[1] << 16); L.y = lo[2] | ((uint32_t)lo[3] << 16);
    L.z = lo[4] | ((uint32_t)lo[5] << 16); L.w = lo[6] | ((uint32_t)lo[7] << 16);
    uint4 *base = reinterpret_cast<uint4 *>(t.p);
    size_t plane = (size_t)t.H * t.W;
    size_t o = ((size_t)n * npl + split_plane(chunk, 0)) * plane + (size_t)y * t.W + x;
    base[o] = H;
    base[o + 2 * plane] = L;
}
__device__ __forceinline__ void load_chunk_split(const Act &t, int n, int chunk, int y, int x, float v[8])
{
    const int npl = t.Cp >> 2;
    const uint4 *base = reinterpret_cast<const uint4 *>(t.p);
    size_t plane = (size_t)t.H * t.W;
    size_t o = ((size_t)n * npl + split_plane(chunk, 0)) * plane + (size_t)y * t.W + x;
    uint4 H = base[o], L = base[o + 2 * plane];
    uint32_t hw[4] = {H.x, H.y, H.z, H.w}, lw[4] = {L.x, L.y, L.z, L.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        v[2 * e] = join16((uint16_t)(hw[e] & 0xffff), (uint16_t)(lw[e] & 0xffff), t.bf16);
        v[2 * e + 1] = join16((uint16_t)(hw[e] >> 16), (uint16_t)(lw[e] >> 16), t.bf16);
    }
}

// scalar element access (slow path: tile loads of the SIMT kernels, element-wise kernels) -------------
__device__ __forceinline__ float load_elem(const Act &t, int n, int c, int y, int x)
{
    if (t.fmt == FMT_F32) return reinterpret_cast<const float *>(t.p)[(((size_t)n * t.C + c) * t.H + y) * t.W + x];
    if (t.fmt == FMT_U8) return (float)reinterpret_cast<const uint8_t *>(t.p)[(((size_t)n * t.C + c) * t.H + y) * t.W + x];
    if (t.fmt == FMT_SPLIT) {
        const int npl = t.Cp >> 2;
        const uint16_t *b = reinterpret_cast<const uint16_t *>(t.p);
        size_t plane = (size_t)t.H * t.W;
        size_t o = ((((size_t)n * npl + split_plane(c >> 3, 0)) * plane + (size_t)y * t.W + x) << 3) + (c & 7);
        return join16(b[o], b[o + (plane << 4)], t.bf16);
    }
    const float *b = reinterpret_cast<const float *>(c == 0 ? t.p : t.p2);       // FMT_PAIR
    return b[(size_t)n * t.bstride + (size_t)y * t.W + x];
}
__device__ __forceinline__ void store_elem_f32(const Act &t, int n, int c, int y, int x, float v)
{
    if (t.fmt == FMT_F32) reinterpret_cast<float *>(t.p)[(((size_t)n * t.C + c) * t.H + y) * t.W + x] = v;
    else {
        float *b = reinterpret_cast<float *>(c == 0 ? t.p : t.p2);
        b[(size_t)n * t.bstride + (size_t)y * t.W + x] = v;
    }
}

inline size_t act_bytes(int fmt, int B, int C, int H, int W)
{
    if (fmt == FMT_SPLIT) return (size_t)B * pad16(C) * H * W * 4;        // hi + lo, 2 bytes each
    if (fmt == FMT_U8) return (size_t)B * C * H * W;
    return (size_t)B * C * H * W * 4;
}

}  // namespace pmp
