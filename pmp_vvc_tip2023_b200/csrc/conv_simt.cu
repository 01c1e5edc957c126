// SIMT engine: exact-fp32 CUDA-core direct convolution with fused epilogue, plus the element-wise
// glue kernels of the four nets (pyramid pooling, up-sample + concat, stem input assembly).
//
// Reference semantics: nn.Conv2d stride 1 with explicit zero padding offsets (Model_QBD.py:27-37,
// :68-76,:108-121) followed by the ResidualBlock tail relu(conv2 + shortcut) (:40-44), optional
// F.max_pool2d(.,2) (:81,:82,:89,:136,:137,:151) and the attention product (:143,:150), fused.
//
// This kernel is (a) the whole conv path of PMP_ENGINE_SIMT (fp32 NCHW everywhere) and (b) inside
// PMP_ENGINE_TC the path of the layers that do not fit tensor cores: the stems (Cin <= 4, u8/f32 in,
// split out), the 8x8 tail of the Q nets and the Cout <= 2 output convs (split in, fp32 out).
#include <cstdlib>
#include "handle.cuh"
#include "tensor.cuh"
#include "kernels.cuh"

namespace pmp {

constexpr int ST_TILE = 16;       // output tile edge
constexpr int ST_CO = 32;         // output channels per CTA
constexpr int ST_TWP = 24;        // padded tile row stride (floats): rows land 8 banks apart

template <int KH, int KW, int ST_CI>       // ST_CI: input channels per smem stage
__global__ void __launch_bounds__(256)
conv_simt_kernel(SimtConvArgs a)
{
    constexpr int TH = ST_TILE + KH - 1, TW = ST_TILE + KW - 1;
    static_assert(TW <= ST_TWP, "tile row too wide");
    __shared__ float s_in[ST_CI][TH][ST_TWP];
    __shared__ __align__(16) float s_w[ST_CI][KH * KW][ST_CO];

    const int tid = threadIdx.x;
    const int cg = tid >> 6;                 // channel group: 8 output channels
    const int t = tid & 63, ty = t >> 3, tx = t & 7;
    const int tiles_x = (a.Wo + ST_TILE - 1) / ST_TILE;
    const int oy0 = (blockIdx.x / tiles_x) * ST_TILE, ox0 = (blockIdx.x % tiles_x) * ST_TILE;
    const int co0 = blockIdx.y * ST_CO;
    const int n = blockIdx.z;

    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[p][c] = 0.f;

    for (int ci0 = 0; ci0 < a.cin; ci0 += ST_CI) {
        // ---- stage input halo tile (zero outside the image == the layer's zero padding) ----
        for (int k = tid; k < ST_CI * TH * TW; k += 256) {
            int ci = k / (TH * TW), r = (k / TW) % TH, c = k % TW;
            int iy = oy0 + r - a.pad_t, ix = ox0 + c - a.pad_l;
            float v = 0.f;
            if (ci0 + ci < a.cin && iy >= 0 && iy < a.in.H && ix >= 0 && ix < a.in.W)
                v = load_elem(a.in, n, ci0 + ci, iy, ix);
            s_in[ci][r][c] = v;
        }
        // ---- stage weights [ci][tap][co] (global layout [cin][kh*kw][coutw], zero padded) ----
        for (int k = tid; k < ST_CI * KH * KW * ST_CO; k += 256) {
            int ci = k / (KH * KW * ST_CO), rem = k % (KH * KW * ST_CO);
            int tap = rem / ST_CO, co = rem % ST_CO;
            float v = 0.f;
            if (ci0 + ci < a.cin && co0 + co < a.coutw) v = a.w[((size_t)(ci0 + ci) * KH * KW + tap) * a.coutw + co0 + co];
            s_w[ci][tap][co] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < ST_CI; ci++) {
#pragma unroll
            for (int ky = 0; ky < KH; ky++) {
#pragma unroll
                for (int kx = 0; kx < KW; kx++) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(&s_w[ci][ky * KW + kx][cg * 8]);
                    const float4 w1 = *reinterpret_cast<const float4 *>(&s_w[ci][ky * KW + kx][cg * 8 + 4]);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const float v = s_in[ci][ty + 8 * (p >> 1) + ky][tx + 8 * (p & 1) + kx];
#pragma unroll
                        for (int c = 0; c < 8; c++) acc[p][c] = fmaf(v, wv[c], acc[p][c]);
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue: bias, residual, ReLU, max-pool 2x2, attention product, store ----
    const int cbase = co0 + cg * 8;
    const int lane = tid & 31;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int oy = oy0 + ty + 8 * (p >> 1), ox = ox0 + tx + 8 * (p & 1);
        const bool inb = (oy < a.Ho && ox < a.Wo);
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            v[c] = acc[p][c];
            if (a.bias && cbase + c < a.cout) v[c] += a.bias[cbase + c];
        }
        if (a.res.p && inb) {
            if (a.res.fmt == FMT_SPLIT) {
                if (cbase < a.res.Cp) {
                    float r[8];
                    load_chunk_split(a.res, n, cbase >> 3, oy, ox, r);
#pragma unroll
                    for (int c = 0; c < 8; c++) v[c] += r[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (cbase + c < a.cout) v[c] += load_elem(a.res, n, cbase + c, oy, ox);
            }
        }
        if (a.add0 && inb && cbase == 0) v[0] += a.add0[(size_t)n * a.add0_bstride + (size_t)oy * a.Wo + ox];
        if (a.relu) {
#pragma unroll
            for (int c = 0; c < 8; c++) v[c] = fmaxf(v[c], 0.f);
        }
        int sy = oy, sx = ox;
        bool writer = inb;
        if (a.pool == 2) {
            // the 2x2 window lives in lanes {l, l^1, l^8} (tx bit 0, ty bit 0); all lanes shuffle
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float o = inb ? v[c] : -3.0e38f;
                o = fmaxf(o, __shfl_xor_sync(0xffffffffu, o, 1));
                o = fmaxf(o, __shfl_xor_sync(0xffffffffu, o, 8));
                v[c] = o;
            }
            writer = inb && !(lane & 1) && !(lane & 8);
            sy = oy >> 1; sx = ox >> 1;
        }
        if (!writer) continue;
        if (a.mul.p) {
            if (a.mul.fmt == FMT_SPLIT) {
                if (cbase < a.mul.Cp) {
                    float r[8];
                    load_chunk_split(a.mul, n, cbase >> 3, sy, sx, r);
#pragma unroll
                    for (int c = 0; c < 8; c++) v[c] *= r[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (cbase + c < a.cout) v[c] *= load_elem(a.mul, n, cbase + c, sy, sx);
            }
        }
        if (a.out.fmt == FMT_SPLIT) {
            if (cbase < ((a.out.C == a.cout) ? a.out.Cp : a.cout)) {
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (cbase + c >= a.cout) v[c] = 0.f;          // padding channels stay exactly zero
                store_chunk_split(a.out, n, (a.out_c_off + cbase) >> 3, sy, sx, v);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (cbase + c < a.cout) store_elem_f32(a.out, n, a.out_c_off + cbase + c, sy, sx, v[c]);
        }
    }
}

template <int KH, int KW, int ST_CI>
static int launch_simt(Handle *h, const SimtConvArgs &a, int B, cudaStream_t s)
{
    // split outputs also get their zero padding channels written, but only when this conv owns the
    // whole tensor (concat slots must not clobber their neighbours)
    int cover = (a.out.fmt == FMT_SPLIT && a.out.C == a.cout) ? pad16(a.cout) : a.cout;
    dim3 grid(cdiv(a.Ho, ST_TILE) * cdiv(a.Wo, ST_TILE), cdiv(cover, ST_CO), B);
    double flops = 2.0 * B * a.Ho * a.Wo * (double)a.cout * a.cin * KH * KW;
    ProfScope ps(h, PROF_CONV_SIMT, s, flops, 0);
    conv_simt_kernel<KH, KW, ST_CI><<<grid, 256, 0, s>>>(a);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// The output convs of the TC engine's nets (conv_q2, conv_B1..3: 3x3 "same", 8 -> 1 or 2 channels on a 8x8 / 16x16 map, bias,
// optional running sum into channel 0, fp32 outputs): the general kernel above would spend 15/16 of its lanes on channels
// that do not exist.  One CTA per image, one thread per output pixel, the 8-channel input chunk staged in shared memory.
// Same accumulation order as conv_simt_kernel (ci, ky, kx; bias and add0 afterwards): bit-identical results.
__global__ void __launch_bounds__(256)
outconv3x3_kernel(SimtConvArgs a)
{
    __shared__ float s_in[8][18][19];
    __shared__ float s_w[8 * 9 * 2];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int H = a.in.H, W = a.in.W;
    for (int k = tid; k < (H + 2) * (W + 2); k += 256) {
        const int r = k / (W + 2), c = k - r * (W + 2), iy = r - 1, ix = c - 1;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) load_chunk_split(a.in, n, 0, iy, ix, v);
#pragma unroll
        for (int ci = 0; ci < 8; ci++) s_in[ci][r][c] = v[ci];
    }
    for (int k = tid; k < 8 * 9 * 2; k += 256) {
        const int co = k & 1, t = (k >> 1) % 9, ci = k / 18;
        s_w[k] = (ci < a.cin && co < a.cout) ? a.w[((size_t)ci * 9 + t) * a.coutw + co] : 0.f;
    }
    __syncthreads();
    if (tid >= H * W) return;
    const int oy = tid / W, ox = tid - oy * W;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1
    for (int ci = 0; ci < a.cin; ci++)
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const float v = s_in[ci][oy + ky][ox + kx];
                acc0 = fmaf(v, s_w[(ci * 9 + ky * 3 + kx) * 2], acc0);
                acc1 = fmaf(v, s_w[(ci * 9 + ky * 3 + kx) * 2 + 1], acc1);
            }
    if (a.bias) { acc0 += a.bias[0]; if (a.cout > 1) acc1 += a.bias[1]; }
    if (a.add0) acc0 += a.add0[(size_t)n * a.add0_bstride + (size_t)oy * a.Wo + ox];
    store_elem_f32(a.out, n, 0, oy, ox, acc0);
    if (a.cout > 1) store_elem_f32(a.out, n, 1, oy, ox, acc1);
}

int conv_simt(Handle *h, const SimtConvArgs &a, int kh, int kw, int B, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    static const int env_oc = [] { const char *e = getenv("PMP_OUTCONV"); return e ? atoi(e) : 1; }();      // A/B knob
    if (env_oc && kh == 3 && kw == 3 && a.cout <= 2 && a.cin <= 8 && a.in.fmt == FMT_SPLIT && a.in.H <= 16 && a.in.W <= 16 &&
        a.Ho == a.in.H && a.Wo == a.in.W && a.pad_t == 1 && a.pad_l == 1 && !a.res.p && !a.mul.p && !a.relu && a.pool == 1 &&
        a.out_c_off == 0 && (a.out.fmt == FMT_F32 || a.out.fmt == FMT_PAIR)) {
        ProfScope ps(h, PROF_CONV_SIMT, s, 2.0 * B * a.Ho * a.Wo * a.cout * a.cin * 9, 0);
        outconv3x3_kernel<<<B, 256, 0, s>>>(a);
        h->launches++;
        PMP_CUDA(cudaGetLastError());
        return PMP_OK;
    }
#define PMP_K(H_, W_, C_) if (kh == H_ && kw == W_) return launch_simt<H_, W_, C_>(h, a, B, s)
    PMP_K(1, 1, 8); PMP_K(3, 3, 8); PMP_K(5, 5, 8); PMP_K(9, 9, 2); PMP_K(5, 9, 2); PMP_K(9, 5, 2); PMP_K(3, 5, 4);
    PMP_K(5, 3, 4);
#undef PMP_K
    set_error("conv_simt: unsupported kernel size %dx%d", kh, kw);
    return PMP_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// element-wise glue
// ------------------------------------------------------------------------------------------------

// x2 = cat[x, pad_lu(nearest_up(qt))]  (Model_QBD.py:130-131 / :228-229) -> fp32 [B, cx+1, S, S]
__global__ void stem_input_kernel(Act x, const float *__restrict__ qt, int up, int ov, Act out, int B)
{
    const int S = out.H, C = out.C;
    size_t total = (size_t)B * C * S * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int xx = (int)(i % S), yy = (int)((i / S) % S), c = (int)((i / ((size_t)S * S)) % C), n = (int)(i / ((size_t)S * S * C));
        float v;
        if (c < C - 1) v = load_elem(x, n, c, yy, xx);
        else v = (yy >= ov && xx >= ov) ? qt[(size_t)n * 64 + ((yy - ov) / up) * 8 + (xx - ov) / up] : 0.f;
        reinterpret_cast<float *>(out.p)[i] = v;
    }
}

int stem_input(Handle *h, const Act &x, const float *qt, int up, int ov, const Act &out, int B, cudaStream_t s)
{
    size_t total = (size_t)B * out.C * out.H * out.W;
    int grid = (int)((total + 255) / 256);
    ProfScope ps(h, PROF_ELEMWISE, s, 0, (double)total * 5);
    stem_input_kernel<<<grid, 256, 0, s>>>(x, qt, up, ov, out, B);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// x6 = cat[x5, up2(pool2 x5), up4(pool4 x5), up8(pool8 x5)]  (Model_QBD.py:84-87): [B,32,16,16] -> [B,128,16,16]
// One CTA per (8-channel chunk, image), one thread per pixel of the 16x16 map; the 2x2 / 4x4 / 8x8 block maxima are
// built hierarchically in shared memory (each level reads 4 entries of the previous one).
__global__ void __launch_bounds__(256) pyramid_kernel(Act in, Act out)
{
    __shared__ float lvl[3][8][256];
    const int ch = blockIdx.x, n = blockIdx.y, t = threadIdx.x, y = t >> 4, x = t & 15;
    float v[8];
    if (in.fmt == FMT_SPLIT) load_chunk_split(in, n, ch, y, x, v);
    else {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = load_elem(in, n, ch * 8 + e, y, x);
    }
    auto emit = [&](int g, const float m[8]) {
        const int oc = g * in.C + ch * 8;
        if (out.fmt == FMT_SPLIT) store_chunk_split(out, n, oc >> 3, y, x, m);
        else {
#pragma unroll
            for (int e = 0; e < 8; e++) store_elem_f32(out, n, oc + e, y, x, m[e]);
        }
    };
    emit(0, v);
#pragma unroll
    for (int e = 0; e < 8; e++) lvl[0][e][t] = v[e];
    __syncthreads();
#pragma unroll
    for (int g = 1; g <= 3; g++) {
        const int half = 1 << (g - 1), y0 = (y >> g) << g, x0 = (x >> g) << g;     // children at stride `half`
        const float (*src)[256] = lvl[g - 1];
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float a = src[e][y0 * 16 + x0], b = src[e][y0 * 16 + x0 + half];
            const float c = src[e][(y0 + half) * 16 + x0], d = src[e][(y0 + half) * 16 + x0 + half];
            m[e] = fmaxf(fmaxf(a, b), fmaxf(c, d));
        }
        emit(g, m);
        if (g < 3) {
            // level g maxima live at the block origins of lvl[g]; every thread writes its own block's value (same
            // value from all threads of a block: benign)
#pragma unroll
            for (int e = 0; e < 8; e++) lvl[g][e][y0 * 16 + x0] = m[e];
            __syncthreads();
        }
    }
}

int pyramid(Handle *h, const Act &in, const Act &out, int B, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    if (in.H != 16 || in.W != 16 || (in.C & 7)) {
        set_error("pyramid: expected a [B,8k,16,16] input");
        return PMP_ERR_UNSUPPORTED;
    }
    dim3 grid(in.C >> 3, B);
    ProfScope ps(h, PROF_ELEMWISE, s, 0, (double)B * in.C * 256 * 4 * 5);
    pyramid_kernel<<<grid, 256, 0, s>>>(in, out);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// attention-trunk input (Model_QBD.py:140,:147): [B,3,S,S] = cat[up(qt, S/8), up(prev ch0, S/16), up(prev ch1, S/16)]
// prev = out0 / accumulated out1 (FMT_PAIR, 16x16).  Output fp32 [B,3,S,S] or split [B,16(pad),S,S].
__global__ void att_input_kernel(const float *__restrict__ qt, Act prev, Act out, int B)
{
    const int S = out.H;
    const int qs = S / 8, ps = S / 16;
    size_t total = (size_t)B * S * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % S), y = (int)((i / S) % S), n = (int)(i / ((size_t)S * S));
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        v[0] = qt[(size_t)n * 64 + (y / qs) * 8 + (x / qs)];
        v[1] = load_elem(prev, n, 0, y / ps, x / ps);
        v[2] = load_elem(prev, n, 1, y / ps, x / ps);
        if (out.fmt == FMT_SPLIT) {
            store_chunk_split(out, n, 0, y, x, v);
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            store_chunk_split(out, n, 1, y, x, z);
        } else {
            for (int c = 0; c < 3; c++) store_elem_f32(out, n, c, y, x, v[c]);
        }
    }
}

int att_input(Handle *h, const float *qt, const Act &prev, const Act &out, int B, cudaStream_t s)
{
    size_t total = (size_t)B * out.H * out.W;
    int grid = (int)((total + 255) / 256);
    ProfScope ps(h, PROF_ELEMWISE, s, 0, (double)total * 64);
    att_input_kernel<<<grid, 256, 0, s>>>(qt, prev, out, B);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

}  // namespace pmp
