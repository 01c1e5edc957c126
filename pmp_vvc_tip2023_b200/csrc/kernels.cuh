// Kernel-launch entry points shared between translation units (internal).
#pragma once
#include "handle.cuh"
#include "tensor.cuh"

namespace pmp {

struct SimtConvArgs {
    Act in, out, res, mul;        // res / mul optional (p == nullptr)
    const float *w = nullptr;     // [cin][kh*kw][coutw] fp32, zero padded
    const float *bias = nullptr;  // [cout] or nullptr
    int cin = 0, cout = 0, coutw = 0;
    int pad_t = 0, pad_l = 0;     // input coordinate = output coordinate + tap - pad
    int Ho = 0, Wo = 0;           // conv output size (before pooling)
    int relu = 0, pool = 1;
    int out_c_off = 0;            // channel offset inside `out` (concat writes)
    const float *add0 = nullptr;  // added to output channel 0 (Model_QBD.py:146,:153)
    long long add0_bstride = 0;
};

int conv_simt(Handle *h, const SimtConvArgs &a, int kh, int kw, int B, cudaStream_t s);
int stem_input(Handle *h, const Act &x, const float *qt, int up, int ov, const Act &out, int B, cudaStream_t s);
int pyramid(Handle *h, const Act &in, const Act &out, int B, cudaStream_t s);
int att_input(Handle *h, const float *qt, const Act &prev, const Act &out, int B, cudaStream_t s);
// out = maxpool2(in) [* mul]; split -> split
int pool2_split(Handle *h, const Act &in, const Act &out, const Act &mul, int B, cudaStream_t s);

// ---- TC engine (conv_tc.cu) ------------------------------------------------------------------
struct TcConvArgs {
    Act in;                 // FMT_SPLIT [B, cin_pad, H, W]
    Act out;                // FMT_SPLIT
    Act res;                // optional identity residual (FMT_SPLIT, same shape as the conv output)
    Act mul;                // optional attention product (FMT_SPLIT, shape of the stored output)
    const uint16_t *w = nullptr;    // packed main weights (see pack_tc_weights)
    const uint16_t *w_pair = nullptr;   // CTA-pair operand image or nullptr
    const float *bias = nullptr;    // [cout_pad] fp32 or nullptr (stems only)
    // fused 1x1 shortcut (ResidualBlock with Cin != Cout, Model_QBD.py:34-38,43): out = epilogue(conv(in) + W_sc * sc_in);
    // sc_in has the conv's H x W, w_pair_sc is packed by pack_tc_pair_fused_sc.  CTA-pair kernel only, excludes `res`.
    Act sc_in;
    const uint16_t *w_pair_sc = nullptr;
    int sc_cin_pad = 0;
    int cin_pad = 0, cout_pad = 0, kh = 1, kw = 1;
    int pad_t = 0, pad_l = 0;       // input coordinate = output coordinate + tap - pad
    int Ho = 0;                     // output rows (0: same as the input)
    int relu = 0, pool = 1;
    // first-layer ("stem") mode: the K chunks of the kx-unrolled input are assembled by TMA from the pre-shifted copies
    // written by stem_shift (layout there); `in` only describes the shape (Cp = 16 * groups, H = source rows, W = out W).
    const void *stem_src = nullptr;
    int stem_planes = 0, stem_rows = 0, stem_uw = 0, stem_nchunk = 0;
    signed char stem_chunk_plane[8] = {0}, stem_chunk_u0[8] = {0};
    int hpool = 0;                  // epilogue takes the horizontal maximum of pixel pairs: `out` is H x W/2 (CTA-pair kernel;
                                    // pool2_split then finishes the 2x2 max-pool with the vertical half)
    double flops_override = 0;      // algorithmic FLOPs per image for the profile (0: from the shapes)
};
int conv_tc(Handle *h, const TcConvArgs &a, int B, cudaStream_t s);
bool tc_supported(int cin_pad, int cout_pad, int kh, int kw, int H, int W);
// host-side packing of one conv's weights into the TC operand image; returns number of uint16 written
size_t tc_packed_elems(int cin_pad, int cout_pad, int kh, int kw);
void pack_tc_weights(const float *w, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, bool bf16, uint16_t *dst);
size_t tc_pair_packed_elems(int cin_pad, int cout_pad, int kh, int kw);
void pack_tc_pair_weights(const float *w, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, bool bf16, uint16_t *dst);
void pack_tc_pair_weights_scheme(const float *w, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, bool bf16, bool st,
                                 uint16_t *dst);
size_t tc_pair_fused_sc_elems(int sc_cin_pad, int cout_pad, int kh, int kw);
void pack_tc_pair_fused_sc(const float *w_sc, int cout, int sc_cin, int sc_cin_pad, int cout_pad, int kh, int kw, bool bf16, uint16_t *dst);
bool tc_fusion_available();      // the CTA-pair kernel is the active TC path (fused shortcuts need it)
// U[c*kw + j][y][x] = X[c][y][x + j] (x2 = cat[x, pad_lu(up(qt))] when qt != nullptr): the kx taps of a first-layer conv
// unrolled into channels so that the tensor-core kernel can run it as a kh x 1 conv.  out: FMT_SPLIT [B, (cx+1?)*kw, S0, S1]
int stem_unroll(Handle *h, const Act &x, const float *qt, int up, int ov, int kw, const Act &out, int B, cudaStream_t s);
// 8 pre-shifted 16-bit copies of every source plane of a first-layer conv: dst [B * planes][S0][uw][8 shifts][8 pixels],
// unit (y, u, s) = pixels 8u+s .. 8u+s+7 of row y (zero beyond the block).  planes = x.C pixel planes, then (qt != nullptr)
// hi and lo halves of cat's extra channel pad_lu(up(qt)) (Model_QBD.py:130-131).
int stem_shift(Handle *h, const Act &x, const float *qt, int up, int ov, int uw, bool bf16, void *dst, int B, cudaStream_t s);

// FMT_SPLIT [B, C, H, W] -> fp32 NCHW (test hooks)
int split_to_f32(Handle *h, const Act &src, float *dst, int B, cudaStream_t s);
// first-layer conv(s) of a net on the TC engine alone (test hook, pmp_debug_stem): out [B,32,S1,S1] fp32
int debug_stem(Handle *h, int wset, const void *blocks, int in_dtype, const float *qt, int B, float *out, cudaStream_t s);

}  // namespace pmp
