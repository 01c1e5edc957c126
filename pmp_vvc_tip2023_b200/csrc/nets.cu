// The four Down-Up-CNN forwards as launch sequences over the conv engines, plus weight packing.
//
// Reference semantics (paths relative to /root/reference):
//   ResidualBlock.forward   Model_QBD.py:40-44     relu(conv2(relu(conv1 x)) + shortcut(x)), no bias
//   Luma_Q_Net.forward      Model_QBD.py:78-98     Chroma_Q_Net.forward      :176-196
//   Luma_MSBD_Net.forward   Model_QBD.py:127-155   Chroma_MSBD_Net.forward   :225-253
// Parameter order/names/shapes are the reference's state_dict contract; the same table
// lives in pmp_vvc_tip2023_b200/netspec.py for the Python side.
//
// Engine selection per layer: PMP_ENGINE_TC runs every square 1x1/3x3/5x5 conv whose input is a split-precision
// activation on the tcgen05 kernels (conv_tc.cu) and keeps activations in FMT_SPLIT; the first-layer valid convs run
// there too after their kx taps are unrolled into channels (stem_tc below); the Cout <= 2 output convs (bias, fp32
// outputs) run on the exact fp32 SIMT kernel, which reads/writes the same activation formats.  PMP_ENGINE_SIMT runs
// everything in fp32 NCHW.
#include "handle.cuh"
#include "kernels.cuh"

#include <cstring>
#include <string>

namespace pmp {

// ---------------------------------------------------------------------------------------------------
// parameter tables
// ---------------------------------------------------------------------------------------------------
struct ParamSpec { std::string name; int d[4]; int nd; };

static void add_conv(std::vector<ParamSpec> &v, const std::string &name, int co, int ci, int kh, int kw, bool bias)
{
    v.push_back({name + ".weight", {co, ci, kh, kw}, 4});
    if (bias) v.push_back({name + ".bias", {co, 0, 0, 0}, 1});
}
static void add_rb(std::vector<ParamSpec> &v, const std::string &p, int ci, int co, int k)
{
    add_conv(v, p + ".left.0", co, ci, k, k, false);
    add_conv(v, p + ".left.2", co, co, k, k, false);
    if (ci != co) add_conv(v, p + ".shortcut.0", co, ci, 1, 1, false);
}

static std::vector<ParamSpec> param_spec(int net)
{
    std::vector<ParamSpec> v;
    const bool luma = (net == PMP_NET_LUMA_Q || net == PMP_NET_LUMA_MSBD);
    if (net == PMP_NET_LUMA_Q || net == PMP_NET_CHROMA_Q) {
        const int cin = luma ? 1 : 3, k1 = luma ? 9 : 5, k12 = luma ? 5 : 3;
        add_conv(v, "conv_q1", 32, cin, k1, k1, true);
        add_rb(v, "resblock_q1", 32, 64, k12);
        add_rb(v, "resblock_q2", 64, 64, k12);
        add_rb(v, "resblock_q3", 64, 32, 3);
        add_rb(v, "resblock_q4", 128, 32, 3);
        add_rb(v, "resblock_q5", 32, 32, 3);
        add_rb(v, "resblock_q6", 32, 8, 3);
        add_conv(v, "conv_q2", 1, 8, 3, 3, true);
    } else {
        const int cin = luma ? 2 : 4, kb = luma ? 9 : 5, ks = luma ? 5 : 3;
        add_conv(v, "conv_b1_1", 16, cin, kb, kb, true);
        add_conv(v, "conv_b1_2", 8, cin, ks, kb, true);
        add_conv(v, "conv_b1_3", 8, cin, kb, ks, true);
        add_rb(v, "trunk_M1.0", 32, 64, 5);
        for (int i = 1; i < 6; i++) add_rb(v, "trunk_M1." + std::to_string(i), 64, 64, 3);
        for (int i = 0; i < 4; i++) add_rb(v, "trunk_M2." + std::to_string(i), 64, 64, 3);
        for (const char *b : {"B1", "B2", "B3"}) {
            add_rb(v, std::string("trunk_") + b + ".0", 64, 32, 3);
            add_rb(v, std::string("trunk_") + b + ".1", 32, 16, 3);
            add_rb(v, std::string("trunk_") + b + ".2", 16, 8, 3);
        }
        for (const char *b : {"B1", "B2", "B3"}) add_conv(v, std::string("conv_") + b, 2, 8, 3, 3, true);
        for (const char *a : {"Att1", "Att2"}) {
            add_rb(v, std::string("trunk_") + a + ".0", 3, 32, 3);
            add_rb(v, std::string("trunk_") + a + ".1", 32, 64, 3);
        }
    }
    return v;
}

template <typename T>
static int dev_upload(WeightSet &ws, const std::vector<T> &host, T **dev)
{
    void *p = nullptr;
    PMP_CUDA(cudaMalloc(&p, host.size() * sizeof(T)));
    ws.allocs.push_back(p);
    PMP_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = reinterpret_cast<T *>(p);
    return PMP_OK;
}

// Tensor-core form of the first-layer convs: the kx taps are unrolled into input channels (cu = c*kw + j, see
// stem_unroll in kernels.cuh), which turns the kh x kw valid conv on 1..4 channels into a kh x 1 conv on cin*kw channels.
// The three MSBD stems (Model_QBD.py:108-110: kb x kb -> 16, ks x kb -> 8, kb x ks -> 8 channels, concatenated :135)
// become ONE 32-output conv by embedding the smaller kernels top/left-aligned in kb x kb (their zero pads are the
// corresponding sides, :104-106).  Stored as ConvW "stem_tc": cin = cin*kb, kh = kb, kw = 1, bias padded to cout_pad.
static int build_stem_tc(WeightSet &ws, int net, const std::vector<ParamSpec> &spec, const float *const *tensors)
{
    auto find = [&](const char *name) -> int {
        for (size_t i = 0; i < spec.size(); i++)
            if (spec[i].name == name) return (int)i;
        return -1;
    };
    const bool q = (net == PMP_NET_LUMA_Q || net == PMP_NET_CHROMA_Q);
    struct Part { int iw, ib, co_off; };
    std::vector<Part> parts;
    if (q) parts.push_back({find("conv_q1.weight"), find("conv_q1.bias"), 0});
    else {
        parts.push_back({find("conv_b1_1.weight"), find("conv_b1_1.bias"), 0});
        parts.push_back({find("conv_b1_2.weight"), find("conv_b1_2.bias"), 16});
        parts.push_back({find("conv_b1_3.weight"), find("conv_b1_3.bias"), 24});
    }
    const int cin = spec[parts[0].iw].d[1], kb = spec[parts[0].iw].d[2];
    const int cu = cin * kb, cout = 32;
    std::vector<float> wm((size_t)cout * cu * kb, 0.f), bias(pad16(cout), 0.f);      // [co][cu][ky]  (kw = 1)
    for (const Part &pt : parts) {
        if (pt.iw < 0 || pt.ib < 0) { set_error("stem parameters missing"); return PMP_ERR_STATE; }
        const int co = spec[pt.iw].d[0], kh = spec[pt.iw].d[2], kw = spec[pt.iw].d[3];
        const float *w = tensors[pt.iw];
        for (int o = 0; o < co; o++) {
            bias[pt.co_off + o] = tensors[pt.ib][o];
            for (int c = 0; c < cin; c++)
                for (int ky = 0; ky < kh; ky++)
                    for (int j = 0; j < kw; j++)
                        wm[((size_t)(pt.co_off + o) * cu + (c * kb + j)) * kb + ky] = w[(((size_t)o * cin + c) * kh + ky) * kw + j];
        }
    }
    ConvW &cw = ws.convs["stem_tc"];
    cw.cout = cout; cw.cin = cu; cw.kh = kb; cw.kw = 1; cw.cin_pad = pad16(cu); cw.cout_pad = pad16(cout);
    int rc = dev_upload(ws, bias, &cw.bias);
    if (rc) return rc;
    std::vector<uint16_t> pk(tc_packed_elems(cw.cin_pad, cw.cout_pad, kb, 1));
    pack_tc_weights(wm.data(), cout, cu, kb, 1, cw.cin_pad, cw.cout_pad, false, pk.data());
    rc = dev_upload(ws, pk, &cw.w_tc_f16);
    if (rc) return rc;
    pack_tc_weights(wm.data(), cout, cu, kb, 1, cw.cin_pad, cw.cout_pad, true, pk.data());
    rc = dev_upload(ws, pk, &cw.w_tc_bf16);
    if (rc) return rc;
    std::vector<uint16_t> pp(tc_pair_packed_elems(cw.cin_pad, cw.cout_pad, kb, 1));
    pack_tc_pair_weights(wm.data(), cout, cu, kb, 1, cw.cin_pad, cw.cout_pad, false, pp.data());
    rc = dev_upload(ws, pp, &cw.w_pair_f16);
    if (rc) return rc;
    pack_tc_pair_weights(wm.data(), cout, cu, kb, 1, cw.cin_pad, cw.cout_pad, true, pp.data());
    rc = dev_upload(ws, pp, &cw.w_pair_bf16);
    if (rc) return rc;

    // "stem2": the same conv with its K dimension laid out in chunks of 8 unrolled channels = (source plane, 8 consecutive
    // kx shifts), the form the pair kernel's stem mode assembles by TMA from pre-shifted copies (no unrolled tensor in
    // HBM).  Source planes: the cin - (q ? 0 : 1) pixel planes, then for the MSBD nets the hi and lo halves of the qt
    // channel (both multiply the qt channel's weights: W * (hi + lo)).
    const int cx = q ? cin : cin - 1;
    StemLayout &sl = ws.stem;
    sl.planes = cx + (q ? 0 : 2); sl.nchunk = 0;
    for (int pl = 0; pl < sl.planes; pl++)
        for (int j0 = 0; j0 < kb; j0 += 8) {
            if (sl.nchunk >= 8) { set_error("stem layout needs more than 8 K chunks"); return PMP_ERR_UNSUPPORTED; }
            sl.chunk_plane[sl.nchunk] = (signed char)pl; sl.chunk_u0[sl.nchunk] = (signed char)(j0 / 8); sl.nchunk++;
        }
    const int groups2 = (sl.nchunk + 1) / 2, cu2 = 16 * groups2;
    std::vector<float> wm2((size_t)cout * cu2 * kb, 0.f);
    for (int c = 0; c < sl.nchunk; c++) {
        const int src_c = sl.chunk_plane[c] < cx ? sl.chunk_plane[c] : cx;      // both qt planes read the qt channel's weights
        for (int e = 0; e < 8; e++) {
            const int j = 8 * sl.chunk_u0[c] + e;
            if (j >= kb) continue;
            for (int o = 0; o < cout; o++)
                for (int ky = 0; ky < kb; ky++)
                    wm2[((size_t)o * cu2 + (8 * c + e)) * kb + ky] = wm[((size_t)o * cu + (src_c * kb + j)) * kb + ky];
        }
    }
    ConvW &c2 = ws.convs["stem2"];
    c2.cout = cout; c2.cin = cu2; c2.kh = kb; c2.kw = 1; c2.cin_pad = cu2; c2.cout_pad = pad16(cout);
    c2.bias = cw.bias;
    std::vector<uint16_t> p2(tc_pair_packed_elems(cu2, c2.cout_pad, kb, 1));
    pack_tc_pair_weights(wm2.data(), cout, cu2, kb, 1, cu2, c2.cout_pad, false, p2.data());
    rc = dev_upload(ws, p2, &c2.w_pair_f16);
    if (rc) return rc;
    pack_tc_pair_weights(wm2.data(), cout, cu2, kb, 1, cu2, c2.cout_pad, true, p2.data());
    return dev_upload(ws, p2, &c2.w_pair_bf16);
}

int weights_create(Handle *h, int net, const float *const *tensors, const int64_t *numel, int n, int *wset)
{
    std::vector<ParamSpec> spec = param_spec(net);
    if ((int)spec.size() != n) {
        set_error("net %d expects %d parameter tensors, got %d", net, (int)spec.size(), n);
        return PMP_ERR_ARG;
    }
    for (int i = 0; i < n; i++) {
        int64_t want = 1;
        for (int k = 0; k < spec[i].nd; k++) want *= spec[i].d[k];
        if (numel[i] != want || !tensors[i]) {
            set_error("parameter %d (%s): expected %lld elements, got %lld", i, spec[i].name.c_str(), (long long)want,
                      (long long)numel[i]);
            return PMP_ERR_ARG;
        }
    }
    const int id = h->next_wset++;
    WeightSet &ws = h->wsets[id];
    ws.net = net;
    int rc = PMP_OK;
    for (int i = 0; i < n && rc == PMP_OK; i++) {
        const std::string &nm = spec[i].name;
        const size_t dot = nm.rfind('.');
        const std::string prefix = nm.substr(0, dot), kind = nm.substr(dot + 1);
        ConvW &cw = ws.convs[prefix];
        if (kind == "bias") {
            std::vector<float> b(tensors[i], tensors[i] + numel[i]);
            rc = dev_upload(ws, b, &cw.bias);
            continue;
        }
        const int co = spec[i].d[0], ci = spec[i].d[1], kh = spec[i].d[2], kw = spec[i].d[3];
        cw.cout = co; cw.cin = ci; cw.kh = kh; cw.kw = kw;
        cw.cin_pad = pad16(ci); cw.cout_pad = pad16(co);
        // SIMT operand: [cin][kh*kw][coutw] fp32, cout innermost (float4 smem reads), zero padded to 4
        const int coutw = (co + 3) & ~3;
        std::vector<float> ps((size_t)ci * kh * kw * coutw, 0.f);
        for (int o = 0; o < co; o++)
            for (int c = 0; c < ci; c++)
                for (int t = 0; t < kh * kw; t++)
                    ps[((size_t)c * kh * kw + t) * coutw + o] = tensors[i][((size_t)o * ci + c) * kh * kw + t];
        rc = dev_upload(ws, ps, &cw.w_simt);
        if (rc) break;
        // TC operand images (both 16-bit formats) for the square kernels the tcgen05 engine covers
        if (kh == kw && (kh == 1 || kh == 3 || kh == 5) && ci >= 3) {
            std::vector<uint16_t> pk(tc_packed_elems(cw.cin_pad, cw.cout_pad, kh, kw));
            if (pk.empty()) continue;
            pack_tc_weights(tensors[i], co, ci, kh, kw, cw.cin_pad, cw.cout_pad, false, pk.data());
            rc = dev_upload(ws, pk, &cw.w_tc_f16);
            if (rc) break;
            pack_tc_weights(tensors[i], co, ci, kh, kw, cw.cin_pad, cw.cout_pad, true, pk.data());
            rc = dev_upload(ws, pk, &cw.w_tc_bf16);
            if (!rc) {
                std::vector<uint16_t> pp(tc_pair_packed_elems(cw.cin_pad, cw.cout_pad, kh, kw));
                pack_tc_pair_weights(tensors[i], co, ci, kh, kw, cw.cin_pad, cw.cout_pad, false, pp.data());
                rc = dev_upload(ws, pp, &cw.w_pair_f16);
                if (rc) break;
                pack_tc_pair_weights(tensors[i], co, ci, kh, kw, cw.cin_pad, cw.cout_pad, true, pp.data());
                rc = dev_upload(ws, pp, &cw.w_pair_bf16);
            }
        }
    }
    // 1x1 shortcuts (Model_QBD.py:34-38) packed for fusion into the second conv of their block (conv_tc.cu, TcConvArgs::sc_in)
    for (int i = 0; i < n && rc == PMP_OK; i++) {
        const std::string &nm = spec[i].name;
        const std::string tail = ".shortcut.0.weight";
        if (nm.size() <= tail.size() || nm.compare(nm.size() - tail.size(), tail.size(), tail) != 0) continue;
        auto it = ws.convs.find(nm.substr(0, nm.size() - tail.size()) + ".left.2");
        if (it == ws.convs.end() || !it->second.w_pair_f16) continue;
        ConvW &l2 = it->second;
        const int co = spec[i].d[0], ci = spec[i].d[1];
        if (co != l2.cout) continue;
        l2.sc_cin = ci; l2.sc_cin_pad = pad16(ci);
        std::vector<uint16_t> pf(tc_pair_fused_sc_elems(l2.sc_cin_pad, l2.cout_pad, l2.kh, l2.kw));
        pack_tc_pair_fused_sc(tensors[i], co, ci, l2.sc_cin_pad, l2.cout_pad, l2.kh, l2.kw, false, pf.data());
        rc = dev_upload(ws, pf, &l2.w_pair_sc_f16);
        if (rc) break;
        pack_tc_pair_fused_sc(tensors[i], co, ci, l2.sc_cin_pad, l2.cout_pad, l2.kh, l2.kw, true, pf.data());
        rc = dev_upload(ws, pf, &l2.w_pair_sc_bf16);
    }
    if (rc == PMP_OK) rc = build_stem_tc(ws, net, spec, tensors);
    if (rc != PMP_OK) {
        weights_destroy(h, id);
        return rc;
    }
    *wset = id;
    return PMP_OK;
}

int weights_destroy(Handle *h, int wset)
{
    auto it = h->wsets.find(wset);
    if (it == h->wsets.end()) {
        set_error("unknown weight set %d", wset);
        return PMP_ERR_STATE;
    }
    cudaDeviceSynchronize();
    for (void *p : it->second.allocs) cudaFree(p);
    h->wsets.erase(it);
    return PMP_OK;
}

// ---------------------------------------------------------------------------------------------------
// forward-pass builder
// ---------------------------------------------------------------------------------------------------
struct ConvOpts {
    int relu = 0, pool = 1;
    Act res, mul;
    Act sc_in;                        // fuse the block's 1x1 shortcut conv on this input (instead of `res`)
    int pad_t = -1, pad_l = -1;       // -1: "same" padding k/2
    int Ho = 0, Wo = 0;               // 0: same as input
    int out_c_off = 0;
    const float *add0 = nullptr;
    long long add0_bstride = 0;
};

struct Net {
    Handle *h;
    WeightSet *ws;
    int B;
    cudaStream_t s;
    bool dry;          // sizing pass: no launches, arena offsets only
    bool tc;
    size_t off = 0, peak = 0;
    int rc = PMP_OK;

    Act alloc(int C, int H, int W, int fmt)
    {
        Act a;
        a.fmt = fmt; a.C = C; a.Cp = (fmt == FMT_SPLIT) ? pad16(C) : C; a.H = H; a.W = W;
        a.bf16 = (h->tc_dtype == PMP_TC_BF16);
        a.bytes = (act_bytes(fmt, B, C, H, W) + 1023) & ~(size_t)1023;
        a.p = dry ? nullptr : (void *)(h->arena + off);
        off += a.bytes;
        if (off > peak) peak = off;
        return a;
    }
    Act act(int C, int H, int W) { return alloc(C, H, W, tc ? FMT_SPLIT : FMT_F32); }

    const ConvW *weights(const std::string &name)
    {
        auto it = ws->convs.find(name);
        if (it == ws->convs.end()) {
            set_error("weight set has no conv '%s'", name.c_str());
            rc = PMP_ERR_STATE;
            return nullptr;
        }
        return &it->second;
    }

    void conv(const std::string &name, const Act &in, const Act &out, const ConvOpts &o)
    {
        if (rc) return;
        const ConvW *w = weights(name);
        if (!w) return;
        const int pad_t = o.pad_t < 0 ? w->kh / 2 : o.pad_t, pad_l = o.pad_l < 0 ? w->kw / 2 : o.pad_l;
        const int Ho = o.Ho ? o.Ho : in.H, Wo = o.Wo ? o.Wo : in.W;
        const bool same = (w->kh == w->kw) && pad_t == w->kh / 2 && pad_l == w->kw / 2 && Ho == in.H && Wo == in.W;
        const bool use_tc = tc && same && in.fmt == FMT_SPLIT && out.fmt == FMT_SPLIT && o.out_c_off == 0 && !o.add0 &&
                            !w->bias && out.C == w->cout && in.C == w->cin && (!o.res.p || o.res.fmt == FMT_SPLIT) &&
                            (!o.mul.p || o.mul.fmt == FMT_SPLIT) && (in.bf16 ? w->w_tc_bf16 : w->w_tc_f16) &&
                            tc_supported(w->cin_pad, w->cout_pad, w->kh, w->kw, in.H, in.W);
        if (o.sc_in.C && !use_tc) {
            set_error("conv '%s': fused shortcut requested on a layer the TC engine does not run", name.c_str());
            rc = PMP_ERR_STATE;
            return;
        }
        if (use_tc) {
            TcConvArgs a;
            a.in = in; a.res = o.res; a.mul = o.mul;
            a.w = in.bf16 ? w->w_tc_bf16 : w->w_tc_f16;
            a.w_pair = in.bf16 ? w->w_pair_bf16 : w->w_pair_f16;
            a.cin_pad = w->cin_pad; a.cout_pad = w->cout_pad; a.kh = w->kh; a.kw = w->kw; a.pad_t = pad_t; a.pad_l = pad_l;
            a.relu = o.relu; a.pool = 1;
            if (o.sc_in.p || (dry && o.sc_in.C)) {
                a.sc_in = o.sc_in; a.sc_cin_pad = w->sc_cin_pad;
                a.w_pair_sc = in.bf16 ? w->w_pair_sc_bf16 : w->w_pair_sc_f16;
            }
            if (o.pool == 2) {
                // the conv's epilogue takes the horizontal maximum (half-width temporary), a pooling pass the vertical
                // one (mul applies after pooling); without the CTA-pair kernel: un-pooled temporary + full 2x2 pass
                static const int env_hp = [] { const char *e = getenv("PMP_TC_HPOOL"); return e ? atoi(e) : 1; }();   // A/B knob
                const bool hp = env_hp && tc_fusion_available() && !(in.W & 1);
                Act tmp = act(out.C, in.H, hp ? in.W / 2 : in.W);
                a.out = tmp; a.mul = Act(); a.hpool = hp ? 1 : 0;
                if (!dry) {
                    rc = conv_tc(h, a, B, s);
                    if (!rc) rc = pool2_split(h, tmp, out, o.mul, B, s);
                }
                off -= tmp.bytes;
            } else if (!dry) {
                a.out = out;
                rc = conv_tc(h, a, B, s);
            }
            return;
        }
        if (dry) return;
        SimtConvArgs a;
        a.in = in; a.out = out; a.res = o.res; a.mul = o.mul;
        a.w = w->w_simt; a.bias = w->bias;
        a.cin = w->cin; a.cout = w->cout; a.coutw = (w->cout + 3) & ~3;
        a.pad_t = pad_t; a.pad_l = pad_l; a.Ho = Ho; a.Wo = Wo;
        a.relu = o.relu; a.pool = o.pool; a.out_c_off = o.out_c_off;
        a.add0 = o.add0; a.add0_bstride = o.add0_bstride;
        rc = conv_simt(h, a, w->kh, w->kw, B, s);
    }

    // first-layer conv(s) on the tensor cores: unroll kx into channels, then one kh x 1 conv with bias + ReLU.
    // Returns false when the TC path is not usable (engine SIMT): the caller runs the SIMT stems instead.
    bool stem_tc(const Act &x, const float *qt, int up, int ov, const Act &out, double flops_per_image)
    {
        if (!tc || rc) return false;
        const ConvW *w = weights("stem_tc");
        if (!w) return false;
        const uint16_t *wp = (h->tc_dtype == PMP_TC_BF16) ? w->w_tc_bf16 : w->w_tc_f16;
        if (!wp || !tc_supported(w->cin_pad, w->cout_pad, w->kh, 1, out.H, out.W)) return false;
        const size_t mark = off;
        static const int env_stem2 = [] { const char *e = getenv("PMP_TC_STEM2"); return e ? atoi(e) : 1; }();      // A/B knob
        const ConvW *w2 = env_stem2 && tc_fusion_available() && !(out.W & 7) ? weights("stem2") : nullptr;
        if (w2 && ws->stem.nchunk > 0) {
            // pair kernel's stem mode: 8 pre-shifted 16-bit copies of each source plane (0.08 MB per luma block instead of
            // a 0.3-0.6 MB unrolled tensor written and read back), the K chunks are assembled by TMA
            const StemLayout &sl = ws->stem;
            const int uw = out.W / 8 + (w2->kh > 8 ? 1 : 0);
            Act cp;
            cp.fmt = FMT_U8; cp.C = cp.Cp = 1; cp.H = 1; cp.W = 1;
            cp.bytes = (((size_t)B * sl.planes * x.H * uw * 128) + 1023) & ~(size_t)1023;
            cp.p = dry ? nullptr : (void *)(h->arena + off);
            off += cp.bytes;
            if (off > peak) peak = off;
            if (!dry) {
                const bool bf = h->tc_dtype == PMP_TC_BF16;
                rc = stem_shift(h, x, qt, up, ov, uw, bf, cp.p, B, s);
                TcConvArgs a;
                a.in.fmt = FMT_SPLIT; a.in.C = w2->cin; a.in.Cp = w2->cin_pad; a.in.H = x.H; a.in.W = out.W; a.in.bf16 = bf; a.in.p = cp.p;
                a.out = out; a.bias = w2->bias;
                a.w_pair = bf ? w2->w_pair_bf16 : w2->w_pair_f16;
                a.w = a.w_pair;         // unused: stem mode runs on the pair kernel only
                a.cin_pad = w2->cin_pad; a.cout_pad = w2->cout_pad; a.kh = w2->kh; a.kw = 1; a.pad_t = 0; a.pad_l = 0;
                a.Ho = out.H; a.relu = 1; a.pool = 1; a.flops_override = flops_per_image;
                a.stem_src = cp.p; a.stem_planes = sl.planes; a.stem_rows = x.H; a.stem_uw = uw; a.stem_nchunk = sl.nchunk;
                for (int c = 0; c < 8; c++) { a.stem_chunk_plane[c] = sl.chunk_plane[c]; a.stem_chunk_u0[c] = sl.chunk_u0[c]; }
                if (!rc) rc = conv_tc(h, a, B, s);
            }
            off = mark;
            return true;
        }
        Act u = alloc(w->cin, x.H, out.W, FMT_SPLIT);
        if (!dry) {
            rc = stem_unroll(h, x, qt, up, ov, w->kh, u, B, s);
            TcConvArgs a;
            a.in = u; a.out = out; a.w = wp; a.bias = w->bias;
            a.w_pair = (h->tc_dtype == PMP_TC_BF16) ? w->w_pair_bf16 : w->w_pair_f16;
            a.cin_pad = w->cin_pad; a.cout_pad = w->cout_pad; a.kh = w->kh; a.kw = 1; a.pad_t = 0; a.pad_l = 0;
            a.Ho = out.H; a.relu = 1; a.pool = 1; a.flops_override = flops_per_image;
            if (!rc) rc = conv_tc(h, a, B, s);
        }
        off = mark;
        return true;
    }

    // ResidualBlock (Model_QBD.py:23-44) into `out`; scratch activations are released on return
    void resblock(const std::string &p, const Act &in, int cout, const Act &out, int pool, const Act &mul)
    {
        if (rc) return;
        const size_t mark = off;
        Act mid = act(cout, in.H, in.W);
        ConvOpts o1;
        o1.relu = 1;
        conv(p + ".left.0", in, mid, o1);
        ConvOpts o2;
        o2.relu = 1; o2.pool = pool; o2.mul = mul;
        if (in.C != cout) {
            // the 1x1 shortcut conv: fused into the second conv as extra K groups on the TC engine (no `sc` tensor through
            // HBM, no bandwidth-bound launch), a separate conv + residual add otherwise
            const ConvW *w2 = weights(p + ".left.2");
            const bool fuse = tc && tc_fusion_available() && w2 && w2->w_pair_sc_f16 && in.fmt == FMT_SPLIT && w2->sc_cin == in.C &&
                              tc_supported(w2->cin_pad, w2->cout_pad, w2->kh, w2->kw, in.H, in.W);
            if (fuse) {
                o2.sc_in = in;
            } else {
                Act sc = act(cout, in.H, in.W);
                conv(p + ".shortcut.0", in, sc, ConvOpts());
                o2.res = sc;
            }
        } else {
            o2.res = in;
        }
        conv(p + ".left.2", mid, out, o2);
        off = mark;
    }

    // nn.Sequential of ResidualBlocks with ping-pong buffers; the last block writes `out` (pool/mul fused there)
    void trunk(const std::string &p, const Act &in, const std::vector<int> &couts, const Act &out, int pool,
               const Act &mul)
    {
        if (rc) return;
        const size_t mark = off;
        Act cur = in;
        const int n = (int)couts.size();
        Act pp[2];
        for (int i = 0; i < n; i++) {
            const bool last = (i == n - 1);
            Act dst = out;
            if (!last) {
                Act &slot = pp[i & 1];
                if (!slot.bytes || slot.C != couts[i]) slot = act(couts[i], in.H, in.W);
                dst = slot;
            }
            resblock(p + "." + std::to_string(i), cur, couts[i], dst, last ? pool : 1, last ? mul : Act());
            cur = dst;
        }
        off = mark;
    }
};

static Act input_act(const void *blocks, int in_dtype, int C, int S)
{
    Act x;
    x.p = const_cast<void *>(blocks);
    x.fmt = (in_dtype == PMP_IN_U8) ? FMT_U8 : FMT_F32;
    x.C = x.Cp = C; x.H = x.W = S;
    return x;
}

static void run_q(Net &n, bool luma, const void *blocks, int in_dtype, float *qt_out)
{
    const int S0 = luma ? 68 : 34, S1 = luma ? 64 : 32;
    Act x = input_act(blocks, in_dtype, luma ? 1 : 3, S0);
    Act x2 = n.act(32, S1, S1);
    const int k1 = luma ? 9 : 5;
    if (!n.stem_tc(x, nullptr, 1, 0, x2, 2.0 * S1 * S1 * 32.0 * x.C * k1 * k1)) {
        ConvOpts c1;
        c1.relu = 1; c1.pad_t = 0; c1.pad_l = 0; c1.Ho = S1; c1.Wo = S1;   // padding_rb + valid conv (:79-80)
        n.conv("conv_q1", x, x2, c1);
    }
    const int S3 = 32;
    Act x3 = n.act(64, S3, S3);
    n.resblock("resblock_q1", x2, 64, x3, luma ? 2 : 1, Act());           // :81 pools, :179 does not
    Act x4 = n.act(64, 16, 16);
    n.resblock("resblock_q2", x3, 64, x4, 2, Act());
    Act x5 = n.act(32, 16, 16);
    n.resblock("resblock_q3", x4, 32, x5, 1, Act());
    Act x6 = n.act(128, 16, 16);
    if (!n.dry && !n.rc) n.rc = pyramid(n.h, x5, x6, n.B, n.s);           // :84-87
    Act x7 = n.act(32, 16, 16);
    n.resblock("resblock_q4", x6, 32, x7, 1, Act());
    Act x8 = n.act(32, 8, 8);
    n.resblock("resblock_q5", x7, 32, x8, 2, Act());
    Act x9 = n.act(8, 8, 8);
    n.resblock("resblock_q6", x8, 8, x9, 1, Act());
    Act out;
    out.p = qt_out; out.fmt = FMT_F32; out.C = out.Cp = 1; out.H = out.W = 8;
    n.conv("conv_q2", x9, out, ConvOpts());
}

static Act pair_act(float *c0, float *c1, long long bstride)
{
    Act a;
    a.p = c0; a.p2 = c1; a.fmt = FMT_PAIR; a.C = a.Cp = 2; a.H = a.W = 16; a.bstride = bstride;
    return a;
}

static void run_msbd(Net &n, bool luma, const void *blocks, int in_dtype, const float *qt, float *o0c0, float *o0c1,
                     float *o1c0, float *o1c1, float *o2c0, float *o2c1, long long bstride)
{
    const int S0 = luma ? 68 : 34, S1 = luma ? 64 : 32, ov = luma ? 4 : 2, up = luma ? 8 : 4;
    const int cx = luma ? 1 : 3;
    Act x = input_act(blocks, in_dtype, cx, S0);
    Act x3 = n.act(32, S1, S1);
    const int kb = luma ? 9 : 5, ks = luma ? 5 : 3;
    const double stem_flops = 2.0 * S1 * S1 * (cx + 1) * (16.0 * kb * kb + 16.0 * kb * ks);
    if (!n.stem_tc(x, qt, up, ov, x3, stem_flops)) {
        Act x2 = n.alloc(cx + 1, S0, S0, FMT_F32);                        // cat[x, pad_lu(up(qt))] (:130-131)
        if (!n.dry && !n.rc) n.rc = stem_input(n.h, x, qt, up, ov, x2, n.B, n.s);
        ConvOpts st;
        st.relu = 1; st.pad_t = 0; st.pad_l = 0; st.Ho = S1; st.Wo = S1;  // asymmetric zero pads == OOB reads (:132-134)
        st.out_c_off = 0;  n.conv("conv_b1_1", x2, x3, st);
        st.out_c_off = 16; n.conv("conv_b1_2", x2, x3, st);
        st.out_c_off = 24; n.conv("conv_b1_3", x2, x3, st);
    }
    const int S4 = 32;
    Act x4 = n.act(64, S4, S4);
    n.trunk("trunk_M1", x3, {64, 64, 64, 64, 64, 64}, x4, luma ? 2 : 1, Act());      // :136 / :234
    Act x5 = n.act(64, 16, 16);
    n.trunk("trunk_M2", x4, {64, 64, 64, 64}, x5, 2, Act());                         // :137
    Act x6 = n.act(8, 16, 16);
    n.trunk("trunk_B1", x5, {32, 16, 8}, x6, 1, Act());
    Act out0 = pair_act(o0c0, o0c1, bstride);
    n.conv("conv_B1", x6, out0, ConvOpts());                                         // :139
    Act a0 = n.act(3, 16, 16);
    if (!n.dry && !n.rc) n.rc = att_input(n.h, qt, out0, a0, n.B, n.s);              // :140
    Act xb1 = n.act(64, 16, 16);
    n.trunk("trunk_Att1", a0, {32, 64}, xb1, 1, x5);                                 // :141-143 (x5 * att fused)
    Act xb2 = n.act(8, 16, 16);
    n.trunk("trunk_B2", xb1, {32, 16, 8}, xb2, 1, Act());
    Act out1 = pair_act(o1c0, o1c1, bstride);
    ConvOpts c2;
    c2.add0 = o0c0; c2.add0_bstride = bstride;                                       // :146
    n.conv("conv_B2", xb2, out1, c2);
    Act a1 = n.act(3, 32, 32);
    if (!n.dry && !n.rc) n.rc = att_input(n.h, qt, out1, a1, n.B, n.s);              // :147 (accumulated out1)
    Act xb3 = n.act(64, 32, 32);
    n.trunk("trunk_Att2", a1, {32, 64}, xb3, 1, x4);                                 // :148-150
    Act xb4 = n.act(8, 16, 16);
    n.trunk("trunk_B3", xb3, {32, 16, 8}, xb4, 2, Act());                            // :151
    Act out2 = pair_act(o2c0, o2c1, bstride);
    ConvOpts c3;
    c3.add0 = o1c0; c3.add0_bstride = bstride;                                       // :153
    n.conv("conv_B3", xb4, out2, c3);
}

static int find_wset(Handle *h, int wset, int kind_a, int kind_b, WeightSet **out)
{
    auto it = h->wsets.find(wset);
    if (it == h->wsets.end()) {
        set_error("unknown weight set %d", wset);
        return PMP_ERR_STATE;
    }
    if (it->second.net != kind_a && it->second.net != kind_b) {
        set_error("weight set %d holds net kind %d, not usable here", wset, it->second.net);
        return PMP_ERR_STATE;
    }
    *out = &it->second;
    return PMP_OK;
}

int forward_q(Handle *h, int wset, const void *blocks, int in_dtype, int B, float *qt_out, cudaStream_t s)
{
    WeightSet *ws = nullptr;
    int rc = find_wset(h, wset, PMP_NET_LUMA_Q, PMP_NET_CHROMA_Q, &ws);
    if (rc) return rc;
    const bool luma = ws->net == PMP_NET_LUMA_Q;
    Net dry{h, ws, B, s, true, h->engine == PMP_ENGINE_TC};
    run_q(dry, luma, blocks, in_dtype, qt_out);
    if (dry.rc) return dry.rc;
    rc = ensure_arena(h, dry.peak);
    if (rc) return rc;
    Net n{h, ws, B, s, false, h->engine == PMP_ENGINE_TC};
    run_q(n, luma, blocks, in_dtype, qt_out);
    return n.rc;
}

int forward_msbd(Handle *h, int wset, const void *blocks, int in_dtype, const float *qt, int B, float *o0c0, float *o0c1,
                 float *o1c0, float *o1c1, float *o2c0, float *o2c1, int out_bstride, cudaStream_t s)
{
    WeightSet *ws = nullptr;
    int rc = find_wset(h, wset, PMP_NET_LUMA_MSBD, PMP_NET_CHROMA_MSBD, &ws);
    if (rc) return rc;
    const bool luma = ws->net == PMP_NET_LUMA_MSBD;
    Net dry{h, ws, B, s, true, h->engine == PMP_ENGINE_TC};
    run_msbd(dry, luma, blocks, in_dtype, qt, o0c0, o0c1, o1c0, o1c1, o2c0, o2c1, out_bstride);
    if (dry.rc) return dry.rc;
    rc = ensure_arena(h, dry.peak);
    if (rc) return rc;
    Net n{h, ws, B, s, false, h->engine == PMP_ENGINE_TC};
    run_msbd(n, luma, blocks, in_dtype, qt, o0c0, o0c1, o1c0, o1c1, o2c0, o2c1, out_bstride);
    return n.rc;
}

// Test hook: only the first-layer conv(s) of the net (conv_q1, or conv_b1_1..3 concatenated; Model_QBD.py:79-80,:130-135)
// through the TC engine's stem path, converted back to fp32 NCHW [B,32,S1,S1].
int debug_stem(Handle *h, int wset, const void *blocks, int in_dtype, const float *qt, int B, float *out, cudaStream_t s)
{
    auto it = h->wsets.find(wset);
    if (it == h->wsets.end()) { set_error("unknown weight set %d", wset); return PMP_ERR_STATE; }
    WeightSet *ws = &it->second;
    const bool luma = ws->net == PMP_NET_LUMA_Q || ws->net == PMP_NET_LUMA_MSBD;
    const bool msbd = ws->net == PMP_NET_LUMA_MSBD || ws->net == PMP_NET_CHROMA_MSBD;
    if (msbd && !qt) { set_error("debug_stem: the MSBD stems need the qt map"); return PMP_ERR_ARG; }
    const int S0 = luma ? 68 : 34, S1 = luma ? 64 : 32, ov = luma ? 4 : 2, up = luma ? 8 : 4;
    int rc = PMP_OK;
    for (int pass = 0; pass < 2 && !rc; pass++) {
        Net n{h, ws, B, s, pass == 0, true};
        Act x = input_act(blocks, in_dtype, luma ? (msbd ? 1 : 1) : 3, S0);
        Act x2 = n.act(32, S1, S1);
        if (!n.stem_tc(x, msbd ? qt : nullptr, msbd ? up : 1, msbd ? ov : 0, x2, 0.0)) {
            set_error("debug_stem: the TC stem path is not available for this net");
            return n.rc ? n.rc : PMP_ERR_UNSUPPORTED;
        }
        rc = n.rc;
        if (!rc && pass == 0) rc = ensure_arena(h, n.peak);
        if (!rc && pass == 1) rc = split_to_f32(h, x2, out, B, s);
    }
    return rc;
}

}  // namespace pmp
