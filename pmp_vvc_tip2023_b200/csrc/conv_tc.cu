// TC engine: stride-1 "same" convolution (1x1 / 3x3 / 5x5, and the kh x 1 form of the first-layer convs) as an implicit
// GEMM on the 5th-gen tensor cores.
//
//   D[pixels, cout] += A[pixels, cin] * W[cin, cout]   per filter tap, fp32 accumulation in TMEM.
//
// Split precision (single-pass 16-bit operands miss the 1e-2 parity bar): activations and
// weights are stored as hi + lo 16-bit pairs and three products are accumulated,
//   a_hi*w_hi + a_hi*w_lo + a_lo*w_hi        (error ~ a_lo*w_lo ~ 2^-22 relative with fp16),
// either as three tcgen05.mma with N = Cout into the same columns (unstacked) or as a_hi x [w_hi | w_lo] (N = 2*Cout) plus
// a_lo x w_hi (N = Cout) with the epilogue adding the column halves (stacked: fewer shared-memory operand reads, twice the
// TMEM columns; the default wherever the layer is not bandwidth-bound, see tc_pair_stacked_layout).
//
// Tiling ("halo-resident linear tile"): a CTA tile is up to 4 M-tiles of 128 consecutive positions of ONE image, counted on
// the zero-padded pitch P = W + k - 1.  One TMA box per 16-channel group (8 x P x rows x 4 planes) brings the halo tile
// of the FMT_SPLIT activation (tensor.cuh) into shared memory exactly in the no-swizzle K-major UMMA layout, TMA's
// out-of-bounds zero fill providing the convolution's zero padding.  Because the tile is linear on pitch P, the A
// operand of filter tap (ky,kx) is the same descriptor advanced by (ky*P + kx)*16 bytes; positions that fall into the
// P-W pad columns compute garbage that the epilogue drops (W/P = 94..97 % efficiency).  Weights stream through a
// shared-memory ring of pre-packed slabs.
//
// Kernels:
//   conv_tc_pair_kernel  the production path (every layer, every batch size): 2-CTA clusters issuing cta_group::2 MMAs
//                        with M = 256 over two images, row-granular weight stages, a lean warp-uniform issue loop
//                        (pair_issuer), accumulator and activation slot rings, fused 1x1 shortcut (second input tensor),
//                        fused horizontal max-pool, TMA-assembled first-layer ("stem") mode, programmatic dependent launch.
//   conv_tc_kernel       the first-generation single-CTA kernel (per-tap stages), kept as the PMP_TC_PAIR=0 / self-test
//                        A-B reference; same arithmetic per accumulator, so both produce the same bits.
// Warp roles, persistent grid.  conv_tc_kernel (448 threads): warp 0 = weight producer, warp 1 = activation producer + TMEM
// owner, warps 2-5 = MMA issuers (one per M-tile), warps 6-13 = epilogue.  conv_tc_pair_kernel (768 threads, roles
// aligned to warpgroups for setmaxnreg: 32 / 64 / 96 registers): warps 0-1 = producers (2-3 idle), warps 4-7 = MMA issuers,
// warps 8-23 = two epilogue groups of 8 warps draining alternate accumulator slots (TMEM -> registers -> bias / residual
// add / ReLU / attention product / pair maximum -> hi/lo split -> coalesced 16-byte stores; residual / attention operands
// are fetched ahead of the accumulator wait).
// What was measured and why each piece exists: profiles/r01_conv_tc_ncu_summary.md, profiles/r01b_conv_tc_pair_summary.md,
// profiles/r02_epilogue_groups.md.
#include "handle.cuh"
#include "kernels.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>

namespace pmp {

constexpr int TC_THREADS = 448;
constexpr int TC_MMA_WARPS = 4;
constexpr int TC_EPI_WARPS = 8;
// CTA-pair kernel: two epilogue groups of TC_EPI_WARPS warps each; consecutive accumulator uses (M-tiles) alternate between
// them, so two M-tiles drain concurrently (measured: with one group the epilogue paced every Cout <= 32 layer and every
// layer with a residual / attention operand, profiles/r02_epilogue_groups.md)
constexpr int TC_PAIR_EPI_GROUPS = 2;
// Warp roles of the pair kernel by warpgroup (setmaxnreg works on whole warpgroups): WG0 = warp 0 weight producer, warp 1
// activation producer + TMEM owner, warps 2-3 idle; WG1 = warps 4-7 MMA issuers; WG2-5 = warps 8-23 epilogue.  The
// kernel launches with 80 registers per thread (768 threads); the producers drop to 32, the issuers to 64 and the
// epilogue warps grow to 96 (128*32 + 128*64 + 512*96 = 768*80: setmaxnreg trades registers inside the pool the CTA launched with).
constexpr int TC_PAIR_ISSUER_WARP0 = 4;
constexpr int TC_PAIR_EPI_WARP0 = 8;
constexpr int TC_PAIR_THREADS = 32 * (TC_PAIR_EPI_WARP0 + TC_PAIR_EPI_GROUPS * TC_EPI_WARPS);
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int TC_MAX_STAGES = 24;
constexpr int TC_MAX_GROUPS = 8;
constexpr uint32_t TC_SMEM_HEADER = 1024;
constexpr uint32_t TC_SMEM_MAX = 232448;          // 227 KB dynamic shared memory per CTA

struct TcGeom {
    int P, MT, total_mt, tiles, Rbox, groups, N1, coutp, nstages, stacked, pairbuf, nbuf;
    uint32_t plane_bytes, group_bytes, act_bytes, stage_bytes, smem_bytes, tmem_cols;
};

// Accumulator schemes (both double-buffer the accumulators so the epilogue overlaps the next tile's MMAs):
//   unstacked : 3 MMAs (N = Cout) per tap; 4 M-tiles x 2 buffers x Cout columns.       Default for Cout = 64.
//   stacked   : 2 MMAs (N = 2*Cout, Cout) per tap; fewer instructions and operand reads, 2*Cout columns per M-tile:
//               4 M-tiles x 2 buffers when 16*Cout <= 512 (default for Cout <= 32, which are issue-bound), else
//               2 M-tiles x 2 buffers with issuer pairs owning alternate tiles ("pairbuf"; measured slower, kept as an option).
static int g_tc_scheme = -1;     // -1 auto, 0 unstacked, 1 stacked
static int g_tc_pair = -1;       // -1 library default / env PMP_TC_PAIR, 0 single-CTA kernel, 1 CTA-pair kernel where applicable
constexpr int TC_DEFAULT_PAIR = 1;

// Accumulator scheme (baked into the CTA-pair kernel's weight image): stacked (2 MMAs per tap, 11 instead of 15 KB of
// shared-memory operand reads per K step) for everything except the 1x1 Cout = 64 shortcuts (bandwidth-bound).  The
// Cout = 64 layers then need 128 accumulator columns per M-tile: PMP_TC_ST64 = 0 keeps them all unstacked, 1 stacks the
// 3x3 ones only, 2 (default) the 3x3 and 5x5 ones (A/B knob).
static int g_tc_st64 = -1;       // -1 library default / env PMP_TC_ST64
static inline bool tc_pair_stacked_layout(int cout_pad, int kh, int kw)
{
    static const int env_st64 = [] { const char *e = getenv("PMP_TC_ST64"); return e ? atoi(e) : 2; }();
    const int st64 = g_tc_st64 < 0 ? env_st64 : g_tc_st64;
    if (cout_pad <= 32) return true;
    if (cout_pad != 64 || kh != kw) return false;
    return (st64 >= 1 && kh == 3) || (st64 >= 2 && kh == 5);
}

static bool tc_pair_default()
{
    static const int env_pair = [] { const char *e = getenv("PMP_TC_PAIR"); return e ? atoi(e) : TC_DEFAULT_PAIR; }();
    return (g_tc_pair < 0 ? env_pair : g_tc_pair) != 0;
}

// Shared-memory split between activation buffers and the weight ring for ring stages of `slab` bytes.  Small images leave
// room for several whole-tile activation buffers (nbuf): the producer then runs up to nbuf - 1 tiles ahead, which hides
// the TMA latency that a single buffer per channel group exposes once per tile (the small layers are latency-bound).
static int g_tc_nbuf_max = -1;   // -1 library default / env PMP_TC_NBUF
static void tc_ring_layout(TcGeom &g, uint32_t slab, int taps)
{
    static const int env_nbuf = [] { const char *e = getenv("PMP_TC_NBUF"); return e ? atoi(e) : 4; }();
    const int nbuf_max = g_tc_nbuf_max > 0 ? g_tc_nbuf_max : (env_nbuf > 0 ? env_nbuf : 1);
    int want = g.groups * taps;                       // one tile's worth of slabs in flight is plenty
    if (want > TC_MAX_STAGES) want = TC_MAX_STAGES;
    int nb = 1;
    while (nb < nbuf_max && (nb + 1) * g.groups <= 16 &&
           TC_SMEM_HEADER + (uint32_t)(nb + 1) * g.act_bytes + (uint32_t)want * slab <= TC_SMEM_MAX)
        nb++;
    g.nbuf = nb;
    int ns = (int)((TC_SMEM_MAX - TC_SMEM_HEADER - (uint32_t)nb * g.act_bytes) / slab);
    g.nstages = ns < TC_MAX_STAGES ? ns : TC_MAX_STAGES;
    g.smem_bytes = TC_SMEM_HEADER + (uint32_t)nb * g.act_bytes + (uint32_t)g.nstages * slab;
}

static bool tc_geometry(int cin_pad, int cout_pad, int kh, int kw, int H, int W, TcGeom &g, int scheme_override = -1, bool pair = false,
                        int ppg = 4)        // planes per channel group in shared memory: hi + lo k8 halves, or hi only (stems)
{
    if (kh < 1 || kh > 9 || kw < 1 || kw > 5) return false;
    if (cin_pad % 16 || cout_pad % 16 || cin_pad < 16 || cout_pad < 16 || cout_pad > 64) return false;
    if (H < 8 || W < 8 || (W & 1)) return false;
    g.groups = cin_pad / 16;
    if (g.groups > TC_MAX_GROUPS) return false;
    g.coutp = cout_pad;
    g.N1 = 2 * cout_pad;
    g.P = W + kw - 1;
    g.total_mt = (H * g.P + 127) / 128;
    g.stage_bytes = 32u * g.N1;
    static const int env_scheme = [] {
        const char *e = getenv("PMP_TC_SCHEME");       // tuning/A-B knob: 0 unstacked, 1 stacked
        return e ? atoi(e) : -1;
    }();
    const int scheme = scheme_override >= 0 ? scheme_override : (g_tc_scheme < 0 ? env_scheme : g_tc_scheme);
    // auto: stacked for Cout <= 32 (those layers are bound by the A-operand reads: 2 instead of 3 per tap), unstacked for
    // Cout = 64.  Both kernels follow the same rule, so results do not depend on batch composition bit for bit.
    g.stacked = scheme == 1 || (scheme < 0 && tc_pair_stacked_layout(cout_pad, kh, kw));
    // stacked Cout = 64 needs 128 columns per accumulator: the single-CTA kernel runs 2 M-tiles x 2 buffers ("pairbuf"),
    // the pair kernel 3 M-tiles on a ring of 4 accumulator slots
    g.pairbuf = !pair && g.stacked && 16 * cout_pad > 512;
    const int mt_max = g.pairbuf ? 2 : (pair && g.stacked && 16 * cout_pad > 512 ? 3 : 4);
    for (int mt = (g.total_mt < mt_max ? g.total_mt : mt_max); mt >= 1; mt--) {
        int maxidx = g.P - 1 + mt * 128 - 1 + (kh - 1) * g.P + (kw - 1);
        int rbox = maxidx / g.P + 1;
        if (rbox > 256) continue;
        uint32_t plane = (uint32_t)rbox * g.P * 16;
        uint32_t act = plane * (uint32_t)ppg * g.groups;
        if (TC_SMEM_HEADER + act + (uint32_t)(kw + 3) * g.stage_bytes > TC_SMEM_MAX) continue;  // ring >= one filter row + 2
        g.MT = mt; g.Rbox = rbox; g.plane_bytes = plane; g.group_bytes = plane * (uint32_t)ppg; g.act_bytes = act;
        tc_ring_layout(g, g.stage_bytes, kh * kw);
        g.tiles = (g.total_mt + mt - 1) / mt;
        uint32_t cols = (g.stacked && !g.pairbuf ? 16u : 8u) * g.coutp, pc = 32;
        if (cols > 512) cols = 512;                  // pair kernel, stacked Cout = 64: 4 slots of 128 columns
        while (pc < cols) pc <<= 1;
        g.tmem_cols = pc;
        return pc <= 512;
    }
    return false;
}

bool tc_supported(int cin_pad, int cout_pad, int kh, int kw, int H, int W)
{
    TcGeom g;
    return tc_geometry(cin_pad, cout_pad, kh, kw, H, W, g);
}

size_t tc_packed_elems(int cin_pad, int cout_pad, int kh, int kw)
{
    return (size_t)kh * kw * (cin_pad / 16) * 2 * (2 * cout_pad) * 8;
}

static inline void host_split(float w, bool bf16, uint16_t &hi, uint16_t &lo)
{
    if (bf16) {
        __nv_bfloat16 h = __float2bfloat16_rn(w);
        __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        hi = *reinterpret_cast<uint16_t *>(&h); lo = *reinterpret_cast<uint16_t *>(&l);
    } else {
        __half h = __float2half_rn(w);
        __half l = __float2half_rn(w - __half2float(h));
        hi = *reinterpret_cast<uint16_t *>(&h); lo = *reinterpret_cast<uint16_t *>(&l);
    }
}

// [group][tap][k8 (2)][n (2*cout_pad: hi couts then lo couts)][8 cin] 16-bit; w is the reference's [cout][cin][kh][kw] fp32
void pack_tc_weights(const float *w, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, bool bf16, uint16_t *dst)
{
    const int groups = cin_pad / 16, N1 = 2 * cout_pad, taps = kh * kw;
    for (int t = 0; t < taps; t++)
        for (int g = 0; g < groups; g++)
            for (int k8 = 0; k8 < 2; k8++)
                for (int n = 0; n < cout_pad; n++)
                    for (int e = 0; e < 8; e++) {
                        const int c = g * 16 + k8 * 8 + e;
                        uint16_t hi = 0, lo = 0;
                        if (n < cout && c < cin) host_split(w[((size_t)n * cin + c) * taps + t], bf16, hi, lo);
                        const size_t base = (((size_t)g * taps + t) * 2 + k8) * N1 * 8;
                        dst[base + (size_t)n * 8 + e] = hi;
                        dst[base + (size_t)(cout_pad + n) * 8 + e] = lo;
                    }
}

// CTA-pair operand image: [rank (2)][group][tap] slabs, one per CTA and (group, tap):
//   unstacked : [k8 (2)][w_hi[h*r .. h*r+h) | w_lo[h*r .. h*r+h)][8], h = Cout/2                        (32*Cout bytes)
//   stacked   : [k8 (2)][this CTA's half of the N = 2*Cout operand [w_hi | w_lo] (rank 0: w_hi, rank 1:
//                           w_lo; Cout rows) | this CTA's half of the N = Cout operand w_hi[h*r .. h*r+h)][8] (48*Cout bytes)
static inline int tc_pair_slab_rows(int cout_pad, int kh, int kw)
{
    return tc_pair_stacked_layout(cout_pad, kh, kw) ? cout_pad + cout_pad / 2 : cout_pad;
}

size_t tc_pair_packed_elems(int cin_pad, int cout_pad, int kh, int kw)
{
    return (size_t)2 * (cin_pad / 16) * kh * kw * 2 * tc_pair_slab_rows(cout_pad, kh, kw) * 8;
}

void pack_tc_pair_weights(const float *w, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, bool bf16, uint16_t *dst)
{
    pack_tc_pair_weights_scheme(w, cout, cin, kh, kw, cin_pad, cout_pad, bf16, tc_pair_stacked_layout(cout_pad, kh, kw), dst);
}

// 1x1 shortcut conv fused into a kh x kw conv with the same Cout: its slabs use the HOST conv's accumulator scheme
size_t tc_pair_fused_sc_elems(int sc_cin_pad, int cout_pad, int kh, int kw)
{
    return (size_t)2 * (sc_cin_pad / 16) * 2 * tc_pair_slab_rows(cout_pad, kh, kw) * 8;
}
void pack_tc_pair_fused_sc(const float *w_sc, int cout, int sc_cin, int sc_cin_pad, int cout_pad, int kh, int kw, bool bf16, uint16_t *dst)
{
    pack_tc_pair_weights_scheme(w_sc, cout, sc_cin, 1, 1, sc_cin_pad, cout_pad, bf16, tc_pair_stacked_layout(cout_pad, kh, kw), dst);
}
bool tc_fusion_available() { return tc_pair_default(); }

void pack_tc_pair_weights_scheme(const float *w, int cout, int cin, int kh, int kw, int cin_pad, int cout_pad, bool bf16, bool st,
                                 uint16_t *dst)
{
    const int groups = cin_pad / 16, taps = kh * kw, half = cout_pad / 2, rows = st ? cout_pad + cout_pad / 2 : cout_pad;
    auto weight = [&](int co, int c, int t, uint16_t &hi, uint16_t &lo) {
        hi = lo = 0;
        if (co < cout && c < cin) host_split(w[((size_t)co * cin + c) * taps + t], bf16, hi, lo);
    };
    for (int r = 0; r < 2; r++)
        for (int g = 0; g < groups; g++)
            for (int t = 0; t < taps; t++)
                for (int k8 = 0; k8 < 2; k8++) {
                    uint16_t *slab = dst + ((((size_t)r * groups + g) * taps + t) * 2 + k8) * rows * 8;
                    for (int e = 0; e < 8; e++) {
                        const int c = g * 16 + k8 * 8 + e;
                        uint16_t hi, lo;
                        if (st) {
                            for (int n = 0; n < cout_pad; n++) {
                                weight(n, c, t, hi, lo);
                                slab[(size_t)n * 8 + e] = r == 0 ? hi : lo;
                            }
                            for (int n = 0; n < half; n++) {
                                weight(half * r + n, c, t, hi, lo);
                                slab[(size_t)(cout_pad + n) * 8 + e] = hi;
                            }
                        } else {
                            for (int n = 0; n < half; n++) {
                                weight(half * r + n, c, t, hi, lo);
                                slab[(size_t)n * 8 + e] = hi;
                                slab[(size_t)(half + n) * 8 + e] = lo;
                            }
                        }
                    }
                }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// issuers waiting for a weight stage (they run ahead of the tensor pipe until the ring is exhausted, 25-50 % of their time):
// build-time A/B knob PMP_ISSUER_BACKOFF=<ns> sleeps between probes instead of spinning
__device__ __forceinline__ void mbar_wait_issuer(uint32_t bar, uint32_t parity)
{
#ifdef PMP_ISSUER_BACKOFF
    while (!mbar_try_wait(bar, parity)) __nanosleep(PMP_ISSUER_BACKOFF);
#else
    while (!mbar_try_wait(bar, parity)) {}
#endif
}
// for the roles that wait long (epilogue for a whole MMA phase, producers for a free slot): back off between probes
// so the polling does not take issue slots (and power) from the MMA issuers
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(100);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// element (row, k) lives at start + (row/8)*SBO + (row%8)*16 + (k/8)*LBO + (k%8)*2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t elect_one_sync()
{
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
                 : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

struct TcParams {
    const uint16_t *w;
    const float *bias;
    Act out, res, mul;
    int H, W, P, kh, kw, pady, padx, groups, total_mt, tiles, N1, coutp, nstages, items;
    uint32_t plane_bytes, group_bytes, stage_bytes, tmem_cols, idesc1, idesc2;
    int relu, stacked, pairbuf, nbuf, dbg, hpool;
    int nslot, mt_alloc, acc_cols;      // CTA-pair kernel: accumulator slot ring (slots, M-tiles per tile, columns per slot)
    int epi_groups;                     // CTA-pair kernel: epilogue groups draining accumulator uses round-robin (2 x 8 warps or 4 x 4 warps)
    int aslots, groups2;                // CTA-pair kernel: activation slot ring (one channel-group box each); fused shortcut groups
    // CTA-pair kernel, first-layer ("stem") mode: the A operand is assembled by TMA from 8 pre-shifted copies of each source
    // plane (stem_shift_kernel); a K chunk of 8 unrolled channels = (plane, 8 consecutive kx shifts).  hi planes only.
    int stem, stem_planes;
    signed char stem_chunk_plane[8], stem_chunk_u0[8];
    int B, pair_items;      // CTA-pair kernel: images in the batch, work items = tiles * ceil(B/2)
    uint32_t pair_slab;     // CTA-pair kernel: bytes of one per-CTA weight slab
    int pair_split;         // CTA-pair kernel: rows of the weight tensor map per slab (1, or 2 for 3 KB slabs)
    int kws;                // CTA-pair kernel: taps per ring stage (kw = one filter row, or 1 when rows do not fit)
    unsigned int *sat;      // handle status word 0: count of epilogue threads that produced |value| > 65504 with fp16 operands
                            // (the hi/lo split clamps there: results beyond are wrong) -- sticky, read by pmp_saturation_count
};

struct TileGeom { int n, mt_count, q0, row0, qoff; };

// Stall profile of the pair kernel (PMP_TC_DBG bit 6): cycles each role spent waiting on each barrier class, per CTA.
// slots: 0 total (issuer 0), 1 issuer acc_empty, 2 issuer act_full, 3 issuer w_full, 4 weight producer w_empty,
// 5 activation producer act_empty, 6 epilogue acc_full, 7 epilogue body, 8 items, 9 epilogue total
constexpr int TC_PROF_SLOTS = 16;
__device__ unsigned long long g_tc_stalls[160 * TC_PROF_SLOTS];
#define TC_PROF_BEGIN(flag) long long _pt0 = (flag) ? clock64() : 0
#define TC_PROF_END(flag, acc) do { if (flag) { long long _pt1 = clock64(); (acc) += _pt1 - _pt0; } } while (0)

__device__ __forceinline__ TileGeom tile_geom(const TcParams &p, int item)
{
    TileGeom t;
    t.n = item / p.tiles;
    const int tl = item - t.n * p.tiles;
    const int mt_begin = (tl * p.total_mt) / p.tiles, mt_end = ((tl + 1) * p.total_mt) / p.tiles;
    t.mt_count = mt_end - mt_begin;
    t.q0 = mt_begin * 128;
    t.row0 = t.q0 / p.P;
    t.qoff = t.q0 - t.row0 * p.P;
    return t;
}

// packed pair versions of split16 / join16 (tensor.cuh): identical arithmetic, two elements per instruction
__device__ __forceinline__ void split2(float a, float b, bool bf, uint32_t &hi, uint32_t &lo)
{
    if (bf) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        float2 hf = __bfloat1622float2(h);
        __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
        hi = *reinterpret_cast<uint32_t *>(&h); lo = *reinterpret_cast<uint32_t *>(&l);
    } else {
        a = fminf(fmaxf(a, -65504.f), 65504.f);
        b = fminf(fmaxf(b, -65504.f), 65504.f);
        __half2 h = __floats2half2_rn(a, b);
        float2 hf = __half22float2(h);
        __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        hi = *reinterpret_cast<uint32_t *>(&h); lo = *reinterpret_cast<uint32_t *>(&l);
    }
}
__device__ __forceinline__ float2 join2(uint32_t hi, uint32_t lo, bool bf)
{
    float2 h, l;
    if (bf) {
        h = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&hi));
        l = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&lo));
    } else {
        h = __half22float2(*reinterpret_cast<__half2 *>(&hi));
        l = __half22float2(*reinterpret_cast<__half2 *>(&lo));
    }
    return make_float2(h.x + l.x, h.y + l.y);
}
__device__ __forceinline__ void unpack_split(const uint4 &H, const uint4 &L, bool bf, float v[8])
{
    const uint32_t hw[4] = {H.x, H.y, H.z, H.w}, lw[4] = {L.x, L.y, L.z, L.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float2 t = join2(hw[e], lw[e], bf);
        v[2 * e] = t.x; v[2 * e + 1] = t.y;
    }
}

// fp16 range guard of the hi/lo split (split2 clamps to +-65504): report instead of clamping silently
template <int CH>
__device__ __forceinline__ void saturation_check(const TcParams &p, const float (&v)[CH][8])
{
    float am = 0.f;
#pragma unroll
    for (int j = 0; j < CH; j++)
#pragma unroll
        for (int e = 0; e < 8; e++) am = fmaxf(am, fabsf(v[j][e]));
    if (am > 65504.f && !p.out.bf16 && p.sat) atomicAdd(p.sat, 1u);
}

// Epilogue operand fetch of one M-tile position: the residual (or, when there is no residual -- the Att trunks' last block
// has a fused shortcut -- the attention product's operand, which then takes the residual's registers) for CH consecutive
// 8-channel chunks.  Issued ahead of the accumulator wait / the TMEM loads so that its latency overlaps them.
template <int CH>
__device__ __forceinline__ void epilogue_fetch(const TcParams &p, int ch0, int n, int r, int c, bool valid, uint4 *rh, uint4 *rl)
{
    const bool mul_early = p.mul.p && !p.res.p;
    if ((p.res.p || mul_early) && valid) {
        const size_t plane = (size_t)p.H * p.W;
        const Act &t = p.res.p ? p.res : p.mul;
        const uint4 *rb = reinterpret_cast<const uint4 *>(t.p) + (size_t)n * (t.Cp >> 2) * plane + (size_t)r * p.W + c;
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const uint4 *q = rb + (size_t)split_plane(ch0 + j, 0) * plane;
            rh[j] = ldg_stream(q);
            rl[j] = ldg_stream(q + 2 * plane);
        }
    }
}

// Epilogue of one M-tile for CH consecutive 8-channel chunks starting at chunk ch0 (one thread = one position); rh / rl
// hold what epilogue_fetch<CH> loaded for the same chunks.
template <int CH>
__device__ __forceinline__ void epilogue_process(const TcParams &p, uint32_t taddr, int ch0, int n, int r, int c, bool valid,
                                                 const uint4 *rh, const uint4 *rl)
{
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)r * p.W + c;
    const bool mul_early = p.mul.p && !p.res.p;
    // The tcgen05.ld / wait::ld below are .sync.aligned: the whole warp must arrive together.  A previous call of this
    // function for the same M-tile ends in a divergent tail (invalid positions return early, the others still store and,
    // with a residual AND an attention operand, wait for late loads), and nothing forces reconvergence between the two
    // calls: measured as a hang of 3x3 64->64 @16x16 with both operands (75 % invalid lanes in the last M-tile).
    __syncwarp();
    float v[CH][8];
    if constexpr (CH <= 2) {
        // pair kernel (CH <= 2 per call, 104 registers): every TMEM load of the call in flight before the one wait
        uint32_t a[CH][8], b[CH][8];
#pragma unroll
        for (int j = 0; j < CH; j++) tmem_ld8(taddr + 8 * (ch0 + j), a[j]);
        if (p.stacked) {        // columns [Cout, 2*Cout) hold a_hi * w_lo
#pragma unroll
            for (int j = 0; j < CH; j++) tmem_ld8(taddr + p.coutp + 8 * (ch0 + j), b[j]);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < CH; j++)
#pragma unroll
                for (int e = 0; e < 8; e++) v[j][e] = __uint_as_float(a[j][e]) + __uint_as_float(b[j][e]);
        } else {
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < CH; j++)
#pragma unroll
                for (int e = 0; e < 8; e++) v[j][e] = __uint_as_float(a[j][e]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < CH; j++) {         // per chunk: small transient register footprint (TMEM latency is short)
            uint32_t a[8];
            tmem_ld8(taddr + 8 * (ch0 + j), a);
            if (p.stacked) {
                uint32_t b[8];
                tmem_ld8(taddr + p.coutp + 8 * (ch0 + j), b);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; e++) v[j][e] = __uint_as_float(a[e]) + __uint_as_float(b[e]);
            } else {
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; e++) v[j][e] = __uint_as_float(a[e]);
            }
        }
    }
    if (p.hpool) {
        // Horizontal half of a fused 2x2 max-pool (pooled ResidualBlocks have no bias / attention product here and an
        // identity-or-fused shortcut): the pair (c, c+1), c even, sits in adjacent lanes (the pitch is even), so the
        // exchange is one shuffle per value -- every lane takes part, also the ones on pad columns.  relu(max) == max(relu),
        // but the residual must be added before the max: done here, ahead of the common path.
        if (p.res.p) {
            const bool bfr = p.out.bf16 != 0;
#pragma unroll
            for (int j = 0; j < CH; j++) {
                float rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) unpack_split(rh[j], rl[j], bfr, rv);
#pragma unroll
                for (int e = 0; e < 8; e++) v[j][e] += rv[e];
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++)
#pragma unroll
            for (int e = 0; e < 8; e++) v[j][e] = fmaxf(v[j][e], __shfl_xor_sync(0xffffffffu, v[j][e], 1));
        if (!valid || (c & 1)) return;
        saturation_check<CH>(p, v);
        const bool bfh = p.out.bf16 != 0;
        const size_t hplane = (size_t)p.H * (p.W >> 1);
        uint4 *obh = reinterpret_cast<uint4 *>(p.out.p) + (size_t)n * (p.out.Cp >> 2) * hplane + (size_t)r * (p.W >> 1) + (c >> 1);
#pragma unroll
        for (int j = 0; j < CH; j++) {
            if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; e++) v[j][e] = fmaxf(v[j][e], 0.f);
            }
            uint4 Hh, Ll;
            split2(v[j][0], v[j][1], bfh, Hh.x, Ll.x);
            split2(v[j][2], v[j][3], bfh, Hh.y, Ll.y);
            split2(v[j][4], v[j][5], bfh, Hh.z, Ll.z);
            split2(v[j][6], v[j][7], bfh, Hh.w, Ll.w);
            uint4 *q = obh + (size_t)split_plane(ch0 + j, 0) * hplane;
            q[0] = Hh;
            q[2 * hplane] = Ll;
        }
        return;
    }
    if (!valid) return;
    const bool bf = p.out.bf16 != 0;
    uint4 *ob = reinterpret_cast<uint4 *>(p.out.p) + (size_t)n * (p.out.Cp >> 2) * plane + pix;
#pragma unroll
    for (int j = 0; j < CH; j++) {
        if (p.bias) {       // stems only; padded to Cout_pad on the host
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(p.bias) + 2 * (ch0 + j));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(p.bias) + 2 * (ch0 + j) + 1);
            v[j][0] += b0.x; v[j][1] += b0.y; v[j][2] += b0.z; v[j][3] += b0.w;
            v[j][4] += b1.x; v[j][5] += b1.y; v[j][6] += b1.z; v[j][7] += b1.w;
        }
        if (p.res.p) {
            float rv[8];
            unpack_split(rh[j], rl[j], bf, rv);
#pragma unroll
            for (int e = 0; e < 8; e++) v[j][e] += rv[e];
        }
        if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; e++) v[j][e] = fmaxf(v[j][e], 0.f);
        }
        if (p.mul.p) {      // attention product: only the last conv of the two Att trunks
            float mv[8];
            if (mul_early) {
                unpack_split(rh[j], rl[j], bf, mv);
            } else {
                const uint4 *q = reinterpret_cast<const uint4 *>(p.mul.p) + (size_t)n * (p.mul.Cp >> 2) * plane + pix +
                                 (size_t)split_plane(ch0 + j, 0) * plane;
                unpack_split(ldg_stream(q), ldg_stream(q + 2 * plane), bf, mv);
            }
#pragma unroll
            for (int e = 0; e < 8; e++) v[j][e] *= mv[e];
        }
        uint4 Hh, Ll;
        split2(v[j][0], v[j][1], bf, Hh.x, Ll.x);
        split2(v[j][2], v[j][3], bf, Hh.y, Ll.y);
        split2(v[j][4], v[j][5], bf, Hh.z, Ll.z);
        split2(v[j][6], v[j][7], bf, Hh.w, Ll.w);
        uint4 *q = ob + (size_t)split_plane(ch0 + j, 0) * plane;
        q[0] = Hh;
        q[2 * plane] = Ll;
    }
    saturation_check<CH>(p, v);
}

template <int CH>
__device__ __forceinline__ void epilogue_chunks(const TcParams &p, uint32_t taddr, int ch0, int n, int r, int c, bool valid)
{
    uint4 rh[CH], rl[CH];
    epilogue_fetch<CH>(p, ch0, n, r, c, valid, rh, rl);
    epilogue_process<CH>(p, taddr, ch0, n, r, c, valid, rh, rl);
}

// Persistent, warp-specialised: warp 0 weight producer, warp 1 activation producer + TMEM owner, warps 2..5 MMA issuers
// (one per M-tile of the CTA tile: a single issuing thread cannot keep the tensor pipe fed), warps 6..13 epilogue
// (TMEM lane quarter = warp % 4, channel half = (warp - 6) / 4).
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    // barrier slots: [0,16) act_full and [16,32) act_empty per (activation buffer, channel group), [32,56) w_full,
    // [56,80) w_empty, [80,82) acc_full per accumulator buffer, [82,90) acc_empty per accumulator (buffer, M-tile);
    // TMEM base address at byte 1008
    const uint32_t bar_afull = smem_u32(bars), bar_aempty = smem_u32(bars + 16), bar_wfull = smem_u32(bars + 32),
                   bar_wempty = smem_u32(bars + 56), bar_acc = smem_u32(bars + 80), bar_accempty = smem_u32(bars + 82);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + 1008);
    uint8_t *act = smem + TC_SMEM_HEADER;
    uint8_t *ring = act + (size_t)p.nbuf * p.groups * p.group_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.kh * p.kw;

    if (threadIdx.x == 0) {
        for (int g = 0; g < p.nbuf * p.groups; g++) { mbar_init(bar_afull + 8 * g, 1); mbar_init(bar_aempty + 8 * g, TC_MMA_WARPS); }
        for (int s = 0; s < p.nstages; s++) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, TC_MMA_WARPS); }
        mbar_init(bar_acc, p.pairbuf ? 2 : TC_MMA_WARPS);
        mbar_init(bar_acc + 8, p.pairbuf ? 2 : TC_MMA_WARPS);
        for (int m = 0; m < 8; m++) mbar_init(bar_accempty + 8 * m, TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + 1008)),
                     "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== weight producer: one (group, tap) slab [2][N1][8] per ring stage, same order for every tile =====
            const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(p.w);
            const int per_item = taps * p.groups;
            uint32_t s = 0, ph = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const uint8_t *src = wsrc;
                for (int it = 0; it < per_item; it++, src += p.stage_bytes) {
                    mbar_wait_relaxed(bar_wempty + 8 * s, ph ^ 1u);
                    mbar_expect_tx(bar_wfull + 8 * s, p.stage_bytes);
                    bulk_g2s(smem_u32(ring) + s * p.stage_bytes, src, p.stage_bytes, bar_wfull + 8 * s);
                    if (++s == (uint32_t)p.nstages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== activation producer: one TMA box per 16-channel group; buffer g is refilled for the next tile as
            //       soon as the MMAs of group g of the current tile have drained it.
            //       With nbuf > 1 (small images) whole tiles are loaded nbuf - 1 items ahead. =====
            uint32_t ab = 0, aph = 0;                 // activation buffer of this item, parity of its use count
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const TileGeom t = tile_geom(p, item);
                for (int g = 0; g < p.groups; g++) {
                    const uint32_t slot = ab * (uint32_t)p.groups + (uint32_t)g;
                    mbar_wait_relaxed(bar_aempty + 8 * slot, aph ^ 1u);
                    mbar_expect_tx(bar_afull + 8 * slot, p.group_bytes);
                    tma_load_4d(smem_u32(act + (size_t)slot * p.group_bytes), &tmap, bar_afull + 8 * slot, -2 * p.padx,
                                t.row0 - p.pady, g * 4, t.n);
                }
                if (++ab == (uint32_t)p.nbuf) { ab = 0; aph ^= 1u; }
            }
        }
    } else if (warp < 2 + TC_MMA_WARPS) {
        // ===== MMA issuers: warp 2+m owns M-tile m (accumulator columns [m*N1, (m+1)*N1)).  The whole warp runs the
        //       warp-uniform loop and one elected lane issues.  Descriptors are linear in the start address (16-byte
        //       units in the low 14 bits), so the constant parts are built once. =====
        const int m = warp - 2;
        const uint64_t desc_c = ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);             // SBO 128, version 1
        const uint64_t adesc_c = desc_c | ((uint64_t)(p.plane_bytes >> 4) << 16);                // LBO = plane stride
        const uint64_t bdesc_c = desc_c | ((uint64_t)(((uint32_t)p.N1 * 16u) >> 4) << 16);       // LBO = N1 rows
        const uint32_t act16 = smem_u32(act) >> 4, ring16 = smem_u32(ring) >> 4;
        const uint32_t group16 = p.group_bytes >> 4, stage16 = p.stage_bytes >> 4, lo16 = (2u * p.plane_bytes) >> 4;
        const uint32_t idesc = p.idesc2, idesc_st = p.idesc1, wlo16 = (uint32_t)p.coutp;        // w_lo rows follow w_hi rows
        const int KH = p.kh, K = p.kw, P = p.P, NS = p.nstages, G = p.groups;
        uint32_t s = 0, ph = 0, idx = 0, ab = 0, aph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, idx++) {
            const TileGeom t = tile_geom(p, item);
            const uint32_t buf = idx & 1u;                      // tiles alternate between the two accumulator buffers
            // default : issuer m owns M-tile m of every tile (accumulator buf*4+m, `accw` columns wide).
            // pairbuf : issuer m owns M-tile m&1 of the tiles whose buffer is m>>1 (accumulator m); the other two
            //           issuers only walk the barriers (parity waits must observe every phase).
            const bool owner = !p.pairbuf || (uint32_t)(m >> 1) == buf;
            const int mym = p.pairbuf ? (m & 1) : m;
            const bool mine = owner && mym < t.mt_count;
            const uint32_t accidx = p.pairbuf ? (uint32_t)m : buf * 4u + (uint32_t)m;
            const uint32_t d_tmem = tmem_base + accidx * (uint32_t)(p.stacked ? p.N1 : p.coutp);
            if (owner) {
                mbar_wait(bar_accempty + 8 * accidx, ((idx >> 1) & 1u) ^ 1u);   // epilogue drained it (tile idx-2)
                tc_fence_after();
            }
            const uint32_t am16 = (uint32_t)mym * 128u;
            uint32_t acc = 0;
            for (int g = 0; g < G; g++) {
                const uint32_t slot = ab * (uint32_t)G + (uint32_t)g;
                mbar_wait(bar_afull + 8 * slot, aph);
                const uint32_t ag = act16 + am16 + slot * group16 + (uint32_t)t.qoff;
                for (int ky = 0; ky < KH; ky++) {
                    // one filter row = K weight slabs: probe all their barriers back to back, then spin on stragglers
                    uint32_t sj[5], pj[5];
                    bool ok[5];
                    {
                        uint32_t s2 = s, ph2 = ph;
#pragma unroll
                        for (int j = 0; j < 5; j++) {
                            if (j < K) {
                                sj[j] = s2; pj[j] = ph2;
                                ok[j] = mbar_try_wait(bar_wfull + 8 * s2, ph2);
                                if (++s2 == (uint32_t)NS) { s2 = 0; ph2 ^= 1u; }
                            }
                        }
                        s = s2; ph = ph2;
                    }
#pragma unroll
                    for (int j = 0; j < 5; j++)
                        if (j < K) while (!ok[j]) ok[j] = mbar_try_wait(bar_wfull + 8 * sj[j], pj[j]);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        if (mine) {
                            const uint64_t ad0 = adesc_c | (uint64_t)(ag + (uint32_t)(ky * P));
#pragma unroll
                            for (int j = 0; j < 5; j++) {
                                if (j < K) {
                                    const uint64_t ad = ad0 + (uint64_t)j;
                                    const uint64_t bd = bdesc_c | (uint64_t)(ring16 + sj[j] * stage16);
                                    if (p.stacked) {
                                        umma_f16(d_tmem, ad, bd, idesc_st, acc);                    // a_hi * [w_hi | w_lo]
                                        umma_f16(d_tmem, ad + (uint64_t)lo16, bd, idesc, 1u);       // a_lo * w_hi
                                    } else {
                                        umma_f16(d_tmem, ad, bd, idesc, acc);                       // a_hi * w_hi
                                        umma_f16(d_tmem, ad, bd + (uint64_t)wlo16, idesc, 1u);      // a_hi * w_lo
                                        umma_f16(d_tmem, ad + (uint64_t)lo16, bd, idesc, 1u);       // a_lo * w_hi
                                    }
                                    umma_commit(bar_wempty + 8 * sj[j]);                        // slab free when read
                                    acc = 1u;
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 5; j++)
                                if (j < K) mbar_arrive(bar_wempty + 8 * sj[j]);
                        }
                    }
                    __syncwarp();
                    acc = 1u;
                }
                if (elect_one_sync()) {                          // activation buffer free for a later tile
                    if (mine) umma_commit(bar_aempty + 8 * slot);
                    else mbar_arrive(bar_aempty + 8 * slot);
                }
                __syncwarp();
            }
            if (owner && elect_one_sync()) {                     // accumulators of this tile complete
                if (mine) umma_commit(bar_acc + 8 * buf);
                else mbar_arrive(bar_acc + 8 * buf);
            }
            __syncwarp();
            if (++ab == (uint32_t)p.nbuf) { ab = 0; aph ^= 1u; }
        }
    } else {
        // ===== epilogue =====
        const int quarter = warp & 3, half = (warp - 2 - TC_MMA_WARPS) >> 2;
        const int nchunk = p.coutp >> 3, chh = nchunk >> 1, ch0 = half * chh;
        uint32_t idx = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, idx++) {
            const TileGeom t = tile_geom(p, item);
            // while the MMAs run: pull this tile's residual / attention operand lines into L2
            if ((p.res.p || p.mul.p) && (lane & 7) == 0) {
                const size_t plane = (size_t)p.H * p.W;
                for (int mt = 0; mt < t.mt_count; mt++) {
                    const int pos = t.q0 + mt * 128 + quarter * 32 + lane;
                    const int r = pos / p.P, c = pos - r * p.P;
                    if (c < p.W && r < p.H) {
                        for (int j = 0; j < chh; j++) {
                            const size_t o = ((size_t)t.n * (p.out.Cp >> 2) + split_plane(ch0 + j, 0)) * plane + (size_t)r * p.W + c;
                            if (p.res.p) { prefetch_l2(reinterpret_cast<const uint4 *>(p.res.p) + o); prefetch_l2(reinterpret_cast<const uint4 *>(p.res.p) + o + 2 * plane); }
                            if (p.mul.p) { prefetch_l2(reinterpret_cast<const uint4 *>(p.mul.p) + o); prefetch_l2(reinterpret_cast<const uint4 *>(p.mul.p) + o + 2 * plane); }
                        }
                    }
                }
            }
            const uint32_t buf = idx & 1u;
            mbar_wait_relaxed(bar_acc + 8 * buf, (idx >> 1) & 1u);
            tc_fence_after();
            for (int mt = 0; mt < t.mt_count; mt++) {
                const int pos = t.q0 + mt * 128 + quarter * 32 + lane;
                const int r = pos / p.P, c = pos - r * p.P;
                const bool valid = (c < p.W) && (r < p.H);
                const uint32_t accidx = p.pairbuf ? 2u * buf + (uint32_t)mt : 4u * buf + (uint32_t)mt;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + accidx * (uint32_t)(p.stacked ? p.N1 : p.coutp);
                if (chh == 4) epilogue_chunks<4>(p, taddr, ch0, t.n, r, c, valid);
                else
                    for (int j = 0; j < chh; j++) epilogue_chunks<1>(p, taddr, ch0 + j, t.n, r, c, valid);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty + 8 * accidx);      // this accumulator may be overwritten
            }
            // M-tiles this tile did not use still owe their issuer warp an arrival
            for (int mt = t.mt_count; mt < (p.pairbuf ? 2 : 4); mt++) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty + 8 * ((p.pairbuf ? 2 : 4) * buf + mt));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): the two CTAs of a cluster process the same tile index of two consecutive images.
// One tcgen05.mma with M = 256 drives both SMs: each SM reads its own 128-position A tile and only HALF of the weight
// operand (N/2 = 32 rows) from its shared memory -- the single-CTA kernel is bound by shared-memory operand reads
// (18 KB per K=16 step, measured), the pair reads 15 KB.  Unstacked scheme only (Cout = 64), same tiling, TMEM double
// buffering and warp roles as conv_tc_kernel; differences:
//   * operands of BOTH CTAs are announced on the LEADER's (cluster rank 0) full barriers: the peer's TMA loads use
//     .cta_group::2 with the leader's barrier address, the leader expects twice the bytes;
//   * only the leader's four issuer warps issue MMAs; completion is multicast-committed to the empty / acc_full barriers
//     of both CTAs; the epilogues of both CTAs report the TMEM drain to the leader's acc_empty barriers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr)     // no fence: orders nothing but the count
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap *tmap, uint32_t leader_bar, int c0, int c1,
                                                 int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap *tmap, uint32_t leader_bar, int c0, int c1,
                                                 int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *tmap, uint32_t leader_bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar)       // arrives on `bar` (same offset) in both CTAs
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct PairGeom { int n, mt_count, q0, row0, qoff; bool store; };

// per-CTA weight slab: [2 k8][Cout/2 w_hi rows | Cout/2 w_lo rows][8] 16-bit = 32 * Cout bytes (p.pair_slab).  A ring stage
// holds one FILTER ROW of one channel group (kw consecutive slabs, contiguous in the packed image): one full/empty barrier
// pair, one TMA request and one commit per row instead of per tap.

__device__ __forceinline__ PairGeom pair_geom(const TcParams &p, int item, int rank)
{
    PairGeom t;
    const int pr = item / p.tiles, tl = item - pr * p.tiles;
    const int n = 2 * pr + rank;
    t.store = n < p.B;
    t.n = t.store ? n : p.B - 1;          // odd batch: the peer recomputes the last image and drops the result
    const int mt_begin = (tl * p.total_mt) / p.tiles, mt_end = ((tl + 1) * p.total_mt) / p.tiles;
    t.mt_count = mt_end - mt_begin;
    t.q0 = mt_begin * 128;
    t.row0 = t.q0 / p.P;
    t.qoff = t.q0 - t.row0 * p.P;
    return t;
}

struct PairBars { uint32_t afull, aempty, wfull, wempty, acc, accempty; };
constexpr uint32_t TC_PAIR_ISSUED_OFF = 960;     // shared-memory header: uint32 issue tickets per accumulator slot (8 slots)

// MMA issuer of the leader CTA: warp 2+m owns M-tile m of BOTH CTAs' tiles (accumulator buf*4+m).  Measured
// (tools/stall_prof.py): the issuers spent 85 % of their time executing the issue loop itself, not waiting -- ~200
// instructions per filter row at IPC 0.1 -- so the loop is kept to the bare minimum: one barrier probe and one commit
// per filter row (row-granular ring stages), the row's kw taps unrolled at compile time with descriptors that differ by
// compile-time constants, every operand warp-uniform so that it lives in uniform registers.
template <int KW, bool ST>
__device__ __forceinline__ void pair_issuer(const TcParams &p, const PairBars &b, uint32_t tmem_base, uint32_t act_addr,
                                            uint32_t ring_addr, int m, int cid, int ncl, bool prof, long long t_begin,
                                            volatile uint32_t *issued)
{
    const uint64_t desc_c = ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);                 // SBO 128, version 1
    const uint64_t adesc_c = desc_c | ((uint64_t)(p.plane_bytes >> 4) << 16);                    // LBO = plane stride
    const uint64_t bdesc_c = desc_c | ((uint64_t)(p.pair_slab >> 5) << 16);                      // LBO = slab rows per k8
    const uint32_t act16 = (act_addr >> 4) + (uint32_t)m * 128u, ring16 = ring_addr >> 4;
    const uint32_t group16 = p.group_bytes >> 4, slab16 = p.pair_slab >> 4, row16 = (uint32_t)KW * slab16;
    const int RS = p.kw / KW;                    // ring stages per filter row (1, or kw single-tap stages)
    const uint64_t lo16 = (2u * p.plane_bytes) >> 4;            // a_lo planes follow the two a_hi planes
    // unstacked: this CTA's w_lo rows follow its w_hi rows; stacked: its rows of the N = Cout operand follow its Cout
    // rows of the N = 2*Cout operand
    const uint64_t wlo16 = (uint64_t)(ST ? p.coutp : (p.coutp >> 1));
    const uint32_t idesc = p.idesc1, idesc_st = p.idesc2;
    const uint32_t acc_cols = (uint32_t)p.acc_cols, NSLOT = (uint32_t)p.nslot, MTA = (uint32_t)p.mt_alloc;
    const bool has_slot = (uint32_t)m < MTA;
    uint32_t slot = (uint32_t)m, sph = 0;       // this issuer's accumulator slot for the current tile, parity of its use count
    const uint32_t peer_wempty = mapa_cluster(b.wempty, 1), peer_aempty = mapa_cluster(b.aempty, 1), peer_acc = mapa_cluster(b.acc, 1);
    const int KH = p.kh, P = p.P, NS = p.nstages, G = p.groups;
    long long st_a = 0, st_b = 0, st_c = 0;
    const bool hi_only = p.stem != 0;            // stems: 8-bit pixels (and the pre-split qt planes) are exact in 16 bits
    const int G2 = p.groups2;                    // fused 1x1 shortcut: extra channel groups of a second input, centre tap only
    const uint32_t ASLOTS = (uint32_t)p.aslots, ctr16 = (uint32_t)(p.pady * p.P + p.padx);
    uint32_t s = 0, ph = 0, idx = 0, as = 0, aph = 0;   // weight stage / activation slot of the ring and their parities
    for (int item = cid; item < p.pair_items; item += ncl, idx++) {
        const PairGeom t = pair_geom(p, item, 0);
        const bool mine = m < t.mt_count;
        const uint32_t d_tmem = tmem_base + slot * acc_cols;
        if (has_slot) {
            TC_PROF_BEGIN(prof);
            // With 4 slots and 3 M-tiles per item the slots rotate through the issuers: consecutive uses of a slot belong to
            // DIFFERENT warps.  A parity wait only tells "the phase before the current one is complete"; an issuer that ran
            // two uses of the slot ahead of the one that owns the use in between (the epilogue groups drain out of global
            // order, the weight ring allows one to two items of skew) would see the still-incomplete phase k-2 as "k-1
            // complete" and overwrite a live accumulator (measured: hang of 3x3 32->64 / 64->64 @16x16 with slow
            // epilogues).  Issue tickets serialise the uses of a slot: use k waits until the owner of use k-1 has passed its
            // own wait.  Cycle-free: that owner is behind in the weight ring, never waiting on stages this warp holds.
            const uint32_t k_use = (idx * MTA + (uint32_t)m) >> (NSLOT == 8u ? 3 : 2);
            while (issued[slot] < k_use) {}
            mbar_wait(b.accempty + 8 * slot, sph ^ 1u);          // both CTAs drained its previous use
            if ((threadIdx.x & 31) == 0) issued[slot] = k_use + 1u;
            TC_PROF_END(prof, st_a);
            tc_fence_after();
        }
        uint32_t acc = 0;
        for (int g = 0; g < G; g++) {
            { TC_PROF_BEGIN(prof); mbar_wait(b.afull + 8 * as, aph); TC_PROF_END(prof, st_b); }
            uint32_t arow = act16 + as * group16 + (uint32_t)t.qoff;
            for (int kr = 0, sx = 0; kr < KH * RS; kr++) {
                { TC_PROF_BEGIN(prof); mbar_wait_issuer(b.wfull + 8 * s, ph); TC_PROF_END(prof, st_c); }
                tc_fence_after();
                const uint64_t ad0 = adesc_c | (uint64_t)(arow + (uint32_t)(sx * KW));
                const uint64_t bd0 = bdesc_c | (uint64_t)(ring16 + s * row16);
                if (elect_one_sync()) {
                    if (mine) {
#pragma unroll
                        for (int j = 0; j < KW; j++) {
                            const uint64_t ad = ad0 + (uint64_t)j, bd = bd0 + (uint64_t)((uint32_t)j * slab16);
                            if (ST) {
                                umma_f16_pair(d_tmem, ad, bd, idesc_st, j == 0 ? acc : 1u);     // a_hi * [w_hi | w_lo]
                                if (!hi_only) umma_f16_pair(d_tmem, ad + lo16, bd + wlo16, idesc, 1u);   // a_lo * w_hi
                            } else {
                                umma_f16_pair(d_tmem, ad, bd, idesc, j == 0 ? acc : 1u);        // a_hi * w_hi
                                umma_f16_pair(d_tmem, ad, bd + wlo16, idesc, 1u);               // a_hi * w_lo
                                umma_f16_pair(d_tmem, ad + lo16, bd, idesc, 1u);                // a_lo * w_hi
                            }
                        }
                        umma_commit_pair(b.wempty + 8 * s);                             // row free when read
                    } else {
                        mbar_arrive(b.wempty + 8 * s);
                        mbar_arrive_cluster_relaxed(peer_wempty + 8 * s);
                    }
                }
                __syncwarp();
                acc = 1u;
                if (++s == (uint32_t)NS) { s = 0; ph ^= 1u; }
                if (++sx == RS) { sx = 0; arow += (uint32_t)P; }
            }
            if (elect_one_sync()) {                              // activation buffer free for a later tile
                if (mine) umma_commit_pair(b.aempty + 8 * as);
                else { mbar_arrive(b.aempty + 8 * as); mbar_arrive_cluster_relaxed(peer_aempty + 8 * as); }
            }
            __syncwarp();
            if (++as == ASLOTS) { as = 0; aph ^= 1u; }
        }
        for (int g = 0; g < G2; g++) {           // out += W_sc * x: one slab per group, A = the tile's own positions
            { TC_PROF_BEGIN(prof); mbar_wait(b.afull + 8 * as, aph); TC_PROF_END(prof, st_b); }
            { TC_PROF_BEGIN(prof); mbar_wait_issuer(b.wfull + 8 * s, ph); TC_PROF_END(prof, st_c); }
            tc_fence_after();
            const uint64_t ad = adesc_c | (uint64_t)(act16 + as * group16 + (uint32_t)t.qoff + ctr16);
            const uint64_t bd = bdesc_c | (uint64_t)(ring16 + s * row16);
            if (elect_one_sync()) {
                if (mine) {
                    if (ST) {
                        umma_f16_pair(d_tmem, ad, bd, idesc_st, 1u);
                        umma_f16_pair(d_tmem, ad + lo16, bd + wlo16, idesc, 1u);
                    } else {
                        umma_f16_pair(d_tmem, ad, bd, idesc, 1u);
                        umma_f16_pair(d_tmem, ad, bd + wlo16, idesc, 1u);
                        umma_f16_pair(d_tmem, ad + lo16, bd, idesc, 1u);
                    }
                    umma_commit_pair(b.wempty + 8 * s);
                    umma_commit_pair(b.aempty + 8 * as);
                } else {
                    mbar_arrive(b.wempty + 8 * s); mbar_arrive_cluster_relaxed(peer_wempty + 8 * s);
                    mbar_arrive(b.aempty + 8 * as); mbar_arrive_cluster_relaxed(peer_aempty + 8 * as);
                }
            }
            __syncwarp();
            if (++s == (uint32_t)NS) { s = 0; ph ^= 1u; }
            if (++as == ASLOTS) { as = 0; aph ^= 1u; }
        }
        if (has_slot) {
            if (elect_one_sync()) {                              // this M-tile's accumulator is complete
                if (mine) umma_commit_pair(b.acc + 8 * slot);
                else { mbar_arrive(b.acc + 8 * slot); mbar_arrive_cluster_relaxed(peer_acc + 8 * slot); }
            }
            __syncwarp();
            slot += MTA;
            if (slot >= NSLOT) { slot -= NSLOT; sph ^= 1u; }
        }
    }
    if (prof && m == 0 && (threadIdx.x & 31) == 0) {
        unsigned long long *o = g_tc_stalls + blockIdx.x * TC_PROF_SLOTS;
        o[0] = clock64() - t_begin; o[1] = st_a; o[2] = st_b; o[3] = st_c; o[8] = idx;
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_PAIR_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap2, const __grid_constant__ CUtensorMap tmap_w2, const TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    // same barrier slots as conv_tc_kernel; a weight stage is a filter row here
    PairBars b;
    b.afull = smem_u32(bars); b.aempty = smem_u32(bars + 16); b.wfull = smem_u32(bars + 32);
    b.wempty = smem_u32(bars + 56); b.acc = smem_u32(bars + 80); b.accempty = smem_u32(bars + 88);     // 8 + 8 accumulator slots
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + 1008);
    uint8_t *act = smem + TC_SMEM_HEADER;
    uint8_t *ring = act + (size_t)p.aslots * p.group_bytes;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const bool prof = (p.dbg & 64) != 0;
    const long long t_begin = prof ? clock64() : 0;

    if (threadIdx.x == 0) {
        for (int g = 0; g < p.aslots; g++) { mbar_init(b.afull + 8 * g, 1); mbar_init(b.aempty + 8 * g, TC_MMA_WARPS); }
        for (int s = 0; s < p.nstages; s++) { mbar_init(b.wfull + 8 * s, 1); mbar_init(b.wempty + 8 * s, TC_MMA_WARPS); }
        for (int m = 0; m < p.nslot; m++) { mbar_init(b.acc + 8 * m, 1); mbar_init(b.accempty + 8 * m, 2 * (TC_PAIR_EPI_GROUPS * TC_EPI_WARPS / p.epi_groups)); }
        for (int m = 0; m < 8; m++) reinterpret_cast<volatile uint32_t *>(smem + TC_PAIR_ISSUED_OFF)[m] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + 1008)),
                     "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, cluster rendezvous) may overlap the
    // tail of the previous kernel in the stream; nothing below touches global memory before that kernel has completed.
    // The next kernel's CTAs may in turn be scheduled onto SMs as this grid's CTAs retire.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // (setmaxnreg sits inside each role's branch: ptxas budgets a region by the setmaxnreg that dominates it)
    if (warp < TC_PAIR_ISSUER_WARP0) {
      reg_dealloc<32>();
      if (warp == 0) {
        if (lane == 0) {
            // ===== weight producer: this CTA's half of every (group, filter row) stage; bytes are counted on the leader =====
            const int rows_per_item = p.kh * (p.kw / p.kws) * p.groups;          // ring stages per tile
            const uint32_t row_bytes = (uint32_t)p.kws * p.pair_slab;
            const uint32_t lead_wfull = mapa_cluster(b.wfull, 0);
            const int slab0 = (int)rank * rows_per_item * p.kws;                   // this CTA's half of the image
            long long st = 0;
            uint32_t s = 0, ph = 0;
            for (int item = cid; item < p.pair_items; item += ncl) {
                for (int it = 0; it < rows_per_item; it++) {
                    { TC_PROF_BEGIN(prof); mbar_wait_relaxed(b.wempty + 8 * s, ph ^ 1u); TC_PROF_END(prof, st); }
                    if (rank == 0) mbar_expect_tx(b.wfull + 8 * s, 2u * row_bytes);
                    tma_load_2d_pair(smem_u32(ring) + s * row_bytes, &tmap_w, lead_wfull + 8 * s, 0, (slab0 + it * p.kws) * p.pair_split);
                    if (++s == (uint32_t)p.nstages) { s = 0; ph ^= 1u; }
                }
                for (int g2 = 0; g2 < p.groups2; g2++) {         // fused shortcut: one slab per group of the second input
                    { TC_PROF_BEGIN(prof); mbar_wait_relaxed(b.wempty + 8 * s, ph ^ 1u); TC_PROF_END(prof, st); }
                    if (rank == 0) mbar_expect_tx(b.wfull + 8 * s, 2u * p.pair_slab);
                    tma_load_2d_pair(smem_u32(ring) + s * row_bytes, &tmap_w2, lead_wfull + 8 * s, 0,
                                     ((int)rank * p.groups2 + g2) * p.pair_split);
                    if (++s == (uint32_t)p.nstages) { s = 0; ph ^= 1u; }
                }
            }
            if (prof) g_tc_stalls[blockIdx.x * TC_PROF_SLOTS + 4] = st;
        }
      } else if (warp == 1) {
        if (lane == 0) {
            // ===== activation producer (own image), bytes counted on the leader's act_full =====
            const uint32_t lead_afull = mapa_cluster(b.afull, 0);
            long long st = 0;
            uint32_t as = 0, aph = 0;
            const int V = p.groups + p.groups2;             // channel groups of the input, then of the fused shortcut's input
            for (int item = cid; item < p.pair_items; item += ncl) {
                const PairGeom t = pair_geom(p, item, (int)rank);
                for (int v = 0; v < V; v++) {
                    { TC_PROF_BEGIN(prof); mbar_wait_relaxed(b.aempty + 8 * as, aph ^ 1u); TC_PROF_END(prof, st); }
                    if (rank == 0) mbar_expect_tx(b.afull + 8 * as, 2u * p.group_bytes);
                    if (p.stem) {           // two K chunks = two boxes of 8-pixel windows: position x = 8u + s
                        for (int kc = 0; kc < 2; kc++)
                            tma_load_5d_pair(smem_u32(act + (size_t)as * p.group_bytes + (size_t)kc * p.plane_bytes), &tmap,
                                             lead_afull + 8 * as, 0, 0, p.stem_chunk_u0[2 * v + kc], t.row0 - p.pady,
                                             t.n * p.stem_planes + p.stem_chunk_plane[2 * v + kc]);
                        if (++as == (uint32_t)p.aslots) { as = 0; aph ^= 1u; }
                        continue;
                    }
                    const bool second = v >= p.groups;
                    tma_load_4d_pair(smem_u32(act + (size_t)as * p.group_bytes), second ? &tmap2 : &tmap, lead_afull + 8 * as,
                                     -2 * p.padx, t.row0 - p.pady, (second ? v - p.groups : v) * 4, t.n);
                    if (++as == (uint32_t)p.aslots) { as = 0; aph ^= 1u; }
                }
            }
            if (prof) g_tc_stalls[blockIdx.x * TC_PROF_SLOTS + 5] = st;
        }
      }     // warps 2-3 of warpgroup 0 are idle: they only donate their registers
    } else if (warp < TC_PAIR_EPI_WARP0) {
        reg_dealloc<64>();
        if (rank == 0) {
            // ===== MMA issuers (leader only) =====
            const int m = warp - TC_PAIR_ISSUER_WARP0;
            const uint32_t aa = smem_u32(act), ra = smem_u32(ring);
            volatile uint32_t *issued = reinterpret_cast<volatile uint32_t *>(smem + TC_PAIR_ISSUED_OFF);
            if (p.stacked) {
                if (p.kws == 3) pair_issuer<3, true>(p, b, tmem_base, aa, ra, m, cid, ncl, prof, t_begin, issued);
                else if (p.kws == 5) pair_issuer<5, true>(p, b, tmem_base, aa, ra, m, cid, ncl, prof, t_begin, issued);
                else pair_issuer<1, true>(p, b, tmem_base, aa, ra, m, cid, ncl, prof, t_begin, issued);
            } else {
                if (p.kws == 3) pair_issuer<3, false>(p, b, tmem_base, aa, ra, m, cid, ncl, prof, t_begin, issued);
                else if (p.kws == 5) pair_issuer<5, false>(p, b, tmem_base, aa, ra, m, cid, ncl, prof, t_begin, issued);
                else pair_issuer<1, false>(p, b, tmem_base, aa, ra, m, cid, ncl, prof, t_begin, issued);
            }
        }
    } else {
        reg_alloc<96>();
        // ===== epilogue (both CTAs, own accumulators; the drain is reported to the leader) =====
        // Accumulator use c (running count over items and M-tile slots) goes to group c % G.  The slot of use c is
        // c % nslot (4 or 8, a multiple of G), so a slot always belongs to the same group and the group sees every phase of
        // its barriers.  G = 2 groups of 8 warps (two warps per TMEM lane quarter share the channel chunks of a position), or,
        // for Cout <= 32, G = 4 groups of 4 warps (one warp per lane quarter takes all chunks): those layers are paced by the
        // epilogue's per-M-tile latency chain, four M-tiles in flight hide more of it than two.
        const int ew = warp - TC_PAIR_EPI_WARP0;
        const int quarter = warp & 3;
        const uint32_t G = (uint32_t)p.epi_groups, gmask = G - 1u;
        const uint32_t grp = G == 4 ? (uint32_t)ew >> 2 : (uint32_t)ew >> 3;
        const int half = G == 4 ? 0 : (ew >> 2) & 1;
        const int nchunk = p.coutp >> 3, chh = G == 4 ? nchunk : nchunk >> 1, ch0 = half * chh;
        const uint32_t lead_accempty = mapa_cluster(b.accempty, 0);
        const uint32_t slot_mask = (uint32_t)p.nslot - 1u, slot_shift = p.nslot == 8 ? 3u : 2u;
        long long st = 0;
        uint32_t c = 0;
        for (int item = cid; item < p.pair_items; item += ncl) {
            const PairGeom t = pair_geom(p, item, (int)rank);
            if ((p.res.p || p.mul.p) && (lane & 7) == 0) {
                const size_t plane = (size_t)p.H * p.W;
                for (int mt = 0; mt < t.mt_count; mt++) {
                    if (((c + (uint32_t)mt) & gmask) != grp) continue;
                    const int pos = t.q0 + mt * 128 + quarter * 32 + lane;
                    const int r = pos / p.P, cc = pos - r * p.P;
                    if (cc < p.W && r < p.H) {
                        for (int j = 0; j < chh; j++) {
                            const size_t o = ((size_t)t.n * (p.out.Cp >> 2) + split_plane(ch0 + j, 0)) * plane + (size_t)r * p.W + cc;
                            if (p.res.p) { prefetch_l2(reinterpret_cast<const uint4 *>(p.res.p) + o); prefetch_l2(reinterpret_cast<const uint4 *>(p.res.p) + o + 2 * plane); }
                            if (p.mul.p) { prefetch_l2(reinterpret_cast<const uint4 *>(p.mul.p) + o); prefetch_l2(reinterpret_cast<const uint4 *>(p.mul.p) + o + 2 * plane); }
                        }
                    }
                }
            }
            for (int mt = 0; mt < p.mt_alloc; mt++, c++) {
                if ((c & gmask) != grp) continue;
                const uint32_t sl = c & slot_mask, par = (c >> slot_shift) & 1u;
                const bool work = mt < t.mt_count;
                const int pos = t.q0 + mt * 128 + quarter * 32 + lane;
                const int r = pos / p.P, cc = pos - r * p.P;
                const bool valid = work && t.store && (cc < p.W) && (r < p.H);
                // residual / attention operands of this thread's chunks: requested before the accumulator wait
                uint4 rh[4], rl[4];
                if (work) {
                    if (chh == 4) epilogue_fetch<4>(p, ch0, t.n, r, cc, valid, rh, rl);
                    else if (chh == 2) epilogue_fetch<2>(p, ch0, t.n, r, cc, valid, rh, rl);
                    else epilogue_fetch<1>(p, ch0, t.n, r, cc, valid, rh, rl);
                }
#ifdef PMP_EPI_SPIN        // build-time A/B knob: epilogue polls without the back-off
                { TC_PROF_BEGIN(prof); mbar_wait(b.acc + 8 * sl, par); TC_PROF_END(prof, st); }
#else
                { TC_PROF_BEGIN(prof); mbar_wait_relaxed(b.acc + 8 * sl, par); TC_PROF_END(prof, st); }
#endif
                if (work) {
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + sl * (uint32_t)p.acc_cols;
                    if (chh == 4) {
                        epilogue_process<2>(p, taddr, ch0, t.n, r, cc, valid, rh, rl);
                        epilogue_process<2>(p, taddr, ch0 + 2, t.n, r, cc, valid, rh + 2, rl + 2);
                    } else if (chh == 2) {
                        epilogue_process<2>(p, taddr, ch0, t.n, r, cc, valid, rh, rl);
                    } else {
                        epilogue_process<1>(p, taddr, ch0, t.n, r, cc, valid, rh, rl);
                    }
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) {        // TMEM reads are complete (wait::ld); nothing else needs ordering with this arrival
                    if (rank == 0) mbar_arrive(b.accempty + 8 * sl);
                    else mbar_arrive_cluster_relaxed(lead_accempty + 8 * sl);
                }
            }
        }
        if (prof && lane == 0 && warp == TC_PAIR_EPI_WARP0) {
            unsigned long long *o = g_tc_stalls + blockIdx.x * TC_PROF_SLOTS;
            o[6] = st; o[9] = clock64() - t_begin;
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encoder(Handle *h, EncodeTiledFn *fn)
{
    if (!h->encode_tiled) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        PMP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !f) {
            set_error("cuTensorMapEncodeTiled is not available from the driver");
            return PMP_ERR_UNSUPPORTED;
        }
        h->encode_tiled = f;
    }
    *fn = reinterpret_cast<EncodeTiledFn>(h->encode_tiled);
    return PMP_OK;
}

int conv_tc(Handle *h, const TcConvArgs &a, int B, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    TcGeom g;
    const int H = a.Ho ? a.Ho : a.in.H, W = a.in.W, Hin = a.in.H;
    // (a batch of one runs the pair kernel too: the peer CTA recomputes the image and drops it -- one arithmetic for
    // every batch composition; the single-CTA kernel remains as the PMP_TC_PAIR=0 / self-test A-B path)
    const bool want_pair = tc_pair_default() && a.w_pair;
    const bool stem = a.stem_src != nullptr;
    if (stem && (!want_pair || a.kw != 1 || (W & 7) || a.stem_nchunk < 1 || a.stem_nchunk > 8 || a.cin_pad != 16 * ((a.stem_nchunk + 1) / 2) ||
                 a.sc_in.p || a.res.p || a.mul.p || a.hpool)) {
        set_error("conv_tc: stem mode needs the CTA-pair kernel, a kh x 1 kernel, W %% 8 == 0 and 1..8 K chunks");
        return PMP_ERR_UNSUPPORTED;
    }
    bool geom_ok = want_pair && tc_geometry(a.cin_pad, a.cout_pad, a.kh, a.kw, H, W, g, -1, true, stem ? 2 : 4) &&
                   (g.stacked != 0) == tc_pair_stacked_layout(a.cout_pad, a.kh, a.kw);
    const bool use_pair = geom_ok;
    if (!geom_ok && !stem) geom_ok = tc_geometry(a.cin_pad, a.cout_pad, a.kh, a.kw, H, W, g);
    if (a.in.fmt != FMT_SPLIT || a.out.fmt != FMT_SPLIT || !geom_ok ||
        a.in.Cp != a.cin_pad || a.out.Cp != a.cout_pad || a.pool != 1 || a.out.H != H || a.out.W != (a.hpool ? W / 2 : W) ||
        (a.hpool && (!use_pair || a.bias || a.mul.p))) {
        set_error("conv_tc: unsupported configuration cin %d cout %d k %dx%d %dx%d", a.cin_pad, a.cout_pad, a.kh, a.kw, H, W);
        return PMP_ERR_UNSUPPORTED;
    }
    EncodeTiledFn enc = nullptr;
    int rc = get_encoder(h, &enc);
    if (rc) return rc;
    CUtensorMap tmap;
    const cuuint64_t planes = (cuuint64_t)a.cin_pad / 4;
    // 8-byte elements: a pixel's 8-channel unit is two elements, so the inner box dimension is a whole P-pixel row
    // (P*16 contiguous bytes) instead of one 16-byte unit (a 16-byte inner dimension makes every request pull a 32-byte
    // sector: measured 2x L2->SM traffic)
    cuuint64_t gdim[4] = {(cuuint64_t)W * 2, (cuuint64_t)Hin, planes, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)Hin * W * 16, planes * Hin * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)g.P * 2, (cuuint32_t)g.Rbox, 4, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr;
    if (stem) {
        // 8 pre-shifted copies per source plane, [B * planes][rows][uw][shift s][8 pixels] 16-bit: the unit (y, u, s) holds
        // pixels 8u+s .. 8u+s+7 of row y.  A box (16 B, 8 shifts, W/8 units, Rbox rows) lands them as positions x = 8u + s,
        // each with its 8-pixel window = 8 consecutive kx shifts: the unrolled K chunk, assembled by the TMA engine.
        cuuint64_t sdim[5] = {2, 8, (cuuint64_t)a.stem_uw, (cuuint64_t)a.stem_rows, (cuuint64_t)B * a.stem_planes};
        cuuint64_t sstr[4] = {16, 128, (cuuint64_t)a.stem_uw * 128, (cuuint64_t)a.stem_rows * a.stem_uw * 128};
        cuuint32_t sbox[5] = {2, 8, (cuuint32_t)W / 8, (cuuint32_t)g.Rbox, 1};
        cuuint32_t sest[5] = {1, 1, 1, 1, 1};
        cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void *>(a.stem_src), sdim, sstr, sbox, sest,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, a.in.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for W %d H %d planes %d B %d box %d x %d", (int)cr, W, H, (int)planes, B,
                  g.P, g.Rbox);
        return PMP_ERR_CUDA;
    }
    TcParams p;
    p.w = a.w; p.bias = a.bias; p.out = a.out; p.res = a.res; p.mul = a.mul;
    p.H = H; p.W = W; p.P = g.P; p.kh = a.kh; p.kw = a.kw; p.pady = a.pad_t; p.padx = a.pad_l; p.groups = g.groups;
    p.total_mt = g.total_mt; p.tiles = g.tiles; p.N1 = g.N1; p.coutp = g.coutp; p.nstages = g.nstages;
    p.items = g.tiles * B;
    p.plane_bytes = g.plane_bytes; p.group_bytes = g.group_bytes; p.stage_bytes = g.stage_bytes; p.tmem_cols = g.tmem_cols;
    const uint32_t fmt = a.in.bf16 ? 1u : 0u;
    const uint32_t idesc_base_nom = (1u << 4) | (fmt << 7) | (fmt << 10);
    const uint32_t idesc_base = idesc_base_nom | ((128u >> 4) << 24);
    p.idesc1 = idesc_base | ((uint32_t)(g.N1 >> 3) << 17);
    p.idesc2 = idesc_base | ((uint32_t)(g.coutp >> 3) << 17);
    p.relu = a.relu; p.stacked = g.stacked; p.pairbuf = g.pairbuf; p.nbuf = g.nbuf; p.hpool = a.hpool;
    p.stem = stem ? 1 : 0; p.stem_planes = a.stem_planes;
    for (int c = 0; c < 8; c++) {           // an odd chunk count repeats chunk 0 (its weights are zero; the operand must be finite)
        const int src = c < a.stem_nchunk ? c : 0;
        p.stem_chunk_plane[c] = stem ? a.stem_chunk_plane[src] : 0;
        p.stem_chunk_u0[c] = stem ? a.stem_chunk_u0[src] : 0;
    }
    static const int env_dbg = [] { const char *e = getenv("PMP_TC_DBG"); return e ? atoi(e) : 0; }();   // bit 6: per-role stall counters
    p.dbg = env_dbg;
    { int rcs = ensure_status(h); if (rcs) return rcs; }
    p.sat = h->d_status;
    if (!h->tc_attr_set) {          // function attributes are per device: one handle per device
        PMP_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX));
        PMP_CUDA(cudaFuncSetAttribute(conv_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX));
        h->tc_attr_set = true;
    }
    const double flops = a.flops_override > 0 ? a.flops_override * B
                                               : 2.0 * B * H * W * (double)a.out.C * ((double)a.in.C * a.kh * a.kw + (a.sc_in.p ? a.sc_in.C : 0));
    // algorithmic HBM bytes of the launch: every operand tensor once at 4 B/element (hi + lo), weights excluded (L2-resident)
    const double hbm_bytes = (stem ? (double)B * a.stem_planes * a.stem_rows * a.stem_uw * 128.0
                                   : (double)B * a.cin_pad * Hin * W * 4.0) +
                             (double)B * a.cout_pad * H * (a.hpool ? W / 2 : W) * 4.0 * (1.0 + (a.res.p ? 1.0 : 0.0) + (a.mul.p ? 1.0 : 0.0)) +
                             (a.sc_in.p ? (double)B * a.sc_cin_pad * Hin * W * 4.0 : 0.0);
    if (a.sc_in.p && (!use_pair || !a.w_pair_sc || a.sc_in.fmt != FMT_SPLIT || a.sc_in.H != Hin || a.sc_in.W != W ||
                      a.sc_in.Cp != a.sc_cin_pad || a.sc_cin_pad % 16 || a.res.p)) {
        set_error("conv_tc: fused shortcut needs the CTA-pair kernel, a split-format second input of the same size and no residual");
        return PMP_ERR_UNSUPPORTED;
    }
    if (use_pair) {
        // CTA-pair kernel: 2-D tensor map over this conv's per-CTA weight slabs (8-byte elements, one slab per map row);
        // one box = one filter row of one channel group = kw consecutive slabs = one ring stage
        CUtensorMap tmap_w;
        const cuuint64_t nslab = (cuuint64_t)2 * g.groups * a.kh * a.kw;
        // (slabs above 2 KB = 256 elements -- stacked Cout = 64: 3 KB -- take two map rows, one per k8 half)
        const uint32_t slab = 32u * (uint32_t)tc_pair_slab_rows(g.coutp, a.kh, a.kw);
        const uint32_t split = slab > 2048u ? 2u : 1u;
        // a ring stage is a whole filter row unless fewer than 4 such stages fit beside the activation tile (5x5 layers
        // with 3 KB slabs): then single taps
        tc_ring_layout(g, slab * (uint32_t)a.kw, a.kh);
        const int kws = (g.nstages >= 4 || a.kw == 1) ? a.kw : 1;
        cuuint64_t wdim[2] = {slab / split / 8, nslab * split};
        cuuint64_t wstr[1] = {slab / split};
        cuuint32_t wbox[2] = {slab / split / 8, (cuuint32_t)kws * split};
        cuuint32_t west[2] = {1, 1};
        cr = enc(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void *)a.w_pair, wdim, wstr, wbox, west,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled (pair weights) failed (%d)", (int)cr);
            return PMP_ERR_CUDA;
        }
        // fused 1x1 shortcut: second activation tensor (same box) and its own one-slab-per-group weight map
        CUtensorMap tmap2 = tmap, tmap_w2 = tmap_w;
        p.groups2 = 0;
        if (a.sc_in.p) {
            const cuuint64_t planes2 = (cuuint64_t)a.sc_cin_pad / 4;
            cuuint64_t gdim2[4] = {(cuuint64_t)W * 2, (cuuint64_t)Hin, planes2, (cuuint64_t)B};
            cuuint64_t gstr2[3] = {(cuuint64_t)W * 16, (cuuint64_t)Hin * W * 16, planes2 * Hin * W * 16};
            cr = enc(&tmap2, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, a.sc_in.p, gdim2, gstr2, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            p.groups2 = a.sc_cin_pad / 16;
            cuuint64_t wdim2[2] = {slab / split / 8, (cuuint64_t)2 * p.groups2 * split};
            cuuint32_t wbox2[2] = {slab / split / 8, split};
            if (cr == CUDA_SUCCESS)
                cr = enc(&tmap_w2, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void *)a.w_pair_sc, wdim2, wstr, wbox2, west,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled (fused shortcut) failed (%d)", (int)cr);
                return PMP_ERR_CUDA;
            }
        }
        p.B = B;
        p.pair_items = g.tiles * ((B + 1) / 2);
        p.pairbuf = 0; p.pair_slab = slab; p.pair_split = (int)split;
        // accumulator slot ring: 8 slots of (2*)Cout columns, tiles of up to 4 M-tiles (= two alternating buffers); stacked
        // Cout = 64: 4 slots of 128 columns, tiles of up to 3 M-tiles -- the 4th slot lets the next tile start while the
        // epilogue drains this one slot by slot
        p.acc_cols = (g.stacked ? 2 : 1) * g.coutp;
        p.nslot = 512 / p.acc_cols < 8 ? 512 / p.acc_cols : 8;
        p.mt_alloc = p.nslot < 8 ? 3 : 4;
        static const int env_epi4 = [] { const char *e = getenv("PMP_TC_EPI4"); return e ? atoi(e) : 1; }();     // A/B knob
        p.epi_groups = (env_epi4 && g.coutp <= 32 && p.nslot == 8) ? 4 : 2;
        p.tmem_cols = g.tmem_cols;
        tc_ring_layout(g, slab * (uint32_t)kws, a.kh * (a.kw / kws));
        p.nstages = g.nstages; p.nbuf = g.nbuf; p.kws = kws;
        p.aslots = g.nbuf * g.groups;           // activation ring: one slot per channel-group box, any number of them
        uint32_t smem_pair = g.smem_bytes;
        if (p.groups2) {
            // the input's and the shortcut input's groups cycle through the slots; take what fits beside a useful ring
            int want = g.groups * a.kh * (a.kw / kws) + p.groups2;
            if (want > TC_MAX_STAGES) want = TC_MAX_STAGES;
            if (want > g.nstages) want = g.nstages;
            int slots = (int)((TC_SMEM_MAX - TC_SMEM_HEADER - (uint32_t)want * slab * (uint32_t)kws) / g.group_bytes);
            if (slots > 16) slots = 16;
            if (slots > 4 * (g.groups + p.groups2)) slots = 4 * (g.groups + p.groups2);
            if (slots < g.groups) slots = g.groups;
            p.aslots = slots;
            int ns = (int)((TC_SMEM_MAX - TC_SMEM_HEADER - (uint32_t)slots * g.group_bytes) / (slab * (uint32_t)kws));
            p.nstages = ns < TC_MAX_STAGES ? ns : TC_MAX_STAGES;
            smem_pair = TC_SMEM_HEADER + (uint32_t)slots * g.group_bytes + (uint32_t)p.nstages * slab * (uint32_t)kws;
        }
        p.idesc1 = idesc_base_nom | ((uint32_t)(g.coutp >> 3) << 17) | ((256u >> 4) << 24);      // M = 256 across the pair
        p.idesc2 = idesc_base_nom | ((uint32_t)(g.N1 >> 3) << 17) | ((256u >> 4) << 24);         // stacked: N = 2*Cout
        int nsm = h->num_sms & ~1;
        int grid = 2 * p.pair_items < nsm ? 2 * p.pair_items : nsm;
        ProfScope ps(h, PROF_CONV_TC, s, flops, hbm_bytes);
        static const int env_pdl = [] { const char *e = getenv("PMP_TC_PDL"); return e ? atoi(e) : 1; }();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_PAIR_THREADS); cfg.dynamicSmemBytes = smem_pair; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = env_pdl ? 1 : 0;
        PMP_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel, tmap, tmap_w, tmap2, tmap_w2, p));
        h->launches++;
        return PMP_OK;
    }
    dim3 grid(p.items < h->num_sms ? p.items : h->num_sms);
    ProfScope ps(h, PROF_CONV_TC, s, flops, hbm_bytes);
    conv_tc_kernel<<<grid, TC_THREADS, g.smem_bytes, s>>>(tmap, p);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// ------------------------------------------------------------------------------------------------
// max-pool 2x2 (+ attention product) on split tensors; format conversion helpers
// ------------------------------------------------------------------------------------------------
// hdone: the producing conv's epilogue already took the horizontal maximum (`in` is H x W/2): vertical half only
__global__ void pool2_split_kernel(Act in, Act out, Act mul, int B, int hdone)
{
    const int Ho = out.H, Wo = out.W, nch = out.Cp >> 3;
    size_t total = (size_t)B * nch * Ho * Wo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), ch = (int)((i / ((size_t)Wo * Ho)) % nch);
        int n = (int)(i / ((size_t)Wo * Ho * nch));
        float m[8], v[8];
        if (hdone) {
            load_chunk_split(in, n, ch, 2 * y, x, m);
            load_chunk_split(in, n, ch, 2 * y + 1, x, v);
#pragma unroll
            for (int e = 0; e < 8; e++) m[e] = fmaxf(m[e], v[e]);
            if (mul.p) {
                load_chunk_split(mul, n, ch, y, x, v);
#pragma unroll
                for (int e = 0; e < 8; e++) m[e] *= v[e];
            }
            store_chunk_split(out, n, ch, y, x, m);
            continue;
        }
        load_chunk_split(in, n, ch, 2 * y, 2 * x, m);
        load_chunk_split(in, n, ch, 2 * y, 2 * x + 1, v);
#pragma unroll
        for (int e = 0; e < 8; e++) m[e] = fmaxf(m[e], v[e]);
        load_chunk_split(in, n, ch, 2 * y + 1, 2 * x, v);
#pragma unroll
        for (int e = 0; e < 8; e++) m[e] = fmaxf(m[e], v[e]);
        load_chunk_split(in, n, ch, 2 * y + 1, 2 * x + 1, v);
#pragma unroll
        for (int e = 0; e < 8; e++) m[e] = fmaxf(m[e], v[e]);
        if (mul.p) {
            load_chunk_split(mul, n, ch, y, x, v);
#pragma unroll
            for (int e = 0; e < 8; e++) m[e] *= v[e];
        }
        store_chunk_split(out, n, ch, y, x, m);
    }
}

int pool2_split(Handle *h, const Act &in, const Act &out, const Act &mul, int B, cudaStream_t s)
{
    size_t total = (size_t)B * (out.Cp >> 3) * out.H * out.W;
    if (!total) return PMP_OK;
    const int hdone = in.W == out.W ? 1 : 0;
    int grid = (int)((total + 255) / 256);
    ProfScope ps(h, PROF_ELEMWISE, s, 0, (double)total * 32 * (hdone ? 3 : 5));
    pool2_split_kernel<<<grid, 256, 0, s>>>(in, out, mul, B, hdone);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// First-layer input with the kx taps unrolled into channels (see kernels.cuh).  grid = (pixel blocks, chunk, image);
// the (channel, shift) pair of each of the chunk's 8 outputs is advanced incrementally (no per-element division).
__global__ void stem_unroll_kernel(Act x, const float *__restrict__ qt, int up, int ov, int kw, Act out)
{
    const int S0 = x.H, W = out.W, H = out.H, cx = x.C;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= H * W) return;
    const int yy = pix / W, xx = pix - yy * W, ch = blockIdx.y, n = blockIdx.z;
    int c = (ch * 8) / kw, j = ch * 8 - c * kw;
    const float *qrow = (qt && yy >= ov) ? qt + (size_t)n * 64 + ((yy - ov) / up) * 8 : nullptr;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int sx = xx + j;
        float val = 0.f;
        if (sx < S0) {
            if (c < cx) val = load_elem(x, n, c, yy, sx);
            else if (c == cx && qrow && sx >= ov) val = qrow[(sx - ov) / up];
        }
        v[e] = val;
        if (++j == kw) { j = 0; c++; }
    }
    store_chunk_split(out, n, ch, yy, xx, v);
}

int stem_unroll(Handle *h, const Act &x, const float *qt, int up, int ov, int kw, const Act &out, int B, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    dim3 grid((out.H * out.W + 255) / 256, out.Cp >> 3, B);
    ProfScope ps(h, PROF_ELEMWISE, s, 0, (double)B * (out.Cp >> 3) * out.H * out.W * 32 + (double)B * x.C * x.H * x.W);
    stem_unroll_kernel<<<grid, 256, 0, s>>>(x, qt, up, ov, kw, out);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// Pre-shifted copies for the stem mode of the pair kernel (kernels.cuh: stem_shift).  One thread per (image, plane, row,
// unit column u): it reads the 15 pixels 8u .. 8u+14 once and writes the 8 windows [8u+s, 8u+s+8) as 128 contiguous bytes.
__global__ void stem_shift_kernel(Act x, const float *__restrict__ qt, int up, int ov, int uw, int bf16, uint4 *__restrict__ dst, int B)
{
    const int S0 = x.H, cx = x.C, planes = cx + (qt ? 2 : 0);
    const size_t total = (size_t)B * planes * S0 * uw;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int u = (int)(r % uw); r /= uw;
        const int y = (int)(r % S0); r /= S0;
        const int pl = (int)(r % planes), n = (int)(r / planes);
        const float *qrow = (pl >= cx && y >= ov) ? qt + (size_t)n * 64 + ((y - ov) / up) * 8 : nullptr;
        uint32_t hv[15];
#pragma unroll
        for (int e = 0; e < 15; e++) {
            const int xx = 8 * u + e;
            float val = 0.f;
            if (xx < S0) {
                if (pl < cx) val = load_elem(x, n, pl, y, xx);
                else if (qrow && xx >= ov) val = qrow[(xx - ov) / up];
            }
            uint16_t hi, lo;
            split16(val, bf16 != 0, hi, lo);
            hv[e] = (pl == cx + 1) ? lo : hi;          // plane cx: hi half of the qt channel, cx + 1: its lo half
        }
#pragma unroll
        for (int sft = 0; sft < 8; sft++) {
            uint4 o;
            o.x = hv[sft] | (hv[sft + 1] << 16); o.y = hv[sft + 2] | (hv[sft + 3] << 16);
            o.z = hv[sft + 4] | (hv[sft + 5] << 16); o.w = hv[sft + 6] | (hv[sft + 7] << 16);
            dst[i * 8 + sft] = o;
        }
    }
}

int stem_shift(Handle *h, const Act &x, const float *qt, int up, int ov, int uw, bool bf16, void *dst, int B, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    const int planes = x.C + (qt ? 2 : 0);
    const size_t total = (size_t)B * planes * x.H * uw;
    const int grid = (int)((total + 255) / 256 < 148 * 64 ? (total + 255) / 256 : 148 * 64);
    ProfScope ps(h, PROF_ELEMWISE, s, 0, (double)total * 128 + (double)B * x.C * x.H * x.W);
    stem_shift_kernel<<<grid, 256, 0, s>>>(x, qt, up, ov, uw, bf16 ? 1 : 0, reinterpret_cast<uint4 *>(dst), B);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

__global__ void f32_to_split_kernel(const float *__restrict__ src, Act dst, int B)
{
    const int nch = dst.Cp >> 3;
    size_t total = (size_t)B * nch * dst.H * dst.W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % dst.W), y = (int)((i / dst.W) % dst.H), ch = (int)((i / ((size_t)dst.W * dst.H)) % nch);
        int n = (int)(i / ((size_t)dst.W * dst.H * nch));
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            int c = ch * 8 + e;
            v[e] = c < dst.C ? src[(((size_t)n * dst.C + c) * dst.H + y) * dst.W + x] : 0.f;
        }
        store_chunk_split(dst, n, ch, y, x, v);
    }
}

__global__ void split_to_f32_kernel(Act src, float *__restrict__ dst, int B)
{
    size_t total = (size_t)B * src.C * src.H * src.W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % src.W), y = (int)((i / src.W) % src.H), c = (int)((i / ((size_t)src.W * src.H)) % src.C);
        int n = (int)(i / ((size_t)src.W * src.H * src.C));
        dst[i] = load_elem(src, n, c, y, x);
    }
}

int split_to_f32(Handle *h, const Act &src, float *dst, int B, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    split_to_f32_kernel<<<1024, 256, 0, s>>>(src, dst, B);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

}  // namespace pmp

// ------------------------------------------------------------------------------------------------
// self test: TC conv against the exact fp32 SIMT conv on the same split-precision inputs
// ------------------------------------------------------------------------------------------------
using namespace pmp;

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return cudaMalloc(&p, bytes) == cudaSuccess ? 0 : -1; }
};
float lcg(uint32_t &s)
{
    s = s * 1664525u + 1013904223u;
    return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
}  // namespace

struct pmp_handle : public pmp::Handle {};

// Debug: copy the pair kernel's per-CTA stall counters (PMP_TC_DBG bit 6) to the host; n = number of uint64 wanted.
extern "C" int pmp_debug_tc_stalls(unsigned long long *out, int n)
{
    if (!out || n <= 0 || n > 160 * pmp::TC_PROF_SLOTS) return PMP_ERR_ARG;
    PMP_CUDA(cudaMemcpyFromSymbol(out, pmp::g_tc_stalls, (size_t)n * sizeof(unsigned long long)));
    return PMP_OK;
}

// flags: bit0 relu, bit1 identity residual, bit2 attention product, bit3 bf16 operands, bits 8..15 kernel variant,
// bit 16: verbose mismatch report on stderr
extern "C" int pmp_selftest_conv(pmp_handle *h, int cin, int cout, int ksize, int hw, int batch, int flags, double *max_err,
                                 double *ref_absmax, double *ms_tc, double *ms_simt)
{
    if (!h) return PMP_ERR_ARG;
    PMP_CUDA(cudaSetDevice(h->device));
    const int B = batch, H = hw, W = hw;
    const bool bf = (flags & 8) != 0;
    struct KnobReset {      // the variant knobs are process-wide: restore the defaults on every exit path
        ~KnobReset() { pmp::g_tc_scheme = -1; pmp::g_tc_pair = -1; pmp::g_tc_nbuf_max = -1; pmp::g_tc_st64 = -1; }
    } knob_reset;
    pmp::g_tc_st64 = (flags >> 12) & 1 ? 0 : -1;         // bit 12: unstacked accumulators for the Cout = 64 layers (A/B)
    const int cinp = pad16(cin), coutp = pad16(cout);
    if (!tc_supported(cinp, coutp, ksize, ksize, H, W)) {
        set_error("selftest: configuration not supported by the TC engine");
        return PMP_ERR_UNSUPPORTED;
    }
    uint32_t seed = 12345u + cin * 7 + cout * 13 + ksize * 31 + hw;
    const size_t n_in = (size_t)B * cin * H * W, n_out = (size_t)B * cout * H * W, n_w = (size_t)cout * cin * ksize * ksize;
    // host data for at most 8 distinct images; larger batches (timing runs) repeat them on the device
    const int Bu = B < 8 ? B : 8;
    const size_t u_in = (size_t)Bu * cin * H * W, u_out = (size_t)Bu * cout * H * W;
    std::vector<float> hin(u_in), hw_(n_w), hres(u_out), hmul(u_out);
    for (auto &v : hin) v = 40.f * fabsf(lcg(seed)) * (lcg(seed) > -0.3f ? 1.f : 0.f);     // relu-like activations
    const float wb = sqrtf(3.0f / (cin * ksize * ksize));
    for (auto &v : hw_) v = wb * lcg(seed);
    for (auto &v : hres) v = 10.f * lcg(seed);
    for (auto &v : hmul) v = 1.5f * lcg(seed);

    DevBuf d_in32, d_res32, d_mul32, d_in, d_res, d_mul, d_out, d_out32, d_ref32, d_wsimt, d_wtc, d_wpair;
    // bit 17: fused 1x1 shortcut on a second input with ((flags >> 20) & 0xff, default 32) channels (excludes bit 1)
    const bool fused = (flags >> 17) & 1;
    const int cin2 = fused ? (((flags >> 20) & 0xff) ? ((flags >> 20) & 0xff) : 32) : 0, cin2p = pad16(cin2);
    DevBuf d_in2_32, d_in2, d_sc32, d_wsimt2, d_wsc;
    if (fused && (flags & 2)) { set_error("selftest: fused shortcut excludes the identity residual"); return PMP_ERR_ARG; }
    const size_t sp_in = act_bytes(FMT_SPLIT, B, cin, H, W), sp_out = act_bytes(FMT_SPLIT, B, cout, H, W);
    if (d_in32.alloc(n_in * 4) || d_res32.alloc(n_out * 4) || d_mul32.alloc(n_out * 4) || d_in.alloc(sp_in) ||
        d_res.alloc(sp_out) || d_mul.alloc(sp_out) || d_out.alloc(sp_out) || d_out32.alloc(n_out * 4) ||
        d_ref32.alloc(n_out * 4)) {
        set_error("selftest: cudaMalloc failed");
        return PMP_ERR_CUDA;
    }
    auto upload_tiled = [&](void *dst, const std::vector<float> &src, size_t total) -> cudaError_t {
        cudaError_t e = cudaMemcpy(dst, src.data(), src.size() * 4, cudaMemcpyHostToDevice);
        for (size_t filled = src.size(); e == cudaSuccess && filled < total;) {
            const size_t n = filled < total - filled ? filled : total - filled;
            e = cudaMemcpy((float *)dst + filled, dst, n * 4, cudaMemcpyDeviceToDevice);
            filled += n;
        }
        return e;
    };
    PMP_CUDA(upload_tiled(d_in32.p, hin, n_in));
    PMP_CUDA(upload_tiled(d_res32.p, hres, n_out));
    PMP_CUDA(upload_tiled(d_mul32.p, hmul, n_out));
    // weights: SIMT layout + TC packing
    const int coutw = (cout + 3) & ~3, taps = ksize * ksize;
    std::vector<float> ps((size_t)cin * taps * coutw, 0.f);
    for (int o = 0; o < cout; o++)
        for (int c = 0; c < cin; c++)
            for (int t = 0; t < taps; t++) ps[((size_t)c * taps + t) * coutw + o] = hw_[((size_t)o * cin + c) * taps + t];
    std::vector<uint16_t> pk(tc_packed_elems(cinp, coutp, ksize, ksize));
    pack_tc_weights(hw_.data(), cout, cin, ksize, ksize, cinp, coutp, bf, pk.data());
    if (d_wsimt.alloc(ps.size() * 4) || d_wtc.alloc(pk.size() * 2)) return PMP_ERR_CUDA;
    PMP_CUDA(cudaMemcpy(d_wsimt.p, ps.data(), ps.size() * 4, cudaMemcpyHostToDevice));
    PMP_CUDA(cudaMemcpy(d_wtc.p, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    {
        std::vector<uint16_t> pp(tc_pair_packed_elems(cinp, coutp, ksize, ksize));
        pack_tc_pair_weights(hw_.data(), cout, cin, ksize, ksize, cinp, coutp, bf, pp.data());
        if (d_wpair.alloc(pp.size() * 2)) return PMP_ERR_CUDA;
        PMP_CUDA(cudaMemcpy(d_wpair.p, pp.data(), pp.size() * 2, cudaMemcpyHostToDevice));
    }

    if (fused) {
        const size_t u2 = (size_t)Bu * cin2 * H * W, n2 = (size_t)B * cin2 * H * W;
        std::vector<float> hin2(u2), hw2((size_t)cout * cin2);
        for (auto &v : hin2) v = 30.f * fabsf(lcg(seed)) * (lcg(seed) > -0.2f ? 1.f : 0.f);
        const float wb2 = sqrtf(3.0f / cin2);
        for (auto &v : hw2) v = wb2 * lcg(seed);
        std::vector<float> ps2((size_t)cin2 * coutw, 0.f);
        for (int o = 0; o < cout; o++)
            for (int c = 0; c < cin2; c++) ps2[(size_t)c * coutw + o] = hw2[(size_t)o * cin2 + c];
        std::vector<uint16_t> pf(tc_pair_fused_sc_elems(cin2p, coutp, ksize, ksize));
        pack_tc_pair_fused_sc(hw2.data(), cout, cin2, cin2p, coutp, ksize, ksize, bf, pf.data());
        if (d_in2_32.alloc(n2 * 4) || d_in2.alloc(act_bytes(FMT_SPLIT, B, cin2, H, W)) || d_sc32.alloc(n_out * 4) ||
            d_wsimt2.alloc(ps2.size() * 4) || d_wsc.alloc(pf.size() * 2))
            return PMP_ERR_CUDA;
        PMP_CUDA(upload_tiled(d_in2_32.p, hin2, n2));
        PMP_CUDA(cudaMemcpy(d_wsimt2.p, ps2.data(), ps2.size() * 4, cudaMemcpyHostToDevice));
        PMP_CUDA(cudaMemcpy(d_wsc.p, pf.data(), pf.size() * 2, cudaMemcpyHostToDevice));
    }

    auto mk = [&](void *p, int C) {
        Act a;
        a.p = p; a.fmt = FMT_SPLIT; a.C = C; a.Cp = pad16(C); a.H = H; a.W = W; a.bf16 = bf;
        return a;
    };
    Act in = mk(d_in.p, cin), res = mk(d_res.p, cout), mul = mk(d_mul.p, cout), out = mk(d_out.p, cout);
    cudaStream_t s = nullptr;
    f32_to_split_kernel<<<1024, 256, 0, s>>>((const float *)d_in32.p, in, B);
    f32_to_split_kernel<<<1024, 256, 0, s>>>((const float *)d_res32.p, res, B);
    f32_to_split_kernel<<<1024, 256, 0, s>>>((const float *)d_mul32.p, mul, B);
    PMP_CUDA(cudaMemsetAsync(d_out.p, 0xFF, sp_out, s));
    Act in2 = mk(d_in2.p, cin2);
    if (fused) f32_to_split_kernel<<<1024, 256, 0, s>>>((const float *)d_in2_32.p, in2, B);
    PMP_CUDA(cudaGetLastError());

    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    // reference: SIMT conv on the same split inputs -> fp32 NCHW
    SimtConvArgs sa;
    sa.in = in;
    sa.out.p = d_ref32.p; sa.out.fmt = FMT_F32; sa.out.C = sa.out.Cp = cout; sa.out.H = H; sa.out.W = W;
    if (flags & 2) sa.res = res;
    if (flags & 4) sa.mul = mul;
    sa.w = (const float *)d_wsimt.p; sa.cin = cin; sa.cout = cout; sa.coutw = coutw;
    sa.pad_t = sa.pad_l = ksize / 2; sa.Ho = H; sa.Wo = W; sa.relu = flags & 1; sa.pool = 1;
    // bit 18: 2x2 max-pool -- horizontal half in the conv epilogue, vertical half by pool2_split -- against the exact conv
    // with its fused 2x2 pooling (excludes the attention product)
    const bool pooled = (flags >> 18) & 1;
    if (pooled) {
        if ((flags & 4) || (W & 1) || (H & 1)) { set_error("selftest: pooled variant needs even sizes and no attention product"); return PMP_ERR_ARG; }
        sa.pool = 2; sa.out.H = H / 2; sa.out.W = W / 2;
    }
    int rc = PMP_OK;
    if (fused) {            // reference: exact fp32 1x1 conv of the second input, added as the residual of the main conv
        SimtConvArgs sc;
        sc.in = in2;
        sc.out.p = d_sc32.p; sc.out.fmt = FMT_F32; sc.out.C = sc.out.Cp = cout; sc.out.H = H; sc.out.W = W;
        sc.w = (const float *)d_wsimt2.p; sc.cin = cin2; sc.cout = cout; sc.coutw = coutw;
        sc.pad_t = sc.pad_l = 0; sc.Ho = H; sc.Wo = W; sc.relu = 0; sc.pool = 1;
        rc = conv_simt(h, sc, 1, 1, B, s);
        if (rc) return rc;
        sa.res = sc.out;
    }
    cudaEventRecord(e0, s);
    rc = conv_simt(h, sa, ksize, ksize, B, s);
    cudaEventRecord(e1, s);
    if (rc) return rc;
    TcConvArgs ta;
    ta.in = in; ta.out = out;
    if (flags & 2) ta.res = res;
    if (flags & 4) ta.mul = mul;
    if (pooled) { ta.hpool = 1; ta.out.W = W / 2; out.W = W / 2; }
    ta.w = (const uint16_t *)d_wtc.p; ta.w_pair = (const uint16_t *)d_wpair.p; ta.cin_pad = cinp; ta.cout_pad = coutp; ta.kh = ta.kw = ksize; ta.pad_t = ta.pad_l = ksize / 2; ta.relu = flags & 1; ta.pool = 1;
    if (fused) { ta.sc_in = in2; ta.w_pair_sc = (const uint16_t *)d_wsc.p; ta.sc_cin_pad = cin2p; }
    pmp::g_tc_scheme = ((flags >> 8) & 3) - 1;          // 0: library default, 1: unstacked, 2: stacked
    pmp::g_tc_pair = (flags >> 10) & 1 ? 1 : ((flags >> 11) & 1 ? 0 : -1);   // bit 10: CTA-pair kernel, bit 11: force single
    pmp::g_tc_nbuf_max = ((flags >> 13) & 7) ? ((flags >> 13) & 7) : -1;      // bits 13..15: cap on activation buffers (0: default)
    if (!tc_supported(cinp, coutp, ksize, ksize, H, W)) { set_error("selftest: scheme not supported"); return PMP_ERR_UNSUPPORTED; }
    rc = conv_tc(h, ta, B, s);          // warm-up (also first-launch overheads)
    cudaEventRecord(e2, s);
    if (!rc) rc = conv_tc(h, ta, B, s);
    cudaEventRecord(e3, s);
    if (rc) return rc;
    if (pooled) {           // `out` holds H x W/2 (d_out is large enough); finish the pool into the d_res buffer (free by now)
        Act pooled_out = mk(d_res.p, cout);
        pooled_out.H = H / 2; pooled_out.W = W / 2;
        rc = pool2_split(h, out, pooled_out, Act(), B, s);
        if (rc) return rc;
        out = pooled_out;
    }
    split_to_f32_kernel<<<1024, 256, 0, s>>>(out, (float *)d_out32.p, B);
    cudaError_t ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) return cuda_fail(ce, "selftest sync", __FILE__, __LINE__);
    float t_simt = 0, t_tc = 0;
    cudaEventElapsedTime(&t_simt, e0, e1);
    cudaEventElapsedTime(&t_tc, e2, e3);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    const size_t n_full = n_out;
    const size_t n_cmp = pooled ? n_full / 4 : n_full;
    std::vector<float> got(n_cmp), want(n_cmp);
    PMP_CUDA(cudaMemcpy(got.data(), d_out32.p, n_cmp * 4, cudaMemcpyDeviceToHost));
    PMP_CUDA(cudaMemcpy(want.data(), d_ref32.p, n_cmp * 4, cudaMemcpyDeviceToHost));
    double me = 0, am = 0;
    size_t nbad = 0, first_bad = (size_t)-1;
    for (size_t i = 0; i < n_cmp; i++) {
        double e = fabs((double)got[i] - (double)want[i]);
        if (!(e == e)) e = 1e30;
        if (e > me) me = e;
        if (fabs(want[i]) > am) am = fabs(want[i]);
        if (e > 1e-3 * (1.0 + fabs(want[i]))) { nbad++; if (first_bad == (size_t)-1) first_bad = i; }
    }
    if ((flags & (1 << 16)) && nbad) {
        fprintf(stderr, "[selftest] cin %d cout %d k %d hw %d B %d flags %x: %zu/%zu bad, max err %g (ref absmax %g)\n", cin,
                cout, ksize, hw, B, flags, nbad, n_cmp, me, am);
        // error map by (y, x) for image 0 / channel 0 and by channel at pixel (H/2, W/2)
        int shown = 0;
        for (size_t i = first_bad; i < n_cmp && shown < 12; i++) {
            double e = fabs((double)got[i] - (double)want[i]);
            if (e > 1e-3 * (1.0 + fabs(want[i])) || !(e == e)) {
                int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((size_t)W * H)) % cout), n = (int)(i / ((size_t)W * H * cout));
                fprintf(stderr, "   n %d c %d y %d x %d: got %g want %g\n", n, c, y, x, got[i], want[i]);
                shown++;
            }
        }
        std::vector<int> by_y(H, 0), by_x(W, 0), by_c(cout, 0);
        for (size_t i = 0; i < n_cmp; i++) {
            double e = fabs((double)got[i] - (double)want[i]);
            if (e > 1e-3 * (1.0 + fabs(want[i])) || !(e == e)) {
                by_x[i % W]++; by_y[(i / W) % H]++; by_c[(i / ((size_t)W * H)) % cout]++;
            }
        }
        fprintf(stderr, "   bad by y:");
        for (int y = 0; y < H; y++) fprintf(stderr, " %d", by_y[y]);
        fprintf(stderr, "\n   bad by x:");
        for (int x = 0; x < W; x++) fprintf(stderr, " %d", by_x[x]);
        fprintf(stderr, "\n   bad by c:");
        for (int c = 0; c < cout; c++) fprintf(stderr, " %d", by_c[c]);
        fprintf(stderr, "\n");
    }
    if (max_err) *max_err = me;
    if (ref_absmax) *ref_absmax = am;
    if (ms_tc) *ms_tc = t_tc;
    if (ms_simt) *ms_simt = t_simt;
    return PMP_OK;
}


// Test hook: one TC-engine convolution on caller-supplied fp32 data (device pointers: NCHW activations; HOST pointers:
// weights in the reference's [cout][cin][k][k] / [cout][cin2] layout), so that tests can compare the tcgen05 kernels with
// an independent convolution (torch F.conv2d on the CPU) instead of this library's own SIMT conv.
// flags: bit0 ReLU, bit3 bf16 operands, bit18 2x2 max-pool (horizontal half fused into the epilogue + pool2_split);
// res / mul: identity residual / attention product operands ([B,cout,H,W], mul at the stored output's size) or NULL;
// in2 + w_sc_host: fused 1x1 shortcut on a second input with cin2 channels, or NULL.  out: [B,cout,Ho,Wo] fp32.
extern "C" int pmp_debug_conv(pmp_handle *h, const float *in, const float *w_host, const float *res, const float *mul,
                              const float *in2, const float *w_sc_host, int cin, int cout, int ksize, int hw, int batch,
                              int cin2, int flags, float *out, void *stream)
{
    if (!h || !in || !w_host || !out) { set_error("pmp_debug_conv: null pointer"); return PMP_ERR_ARG; }
    PMP_CUDA(cudaSetDevice(h->device));
    const int B = batch, H = hw, W = hw;
    const bool bf = (flags & 8) != 0, pooled = (flags >> 18) & 1, fused = in2 != nullptr;
    const int cinp = pad16(cin), coutp = pad16(cout), cin2p = pad16(cin2);
    if (B <= 0 || !tc_supported(cinp, coutp, ksize, ksize, H, W) || (fused && (!w_sc_host || cin2 <= 0 || res)) ||
        (pooled && (mul || (W & 1) || (H & 1)))) {
        set_error("pmp_debug_conv: configuration not supported by the TC engine");
        return PMP_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf d_in, d_res, d_mul, d_in2, d_out, d_tmp, d_wtc, d_wpair, d_wsc;
    const size_t sp_in = act_bytes(FMT_SPLIT, B, cin, H, W), sp_out = act_bytes(FMT_SPLIT, B, cout, H, W);
    std::vector<uint16_t> pk(tc_packed_elems(cinp, coutp, ksize, ksize)), pp(tc_pair_packed_elems(cinp, coutp, ksize, ksize));
    pack_tc_weights(w_host, cout, cin, ksize, ksize, cinp, coutp, bf, pk.data());
    pack_tc_pair_weights(w_host, cout, cin, ksize, ksize, cinp, coutp, bf, pp.data());
    if (d_in.alloc(sp_in) || d_out.alloc(sp_out) || d_tmp.alloc(sp_out) || d_wtc.alloc(pk.size() * 2) || d_wpair.alloc(pp.size() * 2) ||
        (res && d_res.alloc(sp_out)) || (mul && d_mul.alloc(sp_out)) || (fused && d_in2.alloc(act_bytes(FMT_SPLIT, B, cin2, H, W)))) {
        set_error("pmp_debug_conv: cudaMalloc failed");
        return PMP_ERR_CUDA;
    }
    PMP_CUDA(cudaMemcpyAsync(d_wtc.p, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, s));
    PMP_CUDA(cudaMemcpyAsync(d_wpair.p, pp.data(), pp.size() * 2, cudaMemcpyHostToDevice, s));
    std::vector<uint16_t> pf;
    if (fused) {
        pf.resize(tc_pair_fused_sc_elems(cin2p, coutp, ksize, ksize));
        pack_tc_pair_fused_sc(w_sc_host, cout, cin2, cin2p, coutp, ksize, ksize, bf, pf.data());
        if (d_wsc.alloc(pf.size() * 2)) { set_error("pmp_debug_conv: cudaMalloc failed"); return PMP_ERR_CUDA; }
        PMP_CUDA(cudaMemcpyAsync(d_wsc.p, pf.data(), pf.size() * 2, cudaMemcpyHostToDevice, s));
    }
    auto mk = [&](void *p, int C, int hh, int ww) {
        Act a;
        a.p = p; a.fmt = FMT_SPLIT; a.C = C; a.Cp = pad16(C); a.H = hh; a.W = ww; a.bf16 = bf;
        return a;
    };
    const int Ho = pooled ? H / 2 : H, Wo = pooled ? W / 2 : W;
    Act a_in = mk(d_in.p, cin, H, W), a_out = mk(d_out.p, cout, H, pooled ? W / 2 : W);
    f32_to_split_kernel<<<1024, 256, 0, s>>>(in, a_in, B);
    TcConvArgs ta;
    ta.in = a_in; ta.out = a_out;
    if (res) { ta.res = mk(d_res.p, cout, H, W); f32_to_split_kernel<<<1024, 256, 0, s>>>(res, ta.res, B); }
    Act a_mul;
    if (mul) { a_mul = mk(d_mul.p, cout, Ho, Wo); f32_to_split_kernel<<<1024, 256, 0, s>>>(mul, a_mul, B); }
    if (mul && !pooled) ta.mul = a_mul;
    if (fused) {
        ta.sc_in = mk(d_in2.p, cin2, H, W);
        f32_to_split_kernel<<<1024, 256, 0, s>>>(in2, ta.sc_in, B);
        ta.w_pair_sc = (const uint16_t *)d_wsc.p; ta.sc_cin_pad = cin2p;
    }
    ta.w = (const uint16_t *)d_wtc.p; ta.w_pair = (const uint16_t *)d_wpair.p;
    ta.cin_pad = cinp; ta.cout_pad = coutp; ta.kh = ta.kw = ksize; ta.pad_t = ta.pad_l = ksize / 2; ta.relu = flags & 1; ta.pool = 1;
    ta.hpool = pooled ? 1 : 0;
    int rc = conv_tc(h, ta, B, s);
    if (rc) return rc;
    Act fin = a_out;
    if (pooled) {
        fin = mk(d_tmp.p, cout, Ho, Wo);
        rc = pool2_split(h, a_out, fin, Act(), B, s);
        if (rc) return rc;
    }
    split_to_f32_kernel<<<1024, 256, 0, s>>>(fin, out, B);
    PMP_CUDA(cudaGetLastError());
    PMP_CUDA(cudaStreamSynchronize(s));      // the temporaries are freed on return
    return PMP_OK;
}
