// Post-process + map-to-partition decode + frame assembly + text formatting + input prep.
//
// Integer/byte kernels, HBM-bound by design (7.7 KB in / 1.3 KB out per block-component).
//
// Reference semantics (paths relative to /root/reference):
//   qt_postprocess_kernel  == Metrics.py:612-637   (check_square_unity, eli_structual_error)
//   map2partition_kernel   == Map2Partition.py:98-373, but the exhaustive cross-product map tree
//                             (:203-266, Search :53-87) is replaced by the equivalent separable
//                             per-CU recursion: candidate lists
//                             and errors depend only on a CU's own region, so
//                             best(cu,d,cur) = min_m own(cu,d,m) + sum_children best(child,d+1,.)
//                             with strict '<' in ascending mode order == the reference's first-min
//                             over its DFS leaf order.
//   assemble_frames_kernel == Map2Partition.py:389-412 (scatter + per-frame vector order)
//   format_text kernels    == Map2Partition.py:405-412 (str(v) + '\n')
//   cut_blocks_kernel      == Inference_QBD.py:104-149,:194-200
//
// Error arithmetic: the reference sums float32 |int - float32| terms with NumPy's pairwise
// float32 summation; here each float32 term (computed exactly as NumPy does) is converted to
// 2^-32 fixed point and summed in int64 -- order independent and exact for sane inputs.  The
// argmin cost is 5*S_bt + 4*S_dire (== 5 * (S_bt + 0.8*S_dire)).  Decisions whose runner-up lies
// within the reference's float32 evaluation noise are reported in flags bit0.
#include "handle.cuh"

namespace pmp {

// ------------------------------------------------------------------------------------------
// D1/D2: QT-map post-process, one thread per block
// ------------------------------------------------------------------------------------------
constexpr int QT_TPB = 128;

// One thread per block for the 4x4 integer rules; loads and stores go through shared memory so that global accesses
// are coalesced (a block's 64 floats are 256 contiguous bytes; thread-per-block strided accesses wasted 7/8 of every sector).
__global__ void __launch_bounds__(QT_TPB)
qt_postprocess_kernel(const float *__restrict__ qt, int B, float *__restrict__ out_f32, uint8_t *__restrict__ out_u8)
{
    __shared__ float sq[QT_TPB * 65];            // row pitch 65: conflict-free scalar reads by the owning thread
    __shared__ uint32_t sm4[QT_TPB * 4];         // repaired 4x4 map, 16 bytes per block
    const int b0 = blockIdx.x * QT_TPB;
    const int nb = (B - b0) < QT_TPB ? (B - b0) : QT_TPB;
    for (int k = threadIdx.x; k < nb * 64; k += QT_TPB) sq[(k >> 6) * 65 + (k & 63)] = qt[(size_t)b0 * 64 + k];
    __syncthreads();
    if ((int)threadIdx.x < nb) {
        const float *q = sq + threadIdx.x * 65;
        int m[16];
        int n0 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float *r0 = q + (2 * i) * 8 + 2 * j, *r1 = r0 + 8;
                float mx = fmaxf(fmaxf(r0[0], r0[1]), fmaxf(r1[0], r1[1]));  // max_pool2d(.,2)
                float r = rintf(mx);                                         // torch.round: half to even
                r = fminf(fmaxf(r, 0.f), 3.f);                               // clamp [0,3]
                int v = (int)r;
                m[i * 4 + j] = v;
                n0 += (v == 0);
            }
        if (n0 <= 12) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (m[k] == 0) m[k] = 1;
#pragma unroll
            for (int i = 0; i < 4; i += 2)
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    int a = m[i * 4 + j], bq = m[i * 4 + j + 1], c = m[(i + 1) * 4 + j], d = m[(i + 1) * 4 + j + 1];
                    int s = a + bq + c + d;
                    if (s >= 5 && s <= 10) {
                        int n1 = (a == 1) + (bq == 1) + (c == 1) + (d == 1);
                        if (n1 < 3) {
                            if (a == 1) a = 2;
                            if (bq == 1) bq = 2;
                            if (c == 1) c = 2;
                            if (d == 1) d = 2;
                        } else {
                            a = bq = c = d = 1;
                        }
                        m[i * 4 + j] = a; m[i * 4 + j + 1] = bq; m[(i + 1) * 4 + j] = c; m[(i + 1) * 4 + j + 1] = d;
                    }
                }
        } else if (n0 < 16) {
#pragma unroll
            for (int k = 0; k < 16; k++) m[k] = 0;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            sm4[threadIdx.x * 4 + i] = (uint32_t)m[i * 4] | ((uint32_t)m[i * 4 + 1] << 8) | ((uint32_t)m[i * 4 + 2] << 16) |
                                       ((uint32_t)m[i * 4 + 3] << 24);
    }
    __syncthreads();
    // nearest x2 while writing: word k4 of the tile = output row i, column half jh of block bb
    for (int k4 = threadIdx.x; k4 < nb * 16; k4 += QT_TPB) {
        const int bb = k4 >> 4, i = (k4 >> 1) & 7, jh = k4 & 1;
        const uint32_t row = sm4[bb * 4 + (i >> 1)] >> (16 * jh);
        const uint32_t v0 = row & 0xffu, v1 = (row >> 8) & 0xffu;
        if (out_u8) reinterpret_cast<uint32_t *>(out_u8)[(size_t)b0 * 16 + k4] = v0 | (v0 << 8) | (v1 << 16) | (v1 << 24);
        if (out_f32) reinterpret_cast<float4 *>(out_f32)[(size_t)b0 * 16 + k4] = make_float4((float)v0, (float)v0, (float)v1, (float)v1);
    }
}

int qt_postprocess(Handle *h, const float *qt, int B, float *out_f32, uint8_t *out_u8, cudaStream_t s)
{
    if (B <= 0) return PMP_OK;
    PMP_CHECK_ARG((((uintptr_t)out_u8) & 3) == 0 && (((uintptr_t)out_f32) & 15) == 0, "qt_postprocess outputs must be 4/16-byte aligned");
    ProfScope ps(h, PROF_POSTPROC, s, 0, (double)B * (256 + 64 + (out_f32 ? 256 : 0)));
    qt_postprocess_kernel<<<cdiv(B, QT_TPB), QT_TPB, 0, s>>>(qt, B, out_f32, out_u8);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// ------------------------------------------------------------------------------------------
// D3-D8: map -> partition, one warp per block
// ------------------------------------------------------------------------------------------
constexpr int DEC_WARPS = 4;
typedef long long i64;
typedef unsigned long long u64;

struct WarpMaps {              // per-warp shared-memory staging of one block
    float ob[3][256];          // unrounded MTT depth maps       (ori_msbt_map)
    float od[3][256];          // unrounded direction maps       (ori_msdire_map)
    int8_t rb[3][256];         // np.round(bt), clamped to +-100 (only compared with 0..6)
    int8_t rd[3][256];         // th_round(dire, 0.5)
    int8_t outd[3][256];       // out_msdire_map
    uint8_t par[2][17 * 17 + 3];
    uint8_t qt[64];
};

struct Cost {
    i64 e5;       // 5*S_bt + 4*S_dire in 2^-32 fixed point
    i64 sd;       // S_dire alone (tells apart exact ties the reference's float32 may not see)
    i64 mingap;   // smallest positive runner-up gap on the chosen subtree
    u64 modes;    // chosen split modes of the subtree, 3 bits each
};

__device__ __forceinline__ i64 fix32(float a)       // |a| -> 2^-32 fixed point (saturating at 2^16)
{
    return __float2ll_rn(fminf(a, 65536.f) * 4294967296.f);
}

// Exact warp sums of non-negative 64-bit values as three 32-bit REDUX.SUM (redux.sync.add) instead of ten SHFL + ten
// 64-bit adds: limbs narrow enough that 32 lanes cannot overflow 32 bits.
__device__ __forceinline__ i64 warp_sum64(i64 v)           // v in [0, 2^56): 2^-32 fixed-point error sums of <= 8 cells per lane
{
#ifdef PMP_DECODE_SHFL_SUMS
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
#else
    const u64 u = (u64)v;
    const unsigned s0 = __reduce_add_sync(0xffffffffu, (unsigned)(u & 0xFFFFFFu));
    const unsigned s1 = __reduce_add_sync(0xffffffffu, (unsigned)((u >> 24) & 0xFFFFFFu));
    const unsigned s2 = __reduce_add_sync(0xffffffffu, (unsigned)(u >> 48));          // < 2^8 per lane
    return (i64)((u64)s0 + ((u64)s1 << 24) + ((u64)s2 << 48));
#endif
}
__device__ __forceinline__ u64 warp_sumu64(u64 v)          // six 10-bit counters at bit 0, 10, ..., 50, each <= 8 per lane
{
#ifdef PMP_DECODE_SHFL_SUMS
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
#else
    const unsigned s0 = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xFFFFFu));
    const unsigned s1 = __reduce_add_sync(0xffffffffu, (unsigned)((v >> 20) & 0xFFFFFu));
    const unsigned s2 = __reduce_add_sync(0xffffffffu, (unsigned)(v >> 40));
    return (u64)s0 + ((u64)s1 << 20) + ((u64)s2 << 40);
#endif
}

struct Eval {
    unsigned cand;     // bit m set: mode m is a candidate (bit 0 always)
    i64 e5[5];         // own level-d cost of each mode on this CU
    i64 sd[5];
};

__device__ __forceinline__ int sub_of(int off, int len, bool tt)
{
    if (!tt) return off >= (len >> 1);
    return off < (len >> 2) ? 0 : (off < ((len * 3) >> 2) ? 1 : 2);
}

// Decode thresholds lamb1..lamb5 (Map2Partition.py:100,:118-122; Python floats, i.e. doubles) -- kernel parameters.
struct Lamb { double l1, l2, l3, l4, l5; };

// Map2Partition.py:140-201 (candidate list) + the level-d error terms of :307-312 for one CU.
__device__ Eval eval_cu(const WarpMaps &wm, const Lamb &lb, int lane, int x, int y, int h, int w, int d, int cur, int cf)
{
    const int n = h * w;
    // legality (:158-165); horizontal modes split h, vertical modes split w
    bool legal[5];
    legal[0] = true;
    legal[1] = (h / (2 * cf) != 0) && (h % (2 * cf) == 0);
    legal[2] = (w / (2 * cf) != 0) && (w % (2 * cf) == 0);
    legal[3] = (h / (4 * cf) != 0) && (h % (4 * cf) == 0);
    legal[4] = (w / (4 * cf) != 0) && (w % (4 * cf) == 0);

    u64 cnt[5] = {0, 0, 0, 0, 0};      // per mode: sub s -> (minus << 20s) | (zero << (20s+10))
    unsigned misc = 0;                 // zero2 | nh << 10 | nv << 20
    i64 ebt[5] = {0, 0, 0, 0, 0};
    i64 ed_none = 0, ed_hor = 0, ed_ver = 0;

    for (int idx = lane; idx < n; idx += 32) {
        int i = idx / w, j = idx - i * w;
        int cell = (x + i) * 16 + (y + j);
        int rbd = wm.rb[d][cell], rb2 = wm.rb[2][cell], rdd = wm.rd[d][cell];
        float obd = wm.ob[d][cell], odd = wm.od[d][cell];
        misc += (rb2 == cur) + ((rdd == 1) << 10) + ((rdd == -1) << 20);
        ed_none += fix32(fabsf(0.f - odd));
        ed_hor += fix32(fabsf(1.f - odd));
        ed_ver += fix32(fabsf(-1.f - odd));
        ebt[0] += fix32(fabsf((float)cur - obd));
#pragma unroll
        for (int m = 1; m <= 4; m++) {
            if (!legal[m]) continue;
            bool horiz = (m & 1);
            int s = sub_of(horiz ? i : j, horiz ? h : w, m >= 3);
            int v = cur + ((m >= 3 && s != 1) ? 2 : 1);
            ebt[m] += fix32(fabsf((float)v - obd));
            cnt[m] += ((u64)(rbd < v) << (20 * s)) + ((u64)(rbd == v) << (20 * s + 10));
        }
    }
    misc = __reduce_add_sync(0xffffffffu, misc);
    ed_none = warp_sum64(ed_none);
    ed_hor = warp_sum64(ed_hor);
    ed_ver = warp_sum64(ed_ver);
    ebt[0] = warp_sum64(ebt[0]);
#pragma unroll
    for (int m = 1; m <= 4; m++)
        if (legal[m]) { cnt[m] = warp_sumu64(cnt[m]); ebt[m] = warp_sum64(ebt[m]); }

    Eval ev;
    ev.e5[0] = 5 * ebt[0] + 4 * ed_none;
    ev.sd[0] = ed_none;
#pragma unroll
    for (int m = 1; m <= 4; m++) {
        i64 edm = (m & 1) ? ed_hor : ed_ver;
        ev.e5[m] = 5 * ebt[m] + 4 * edm;
        ev.sd[m] = edm;
    }
    ev.cand = 1u;
    int zero2 = misc & 1023, nh = (misc >> 10) & 1023, nv = (misc >> 20) & 1023;
    // (i) already at final depth on >= 70% of the CU: no split (:142-145).  Products are formed in
    // double in the reference's association order: (lamb1*h)*w.
    if ((double)zero2 >= lb.l1 * (double)h * (double)w) return ev;
    // (ii) direction vote (:146-154)
    int direction = 0;
    if ((double)(nv + nh) >= lb.l2 * (double)h * (double)w) {
        if ((double)nh >= lb.l3 * (double)nv) direction = 1;
        else if ((double)nv >= lb.l3 * (double)nh) direction = 2;
    }
#pragma unroll
    for (int m = 1; m <= 4; m++) {
        if (!legal[m]) continue;
        if ((m & 1) && direction == 2) continue;
        if (!(m & 1) && direction == 1) continue;
        int nsub = (m >= 3) ? 3 : 2;
        bool ok = true;
        for (int s = 0; s < nsub; s++) {
            int minus = (int)((cnt[m] >> (20 * s)) & 1023), zero = (int)((cnt[m] >> (20 * s + 10)) & 1023);
            int len = (m & 1) ? h : w, oth = (m & 1) ? w : h;
            int sl = (m >= 3) ? ((s == 1) ? (len >> 1) : (len >> 2)) : (len >> 1);
            int np = sl * oth;
            if (!((double)minus < (double)np * lb.l4 && (double)zero > (double)np * lb.l5)) ok = false;   // (:194)
        }
        if (ok) ev.cand |= 1u << m;
    }
    return ev;
}

__device__ __forceinline__ void child_rect(int x, int y, int h, int w, int m, int c, int &cx, int &cy, int &ch,
                                           int &cw)
{
    cx = x; cy = y; ch = h; cw = w;          // Map2Partition.py:124-138
    if (m == 1) { ch = h >> 1; cx = x + c * (h >> 1); }
    else if (m == 2) { cw = w >> 1; cy = y + c * (w >> 1); }
    else if (m == 3) { ch = (c == 1) ? (h >> 1) : (h >> 2); cx = x + (c == 0 ? 0 : (c == 1 ? (h >> 2) : ((h * 3) >> 2))); }
    else if (m == 4) { cw = (c == 1) ? (w >> 1) : (w >> 2); cy = y + (c == 0 ? 0 : (c == 1 ? (w >> 2) : ((w * 3) >> 2))); }
}

template <int D> struct ModeBits { static constexpr int value = 3 + 3 * ModeBits<D + 1>::value; };
template <> struct ModeBits<2> { static constexpr int value = 3; };

constexpr i64 GAP_INF = (i64)1 << 62;

template <int D>
__device__ Cost best_cu(const WarpMaps &wm, const Lamb &lb, int lane, int x, int y, int h, int w, int cur, int cf)
{
    Eval ev = eval_cu(wm, lb, lane, x, y, h, w, D, cur, cf);
    Cost best;
    best.e5 = GAP_INF; best.sd = 0; best.mingap = GAP_INF; best.modes = 0;
    i64 second_e5 = GAP_INF, second_sd = 0;
    bool have = false;
#pragma unroll 1
    for (int m = 0; m <= 4; m++) {
        if (!((ev.cand >> m) & 1u)) continue;
        Cost r;
        r.e5 = ev.e5[m]; r.sd = ev.sd[m]; r.mingap = GAP_INF; r.modes = (u64)m;
        if constexpr (D < 2) {
            int nch = (m == 0) ? 1 : ((m >= 3) ? 3 : 2);
            for (int c = 0; c < nch; c++) {
                int cx, cy, ch, cw;
                child_rect(x, y, h, w, m, c, cx, cy, ch, cw);
                int v = cur + ((m == 0) ? 0 : ((m >= 3 && c != 1) ? 2 : 1));
                Cost cr = best_cu<D + 1>(wm, lb, lane, cx, cy, ch, cw, v, cf);
                r.e5 += cr.e5; r.sd += cr.sd;
                r.mingap = cr.mingap < r.mingap ? cr.mingap : r.mingap;
                r.modes |= cr.modes << (3 + c * ModeBits<D + 1>::value);
            }
        }
        if (!have || r.e5 < best.e5) {                 // strict '<' : first minimum wins
            if (have) { second_e5 = best.e5; second_sd = best.sd; }
            best = r; have = true;
        } else if (r.e5 < second_e5) {
            second_e5 = r.e5; second_sd = r.sd;
        }
    }
    if (second_e5 < GAP_INF) {
        i64 gap = second_e5 - best.e5;
        if (gap > 0 || second_sd != best.sd) best.mingap = gap < best.mingap ? gap : best.mingap;
    }
    return best;
}

__device__ __forceinline__ void fill_rect_i8(int8_t *map, int lane, int x, int y, int h, int w, int8_t v)
{
    for (int idx = lane; idx < h * w; idx += 32) {
        int i = idx / w, j = idx - i * w;
        map[(x + i) * 16 + (y + j)] = v;
    }
}

__device__ __forceinline__ void draw_cu(WarpMaps &wm, int lane, int x, int y, int h, int w)
{
    for (int j = lane; j < w; j += 32) { wm.par[0][x * 17 + y + j] = 1; wm.par[0][(x + h) * 17 + y + j] = 1; }
    for (int i = lane; i < h; i += 32) { wm.par[1][(x + i) * 17 + y] = 1; wm.par[1][(x + i) * 17 + y + w] = 1; }
}

template <int D>
__device__ void apply_tree(WarpMaps &wm, int lane, int x, int y, int h, int w, u64 modes)
{
    int m = (int)(modes & 7);
    if (m != 0) fill_rect_i8(wm.outd[D], lane, x, y, h, w, (m & 1) ? (int8_t)1 : (int8_t)-1);     // :322-324
    int nch = (m == 0) ? 1 : ((m >= 3) ? 3 : 2);
    for (int c = 0; c < nch; c++) {
        int cx, cy, ch, cw;
        child_rect(x, y, h, w, m, c, cx, cy, ch, cw);
        if constexpr (D < 2) apply_tree<D + 1>(wm, lane, cx, cy, ch, cw, modes >> (3 + c * ModeBits<D + 1>::value));
        else draw_cu(wm, lane, cx, cy, ch, cw);                                                    // :339-346
    }
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
map2partition_kernel(const uint8_t *__restrict__ qt, const float *__restrict__ bt, const float *__restrict__ dire,
                     int B, int cf, const Lamb lb, const float *__restrict__ qt_raw, float near_tol,
                     uint8_t *__restrict__ hor, uint8_t *__restrict__ ver, int8_t *__restrict__ dout,
                     uint32_t *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpMaps &wm = reinterpret_cast<WarpMaps *>(smem_raw)[warp];
    const int nwarps_total = gridDim.x * DEC_WARPS;
    for (int b = blockIdx.x * DEC_WARPS + warp; b < B; b += nwarps_total) {
        // ---- stage: coalesced float4 loads, threshold maps (D3: :104-105) ----
        const float4 *bt4 = reinterpret_cast<const float4 *>(bt + (size_t)b * 768);
        const float4 *di4 = reinterpret_cast<const float4 *>(dire + (size_t)b * 768);
        // near-threshold report (north star: "a reported count of CTUs whose values fall within tolerance of a decision
        // threshold"): bit1 a depth value within near_tol of a rounding threshold k + 0.5 (np.round, :104), bit2 a direction
        // value within near_tol of +-0.5 (th_round, :105), bit3 a 2x2-pooled raw qt value within near_tol of 0.5 / 1.5 / 2.5
        // (Metrics.py:631-632; only when the raw qt map is supplied); bits 4-6 the same three tests at near_tol / 100 (with
        // the default 1e-2 that is 1e-4, twice the largest map error measured for the TC engine)
        const float tight = 0.01f * near_tol;
        unsigned near = 0;
        for (int k = lane; k < 192; k += 32) {
            float4 a = __ldg(bt4 + k), c = __ldg(di4 + k);
            float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                int p = 4 * k + e;
                (&wm.ob[0][0])[p] = av[e];
                (&wm.od[0][0])[p] = cv[e];
                float r = rintf(av[e]);                                   // np.round: half to even
                r = (r == r) ? fminf(fmaxf(r, -100.f), 100.f) : 127.f;
                (&wm.rb[0][0])[p] = (int8_t)(int)r;
                (&wm.rd[0][0])[p] = cv[e] >= 0.5f ? (int8_t)1 : (cv[e] <= -0.5f ? (int8_t)-1 : (int8_t)0);
                (&wm.outd[0][0])[p] = 0;
                const float db = fabsf(av[e] - floorf(av[e]) - 0.5f), dd = fabsf(fabsf(cv[e]) - 0.5f);
                near |= (db < near_tol ? 2u : 0u) | (dd < near_tol ? 4u : 0u) | (db < tight ? 16u : 0u) | (dd < tight ? 32u : 0u);
            }
        }
        if (qt_raw && lane < 16) {
            const float *q = qt_raw + (size_t)b * 64 + (lane >> 2) * 16 + (lane & 3) * 2;
            const float mx = fmaxf(fmaxf(q[0], q[1]), fmaxf(q[8], q[9]));
            const float dq = fabsf(mx - floorf(mx) - 0.5f);
            if (mx > 0.f && mx < 3.f) near |= (dq < near_tol ? 8u : 0u) | (dq < tight ? 64u : 0u);
        }
        near = __reduce_or_sync(0xffffffffu, near);
        for (int k = lane; k < 64; k += 32) wm.qt[k] = qt[(size_t)b * 64 + k];
        for (int k = lane; k < 2 * (17 * 17 + 3); k += 32) (&wm.par[0][0])[k] = 0;
        __syncwarp();

        // ---- D4: QT recursion (:348-362) flattened: a node is visited iff every ancestor had
        //      qt[top-left] > its depth ----
        i64 mingap = GAP_INF, emax = 0;
        int regions = 0;
        for (int depth = 0; depth <= 3; depth++) {
            int size = 8 >> depth, nn = 1 << depth;
            for (int node = 0; node < nn * nn; node++) {
                int qx = (node / nn) * size, qy = (node % nn) * size;
                bool active = true;
                for (int a = 0; a < depth; a++) {
                    int as = 8 >> a;
                    if (!(wm.qt[(qx / as * as) * 8 + (qy / as * as)] > a)) { active = false; break; }
                }
                if (!active) continue;
                int cur = wm.qt[qx * 8 + qy];
                if (cur == depth) {
                    Cost c = best_cu<0>(wm, lb, lane, 2 * qx, 2 * qy, 2 * size, 2 * size, 0, cf);
                    apply_tree<0>(wm, lane, 2 * qx, 2 * qy, 2 * size, 2 * size, c.modes);
                    // float32 evaluation noise of the reference scales with the region total
                    i64 tol = (c.e5 >> 19) + 4096;
                    if (c.mingap <= tol) mingap = 0;
                    emax = c.e5 > emax ? c.e5 : emax;
                    regions++;
                } else if (cur > depth && depth < 3) {
                    for (int i = lane; i < 2 * size; i += 32) {
                        wm.par[0][(2 * qx + size) * 17 + 2 * qy + i] = 1;
                        wm.par[1][(2 * qx + i) * 17 + 2 * qy + size] = 1;
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        // ---- D8: export [:16,:16] ----
        for (int k = lane; k < 256; k += 32) {
            int i = k >> 4, j = k & 15;
            hor[(size_t)b * 256 + k] = wm.par[0][i * 17 + j];
            ver[(size_t)b * 256 + k] = wm.par[1][i * 17 + j];
        }
        const int *od4 = reinterpret_cast<const int *>(&wm.outd[0][0]);
        int *g4 = reinterpret_cast<int *>(dout + (size_t)b * 768);
        for (int k = lane; k < 192; k += 32) g4[k] = od4[k];
        if (flags && lane == 0) flags[b] = (mingap == 0 ? 1u : 0u) | near | ((unsigned)regions << 8);
        __syncwarp();
    }
}

int map2partition(Handle *h, const uint8_t *qt, const float *bt, const float *dire, int B, int cf, uint8_t *hor,
                  uint8_t *ver, int8_t *dout, uint32_t *flags, cudaStream_t s, const double *lamb, const float *qt_raw,
                  float near_tol)
{
    if (B <= 0) return PMP_OK;
    PMP_CHECK_ARG(cf == 1 || cf == 2, "chroma_factor must be 1 or 2");
    Lamb lb{0.7, 0.7, 1.5, 0.3, 0.7};            // Map2Partition.py:100 defaults
    if (lamb) {
        for (int i = 0; i < 5; i++) PMP_CHECK_ARG(lamb[i] == lamb[i], "lamb thresholds must not be NaN");
        lb = Lamb{lamb[0], lamb[1], lamb[2], lamb[3], lamb[4]};
    }
    size_t smem = sizeof(WarpMaps) * DEC_WARPS;
    if (!h->dec_attr_set) {         // function attributes are per device: one handle per device
        PMP_CUDA(cudaFuncSetAttribute(map2partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->dec_attr_set = true;
    }
    int grid = cdiv(B, DEC_WARPS);
    int cap = h->num_sms * 8;
    if (grid > cap) grid = cap;
    ProfScope ps(h, PROF_DECODE, s, 0, (double)B * (64 + 6144 + 1280 + 4));
    map2partition_kernel<<<grid, DEC_WARPS * 32, smem, s>>>(qt, bt, dire, B, cf, lb, qt_raw, near_tol, hor, ver, dout, flags);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// ------------------------------------------------------------------------------------------
// D9: frame assembly -- block results -> per-frame vectors in file order
// ------------------------------------------------------------------------------------------
__global__ void assemble_frames_kernel(const uint8_t *__restrict__ hor, const uint8_t *__restrict__ ver,
                                       const uint8_t *__restrict__ qt, const int8_t *__restrict__ dire, int frames,
                                       int bh, int bw, int8_t *__restrict__ out)
{
    const int R = bh * 16, C = bw * 16;
    const long long per = 2LL * R * C + (long long)(R / 2) * (C / 2) + 3LL * R * C;
    long long total = per * frames;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int f = (int)(t / per);
        long long o = t - (long long)f * per;
        int8_t v;
        if (o < 2LL * R * C) {
            int which = o >= (long long)R * C;
            int p = (int)(o - (long long)which * R * C);
            int r = p / C, c = p - r * C;
            long long blk = ((long long)f * bh + (r >> 4)) * bw + (c >> 4);
            const uint8_t *src = which ? ver : hor;
            v = (int8_t)src[blk * 256 + (r & 15) * 16 + (c & 15)];
        } else if (o < 2LL * R * C + (long long)(R / 2) * (C / 2)) {
            int p = (int)(o - 2LL * R * C);
            int r = p / (C / 2), c = p - r * (C / 2);
            long long blk = ((long long)f * bh + (r >> 3)) * bw + (c >> 3);
            v = (int8_t)qt[blk * 64 + (r & 7) * 8 + (c & 7)];
        } else {
            long long p = o - 2LL * R * C - (long long)(R / 2) * (C / 2);
            int l = (int)(p / ((long long)R * C));
            int q = (int)(p - (long long)l * R * C);
            int r = q / C, c = q - r * C;
            long long blk = ((long long)f * bh + (r >> 4)) * bw + (c >> 4);
            v = dire[blk * 768 + l * 256 + (r & 15) * 16 + (c & 15)];
        }
        out[t] = v;
    }
}

int assemble_frames(Handle *h, const uint8_t *hor, const uint8_t *ver, const uint8_t *qt, const int8_t *dire,
                    int frames, int bh, int bw, int8_t *out, cudaStream_t s)
{
    if (frames <= 0 || bh <= 0 || bw <= 0) return PMP_OK;
    long long total = pmp_frame_values(bh, bw) * frames;
    int grid = (int)((total + 255) / 256);
    if (grid > h->num_sms * 16) grid = h->num_sms * 16;
    ProfScope ps(h, PROF_ASSEMBLE, s, 0, 2.0 * (double)total);
    assemble_frames_kernel<<<grid, 256, 0, s>>>(hor, ver, qt, dire, frames, bh, bw, out);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// ------------------------------------------------------------------------------------------
// Text formatting: value v in {-1,0,1,2,3} -> "v\n" (2 or 3 bytes).  Three passes: per-CTA byte
// counts, single-CTA exclusive scan, scatter.
// ------------------------------------------------------------------------------------------
constexpr int TXT_THREADS = 256, TXT_PER_THREAD = 16, TXT_PER_CTA = TXT_THREADS * TXT_PER_THREAD;

__global__ void text_count_kernel(const int8_t *__restrict__ v, long long n, unsigned long long *__restrict__ cta_bytes)
{
    long long base = (long long)blockIdx.x * TXT_PER_CTA;
    int cnt = 0;
    for (int k = threadIdx.x; k < TXT_PER_CTA; k += TXT_THREADS) {
        long long i = base + k;
        if (i < n) cnt += 2 + (v[i] < 0);
    }
    __shared__ int red[TXT_THREADS / 32];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < TXT_THREADS / 32; w++) t += red[w];
        cta_bytes[blockIdx.x] = (unsigned long long)t;
    }
}

__global__ void text_scan_kernel(unsigned long long *__restrict__ cta_bytes, int nblocks,
                                 unsigned long long *__restrict__ total)
{
    // single CTA, sequential chunks of 1024 with a block-wide Hillis-Steele scan
    __shared__ unsigned long long buf[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        unsigned long long x = (i < nblocks) ? cta_bytes[i] : 0;
        buf[threadIdx.x] = x;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            unsigned long long t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        unsigned long long incl = buf[threadIdx.x];
        if (i < nblocks) cta_bytes[i] = carry + incl - x;        // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void text_write_kernel(const int8_t *__restrict__ v, long long n,
                                  const unsigned long long *__restrict__ cta_off, char *__restrict__ text)
{
    // each thread formats TXT_PER_THREAD consecutive values
    long long base = (long long)blockIdx.x * TXT_PER_CTA + (long long)threadIdx.x * TXT_PER_THREAD;
    int8_t vals[TXT_PER_THREAD];
    int mybytes = 0;
#pragma unroll
    for (int k = 0; k < TXT_PER_THREAD; k++) {
        long long i = base + k;
        vals[k] = (i < n) ? v[i] : (int8_t)127;
        if (i < n) mybytes += 2 + (vals[k] < 0);
    }
    // block exclusive scan of mybytes
    __shared__ int wsum[TXT_THREADS / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = mybytes;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    unsigned long long off = cta_off[blockIdx.x] + (unsigned long long)(woff + incl - mybytes);
    char *p = text + off;
#pragma unroll
    for (int k = 0; k < TXT_PER_THREAD; k++) {
        if (base + k < n) {
            int x = vals[k];
            if (x < 0) { *p++ = '-'; x = -x; }
            *p++ = (char)('0' + x);
            *p++ = '\n';
        }
    }
}

int format_text(Handle *h, const int8_t *values, int64_t n, char *text, int64_t *n_bytes_host, cudaStream_t s)
{
    if (n <= 0) { if (n_bytes_host) *n_bytes_host = 0; return PMP_OK; }
    int nblocks = (int)((n + TXT_PER_CTA - 1) / TXT_PER_CTA);
    int rc = ensure_scratch(h, (size_t)(nblocks + 1) * 8);
    if (rc) return rc;
    unsigned long long *cta = reinterpret_cast<unsigned long long *>(h->scratch);
    unsigned long long *total = cta + nblocks;
    ProfScope ps(h, PROF_TEXT, s, 0, 3.5 * (double)n);
    text_count_kernel<<<nblocks, TXT_THREADS, 0, s>>>(values, n, cta);
    text_scan_kernel<<<1, 1024, 0, s>>>(cta, nblocks, total);
    text_write_kernel<<<nblocks, TXT_THREADS, 0, s>>>(values, n, cta, text);
    h->launches += 3;
    PMP_CUDA(cudaGetLastError());
    unsigned long long tot = 0;
    PMP_CUDA(cudaMemcpyAsync(&tot, total, 8, cudaMemcpyDeviceToHost, s));
    PMP_CUDA(cudaStreamSynchronize(s));
    if (n_bytes_host) *n_bytes_host = (int64_t)tot;
    return PMP_OK;
}

// ------------------------------------------------------------------------------------------
// L0: input prep on device
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int sample8(const T *p, long long i, bool reduce10)
{
    int y = (int)p[i];
    if (!reduce10) return y;
    int q = y >> 2, r = y & 3;                   // np.round(y/4): half to even, then clip 0..255
    int o = q + (r == 3) + ((r == 2) & (q & 1));
    return o > 255 ? 255 : o;
}

template <typename T>
__global__ void cut_blocks_kernel(const T *__restrict__ yp, const T *__restrict__ up, const T *__restrict__ vp,
                                  bool reduce10, int frames, int W, int H, uint8_t *__restrict__ lb,
                                  uint8_t *__restrict__ cb)
{
    const int bw = W / 64, bh = H / 64;
    const int blk = blockIdx.x;                  // global block id: (f*bh + bi)*bw + bj
    const int f = blk / (bh * bw), rem = blk - f * bh * bw, bi = rem / bw, bj = rem - bi * bw;
    const long long ybase = (long long)f * H * W, cbase = (long long)f * (H / 2) * (W / 2);
    __shared__ uint8_t tile[68 * 68];
    for (int k = threadIdx.x; k < 68 * 68; k += blockDim.x) {
        int r = k / 68, c = k - r * 68;
        int gy = bi * 64 + r - 4, gx = bj * 64 + c - 4;            // zero pad top/left (:120-121)
        int val = (gy >= 0 && gx >= 0) ? sample8(yp, ybase + (long long)gy * W + gx, reduce10) : 0;
        tile[k] = (uint8_t)val;
        if (lb) lb[(size_t)blk * 4624 + k] = (uint8_t)val;
    }
    __syncthreads();
    if (!cb) return;
    uint8_t *o = cb + (size_t)blk * 3 * 1156;
    for (int k = threadIdx.x; k < 34 * 34; k += blockDim.x) {
        int r = k / 34, c = k - r * 34;
        int a = tile[(2 * r) * 68 + 2 * c], b2 = tile[(2 * r) * 68 + 2 * c + 1];
        int c2 = tile[(2 * r + 1) * 68 + 2 * c], d = tile[(2 * r + 1) * 68 + 2 * c + 1];
        o[k] = (uint8_t)max(max(a, b2), max(c2, d));               // F.max_pool2d(luma block, 2) (:197)
        int gy = bi * 32 + r - 2, gx = bj * 32 + c - 2;
        bool in = (gy >= 0 && gx >= 0);
        long long ci = cbase + (long long)gy * (W / 2) + gx;
        o[1156 + k] = in ? (uint8_t)sample8(up, ci, reduce10) : 0;
        o[2312 + k] = in ? (uint8_t)sample8(vp, ci, reduce10) : 0;
    }
}

int cut_blocks(Handle *h, const void *y, const void *u, const void *v, int sample_bytes, int frames, int width,
               int height, uint8_t *luma_blocks, uint8_t *chroma_blocks, cudaStream_t s)
{
    PMP_CHECK_ARG(sample_bytes == 1 || sample_bytes == 2, "sample_bytes must be 1 or 2");
    int nb = frames * (width / 64) * (height / 64);
    if (nb <= 0) return PMP_OK;
    ProfScope ps(h, PROF_PREP, s, 0, (double)frames * width * height * 1.5 * sample_bytes + (double)nb * (4624 + 3468));
    if (sample_bytes == 2)
        cut_blocks_kernel<uint16_t><<<nb, 256, 0, s>>>((const uint16_t *)y, (const uint16_t *)u, (const uint16_t *)v,
                                                      true, frames, width, height, luma_blocks, chroma_blocks);
    else
        cut_blocks_kernel<uint8_t><<<nb, 256, 0, s>>>((const uint8_t *)y, (const uint8_t *)u, (const uint8_t *)v, false,
                                                     frames, width, height, luma_blocks, chroma_blocks);
    h->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

}  // namespace pmp

extern "C" int64_t pmp_frame_values(int bh, int bw)
{
    int64_t R = (int64_t)bh * 16, C = (int64_t)bw * 16;
    return 2 * R * C + (R / 2) * (C / 2) + 3 * R * C;
}
