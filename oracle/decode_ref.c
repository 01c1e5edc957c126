/*
 * Plain-C restatement of the reference map-to-partition decode and QT post-process.
 * TEST ORACLE ONLY -- never linked into or called from the product path.
 *
 * Follows /root/reference/Map2Partition.py (th_round :30-35, split_cur_map :124-138,
 * can_split_mode_list :140-201, get_candidate_map_tree :203-266 with Search :53-87,
 * set_bt_partition_vector :287-346, set_partition_vector :348-362,
 * map_to_parititon :368-373) and /root/reference/Metrics.py
 * (check_square_unity :612-628, eli_structual_error :630-637).
 *
 * Same algorithm as the reference (EXHAUSTIVE cross-product map tree, leaves in
 * the reference's DFS order, first-minimum argmin), NOT the separable recursion
 * the CUDA kernel uses.  The leaf error reproduces NumPy >= 2 float32 semantics
 * bit for bit: np.sum over the contiguous (h,w) float32 array is
 * 0 + pairwise_sum(n) with NumPy's 8-way unrolled pairwise summation
 * (blocksize 128), the three per-level sums are added in float32, and
 * 0.8 * (...) is float32(0.8) * float32 (NEP-50 weak scalar).
 * Build with -ffp-contract=off (no FMA contraction).
 *
 * Pinned against the reference itself by tests/golden/decode_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXCU 32          /* 3 levels of <=3-way splits: at most 27 CUs */
#define LEAF_CAP 20000000L

typedef struct { int x, y, h, w; } Rect;

typedef struct {
    const float *obt;     /* [3][256] unrounded MTT depth maps */
    const float *odire;   /* [3][256] unrounded direction maps */
    float rbt[3][256];    /* np.round */
    int8_t rdire[3][256]; /* th_round(.,0.5) */
    int cf;
    double lamb[5];       /* lamb1..lamb5 of the constructor (Map2Partition.py:100,:118-122); Python floats = doubles */
    uint8_t par[2][17][17];
    int8_t out_dire[3][256];
    /* search state */
    int rx, ry, rh, rw;                 /* current QT-leaf region */
    int8_t bt[4][256];                  /* bt[d] = map entering level d; bt[d+1] after level-d splits */
    int8_t dire[3][256];                /* dire[d] = direction map produced by level-d splits */
    float best_err; int have_best;
    int8_t best_dire[3][256];
    Rect best_cus[MAXCU]; int best_ncu;
    long leaves; int overflow;
} Ctx;

static float np_pairwise(const float *a, int n)
{
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
    }
}

/* np.sum(np.abs(int8_map[region] - f32_map[region])) as float32 */
static float region_abs_err(const Ctx *c, const int8_t *m, const float *o)
{
    float tmp[256];
    int n = 0;
    for (int i = c->rx; i < c->rx + c->rh; i++)
        for (int j = c->ry; j < c->ry + c->rw; j++)
            tmp[n++] = fabsf((float)m[i * 16 + j] - o[i * 16 + j]);
    return 0.f + np_pairwise(tmp, n);
}

static int split_rects(Rect r, int mode, Rect *out)
{
    int x = r.x, y = r.y, h = r.h, w = r.w;
    switch (mode) {
    case 0: out[0] = r; return 1;
    case 1: out[0] = (Rect){x, y, h / 2, w}; out[1] = (Rect){x + h / 2, y, h / 2, w}; return 2;
    case 2: out[0] = (Rect){x, y, h, w / 2}; out[1] = (Rect){x, y + w / 2, h, w / 2}; return 2;
    case 3: out[0] = (Rect){x, y, h / 4, w}; out[1] = (Rect){x + h / 4, y, h / 2, w};
            out[2] = (Rect){x + (h * 3) / 4, y, h / 4, w}; return 3;
    default: out[0] = (Rect){x, y, h, w / 4}; out[1] = (Rect){x, y + w / 4, h, w / 2};
            out[2] = (Rect){x, y + (w * 3) / 4, h, w / 4}; return 3;
    }
}

static int depth_inc(int mode, int idx) { return mode == 0 ? 0 : ((mode >= 3 && idx != 1) ? 2 : 1); }

/* Map2Partition.py:140-201 */
static int candidate_modes(const Ctx *c, Rect r, const int8_t *cur, int d, int *modes)
{
    int n = r.h * r.w, zero2 = 0, nh = 0, nv = 0;
    for (int i = r.x; i < r.x + r.h; i++)
        for (int j = r.y; j < r.y + r.w; j++) {
            int k = i * 16 + j;
            if (c->rbt[2][k] - (float)cur[k] == 0.f) zero2++;
            if (c->rdire[d][k] == 1) nh++;
            if (c->rdire[d][k] == -1) nv++;
        }
    modes[0] = 0;
    if ((double)zero2 >= c->lamb[0] * r.h * r.w) return 1;
    int direction = 0;
    if ((double)(nv + nh) >= c->lamb[1] * r.h * r.w) {
        if ((double)nh >= c->lamb[2] * nv) direction = 1;
        else if ((double)nv >= c->lamb[2] * nh) direction = 2;
    }
    int nm = 1;
    (void)n;
    for (int mode = 1; mode <= 4; mode++) {
        int len = (mode == 1 || mode == 3) ? r.h : r.w;
        int unit = (mode <= 2 ? 2 : 4) * c->cf;
        if (len / unit == 0 || len % unit != 0) continue;
        if ((mode == 1 || mode == 3) && direction == 2) continue;
        if ((mode == 2 || mode == 4) && direction == 1) continue;
        Rect sub[3];
        int ns = split_rects(r, mode, sub), ok = 1;
        for (int s = 0; s < ns; s++) {
            int inc = depth_inc(mode, s), minus = 0, zero = 0;
            for (int i = sub[s].x; i < sub[s].x + sub[s].h; i++)
                for (int j = sub[s].y; j < sub[s].y + sub[s].w; j++) {
                    int k = i * 16 + j;
                    float cmp = c->rbt[d][k] - (float)(cur[k] + inc);
                    if (cmp < 0.f) minus++;
                    if (cmp == 0.f) zero++;
                }
            int np_ = sub[s].h * sub[s].w;
            if (!((double)minus < np_ * c->lamb[3] && (double)zero > np_ * c->lamb[4])) ok = 0;
        }
        if (ok) modes[nm++] = mode;
    }
    return nm;
}

/* Map2Partition.py:203-266 + :297-315 -- DFS over the cross-product map tree */
static void enumerate(Ctx *c, int d, const Rect *cus, int ncu)
{
    if (c->overflow) return;
    if (d == 3) {
        if (++c->leaves > LEAF_CAP) { c->overflow = 1; return; }
        float eb = region_abs_err(c, c->bt[1], c->obt) +
                   region_abs_err(c, c->bt[2], c->obt + 256) +
                   region_abs_err(c, c->bt[3], c->obt + 512);
        float ed = region_abs_err(c, c->dire[0], c->odire) +
                   region_abs_err(c, c->dire[1], c->odire + 256) +
                   region_abs_err(c, c->dire[2], c->odire + 512);
        float err = eb + 0.8f * ed;
        if (!c->have_best || err < c->best_err) {
            c->have_best = 1; c->best_err = err;
            memcpy(c->best_dire, c->dire, sizeof c->dire);
            memcpy(c->best_cus, cus, ncu * sizeof(Rect)); c->best_ncu = ncu;
        }
        return;
    }
    int cand[MAXCU][5], ncand[MAXCU], pick[MAXCU];
    for (int i = 0; i < ncu; i++) { ncand[i] = candidate_modes(c, cus[i], c->bt[d], d, cand[i]); pick[i] = 0; }
    for (;;) {
        /* build child maps for this combination */
        memcpy(c->bt[d + 1], c->bt[d], 256);
        memset(c->dire[d], 0, 256);
        Rect child[MAXCU]; int nchild = 0;
        for (int i = 0; i < ncu; i++) {
            int m = cand[i][pick[i]];
            Rect sub[3];
            int ns = split_rects(cus[i], m, sub);
            for (int s = 0; s < ns; s++) child[nchild++] = sub[s];
            if (m == 0) continue;
            int8_t dv = (m == 1 || m == 3) ? 1 : -1;
            for (int a = cus[i].x; a < cus[i].x + cus[i].h; a++)
                for (int b = cus[i].y; b < cus[i].y + cus[i].w; b++) c->dire[d][a * 16 + b] = dv;
            for (int s = 0; s < ns; s++)
                for (int a = sub[s].x; a < sub[s].x + sub[s].h; a++)
                    for (int b = sub[s].y; b < sub[s].y + sub[s].w; b++)
                        c->bt[d + 1][a * 16 + b] += (int8_t)depth_inc(m, s);
        }
        enumerate(c, d + 1, child, nchild);
        /* odometer: last CU varies fastest (first CU most significant) */
        int i = ncu - 1;
        while (i >= 0 && ++pick[i] == ncand[i]) { pick[i] = 0; i--; }
        if (i < 0) break;
    }
}

/* Map2Partition.py:287-346 */
static void decode_mtt_region(Ctx *c, int x, int y, int h, int w)
{
    c->rx = x; c->ry = y; c->rh = h; c->rw = w;
    memset(c->bt[0], 0, 256);
    c->have_best = 0; c->best_ncu = 0;
    Rect root = {x, y, h, w};
    enumerate(c, 0, &root, 1);
    if (!c->have_best) return;
    for (int l = 0; l < 3; l++)
        for (int i = x; i < x + h; i++)
            for (int j = y; j < y + w; j++) c->out_dire[l][i * 16 + j] = c->best_dire[l][i * 16 + j];
    for (int k = 0; k < c->best_ncu; k++) {
        Rect r = c->best_cus[k];
        for (int j = 0; j < r.w; j++) { c->par[0][r.x][r.y + j] = 1; c->par[0][r.x + r.h][r.y + j] = 1; }
        for (int i = 0; i < r.h; i++) { c->par[1][r.x + i][r.y] = 1; c->par[1][r.x + i][r.y + r.w] = 1; }
    }
}

/* Map2Partition.py:348-362 */
static void decode_qt(Ctx *c, const float *qt, int depth, int qx, int qy)
{
    float cur = qt[qx * 8 + qy];
    int size = 8 >> depth;
    if (cur == (float)depth) {
        decode_mtt_region(c, 2 * qx, 2 * qy, 2 * size, 2 * size);
    } else if (cur > (float)depth) {
        for (int i = 0; i < 2 * size; i++) {
            c->par[0][2 * qx + size][2 * qy + i] = 1;
            c->par[1][2 * qx + i][2 * qy + size] = 1;
        }
        for (int io = 0; io < 2; io++)
            for (int jo = 0; jo < 2; jo++)
                decode_qt(c, qt, depth + 1, qx + io * size / 2, qy + jo * size / 2);
    }
}

/* map_to_parititon for a batch.  qt [n][64] f32 (ints), bt/dire [n][3][256] f32.
 * Outputs hor/ver [n][256] u8, dire_out [n][3][256] i8.  Returns the number of blocks
 * whose search overflowed LEAF_CAP (0 normally); leaves_out (optional) gets the leaf count
 * per block. */
int oracle_map_to_partition_lamb(const float *qt, const float *bt, const float *dire, int n, int chroma_factor,
                                 const double *lamb, uint8_t *hor, uint8_t *ver, int8_t *dire_out, long *leaves_out);

int oracle_map_to_partition(const float *qt, const float *bt, const float *dire, int n, int chroma_factor,
                            uint8_t *hor, uint8_t *ver, int8_t *dire_out, long *leaves_out)
{
    static const double def[5] = {0.7, 0.7, 1.5, 0.3, 0.7};      /* Map2Partition.py:100 */
    return oracle_map_to_partition_lamb(qt, bt, dire, n, chroma_factor, def, hor, ver, dire_out, leaves_out);
}

int oracle_map_to_partition_lamb(const float *qt, const float *bt, const float *dire, int n, int chroma_factor,
                                 const double *lamb, uint8_t *hor, uint8_t *ver, int8_t *dire_out, long *leaves_out)
{
    int bad = 0;
    static _Thread_local Ctx ctx;
    for (int b = 0; b < n; b++) {
        Ctx *c = &ctx;
        memset(c, 0, sizeof *c);
        c->obt = bt + (size_t)b * 768;
        c->odire = dire + (size_t)b * 768;
        c->cf = chroma_factor;
        for (int k = 0; k < 5; k++) c->lamb[k] = lamb[k];
        for (int k = 0; k < 768; k++) {
            float v = c->obt[k];
            (&c->rbt[0][0])[k] = rintf(v);                      /* np.round: half to even */
            float dv = c->odire[k];
            (&c->rdire[0][0])[k] = dv >= 0.5f ? 1 : (dv <= -0.5f ? -1 : 0);
        }
        decode_qt(c, qt + (size_t)b * 64, 0, 0, 0);
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 16; j++) {
                hor[(size_t)b * 256 + i * 16 + j] = c->par[0][i][j];
                ver[(size_t)b * 256 + i * 16 + j] = c->par[1][i][j];
            }
        memcpy(dire_out + (size_t)b * 768, c->out_dire, 768);
        if (leaves_out) leaves_out[b] = c->leaves;
        bad += c->overflow;
    }
    return bad;
}

/* Metrics.py:612-637 for a batch: qt [n][64] f32 -> out [n][64] f32 holding ints 0..3 */
void oracle_qt_postprocess(const float *qt, int n, float *out)
{
    for (int b = 0; b < n; b++) {
        const float *q = qt + (size_t)b * 64;
        int m[4][4], n0 = 0;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                float a = q[(2 * i) * 8 + 2 * j], b2 = q[(2 * i) * 8 + 2 * j + 1];
                float c2 = q[(2 * i + 1) * 8 + 2 * j], d = q[(2 * i + 1) * 8 + 2 * j + 1];
                float mx = fmaxf(fmaxf(a, b2), fmaxf(c2, d));
                float r = rintf(mx);
                r = r < 0.f ? 0.f : (r > 3.f ? 3.f : r);
                m[i][j] = (int)r;
                n0 += (m[i][j] == 0);
            }
        if (n0 <= 12) {
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (m[i][j] == 0) m[i][j] = 1;
            for (int i = 0; i < 4; i += 2)
                for (int j = 0; j < 4; j += 2) {
                    int s = m[i][j] + m[i][j + 1] + m[i + 1][j] + m[i + 1][j + 1];
                    if (s >= 5 && s <= 10) {
                        int n1 = (m[i][j] == 1) + (m[i][j + 1] == 1) + (m[i + 1][j] == 1) + (m[i + 1][j + 1] == 1);
                        for (int a = 0; a < 2; a++)
                            for (int bb = 0; bb < 2; bb++) {
                                if (n1 < 3) { if (m[i + a][j + bb] == 1) m[i + a][j + bb] = 2; }
                                else m[i + a][j + bb] = 1;
                            }
                    }
                }
        } else if (n0 < 16) {
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = 0;
        }
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) out[(size_t)b * 64 + i * 8 + j] = (float)m[i / 2][j / 2];
    }
}
