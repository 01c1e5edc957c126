// Handle, weight sets and the activation arena (internal).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace pmp {

// One conv layer's packed weights, in both engine layouts.
struct ConvW {
    int cout = 0, cin = 0, kh = 0, kw = 0;
    float *w_simt = nullptr;     // [cin][kh][kw][cout_pad4]  fp32 (SIMT engine; cout innermost)
    float *bias = nullptr;       // [cout] or nullptr
    // TC engine (split precision) operand images, see pack_tc_weights / pack_tc_pair_weights in conv_tc.cu
    uint16_t *w_tc_f16 = nullptr;
    uint16_t *w_tc_bf16 = nullptr;
    uint16_t *w_pair_f16 = nullptr, *w_pair_bf16 = nullptr;   // CTA-pair kernel images
    int cin_pad = 0, cout_pad = 0;      // channel counts padded to multiples of 16
    // second conv of a ResidualBlock with a 1x1 shortcut: the shortcut's weights packed for fusion into this conv
    uint16_t *w_pair_sc_f16 = nullptr, *w_pair_sc_bf16 = nullptr;
    int sc_cin = 0, sc_cin_pad = 0;
};

// K-chunk layout of a net's first-layer conv for the pair kernel's stem mode (nets.cu: build_stem_tc)
struct StemLayout {
    int planes = 0, nchunk = 0;
    signed char chunk_plane[8] = {0}, chunk_u0[8] = {0};
};

struct WeightSet {
    int net = -1;
    StemLayout stem;
    std::map<std::string, ConvW> convs;     // keyed by reference parameter prefix ("resblock_q1.left.0")
    std::vector<void *> allocs;
};

struct EventPair { cudaEvent_t e0, e1; int cls; double flops, bytes; };

struct Handle {
    int device = 0;
    int engine = PMP_ENGINE_TC;         // the production engine; PMP_ENGINE_SIMT is the exact-fp32 checker engine
    float near_tol = 1e-2f;             // near-threshold report of the decode (pmp_set_near_tol): the map parity bar
    unsigned int *d_status = nullptr;   // device status words: [0] fp16 saturation events of the TC conv epilogues
    int tc_dtype = PMP_TC_FP16;
    int num_sms = 148;
    long long launches = 0;
    bool profiling = false;
    bool tc_attr_set = false, dec_attr_set = false;     // per-device cudaFuncSetAttribute done
    ProfSlot prof[PROF_COUNT];
    std::vector<EventPair> pending;
    std::vector<cudaEvent_t> ev_pool;
    std::map<int, WeightSet> wsets;
    int next_wset = 1;
    // activation arena (grown on demand; never shrinks)
    char *arena = nullptr;
    size_t arena_bytes = 0;
    // scratch for whole-path convenience calls
    char *scratch = nullptr;
    size_t scratch_bytes = 0;
    // driver entry point for tensor-map encoding (resolved lazily; avoids linking libcuda)
    void *encode_tiled = nullptr;
};

int ensure_arena(Handle *h, size_t bytes);
int ensure_scratch(Handle *h, size_t bytes);
int ensure_status(Handle *h);

// ---- nets.cu -----------------------------------------------------------------------------
int weights_create(Handle *h, int net, const float *const *tensors, const int64_t *numel, int n, int *wset);
int weights_destroy(Handle *h, int wset);
int forward_q(Handle *h, int wset, const void *blocks, int in_dtype, int B, float *qt_out, cudaStream_t s);
int forward_msbd(Handle *h, int wset, const void *blocks, int in_dtype, const float *qt, int B,
                 float *o0c0, float *o0c1, float *o1c0, float *o1c1, float *o2c0, float *o2c1, int out_bstride,
                 cudaStream_t s);

// ---- decode.cu ---------------------------------------------------------------------------
int qt_postprocess(Handle *h, const float *qt, int B, float *out_f32, uint8_t *out_u8, cudaStream_t s);
int map2partition(Handle *h, const uint8_t *qt, const float *bt, const float *dire, int B, int cf,
                  uint8_t *hor, uint8_t *ver, int8_t *dout, uint32_t *flags, cudaStream_t s,
                  const double *lamb = nullptr,        // 5 thresholds lamb1..lamb5 or nullptr (reference defaults)
                  const float *qt_raw = nullptr,       // raw (un-rounded) qt maps for the near-threshold report, or nullptr
                  float near_tol = 1e-2f);             // flags bits 1..3: a map value within near_tol of a decision threshold
int assemble_frames(Handle *h, const uint8_t *hor, const uint8_t *ver, const uint8_t *qt, const int8_t *dire,
                    int frames, int bh, int bw, int8_t *out, cudaStream_t s);
int format_text(Handle *h, const int8_t *values, int64_t n, char *text, int64_t *n_bytes_host, cudaStream_t s);
int cut_blocks(Handle *h, const void *y, const void *u, const void *v, int sample_bytes, int frames, int width,
               int height, uint8_t *luma_blocks, uint8_t *chroma_blocks, cudaStream_t s);

}  // namespace pmp
