// C ABI of libpmp_b200 (include/pmp_b200.h): handle lifetime, error reporting, per-kernel-class CUDA-event
// timers, and thin argument-checking wrappers around the engines in nets.cu / decode.cu / conv_*.cu.
#include "handle.cuh"
#include "kernels.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace pmp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PMP_ERR_NO_DEVICE : PMP_ERR_CUDA;
}

static const char *const kProfNames[PROF_COUNT] = {"conv_tc", "conv_simt", "elementwise", "decode", "qt_postprocess",
                                                   "assemble_frames", "cut_blocks", "format_text"};

ProfScope::ProfScope(Handle *h_, int cls_, cudaStream_t s_, double flops_, double bytes_)
    : h(h_), cls(cls_), s(s_), flops(flops_), bytes(bytes_)
{
    if (!h->profiling) {
        h->prof[cls].launches++;
        h->prof[cls].flops += flops;
        h->prof[cls].bytes += bytes;
        return;
    }
    for (cudaEvent_t *e : {&e0, &e1}) {
        if (!h->ev_pool.empty()) { *e = h->ev_pool.back(); h->ev_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(e0, s);
}

ProfScope::~ProfScope()
{
    if (!e0) return;
    cudaEventRecord(e1, s);
    EventPair p{e0, e1, cls, flops, bytes};
    h->pending.push_back(p);
}

static void drain_profile(Handle *h)
{
    for (auto &p : h->pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            h->prof[p.cls].ms += ms;
            h->prof[p.cls].launches++;
            h->prof[p.cls].flops += p.flops;
            h->prof[p.cls].bytes += p.bytes;
        }
        h->ev_pool.push_back(p.e0);
        h->ev_pool.push_back(p.e1);
    }
    h->pending.clear();
}

int ensure_arena(Handle *h, size_t bytes)
{
    if (bytes <= h->arena_bytes) return PMP_OK;
    if (h->arena) {
        PMP_CUDA(cudaDeviceSynchronize());
        PMP_CUDA(cudaFree(h->arena));
        h->arena = nullptr;
        h->arena_bytes = 0;
    }
    bytes = (bytes + ((size_t)1 << 21)) & ~(((size_t)1 << 21) - 1);
    PMP_CUDA(cudaMalloc((void **)&h->arena, bytes));
    h->arena_bytes = bytes;
    return PMP_OK;
}

int ensure_scratch(Handle *h, size_t bytes)
{
    if (bytes <= h->scratch_bytes) return PMP_OK;
    if (h->scratch) {
        PMP_CUDA(cudaDeviceSynchronize());
        PMP_CUDA(cudaFree(h->scratch));
        h->scratch = nullptr;
        h->scratch_bytes = 0;
    }
    bytes = (bytes + 65535) & ~(size_t)65535;
    PMP_CUDA(cudaMalloc((void **)&h->scratch, bytes));
    h->scratch_bytes = bytes;
    return PMP_OK;
}

int ensure_status(Handle *h)
{
    if (h->d_status) return PMP_OK;
    PMP_CUDA(cudaMalloc((void **)&h->d_status, 64));
    PMP_CUDA(cudaMemset(h->d_status, 0, 64));
    return PMP_OK;
}

}  // namespace pmp

using namespace pmp;

#define H_CHECK(h)                                                   \
    do {                                                             \
        if (!(h)) { set_error("null handle"); return PMP_ERR_ARG; }  \
        cudaError_t _e = cudaSetDevice((h)->device);                 \
        if (_e != cudaSuccess) return cuda_fail(_e, "cudaSetDevice", __FILE__, __LINE__); \
    } while (0)

struct pmp_handle : public pmp::Handle {};

extern "C" {

int pmp_version(void) { return PMP_B200_VERSION; }

const char *pmp_last_error(void) { return g_err; }

int pmp_create(int device, pmp_handle **out)
{
    PMP_CHECK_ARG(out != nullptr, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libpmp_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PMP_ERR_NO_DEVICE;
    }
    PMP_CHECK_ARG(device >= 0 && device < count, "device index out of range");
    PMP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PMP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libpmp_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return PMP_ERR_UNSUPPORTED;
    }
    pmp_handle *h = new pmp_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return PMP_OK;
}

void pmp_destroy(pmp_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    drain_profile(h);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    std::vector<int> ids;
    for (auto &kv : h->wsets) ids.push_back(kv.first);
    for (int id : ids) weights_destroy(h, id);
    if (h->arena) cudaFree(h->arena);
    if (h->scratch) cudaFree(h->scratch);
    if (h->d_status) cudaFree(h->d_status);
    delete h;
}

int pmp_set_engine(pmp_handle *h, int engine, int tc_dtype)
{
    H_CHECK(h);
    PMP_CHECK_ARG(engine == PMP_ENGINE_SIMT || engine == PMP_ENGINE_TC, "unknown engine");
    PMP_CHECK_ARG(tc_dtype == PMP_TC_FP16 || tc_dtype == PMP_TC_BF16, "unknown tc dtype");
    h->engine = engine;
    h->tc_dtype = tc_dtype;
    return PMP_OK;
}

int pmp_get_engine(pmp_handle *h) { return h ? h->engine : PMP_ERR_ARG; }

int pmp_set_near_tol(pmp_handle *h, float near_tol)
{
    H_CHECK(h);
    PMP_CHECK_ARG(near_tol >= 0.f && near_tol <= 0.5f, "near_tol must be in [0, 0.5]");
    h->near_tol = near_tol;
    return PMP_OK;
}

int pmp_saturation_count(pmp_handle *h, long long *count_host, int reset)
{
    H_CHECK(h);
    PMP_CHECK_ARG(count_host != nullptr, "null pointer");
    *count_host = 0;
    if (!h->d_status) return PMP_OK;
    unsigned int v = 0;
    PMP_CUDA(cudaDeviceSynchronize());
    PMP_CUDA(cudaMemcpy(&v, h->d_status, sizeof(v), cudaMemcpyDeviceToHost));
    if (reset) PMP_CUDA(cudaMemset(h->d_status, 0, sizeof(v)));
    *count_host = (long long)v;
    return PMP_OK;
}

long long pmp_launch_count(pmp_handle *h) { return h ? h->launches : -1; }

int pmp_profile(pmp_handle *h, int enable)
{
    H_CHECK(h);
    drain_profile(h);
    h->profiling = enable != 0;
    if (enable == 2)
        for (auto &p : h->prof) p = ProfSlot();
    return PMP_OK;
}

int pmp_profile_read(pmp_handle *h, int idx, char *name, int name_len, double *ms, long long *launches, double *flops,
                     double *bytes)
{
    H_CHECK(h);
    if (idx < 0 || idx >= PROF_COUNT) return PMP_ERR_ARG;
    drain_profile(h);
    if (name && name_len > 0) {
        strncpy(name, kProfNames[idx], name_len - 1);
        name[name_len - 1] = 0;
    }
    if (ms) *ms = h->prof[idx].ms;
    if (launches) *launches = h->prof[idx].launches;
    if (flops) *flops = h->prof[idx].flops;
    if (bytes) *bytes = h->prof[idx].bytes;
    return PMP_OK;
}

int pmp_weights_create(pmp_handle *h, int net, const float *const *tensors_host, const int64_t *numel, int n_tensors,
                       int *wset)
{
    H_CHECK(h);
    PMP_CHECK_ARG(tensors_host && numel && wset, "null pointer");
    PMP_CHECK_ARG(net >= 0 && net <= 3, "unknown net kind");
    return weights_create(h, net, tensors_host, numel, n_tensors, wset);
}

int pmp_weights_destroy(pmp_handle *h, int wset)
{
    H_CHECK(h);
    return weights_destroy(h, wset);
}

int pmp_forward_q(pmp_handle *h, int wset, const void *blocks, int in_dtype, int B, float *qt_out, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B >= 0, "negative batch");
    if (B == 0) return PMP_OK;
    PMP_CHECK_ARG(blocks && qt_out, "null pointer");
    PMP_CHECK_ARG(in_dtype == PMP_IN_U8 || in_dtype == PMP_IN_F32, "unknown input dtype");
    return forward_q(h, wset, blocks, in_dtype, B, qt_out, (cudaStream_t)stream);
}

int pmp_forward_msbd(pmp_handle *h, int wset, const void *blocks, int in_dtype, const float *qt, int B, float *out0,
                     float *out1, float *out2, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B >= 0, "negative batch");
    if (B == 0) return PMP_OK;
    PMP_CHECK_ARG(blocks && qt && out0 && out1 && out2, "null pointer");
    PMP_CHECK_ARG(in_dtype == PMP_IN_U8 || in_dtype == PMP_IN_F32, "unknown input dtype");
    return forward_msbd(h, wset, blocks, in_dtype, qt, B, out0, out0 + 256, out1, out1 + 256, out2, out2 + 256, 512,
                        (cudaStream_t)stream);
}

int pmp_predict_maps(pmp_handle *h, int wset_q, int wset_msbd, const void *blocks, int in_dtype, int B, float *qt,
                     float *bt, float *dire, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B >= 0, "negative batch");
    if (B == 0) return PMP_OK;
    PMP_CHECK_ARG(blocks && qt && bt && dire, "null pointer");
    PMP_CHECK_ARG(in_dtype == PMP_IN_U8 || in_dtype == PMP_IN_F32, "unknown input dtype");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = forward_q(h, wset_q, blocks, in_dtype, B, qt, s);
    if (rc) return rc;
    // regrouped outputs of inference_pre_QBD (Metrics.py:399-402): bt[:,k] = out_k[:,0], dire[:,k] = out_k[:,1]
    return forward_msbd(h, wset_msbd, blocks, in_dtype, qt, B, bt, dire, bt + 256, dire + 256, bt + 512, dire + 512, 768,
                        s);
}

int pmp_debug_stem(pmp_handle *h, int wset, const void *blocks, int in_dtype, const float *qt, int B, float *out,
                   void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B > 0 && blocks && out, "bad argument");
    PMP_CHECK_ARG(in_dtype == PMP_IN_U8 || in_dtype == PMP_IN_F32, "unknown input dtype");
    return debug_stem(h, wset, blocks, in_dtype, qt, B, out, (cudaStream_t)stream);
}

int pmp_qt_postprocess(pmp_handle *h, const float *qt, int B, float *out_f32, uint8_t *out_u8, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B >= 0, "negative batch");
    if (B == 0) return PMP_OK;
    PMP_CHECK_ARG(qt && (out_f32 || out_u8), "null pointer");
    return qt_postprocess(h, qt, B, out_f32, out_u8, (cudaStream_t)stream);
}

int pmp_map2partition(pmp_handle *h, const uint8_t *qt_u8, const float *bt, const float *dire, int B, int chroma_factor,
                      uint8_t *hor, uint8_t *ver, int8_t *dire_out, uint32_t *flags, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B >= 0, "negative batch");
    if (B == 0) return PMP_OK;
    PMP_CHECK_ARG(qt_u8 && bt && dire && hor && ver && dire_out, "null pointer");
    return map2partition(h, qt_u8, bt, dire, B, chroma_factor, hor, ver, dire_out, flags, (cudaStream_t)stream, nullptr,
                         nullptr, h->near_tol);
}

int pmp_map2partition_ex(pmp_handle *h, const uint8_t *qt_u8, const float *bt, const float *dire, int B, int chroma_factor,
                         const double *lamb, const float *qt_raw, float near_tol, uint8_t *hor, uint8_t *ver,
                         int8_t *dire_out, uint32_t *flags, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(B >= 0, "negative batch");
    if (B == 0) return PMP_OK;
    PMP_CHECK_ARG(qt_u8 && bt && dire && hor && ver && dire_out, "null pointer");
    PMP_CHECK_ARG(near_tol >= 0.f && near_tol <= 0.5f, "near_tol must be in [0, 0.5]");
    return map2partition(h, qt_u8, bt, dire, B, chroma_factor, hor, ver, dire_out, flags, (cudaStream_t)stream, lamb, qt_raw,
                         near_tol);
}

int pmp_assemble_frames(pmp_handle *h, const uint8_t *hor, const uint8_t *ver, const uint8_t *qt_u8, const int8_t *dire,
                        int frames, int bh, int bw, int8_t *out, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(frames >= 0 && bh >= 0 && bw >= 0, "negative size");
    if (frames == 0 || bh == 0 || bw == 0) return PMP_OK;
    PMP_CHECK_ARG(hor && ver && qt_u8 && dire && out, "null pointer");
    return assemble_frames(h, hor, ver, qt_u8, dire, frames, bh, bw, out, (cudaStream_t)stream);
}

int pmp_format_text(pmp_handle *h, const int8_t *values, int64_t n, char *text, int64_t *n_bytes_host, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(n >= 0, "negative size");
    PMP_CHECK_ARG(n == 0 || (values && text), "null pointer");
    return format_text(h, values, n, text, n_bytes_host, (cudaStream_t)stream);
}

int pmp_cut_blocks(pmp_handle *h, const void *y, const void *u, const void *v, int sample_bytes, int frames, int width,
                   int height, uint8_t *luma_blocks, uint8_t *chroma_blocks, void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(frames >= 0 && width >= 0 && height >= 0, "negative size");
    PMP_CHECK_ARG((width % 2) == 0 && (height % 2) == 0, "4:2:0 planes need even width and height");
    if (frames == 0 || width < 64 || height < 64) return PMP_OK;
    PMP_CHECK_ARG(y && (luma_blocks || chroma_blocks), "null pointer");
    PMP_CHECK_ARG(!chroma_blocks || (u && v), "chroma blocks need u and v planes");
    return cut_blocks(h, y, u, v, sample_bytes, frames, width, height, luma_blocks, chroma_blocks, (cudaStream_t)stream);
}

int pmp_run_component(pmp_handle *h, int wset_q, int wset_msbd, int luma, const uint8_t *blocks, int frames, int bh,
                      int bw, int chunk, int8_t *out, float *qt_raw, float *bt, float *dire, uint32_t *flags,
                      void *stream)
{
    H_CHECK(h);
    PMP_CHECK_ARG(frames >= 0 && bh >= 0 && bw >= 0, "negative size");
    const long long N = (long long)frames * bh * bw;
    if (N == 0) return PMP_OK;
    PMP_CHECK_ARG(blocks && out, "null pointer");
    PMP_CHECK_ARG(N < (1LL << 30), "too many blocks");
    if (chunk <= 0) chunk = 2048;
    cudaStream_t s = (cudaStream_t)stream;
    // scratch: per-chunk maps (unless the caller wants them) + per-sequence integer results
    const size_t per_blk_maps = (64 + 768 + 768) * sizeof(float);
    const size_t off_qt8 = 0, off_hor = off_qt8 + (size_t)N * 64, off_ver = off_hor + (size_t)N * 256,
                 off_dir = off_ver + (size_t)N * 256, off_maps = (off_dir + (size_t)N * 768 + 255) & ~(size_t)255;
    int rc = ensure_scratch(h, off_maps + (size_t)chunk * per_blk_maps + 4096);
    if (rc) return rc;
    char *sc = h->scratch;
    uint8_t *qt8 = (uint8_t *)(sc + off_qt8), *hor = (uint8_t *)(sc + off_hor), *ver = (uint8_t *)(sc + off_ver);
    int8_t *dout = (int8_t *)(sc + off_dir);
    float *cqt = (float *)(sc + off_maps), *cbt = cqt + (size_t)chunk * 64, *cdi = cbt + (size_t)chunk * 768;
    const size_t blk_bytes = luma ? 68 * 68 : 3 * 34 * 34;
    for (long long b0 = 0; b0 < N; b0 += chunk) {
        int nb = (int)((N - b0 < chunk) ? (N - b0) : chunk);
        float *pqt = qt_raw ? qt_raw + b0 * 64 : cqt, *pbt = bt ? bt + b0 * 768 : cbt, *pdi = dire ? dire + b0 * 768 : cdi;
        rc = pmp_predict_maps(h, wset_q, wset_msbd, blocks + b0 * blk_bytes, PMP_IN_U8, nb, pqt, pbt, pdi, s);
        if (rc) return rc;
        rc = qt_postprocess(h, pqt, nb, nullptr, qt8 + b0 * 64, s);
        if (rc) return rc;
        rc = map2partition(h, qt8 + b0 * 64, pbt, pdi, nb, luma ? 1 : 2, hor + b0 * 256, ver + b0 * 256, dout + b0 * 768,
                           flags ? flags + b0 : nullptr, s, nullptr, pqt, h->near_tol);
        if (rc) return rc;
    }
    return assemble_frames(h, hor, ver, qt8, dout, frames, bh, bw, out, s);
}

}  // extern "C"
