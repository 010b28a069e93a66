// fic_api.cu -- the C ABI of libfic_b200.so (see include/fic_b200.h): handle, workspace,
// call sequencing on one CUDA stream, timings, and the .run stream helpers.
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "fic_device.cuh"

using namespace fic;

// ------------------------------------------------------------------------------------
// geometry / argument checking
// ------------------------------------------------------------------------------------
int fic::make_geom(int W, int H, int B, int wk, int mode, Geom *g, const char **why)
{
    const char *dummy;
    if (!why) why = &dummy;
    if (mode < FIC_MODE_GREY || mode > FIC_MODE_GREY_ISO) { *why = "mode must be FIC_MODE_GREY, FIC_MODE_RGB or FIC_MODE_GREY_ISO"; return FIC_E_ARG; }
    // FC:1019 `abstand = blockgroesse / 4` is 0 for B < 4 (ArithmeticException); the
    // reference GUI offers 4, 8, 16 (RLEAppView.fxml:55).  Larger blocks leave the range
    // in which the reference's float covariance is an exact integer, so they are refused.
    if (!(B == 4 || B == 8 || B == 16)) { *why = "blockgroesse must be 4, 8 or 16"; return FIC_E_ARG; }
    if (W <= 0 || H <= 0 || W % B || H % B) { *why = "width and height must be positive multiples of blockgroesse"; return FIC_E_ARG; }
    if (W > 32768 || H > 32768) { *why = "image larger than 32768 pixels on a side"; return FIC_E_ARG; }
    int rpw = W / B, rph = H / B;
    if (rpw < 2 || rph < 2) { *why = "need at least 2x2 range blocks"; return FIC_E_ARG; }
    int dpw = 2 * rpw - 3, dph = 2 * rph - 3;
    if (wk < 1 || wk > dpw || wk > dph) { *why = "widthKernel must be in [1, min(domain blocks per width, per height)]"; return FIC_E_ARG; }
    int64_t ND = (int64_t)dpw * dph;
    // The reference keeps indices in float (FC:124, FC:629): exact only below 2^24.
    if (ND >= (1 << 24) || (int64_t)wk * wk >= (1 << 24)) { *why = "domain pool too large for the reference's float index (>= 2^24)"; return FIC_E_ARG; }
    g->W = W; g->H = H; g->B = B; g->n = B * B; g->wk = wk;
    g->rpw = rpw; g->rph = rph; g->dpw = dpw; g->dph = dph;
    g->sw = W / 2; g->sh = H / 2; g->step = B / 4; g->C = mode == FIC_MODE_RGB ? 3 : 1;
    g->n_iso = mode == FIC_MODE_GREY_ISO ? 8 : 1;
    g->NR = (int64_t)rpw * rph; g->ND = ND;
    return FIC_OK;
}

// ------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------
struct fic_handle {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr};
    Work w;
    int engine_opt = FIC_ENGINE_AUTO;
    int umma_kind = FIC_UMMA_KIND_AUTO;
    int f16_state = 0;  // kind::f16 self-test: 0 not run yet, 1 exact, -1 not exact on this device (use kind::i8)
    fic_timings tm;
    char err[512];
    bool tm_pending_dev = false;
    // pinned scratch for small device->host reads
    unsigned long long *h_acc = nullptr;
};

static char g_create_err[512] = "";

static int set_err(fic_handle *h, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h ? h->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return set_err(h, FIC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

enum Slot { S_ARGB, S_SRC, S_DEC, S_DSUM, S_DSQ, S_RSUM, S_BEST, S_INFO, S_Q, S_OPA, S_OPB, S_IMG, S_DEC2, S_DCODE, S_PERR, S_ACC, S_DEC3 };

template <typename T>
static int ensure(fic_handle *h, T *&p, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (h->w.cap[slot] >= bytes && p) return FIC_OK;
    if (p) cudaFree(p);
    p = nullptr;
    h->w.cap[slot] = 0;
    size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc((void **)&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(h, FIC_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    h->w.cap[slot] = want;
    return FIC_OK;
}

#define ENSURE(ptr, slot, bytes)                       \
    do {                                               \
        int rc_ = ensure(h, ptr, slot, (size_t)(bytes)); \
        if (rc_) return rc_;                           \
    } while (0)

extern "C" {

const char *fic_version(void) { return "fic_b200 0.1 (sm_100a)"; }

int fic_create(int device, fic_handle **out)
{
    fic_handle *h = nullptr;
    if (!out) return set_err(nullptr, FIC_E_ARG, "fic_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_err(nullptr, FIC_E_CUDA, "no CUDA device available (%s); libfic_b200 has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return set_err(nullptr, FIC_E_ARG, "device %d out of range [0,%d)", device, count);
    h = new (std::nothrow) fic_handle();
    if (!h) return set_err(nullptr, FIC_E_NOMEM, "out of host memory");
    h->err[0] = 0;
    memset(&h->tm, 0, sizeof h->tm);
    h->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        set_err(nullptr, FIC_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        delete h;
        return FIC_E_CUDA;
    }
    if (prop.major != 10) {
        set_err(nullptr, FIC_E_CUDA, "device %d is sm_%d%d; libfic_b200 is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
        delete h;
        return FIC_E_CUDA;
    }
    h->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        set_err(nullptr, FIC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        delete h;
        return FIC_E_CUDA;
    }
    h->stream = h->own_stream;
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&h->ev[i]);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&h->h_acc, 64 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_err(nullptr, FIC_E_CUDA, "creating events / pinned scratch: %s", cudaGetErrorString(e));
        fic_destroy(h);
        return FIC_E_CUDA;
    }
    *out = h;
    return FIC_OK;
}

void fic_destroy(fic_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    Work &w = h->w;
    void *ptrs[] = {w.argb, w.src, w.dec, w.dsum, w.dsq, w.rsum, w.best, w.info, w.q, w.opA, w.opB,
                    w.img, w.dec2, w.dcode, w.perr, w.acc, w.avgf, w.dec3};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 8; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->h_acc) cudaFreeHost(h->h_acc);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

const char *fic_last_error(const fic_handle *h) { return h ? h->err : g_create_err; }

int fic_set_option(fic_handle *h, int option, int value)
{
    if (!h) return FIC_E_ARG;
    if (option == FIC_OPT_ENGINE && value >= FIC_ENGINE_AUTO && value <= FIC_ENGINE_UMMA) {
        h->engine_opt = value;
        return FIC_OK;
    }
    if (option == FIC_OPT_UMMA_KIND && value >= FIC_UMMA_KIND_AUTO && value <= FIC_UMMA_KIND_F16) {
        h->umma_kind = value;
        return FIC_OK;
    }
    return set_err(h, FIC_E_ARG, "unknown option %d / value %d", option, value);
}

int fic_get_option(fic_handle *h, int option, int *value)
{
    if (!h || !value) return FIC_E_ARG;
    if (option == FIC_OPT_ENGINE) { *value = h->engine_opt; return FIC_OK; }
    if (option == FIC_OPT_UMMA_KIND) { *value = h->umma_kind; return FIC_OK; }
    if (option == FIC_OPT_F16_EXACT) {
        if (h->f16_state == 0) {
            CU(cudaSetDevice(h->device));
            const char *why = nullptr;
            int ok = umma_f16_selftest(h->num_sms, h->stream, &why);
            if (ok < 0) return set_err(h, FIC_E_CUDA, "kind::f16 self-test failed to run: %s", why ? why : "?");
            h->f16_state = ok ? 1 : -1;
        }
        *value = h->f16_state > 0;
        return FIC_OK;
    }
    return set_err(h, FIC_E_ARG, "unknown option %d", option);
}

int fic_set_stream(fic_handle *h, void *cuda_stream)
{
    if (!h) return FIC_E_ARG;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return FIC_OK;
}

static void collect_timings(fic_handle *h, bool with_copies);

int fic_get_timings(const fic_handle *h, fic_timings *out)
{
    if (!h || !out) return FIC_E_ARG;
    fic_handle *hm = const_cast<fic_handle *>(h);
    if (hm->tm_pending_dev) {  // asynchronous device entry: events are final once the stream has drained
        if (cudaEventQuery(hm->ev[4]) == cudaSuccess) {
            collect_timings(hm, false);
            hm->tm_pending_dev = false;
        }
    }
    *out = h->tm;
    return FIC_OK;
}

int fic_geometry(int W, int H, int B, int wk, int64_t *n_ranges, int64_t *n_domains)
{
    Geom g;
    int rc = make_geom(W, H, B, wk, 0, &g, nullptr);
    if (rc) return rc;
    if (n_ranges) *n_ranges = g.NR;
    if (n_domains) *n_domains = g.ND;
    return FIC_OK;
}

int fic_sync(fic_handle *h)
{
    if (!h) return FIC_E_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return FIC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------

// Pool build + search + solve on device-resident planes.  Records events 1..4.
static int encode_on_device(fic_handle *h, const Geom &g, const uint8_t *d_src, int64_t j0, int64_t j1,
                            float *d_info, int32_t *d_q)
{
    Work &w = h->w;
    cudaStream_t s = h->stream;
    int engine = FIC_ENGINE_DIRECT;
    if (h->engine_opt == FIC_ENGINE_UMMA) {
        if (!umma_applicable(g))
            return set_err(h, FIC_E_ARG, "tcgen05 search needs widthKernel == domain blocks per width == per height (and, for RGB, blockgroesse 4 or 8 without isometries)");
        engine = FIC_ENGINE_UMMA;
    } else if (h->engine_opt == FIC_ENGINE_AUTO && umma_applicable(g) && (j1 - j0) * g.ND >= (int64_t)1 << 22) {
        engine = FIC_ENGINE_UMMA;
    }
    ENSURE(w.dec, S_DEC, (size_t)g.C * g.sw * g.sh);
    ENSURE(w.dsum, S_DSUM, sizeof(int32_t) * g.C * g.ND);
    ENSURE(w.dsq, S_DSQ, sizeof(int32_t) * g.C * g.ND);
    ENSURE(w.rsum, S_RSUM, sizeof(int32_t) * g.C * g.NR);
    ENSURE(w.best, S_BEST, sizeof(int32_t) * g.NR);
    if (g.C == 3) ENSURE(w.dec3, S_DEC3, sizeof(uint16_t) * (size_t)g.sw * g.sh);  // R + G + B of the decimated planes
    int kind = h->umma_kind;
    if (engine == FIC_ENGINE_UMMA) {
        // The first kind::f16 search of a handle verifies, once, that this device's f16 tensor path
        // accumulates the integer covariances exactly; a device that does not runs kind::i8 instead.
        // RGB has a kind::f16 tensor path only; without an exact f16 path it stays on the CUDA-core kernel.
        const bool rgb = g.C == 3;
        const bool wants_f16 = rgb || (g.B != 16 && (kind == FIC_UMMA_KIND_F16 || (kind == FIC_UMMA_KIND_AUTO && umma_default_kind(g) == FIC_UMMA_KIND_F16)));
        if (wants_f16 && h->f16_state == 0) {
            const char *why = nullptr;
            int ok = umma_f16_selftest(h->num_sms, s, &why);
            if (ok < 0) return set_err(h, FIC_E_CUDA, "kind::f16 self-test failed to run: %s", why ? why : "?");
            h->f16_state = ok ? 1 : -1;
        }
        if (wants_f16 && h->f16_state < 0) {
            if (rgb) engine = FIC_ENGINE_DIRECT;
            else kind = FIC_UMMA_KIND_I8;
        }
    }
    if (engine == FIC_ENGINE_UMMA) {
        ENSURE(w.opA, S_OPA, umma_opA_bytes(g, j0, j1, h->num_sms, kind));
        ENSURE(w.opB, S_OPB, umma_opB_bytes(g, kind));
    }
    Work call = w;  // per-call view: the source planes may belong to the caller
    call.src = const_cast<uint8_t *>(d_src);
    int launches = 0;
    CU(cudaEventRecord(h->ev[1], s));
    launches += launch_decimate(d_src, w.dec, g, s);
    launches += launch_domain_stats(w.dec, w.dsum, w.dsq, g, s);
    launches += launch_range_stats(d_src, w.rsum, g, s);
    CU(cudaEventRecord(h->ev[2], s));
    if (engine == FIC_ENGINE_UMMA) {
        const char *why = nullptr;
        int n = launch_search_umma(call, g, j0, j1, h->num_sms, s, &why, kind, h->ev[6], h->ev[7]);
        if (n < 0) return set_err(h, FIC_E_CUDA, "tcgen05 search launch failed: %s", why ? why : "?");
        launches += n;
    } else {
        if (g.C == 3) launches += launch_sum_planes(w.dec, w.dec3, g, s);
        CU(cudaEventRecord(h->ev[6], s));
        launches += launch_search_direct(call, g, j0, j1, s);
        CU(cudaEventRecord(h->ev[7], s));
    }
    CU(cudaEventRecord(h->ev[3], s));
    launches += launch_solve(call, g, j0, j1, d_info, d_q, s);
    CU(cudaEventRecord(h->ev[4], s));
    CU(cudaGetLastError());
    h->tm.engine = engine;
    h->tm.launches += launches;
    h->tm.search_evals = (double)(j1 - j0) * (double)g.wk * (double)g.wk * (double)g.n_iso;
    return FIC_OK;
}

static void collect_timings(fic_handle *h, bool with_copies)
{
    float ms = 0;
    if (with_copies && cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->tm.h2d_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->tm.pool_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->tm.search_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]) == cudaSuccess) h->tm.kernel_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess) h->tm.solve_ms = ms;
    if (with_copies && cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]) == cudaSuccess) h->tm.d2h_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[with_copies ? 0 : 1], h->ev[with_copies ? 5 : 4]) == cudaSuccess)
        h->tm.total_ms = ms;
}

static int encode_host(fic_handle *h, int is_rgb, const int32_t *argb, int W, int H, int B, int wk, int64_t j0,
                       int64_t j1, float *info, int32_t *qcodes)
{
    if (!h) return FIC_E_ARG;
    if (!argb || (!info && !qcodes)) return set_err(h, FIC_E_ARG, "argb and at least one of info/qcodes must be non-NULL");
    Geom g;
    const char *why = "";
    int rc = make_geom(W, H, B, wk, is_rgb, &g, &why);
    if (rc) return set_err(h, rc, "%s", why);
    if (j0 < 0 || j1 > g.NR || j0 > j1) return set_err(h, FIC_E_ARG, "range interval [%lld,%lld) outside [0,%lld]", (long long)j0, (long long)j1, (long long)g.NR);
    CU(cudaSetDevice(h->device));
    memset(&h->tm, 0, sizeof h->tm);
    Work &w = h->w;
    cudaStream_t s = h->stream;
    int S = code_stride(g);
    ENSURE(w.argb, S_ARGB, sizeof(int32_t) * (size_t)W * H);
    ENSURE(w.src, S_SRC, (size_t)g.C * W * H);
    ENSURE(w.info, S_INFO, sizeof(float) * g.NR * S);
    ENSURE(w.q, S_Q, sizeof(int32_t) * g.NR * S);
    CU(cudaEventRecord(h->ev[0], s));
    CU(cudaMemcpyAsync(w.argb, argb, sizeof(int32_t) * (size_t)W * H, cudaMemcpyHostToDevice, s));
    h->tm.launches += launch_unpack(w.argb, w.src, W, H, g.C, s);
    rc = encode_on_device(h, g, w.src, j0, j1, w.info, w.q);
    if (rc) return rc;
    if (j1 > j0) {
        if (info) CU(cudaMemcpyAsync(info + j0 * S, w.info + j0 * S, sizeof(float) * (j1 - j0) * S, cudaMemcpyDeviceToHost, s));
        if (qcodes) CU(cudaMemcpyAsync(qcodes + j0 * S, w.q + j0 * S, sizeof(int32_t) * (j1 - j0) * S, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaEventRecord(h->ev[5], s));
    CU(cudaStreamSynchronize(s));
    collect_timings(h, true);
    return FIC_OK;
}

extern "C" {

int fic_encode_grey(fic_handle *h, const int32_t *argb, int W, int H, int B, int wk, int64_t range_begin,
                    int64_t range_end, float *info, int32_t *qcodes)
{
    return encode_host(h, 0, argb, W, H, B, wk, range_begin, range_end, info, qcodes);
}

int fic_encode_rgb(fic_handle *h, const int32_t *argb, int W, int H, int B, int wk, int64_t range_begin,
                   int64_t range_end, float *info, int32_t *qcodes)
{
    return encode_host(h, 1, argb, W, H, B, wk, range_begin, range_end, info, qcodes);
}

int fic_encode_grey_iso(fic_handle *h, const int32_t *argb, int W, int H, int B, int wk, int64_t range_begin,
                        int64_t range_end, float *info, int32_t *qcodes)
{
    return encode_host(h, FIC_MODE_GREY_ISO, argb, W, H, B, wk, range_begin, range_end, info, qcodes);
}

int fic_encode_planes_dev(fic_handle *h, const uint8_t *d_planes, int is_rgb, int W, int H, int B, int wk,
                          int64_t range_begin, int64_t range_end, float *d_info, int32_t *d_qcodes)
{
    if (!h) return FIC_E_ARG;
    if (!d_planes || (!d_info && !d_qcodes)) return set_err(h, FIC_E_ARG, "d_planes and at least one output must be non-NULL");
    if ((uintptr_t)d_planes & 15) return set_err(h, FIC_E_ARG, "d_planes must be 16-byte aligned (the kernels read it with vector loads)");
    Geom g;
    const char *why = "";
    int rc = make_geom(W, H, B, wk, is_rgb, &g, &why);
    if (rc) return set_err(h, rc, "%s", why);
    if (range_begin < 0 || range_end > g.NR || range_begin > range_end) return set_err(h, FIC_E_ARG, "range interval outside [0,NR]");
    CU(cudaSetDevice(h->device));
    // Timings of the previous asynchronous call are final once its events completed.
    memset(&h->tm, 0, sizeof h->tm);
    rc = encode_on_device(h, g, d_planes, range_begin, range_end, d_info, d_qcodes);
    h->tm_pending_dev = rc == FIC_OK;
    return rc;
}

int fic_pin_host_buffer(fic_handle *h, void *ptr, size_t bytes)
{
    if (!h) return FIC_E_ARG;
    if (!ptr || bytes == 0) return set_err(h, FIC_E_ARG, "fic_pin_host_buffer: empty buffer");
    CU(cudaSetDevice(h->device));
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(h, e == cudaErrorMemoryAllocation ? FIC_E_NOMEM : FIC_E_CUDA, "cudaHostRegister(%zu bytes): %s", bytes,
                       cudaGetErrorString(e));
    }
    return FIC_OK;
}

int fic_unpin_host_buffer(fic_handle *h, void *ptr)
{
    if (!h) return FIC_E_ARG;
    if (!ptr) return set_err(h, FIC_E_ARG, "fic_unpin_host_buffer: NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));  // no copy of this handle may still be reading the buffer
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(h, FIC_E_ARG, "cudaHostUnregister: %s", cudaGetErrorString(e));
    }
    return FIC_OK;
}

int fic_measure_mma_peak(fic_handle *h, int kind, int n_cols, double *tops)
{
    if (!h || !tops || (kind != FIC_UMMA_KIND_I8 && kind != FIC_UMMA_KIND_F16) || (n_cols != 128 && n_cols != 256))
        return FIC_E_ARG;
    CU(cudaSetDevice(h->device));
    const char *why = nullptr;
    double v = measure_mma_peak(h->num_sms, h->stream, 3, kind == FIC_UMMA_KIND_F16, n_cols, &why);
    if (v < 0) return set_err(h, FIC_E_CUDA, "tensor peak measurement failed: %s", why ? why : "?");
    *tops = v;
    return FIC_OK;
}

int fic_measure_int8_peak(fic_handle *h, double *tops) { return fic_measure_mma_peak(h, FIC_UMMA_KIND_I8, 256, tops); }

int fic_build_pool(fic_handle *h, const int32_t *argb, int is_rgb, int W, int H, int B, uint8_t *decimated,
                   int32_t *dom_sum, int32_t *dom_sumsq)
{
    if (!h || !argb) return FIC_E_ARG;
    Geom g;
    const char *why = "";
    int rc = make_geom(W, H, B, 1, is_rgb, &g, &why);
    if (rc) return set_err(h, rc, "%s", why);
    CU(cudaSetDevice(h->device));
    Work &w = h->w;
    cudaStream_t s = h->stream;
    ENSURE(w.argb, S_ARGB, sizeof(int32_t) * (size_t)W * H);
    ENSURE(w.src, S_SRC, (size_t)g.C * W * H);
    ENSURE(w.dec, S_DEC, (size_t)g.C * g.sw * g.sh);
    ENSURE(w.dsum, S_DSUM, sizeof(int32_t) * g.C * g.ND);
    ENSURE(w.dsq, S_DSQ, sizeof(int32_t) * g.C * g.ND);
    CU(cudaMemcpyAsync(w.argb, argb, sizeof(int32_t) * (size_t)W * H, cudaMemcpyHostToDevice, s));
    launch_unpack(w.argb, w.src, W, H, g.C, s);
    launch_decimate(w.src, w.dec, g, s);
    launch_domain_stats(w.dec, w.dsum, w.dsq, g, s);
    CU(cudaGetLastError());
    if (decimated) CU(cudaMemcpyAsync(decimated, w.dec, (size_t)g.C * g.sw * g.sh, cudaMemcpyDeviceToHost, s));
    if (dom_sum) CU(cudaMemcpyAsync(dom_sum, w.dsum, sizeof(int32_t) * g.C * g.ND, cudaMemcpyDeviceToHost, s));
    if (dom_sumsq) CU(cudaMemcpyAsync(dom_sumsq, w.dsq, sizeof(int32_t) * g.C * g.ND, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return FIC_OK;
}

// ------------------------------------------------------------------------------------
// decode / collage
// ------------------------------------------------------------------------------------

int fic_decode(fic_handle *h, int is_rgb, int W, int H, int B, int wk, const int32_t *qcodes, int max_iters,
               int32_t *argb_out, float *avg_error, int *iterations)
{
    if (!h) return FIC_E_ARG;
    if (!qcodes || !argb_out) return set_err(h, FIC_E_ARG, "qcodes and argb_out must be non-NULL");
    if (max_iters < 1) return set_err(h, FIC_E_ARG, "max_iters must be >= 1");
    Geom g;
    const char *why = "";
    int rc = make_geom(W, H, B, wk, is_rgb, &g, &why);
    if (rc) return set_err(h, rc, "%s", why);
    CU(cudaSetDevice(h->device));
    memset(&h->tm, 0, sizeof h->tm);
    Work &w = h->w;
    cudaStream_t s = h->stream;
    int S = code_stride(g);
    size_t plane = (size_t)W * H;
    ENSURE(w.q, S_Q, sizeof(int32_t) * g.NR * S);
    ENSURE(w.dcode, S_DCODE, sizeof(float) * g.NR * (S + 1));
    ENSURE(w.img, S_IMG, g.C * plane);
    ENSURE(w.dec, S_DEC, (size_t)g.C * g.sw * g.sh);
    ENSURE(w.dec2, S_DEC2, (size_t)g.C * g.sw * g.sh);
    ENSURE(w.acc, S_ACC, 64 * sizeof(unsigned long long));
    ENSURE(w.argb, S_ARGB, sizeof(int32_t) * plane);
    if (!w.avgf) CU(cudaMalloc((void **)&w.avgf, 256));
    CU(cudaEventRecord(h->ev[0], s));
    CU(cudaMemcpyAsync(w.q, qcodes, sizeof(int32_t) * g.NR * S, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(w.acc, 0, 64 * sizeof(unsigned long long), s));
    int launches = 0;
    int32_t *doff = (int32_t *)(w.dcode + g.NR * S);
    launches += launch_dequant(w.q, w.dcode, doff, g, 0, nullptr, w.acc, s);
    launches += launch_fill(w.img, g.C * plane, 128, s);  // FC:360, FC:1142-1148
    launches += launch_decimate(w.img, w.dec, g, s);
    CU(cudaMemcpyAsync(h->h_acc, w.acc, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (h->h_acc[1]) return set_err(h, FIC_E_STREAM, "a code indexes outside the domain pool (the reference would throw ArrayIndexOutOfBounds)");

    // avgError bookkeeping (FC:407-417).  The reference adds every squared pixel change
    // to a binary32 running sum, divides by W*H after the sweep and stops below 1.  The
    // sum is an exact integer while it stays below 2^24, so with W*H <= 2^24 the integer
    // total S decides: S < 2^24 -> the float sum equals S; otherwise the float sum is
    // >= 2^24 >= W*H and the sweep did not converge (its value is then discarded,
    // FC:416-417).  Sweeps whose float sum may be both inexact and kept (carry-in on the
    // first sweep, the last allowed sweep, images above 2^24 pixels) replay the float
    // accumulation in order on the device (k_serial_avg).
    const bool big = (int64_t)W * H > ((int64_t)1 << 24);
    float avg = avg_error ? *avg_error : 0.0f;
    const float fwh = (float)(W * H);  // FC:413 (float)(width*height)
    uint8_t *dcur = w.dec, *dnext = w.dec2;
    int done = 0;
    for (int it = 0; it < max_iters; it++) {
        bool last = it == max_iters - 1;
        bool serial = big || last || (it == 0 && avg != 0.0f);
        if (serial) ENSURE(w.perr, S_PERR, sizeof(int32_t) * plane);
        CU(cudaMemsetAsync(w.acc, 0, sizeof(unsigned long long), s));
        launches += launch_decode_sweep(dcur, w.img, dnext, w.dcode, doff, g, w.acc, serial ? w.perr : nullptr, s);
        float sum;
        if (serial) {
            CU(cudaMemcpyAsync(w.avgf, &avg, sizeof(float), cudaMemcpyHostToDevice, s));
            launches += launch_serial_avg(w.perr, (int64_t)plane, w.avgf, s);
            CU(cudaMemcpyAsync(h->h_acc + 8, w.avgf, sizeof(float), cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            memcpy(&sum, h->h_acc + 8, sizeof(float));
        } else {
            CU(cudaMemcpyAsync(h->h_acc, w.acc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            unsigned long long Ssum = h->h_acc[0];
            sum = Ssum < (1ull << 24) ? (float)Ssum : 2.0f * fwh;
        }
        done = it + 1;
        avg = sum / fwh;       // FC:413
        if (avg < 1) break;    // FC:414
        if (!last) avg = 0;    // FC:416-417
        uint8_t *t = dcur; dcur = dnext; dnext = t;
    }
    launches += launch_pack_argb(w.img, w.argb, W, H, g.C, s);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(argb_out, w.argb, sizeof(int32_t) * plane, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(h->ev[5], s));
    CU(cudaStreamSynchronize(s));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]) == cudaSuccess) h->tm.total_ms = ms;
    h->tm.launches = launches;
    if (avg_error) *avg_error = avg;
    if (iterations) *iterations = done;
    return FIC_OK;
}

int fic_collage(fic_handle *h, int is_rgb, const int32_t *argb, int W, int H, int B, int wk, float *info,
                int32_t *argb_out)
{
    if (!h) return FIC_E_ARG;
    if (!argb || !info || !argb_out) return set_err(h, FIC_E_ARG, "argb, info and argb_out must be non-NULL");
    Geom g;
    const char *why = "";
    int rc = make_geom(W, H, B, wk, is_rgb, &g, &why);
    if (rc) return set_err(h, rc, "%s", why);
    CU(cudaSetDevice(h->device));
    Work &w = h->w;
    cudaStream_t s = h->stream;
    int S = code_stride(g);
    size_t plane = (size_t)W * H;
    ENSURE(w.argb, S_ARGB, sizeof(int32_t) * plane);
    ENSURE(w.src, S_SRC, g.C * plane);
    ENSURE(w.dec, S_DEC, (size_t)g.C * g.sw * g.sh);
    ENSURE(w.info, S_INFO, sizeof(float) * g.NR * S);
    ENSURE(w.dcode, S_DCODE, sizeof(float) * g.NR * (S + 1));
    ENSURE(w.img, S_IMG, g.C * plane);
    ENSURE(w.acc, S_ACC, 64 * sizeof(unsigned long long));
    CU(cudaMemcpyAsync(w.argb, argb, sizeof(int32_t) * plane, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(w.info, info, sizeof(float) * g.NR * S, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(w.acc, 0, 64 * sizeof(unsigned long long), s));
    launch_unpack(w.argb, w.src, W, H, g.C, s);
    launch_decimate(w.src, w.dec, g, s);                       // FC:275 codebook of the source
    int32_t *doff = (int32_t *)(w.dcode + g.NR * S);
    launch_dequant(nullptr, w.dcode, doff, g, 1, w.info, w.acc, s);  // FC:273 calculateIndices
    launch_fill(w.img, g.C * plane, 0xa0, s);                  // RasterImage.java:19,31
    launch_decode_sweep(w.dec, w.img, nullptr, w.dcode, doff, g, nullptr, nullptr, s);
    launch_pack_argb(w.img, w.argb, W, H, g.C, s);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->h_acc, w.acc, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(argb_out, w.argb, sizeof(int32_t) * plane, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(info, w.dcode, sizeof(float) * g.NR * S, cudaMemcpyDeviceToHost, s));  // FC:888 in-place
    CU(cudaStreamSynchronize(s));
    if (h->h_acc[1]) return set_err(h, FIC_E_STREAM, "a code indexes outside the domain pool");
    return FIC_OK;
}

// ------------------------------------------------------------------------------------
// .run stream helpers (host only; DataOutputStream.writeInt is big endian)
// ------------------------------------------------------------------------------------
static void put_be32(uint8_t *p, int32_t v)
{
    uint32_t u = (uint32_t)v;
    p[0] = (uint8_t)(u >> 24); p[1] = (uint8_t)(u >> 16); p[2] = (uint8_t)(u >> 8); p[3] = (uint8_t)u;
}
static int32_t get_be32(const uint8_t *p)
{
    return (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]);
}

size_t fic_stream_size(int is_rgb, int W, int H, int B)
{
    if (B <= 0 || W <= 0 || H <= 0) return 0;
    const size_t per_range = is_rgb == FIC_MODE_GREY_ISO ? 16 : (is_rgb ? 20 : 12);
    return 20 + per_range * (size_t)(W / B) * (size_t)(H / B);
}

int fic_stream_write(int is_rgb, int W, int H, int B, int wk, const int32_t *qcodes, uint8_t *out, size_t out_bytes)
{
    size_t need = fic_stream_size(is_rgb, W, H, B);
    if (!qcodes || !out || need == 0 || out_bytes < need) return FIC_E_ARG;
    put_be32(out, is_rgb == FIC_MODE_GREY_ISO ? 2 : (is_rgb ? 1 : 0));  // FC:234-238; 2 = isometry extension
    put_be32(out + 4, W);
    put_be32(out + 8, H);
    put_be32(out + 12, B);
    put_be32(out + 16, wk);
    size_t n = (need - 20) / 4;
    for (size_t i = 0; i < n; i++) put_be32(out + 20 + 4 * i, qcodes[i]);
    return FIC_OK;
}

int fic_stream_read_header(const uint8_t *stream, size_t nbytes, int *is_rgb, int *W, int *H, int *B, int *wk,
                           size_t *qcodes_off)
{
    if (!stream || nbytes < 20) return FIC_E_STREAM;
    // FC:548-552: 0 -> grey, anything else -> RGB; this library's isometry extension writes 2
    int rgb = get_be32(stream) == 2 ? FIC_MODE_GREY_ISO : (get_be32(stream) != 0);
    int w = get_be32(stream + 4), hh = get_be32(stream + 8), b = get_be32(stream + 12), k = get_be32(stream + 16);
    Geom g;
    if (make_geom(w, hh, b, k, rgb, &g, nullptr)) return FIC_E_STREAM;
    if (nbytes < fic_stream_size(rgb, w, hh, b)) return FIC_E_STREAM;
    if (is_rgb) *is_rgb = rgb;
    if (W) *W = w;
    if (H) *H = hh;
    if (B) *B = b;
    if (wk) *wk = k;
    if (qcodes_off) *qcodes_off = 20;
    return FIC_OK;
}

int fic_stream_read_codes(const uint8_t *stream, size_t nbytes, int32_t *qcodes)
{
    int rgb, W, H, B, wk;
    size_t off;
    int rc = fic_stream_read_header(stream, nbytes, &rgb, &W, &H, &B, &wk, &off);
    if (rc) return rc;
    if (!qcodes) return FIC_E_ARG;
    size_t n = (fic_stream_size(rgb, W, H, B) - 20) / 4;
    for (size_t i = 0; i < n; i++) qcodes[i] = get_be32(stream + off + 4 * i);
    return FIC_OK;
}

}  // extern "C"
