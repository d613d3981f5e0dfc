// C ABI of liblp2d.so (see include/lp2d.h).  Host-side orchestration only: which
// line kernels run, in which order, over which chunk of the batch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lp2d.h"
#include "dispatch.h"
#include "pipe2_ops.cuh"
#include "colstep.cuh"
#include "xl_line.cuh"


namespace lp2d {
long long g_launches = 0;

// cuTensorMapEncodeTiled, fetched from the driver at run time (the library is not linked against libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeFn>(sym);
        else
            cudaGetLastError();
    }
    return fn;
}
// 2-D map over `planes` stacked n x n planes of `csz`-byte complex samples, seen as FP32 words:
// dim0 = n * csz / 4 words (one grid row), dim1 = planes * n rows; box = w columns x `box_rows` rows
static bool encode_plane_map(CUtensorMap* tm, void* base, int n, long long planes, int csz, int w, int box_rows, bool swizzle64 = false) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n * (cuuint64_t)(csz / 4), (cuuint64_t)planes * (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)n * (cuuint64_t)csz};
    const cuuint32_t box[2] = {(cuuint32_t)(w * csz / 4), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static thread_local std::string t_err;

static int fail(int code, const std::string& msg) {
    t_err = msg;
    return code;
}
int set_error(int code, const std::string& msg) { return fail(code, msg); }
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(LP2D_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// probe[b] = |field[b][idx]|^2: what a z-sweep loop reads every step (tests/test_analytics.py:43-50)
template <typename T>
__global__ void probe_kernel(const cx<T>* __restrict__ f, T* __restrict__ out, long long plane, long long idx, long long batch) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) {
        const cx<T> v = f[b * plane + idx];
        out[b] = v.x * v.x + v.y * v.y;
    }
}

// |field|^2, vectorised two samples per thread-iteration
template <typename T>
__global__ void intensity_kernel(const cx<T>* __restrict__ f, T* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const cx<T> v = f[i];
        out[i] = v.x * v.x + v.y * v.y;
    }
}
}  // namespace lp2d

using namespace lp2d;

struct lp2d_plan {
    int device;
    int n;
    int dtype;
    const SizeOps<float>* ops32;
    const SizeOps<double>* ops64;
    const XlOps<float>* xl32;      // long lines (cluster kernel): lp2d_pass_lines only, no 2-D step
    const XlOps<double>* xl64;
    void* tw;          // device twiddles (cx<T>)
    double* tabs;      // device: a | b | x2, n doubles each
    long long chunk_bytes;   // intermediate kept per kernel pair (L2 residency)
    bool chunk_explicit;     // LP2D_CHUNK_MB was given: no automatic whole-batch passes for large planes
    int h_groups;            // beam slabs one CTA of the H row pass walks over
    int hints;               // use evict-first accesses for streaming data
    int nstreams;            // auxiliary streams the chunks are fanned out over
    cudaStream_t aux[8];
    cudaEvent_t ev_fork;
    cudaEvent_t ev_join[8];
    int num_sms;
    void* ring;                // L2-resident intermediate planes
    long long ring_bytes;
    unsigned int* counters;    // two sets of: work queue head + per-(stage, group) completion counters; launches
    int counters_cap;          // alternate between the sets and zero what the previous launch used
    int counter_set;           // set the next fused launch uses
    int counters_used[2];      // entries the last launch on each set dirtied
    int pipe_g, pipe_lag, pipe_slots;   // tuning overrides (0 = automatic)
    int pipe_hints;            // P2_HINT_* (L2 eviction priorities)
    int pipe_persist;          // launch the fused step kernel with an L2 persisting window over the ring
    int use_tma;               // final column stage of two-stage fused steps through the TMA unit
    int fused_min_pct;         // ticket form only for launches of at least this many percent of the resident CTA slots per stage
    bool fused_prop;           // the stateful propagate uses the fused step kernel (default: as bit 1 of fused_ops, but see plan_create)
    int fused_ops;             // which steps use the fused step kernel: 1 plain fft2, 2 steps that store state
                               // (propagate, lens, mask + fft2), 4 the fresh field -> field chain
    // fused step kernel (pipe2.cuh)
    const P2Ops<float>* p2_32;
    const P2Ops<double>* p2_64;
    // persistent form of the fused step kernel (pipe3.cuh); used instead of the ticket form when present and enabled
    const Q4Ops<float>* q4_32;
    const Q4Ops<double>* q4_64;
    int q4_ctas;               // resident CTAs per SM (1)
    int q4_lag;                // tuning override (0 = automatic)
    unsigned long long* q4_timing;   // tuning aid: phase clocks of the persistent form (LP2D_Q4_TIMING=1), else NULL
    // long columns in two steps (colstep.cuh): column length n = cs_a * cs_b; 0 = the one-kernel column pass
    int cs_a, cs_b;
    const ColStepOps<float>* cs32[2];
    const ColStepOps<double>* cs64[2];
    void* cs_tw[2];            // twiddles of the two short-line plans
    void* cs_tw4;              // exp(-2 pi i q / n), q < n
    void* cs_scratch;          // intermediate of the two steps (allocated on first use)
    size_t cs_scratch_bytes;
    int cs_ctas[2];
    int cs_sms;
    int cs_grid_mult;          // CTAs per launch = SMs x resident CTAs x this (0: one CTA per item)
    int cs_panel_mb;           // column panel whose scratch stays in L2 between the two steps (0: whole planes)
    // one-kernel-per-pass path: H plane of a repeated step (z-sweeps): built when a z comes a second time in a row
    void* hp_fac;              // [n][n/2 + 2], allocated on first use
    double hp_z, hp_seen_z;    // z the plane holds / z of the previous stateful call
    bool hp_valid, hp_seen;
    cudaStream_t hp_stream;    // stream the plane was built on; hp_event: the build; hp_multi: other streams have read it
    cudaEvent_t hp_event;
    bool hp_multi;
    int hp_mode;               // LP2D_HPLANE: 0 never, 1 automatic (default), 2 always
    // the H plane and the column scratch are per-plan buffers: a call on another stream than the previous one that
    // used them first waits for that stream (calls on one stream are ordered anyway)
    cudaStream_t side_stream;
    bool side_used;
    int spectrum_t;            // spectra of this plan are stored TRANSPOSED (lp2d_plan_set_spectrum_layout)
    int q4_colpriv;            // persistent form: warp-private final column stage (swizzled TMA tile + TMA store)
    int q4_ops;                // which steps use the persistent form (bits as fused_ops): measured per size, r2g_sweep*.log
    int p2_ctas_factor, p2_ctas_plain;
    void* p2_tw;               // twiddles of the P2 (N, E) schedule
    void* p2_fac;              // factor plane [n][fl]
    int fac_kind, fac_by_column;   // what p2_fac currently holds (0 = nothing)
    int fac_full;              // ... stored unfolded (lens on a grid whose X^2 is not even around n/2)
    bool x2_even;              // x2[n/2 - k] == x2[n/2 + k] bit for bit (false after a generator shifted X in place)
    bool h_even;               // a[k] == a[n - k] and b[k] == b[n - k] bit for bit (the fold of the H plane)
    double fac_p0, fac_p1;
    cudaEvent_t p2_done;       // the ring, the counters and the factor plane are plan state: fused launches of one
    cudaStream_t p2_stream;    // plan are serialised across streams (event recorded on the old stream at a switch)
    bool p2_used;
};

#define LP2D_CHECK_PLAN(p)                                                      \
    do {                                                                        \
        if ((p) == nullptr) return fail(LP2D_EINVAL, "plan is NULL");           \
    } while (0)
// entry points that need the 2-D kernel set (plans of more than lp2d_max_npoints points hold line kernels only)
#define LP2D_CHECK_2D(p)                                                                                   \
    do {                                                                                                   \
        if ((p) == nullptr) return fail(LP2D_EINVAL, "plan is NULL");                                      \
        if ((p)->ops32 == nullptr && (p)->ops64 == nullptr)                                                \
            return fail(LP2D_EUNSUPPORTED, "npoints above lp2d_max_npoints: only lp2d_pass_lines (the slab-decomposed grid) is available"); \
    } while (0)

extern "C" {

int lp2d_version(void) { return LP2D_VERSION; }
const char* lp2d_last_error(void) { return t_err.c_str(); }
long long lp2d_launch_count(void) { return g_launches; }
int lp2d_min_npoints(void) { return 16; }
int lp2d_max_npoints(int dtype) { (void)dtype; return 8192; }
int lp2d_max_line_points(int dtype) { (void)dtype; return 32768; }

int lp2d_grid_tables(int n, double area_size, double wl, double* X, double* Kx, double* a,
                     double* b, double* x2) {
    if (n <= 0 || !(area_size > 0) || !(wl > 0)) return fail(LP2D_EINVAL, "grid_tables: bad n/area_size/wl");
    const double dL = area_size / (double)n;              // beam2d.py:299
    const double k0 = 1.0 / wl;                           // :304
    const double k0sq = k0 * k0;
    const double val = 1.0 / ((double)n * dL);            // numpy fftfreq: 1.0/(n*d)
    const int first = -((n + 1) / 2);                     // Python: -n // 2
    const int npos = (n - 1) / 2 + 1;                     // fftfreq: 0..(n-1)//2 then -(n//2)..-1
    for (int k = 0; k < n; ++k) {
        const double x = (double)(first + k) * dL;        // :300-301
        const int idx = k < npos ? k : k - n;
        const double kx = (double)idx * val;              // :305
        if (X) X[k] = x;
        if (x2) x2[k] = x * x;
        if (Kx) Kx[k] = kx;
        if (a) a[k] = k0sq - kx * kx;                     // :307-308, left-to-right
        if (b) b[k] = kx * kx;
    }
    return LP2D_OK;
}

int lp2d_plan_create(lp2d_plan** out, int n, int dtype, const double* host_a, const double* host_b,
                     const double* host_x2) {
    if (out == nullptr) return fail(LP2D_EINVAL, "plan_create: out is NULL");
    *out = nullptr;
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return fail(LP2D_EINVAL, "plan_create: dtype must be LP2D_C64 or LP2D_C128");
    if (host_a == nullptr || host_b == nullptr || host_x2 == nullptr)
        return fail(LP2D_EINVAL, "plan_create: grid tables must not be NULL");
    const char* venv = getenv("LP2D_VARIANT");
    const int variant = venv ? atoi(venv) : 0;
    const SizeOps<float>* o32 = dtype == LP2D_C64 ? find_ops_f32(n, variant) : nullptr;
    const SizeOps<double>* o64 = dtype == LP2D_C128 ? find_ops_f64(n, variant) : nullptr;
    const char* xenv = getenv("LP2D_XL_VARIANT");
    const int xvar = xenv ? atoi(xenv) : 0;
    const XlOps<float>* x32 = (dtype == LP2D_C64 && o32 == nullptr) ? find_xl_f32(n, xvar) : nullptr;
    const XlOps<double>* x64 = (dtype == LP2D_C128 && o64 == nullptr) ? find_xl_f64(n, xvar) : nullptr;
    if (o32 == nullptr && o64 == nullptr && x32 == nullptr && x64 == nullptr) {
        char buf[160];
        snprintf(buf, sizeof buf, "plan_create: npoints=%d not supported for %s (powers of two %d..%d)", n,
                 dtype == LP2D_C64 ? "complex64" : "complex128", lp2d_min_npoints(), lp2d_max_line_points(dtype));
        return fail(LP2D_EUNSUPPORTED, buf);
    }
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if ((e = (o32 ? o32->configure() : o64 ? o64->configure() : x32 ? x32->configure() : x64->configure())) != cudaSuccess)
        return cuda_fail(e, "cudaFuncSetAttribute");

    lp2d_plan* p = new lp2d_plan();
    p->device = dev; p->n = n; p->dtype = dtype; p->ops32 = o32; p->ops64 = o64; p->xl32 = x32; p->xl64 = x64;
    p->tw = nullptr; p->tabs = nullptr;
    const char* env = getenv("LP2D_CHUNK_MB");
    p->chunk_bytes = (env ? atoll(env) : 32) * (1ll << 20);
    p->chunk_explicit = env != nullptr;
    if (p->chunk_bytes <= 0) p->chunk_bytes = 1ll << 62;
    env = getenv("LP2D_HGROUPS");
    p->h_groups = env ? atoi(env) : 8;
    env = getenv("LP2D_HINTS");
    p->hints = env ? atoi(env) : 1;
    env = getenv("LP2D_STREAMS");
    p->nstreams = env ? atoi(env) : 4;
    if (p->nstreams < 1) p->nstreams = 1;
    if (p->nstreams > 8) p->nstreams = 8;
    if (p->nstreams > 1) {
        for (int i = 0; i < p->nstreams; ++i) {
            if ((e = cudaStreamCreateWithFlags(&p->aux[i], cudaStreamNonBlocking)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) {
                delete p;
                return cuda_fail(e, "aux stream/event create");
            }
        }
        if ((e = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming)) != cudaSuccess) { delete p; return cuda_fail(e, "event create"); }
    }

    std::vector<double> w;
    if (o32) o32->twiddles(w); else if (o64) o64->twiddles(w); else if (x32) x32->twiddles(w); else x64->twiddles(w);
    const size_t ntw = w.size() / 2;
    const size_t tw_bytes = (ntw ? ntw : 1) * (dtype == LP2D_C64 ? sizeof(float2) : sizeof(double2));
    if ((e = cudaMalloc(&p->tw, tw_bytes)) != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(twiddles)"); }
    if (ntw) {
        if (dtype == LP2D_C64) {
            std::vector<float> wf(w.size());
            for (size_t i = 0; i < w.size(); ++i) wf[i] = (float)w[i];
            e = cudaMemcpy(p->tw, wf.data(), tw_bytes, cudaMemcpyHostToDevice);
        } else {
            e = cudaMemcpy(p->tw, w.data(), tw_bytes, cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) { cudaFree(p->tw); delete p; return cuda_fail(e, "cudaMemcpy(twiddles)"); }
    }
    if ((e = cudaMalloc(&p->tabs, 3 * sizeof(double) * n)) != cudaSuccess) { cudaFree(p->tw); delete p; return cuda_fail(e, "cudaMalloc(tables)"); }
    if ((e = cudaMemcpy(p->tabs, host_a, sizeof(double) * n, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(p->tabs + n, host_b, sizeof(double) * n, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(p->tabs + 2 * n, host_x2, sizeof(double) * n, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cudaFree(p->tw); cudaFree(p->tabs); delete p;
        return cuda_fail(e, "cudaMemcpy(tables)");
    }
    p->ring = nullptr; p->counters = nullptr; p->ring_bytes = 0; p->counters_cap = 0; p->num_sms = 0;
    // the folds of the factor plane are only valid for tables that are even bit for bit: the reference's
    // gaussian_beam(x0, y0 != 0) shifts X in place (profiles.py:69-70), slab plans pass swapped tables
    p->x2_even = true; p->h_even = true; p->fac_full = 0;
    for (int k = 1; k < n / 2; ++k) {
        if (host_x2[n / 2 - k] != host_x2[n / 2 + k]) p->x2_even = false;
        if (host_a[k] != host_a[n - k] || host_b[k] != host_b[n - k]) p->h_even = false;
    }
    p->counter_set = 0; p->counters_used[0] = p->counters_used[1] = 0;
    p->q4_32 = nullptr; p->q4_64 = nullptr; p->q4_ctas = 0; p->q4_lag = 0; p->q4_ops = 0; p->q4_colpriv = 0; p->q4_timing = nullptr; p->spectrum_t = 0;
    p->side_stream = nullptr; p->side_used = false;
    p->hp_stream = nullptr; p->hp_event = nullptr; p->hp_multi = false;
    p->hp_fac = nullptr; p->hp_z = 0; p->hp_seen_z = 0; p->hp_valid = false; p->hp_seen = false;
    env = getenv("LP2D_HPLANE"); p->hp_mode = env ? atoi(env) : 1;
    p->cs_a = p->cs_b = 0; p->cs32[0] = p->cs32[1] = nullptr; p->cs64[0] = p->cs64[1] = nullptr;
    p->cs_tw[0] = p->cs_tw[1] = nullptr; p->cs_tw4 = nullptr; p->cs_scratch = nullptr; p->cs_scratch_bytes = 0;
    p->p2_32 = nullptr; p->p2_64 = nullptr; p->p2_tw = nullptr; p->p2_fac = nullptr; p->fac_kind = 0; p->p2_done = nullptr; p->p2_used = false; p->p2_stream = nullptr;
    // the fused step kernel is the default wherever an instance exists (LP2D_PIPE=0: one kernel per pass)
    env = getenv("LP2D_PIPE");
    if (env == nullptr || atoi(env) == 2) {
        env = getenv("LP2D_PIPE_VARIANT");
        const int pvar = env ? atoi(env) : 0;
        const P2Ops<float>* q32 = dtype == LP2D_C64 ? find_p2_f32(n, pvar) : nullptr;
        const P2Ops<double>* q64 = dtype == LP2D_C128 ? find_p2_f64(n, pvar) : nullptr;
        if (q32 || q64) {
            env = getenv("LP2D_RING_MB");
            const long long ring_bytes = (env ? atoll(env) : 64) * (1ll << 20);
            const size_t csz = dtype == LP2D_C64 ? sizeof(float2) : sizeof(double2);
            std::vector<double> w2;
            if (q32) q32->twiddles(w2); else q64->twiddles(w2);
            const size_t n2 = w2.size() / 2;
            e = q32 ? q32->configure(&p->p2_ctas_factor, &p->p2_ctas_plain) : q64->configure(&p->p2_ctas_factor, &p->p2_ctas_plain);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
            p->counters_cap = 1 + P2_MAX_STAGES * 16384;
            if (e == cudaSuccess) e = cudaMalloc(&p->ring, (size_t)ring_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&p->counters, 2 * sizeof(unsigned int) * p->counters_cap);
            // zeroed once: every launch zeroes what the previous launch of the plan used (t3_kernel)
            if (e == cudaSuccess) e = cudaMemset(p->counters, 0, 2 * sizeof(unsigned int) * p->counters_cap);
            if (e == cudaSuccess) e = cudaMalloc(&p->p2_tw, n2 * csz);
            // room for an unfolded plane (row stride n) when X^2 is not even
            if (e == cudaSuccess) e = cudaMalloc(&p->p2_fac, (size_t)n * (size_t)(p->x2_even ? (q32 ? q32->fl : q64->fl) : n + 2) * csz);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->p2_done, cudaEventDisableTiming);
            if (e == cudaSuccess) {
                if (dtype == LP2D_C64) {
                    std::vector<float> wf(w2.size());
                    for (size_t i = 0; i < w2.size(); ++i) wf[i] = (float)w2[i];
                    e = cudaMemcpy(p->p2_tw, wf.data(), n2 * csz, cudaMemcpyHostToDevice);
                } else {
                    e = cudaMemcpy(p->p2_tw, w2.data(), n2 * csz, cudaMemcpyHostToDevice);
                }
            }
            if (e != cudaSuccess || p->p2_ctas_factor < 1 || p->p2_ctas_plain < 1) {
                cudaFree(p->ring); cudaFree(p->counters); cudaFree(p->p2_tw); cudaFree(p->p2_fac); cudaFree(p->tw); cudaFree(p->tabs); delete p;
                return e != cudaSuccess ? cuda_fail(e, "fused step kernel setup") : fail(LP2D_ECUDA, "fused step kernel does not fit an SM");
            }
            p->p2_32 = q32; p->p2_64 = q64; p->ring_bytes = ring_bytes;
            // The persistent warp-specialised form (pipe3.cuh) exists where an instance with the same (N, E) schedule
            // does (it shares the twiddle table and the factor plane).  LP2D_Q4_OPS: which steps use it (bit 0 plain
            // fft2, bit 1 steps that store state, bit 2 the fresh chain); default = where it measured faster than the
            // ticket form (profiles/r2g_sweep1024.log: plain fft2 at 1024 points 4.92 vs 5.57 us per beam; the steps
            // with a state store 6.9 vs 6.6).  LP2D_PIPE_FORM=4 / 3 forces it everywhere / nowhere.
            {
                env = getenv("LP2D_Q4_VARIANT");
                const int qvar = env ? atoi(env) : 0;
                const Q4Ops<float>* r32 = dtype == LP2D_C64 ? find_q4_f32(n, qvar) : nullptr;
                const Q4Ops<double>* r64 = dtype == LP2D_C128 ? find_q4_f64(n, qvar) : nullptr;
                const int e_p2 = q32 ? q32->e : q64->e;
                if ((r32 && r32->e == e_p2) || (r64 && r64->e == e_p2)) {
                    int ctas = 0;
                    e = r32 ? r32->configure(&ctas) : r64->configure(&ctas);
                    if (e == cudaSuccess && ctas >= 1) { p->q4_32 = r32; p->q4_64 = r64; p->q4_ctas = ctas; }
                    else cudaGetLastError();
                }
                int q4_default = (dtype == LP2D_C64 && n == 1024) ? 1 : 0;
                env = getenv("LP2D_PIPE_FORM");
                if (env != nullptr) q4_default = atoi(env) == 4 ? 7 : 0;
                env = getenv("LP2D_Q4_OPS");
                p->q4_ops = env ? atoi(env) : q4_default;
            }
            env = getenv("LP2D_Q4_LAG"); p->q4_lag = env ? atoi(env) : 0;
            env = getenv("LP2D_Q4_COLPRIV"); p->q4_colpriv = env ? atoi(env) : 0;
            env = getenv("LP2D_Q4_TIMING");
            if (env && atoi(env) != 0 && cudaMalloc(&p->q4_timing, 16 * sizeof(unsigned long long)) == cudaSuccess)
                cudaMemset(p->q4_timing, 0, 16 * sizeof(unsigned long long));
        }
    }
    // Long columns in two steps (colstep.cuh) where the one-kernel column pass works on 32-byte pieces AND the two
    // streaming steps (4 P B instead of 2 P B) still win: 4096 complex128 (398 -> 372 us per beam) and 8192 complex64
    // (996 -> 819); not 2048 complex128 (83 -> 87) or 4096 complex64 (181 -> 184), profiles/r3n_colstep.log.
    // LP2D_COLSTEP=0 / 1 forces it off / on (1: from 1024 points)
    {
        const SizeOps<float>* c32 = p->ops32;
        const SizeOps<double>* c64 = p->ops64;
        const int wbytes = c32 ? c32->w_col * 8 : (c64 ? c64->w_col * 16 : 1 << 20);
        env = getenv("LP2D_COLSTEP");
        const bool want = env ? (atoi(env) != 0 && n >= 1024) : (wbytes < 64 && n * (dtype == LP2D_C64 ? 8 : 16) >= 65536);
        if (want && (c32 != nullptr || c64 != nullptr) && n >= 1024 && n <= 16384) {
            int a = 32;
            while (a * a < n) a *= 2;                       // a * b = n, a <= b <= 2 a: 1024 = 32 x 32 ... 8192 = 64 x 128
            int b = n / a;
            if (a > b) { const int t = a; a = b; b = t; }
            const int len[2] = {a, b};
            bool ok = true;
            const char* ev = getenv("LP2D_COLSTEP_VARIANT");
            const int csvar = ev ? atoi(ev) : 0;
            ev = getenv("LP2D_COLSTEP_GRID"); p->cs_grid_mult = ev ? atoi(ev) : 8;
            for (int i = 0; i < 2 && ok; ++i) {
                if (dtype == LP2D_C64) { p->cs32[i] = find_colstep_f32(len[i], csvar); ok = p->cs32[i] != nullptr && p->cs32[i]->configure(&p->cs_ctas[i]) == cudaSuccess; }
                else { p->cs64[i] = find_colstep_f64(len[i], csvar); ok = p->cs64[i] != nullptr && p->cs64[i]->configure(&p->cs_ctas[i]) == cudaSuccess; }
                if (ok && p->cs_ctas[i] < 1) ok = false;
            }
            const size_t csz = dtype == LP2D_C64 ? sizeof(float2) : sizeof(double2);
            auto upload = [&](const std::vector<double>& w, void** dst) -> bool {
                const size_t cnt = w.size() / 2;
                if (cudaMalloc(dst, (cnt ? cnt : 1) * csz) != cudaSuccess) return false;
                if (cnt == 0) return true;
                if (dtype == LP2D_C64) {
                    std::vector<float> wf(w.size());
                    for (size_t i = 0; i < w.size(); ++i) wf[i] = (float)w[i];
                    return cudaMemcpy(*dst, wf.data(), cnt * csz, cudaMemcpyHostToDevice) == cudaSuccess;
                }
                return cudaMemcpy(*dst, w.data(), cnt * csz, cudaMemcpyHostToDevice) == cudaSuccess;
            };
            for (int i = 0; i < 2 && ok; ++i) {
                std::vector<double> w;
                if (dtype == LP2D_C64) p->cs32[i]->twiddles(w); else p->cs64[i]->twiddles(w);
                ok = upload(w, &p->cs_tw[i]);
            }
            if (ok) {
                std::vector<double> w(2 * (size_t)n);
                const long double two_pi = 6.283185307179586476925286766559005768L;
                for (int q = 0; q < n; ++q) {
                    const long double ang = -two_pi * (long double)q / (long double)n;
                    w[2 * (size_t)q] = (double)cosl(ang);
                    w[2 * (size_t)q + 1] = (double)sinl(ang);
                }
                ok = upload(w, &p->cs_tw4);
            }
            int dev = 0;
            p->cs_sms = 148;
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&p->cs_sms, cudaDevAttrMultiProcessorCount, dev);
            env = getenv("LP2D_COLSTEP_PANEL_MB"); p->cs_panel_mb = env ? atoi(env) : 0;       // (panels measured slower: launches of a few microseconds, profiles/r3o_colstep_panels.log)
            if (ok) { p->cs_a = a; p->cs_b = b; }
            else { cudaGetLastError(); p->cs_a = p->cs_b = 0; }
        }
    }
    env = getenv("LP2D_PIPE_G");     p->pipe_g = env ? atoi(env) : 0;
    env = getenv("LP2D_PIPE_LAG");   p->pipe_lag = env ? atoi(env) : 0;
    env = getenv("LP2D_PIPE_SLOTS"); p->pipe_slots = env ? atoi(env) : 0;
    // bit 1: evict-last ring stores; bit 2: speculative L2 prefetch of the first-stage rows (-1: for two-stage steps only,
    // where it is worth 2 %; it costs 5 % in the three-stage chain at 512, profiles/r1d_pipe_sweep.log)
    env = getenv("LP2D_PIPE_HINTS"); p->pipe_hints = env ? atoi(env) : -1;
    // measured (profiles/r1c_size_sweep.log): the fused kernel wins on the steps that read AND write state
    // (3*P*B algorithmic bytes); plain fft2 and the three-stage fresh chain stay on the per-pass kernels
    // and, for the small planes, on more of them (profiles/r1c_size_sweep.log, r1c_group_sweep.log):
    //   complex64 256: + plain fft2; complex64 512 and complex128 256: every step
    int ops_default = 2;                            // (also complex64 2048: the steps with a state store only)
    // round 2 (profiles/r2c_sweep_final.log): at 1024 complex64 every step -- plain fft2 on the persistent form
    // (4.83 vs 7.0 us per beam per-pass), the fresh chain on the ticket form (10.8 vs 13.7)
    if (dtype == LP2D_C64 && n == 1024) ops_default = 7;
    if (dtype == LP2D_C64 && n == 256) ops_default = 3;
    if ((dtype == LP2D_C64 && n == 512) || (dtype == LP2D_C128 && n == 256)) ops_default = 7;
    // second half of round 2 (profiles/r3v_pipe_vs_perpass*.log): with whole-batch passes and the H plane of a repeated
    // step the one-kernel-per-pass path beats the fused kernel at 2048 complex64 (propagate 37.3 vs 45.9 us per beam,
    // lens 42.3 vs 45.9) and, for the stateful propagate only, at 512 complex128 (3.53 vs 3.97; lens 3.77 vs 3.50)
    if (dtype == LP2D_C64 && n == 2048) ops_default = 0;
    env = getenv("LP2D_FUSED_OPS"); p->fused_ops = env ? atoi(env) : ops_default;
    p->fused_prop = (p->fused_ops & 2) != 0 && !(dtype == LP2D_C128 && n == 512 && env == nullptr);
    env = getenv("LP2D_FUSED_PROP"); if (env) p->fused_prop = atoi(env) != 0;
    env = getenv("LP2D_FUSED_MIN"); p->fused_min_pct = env ? atoi(env) : 75;
    // opt-in: correct, but 6 % slower than per-thread accesses in the one-tile-per-CTA form
    env = getenv("LP2D_TMA"); p->use_tma = env ? atoi(env) : 0;
    env = getenv("LP2D_PIPE_PERSIST"); p->pipe_persist = env ? atoi(env) : 0;
    if (p->pipe_persist && p->ring != nullptr) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
            size_t want = (size_t)p->ring_bytes;
            if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); p->pipe_persist = 0; }
            if (getenv("LP2D_VERBOSE")) fprintf(stderr, "[lp2d] L2 %d MB, persisting max %d MB, set %zu MB\n", prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, want >> 20);
        }
    }
    // the tables, twiddles and counters were written on the legacy default stream: user streams (non-blocking
    // ones too) must find them complete
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) { lp2d_plan_destroy(p); return cuda_fail(e, "plan_create: synchronise"); }
    *out = p;
    return LP2D_OK;
}

int lp2d_plan_step_kernel(lp2d_plan* p, char* name, int name_len) {
    if (p == nullptr) return 0;
    if (p->ops32 == nullptr && p->ops64 == nullptr) {
        if (name != nullptr && name_len > 0)
            snprintf(name, (size_t)name_len, "xl_line_kernel<%s, N=%d, E=%d, CL=%d> cluster line pass", p->dtype == LP2D_C64 ? "float" : "double",
                     p->n, p->xl32 ? p->xl32->e : p->xl64->e, p->xl32 ? p->xl32->cl : p->xl64->cl);
        return 0;
    }
    const bool fused = (p->p2_32 != nullptr || p->p2_64 != nullptr) && p->fused_prop;
    if (name != nullptr && name_len > 0) {
        if (fused) {
            const bool q4 = (p->q4_32 != nullptr || p->q4_64 != nullptr) && (p->q4_ops & 2);
            const int e = p->p2_32 ? p->p2_32->e : p->p2_64->e;
            const int w = q4 ? (p->q4_32 ? p->q4_32->w : p->q4_64->w) : (p->p2_32 ? p->p2_32->w : p->p2_64->w);
            snprintf(name, (size_t)name_len, "%s<%s, N=%d, E=%d, W=%d> fused step (rows -> L2 ring -> columns, one launch%s)",
                     q4 ? "q4_kernel" : "t3_kernel", p->dtype == LP2D_C64 ? "float" : "double", p->n, e, w,
                     q4 ? ", persistent: TMA / bulk-copy producer + transform warps" : "");
        } else {
            const int e = p->ops32 ? p->ops32->e : p->ops64->e;
            if (p->cs_a != 0)
                snprintf(name, (size_t)name_len, "line_kernel<%s, N=%d, E=%d> row pass + colstep_kernel x 2 (columns as %d x %d)",
                         p->dtype == LP2D_C64 ? "float" : "double", p->n, e, p->cs_a, p->cs_b);
            else
                snprintf(name, (size_t)name_len, "line_kernel<%s, N=%d, E=%d> row pass + column pass", p->dtype == LP2D_C64 ? "float" : "double", p->n, e);
        }
    }
    return fused ? 1 : 0;
}

int lp2d_plan_set_spectrum_layout(lp2d_plan* p, int transposed) {
    if (p == nullptr) return fail(LP2D_EINVAL, "set_spectrum_layout: NULL plan");
    if (transposed) {
        if ((p->p2_32 == nullptr && p->p2_64 == nullptr) || !p->h_even)
            return fail(LP2D_EUNSUPPORTED, "set_spectrum_layout: the transposed layout needs the fused step kernel (no instance for this size / tables)");
        // every two-stage step must fit the ring with its default schedule; run_p2 reports it per call otherwise
    }
    p->spectrum_t = transposed ? 1 : 0;
    return LP2D_OK;
}

int lp2d_plan_get_spectrum_layout(lp2d_plan* p) { return p != nullptr ? p->spectrum_t : 0; }

int lp2d_q4_timing(lp2d_plan* p, unsigned long long* host16, int reset) {
    if (p == nullptr || p->q4_timing == nullptr || host16 == nullptr) return fail(LP2D_EINVAL, "q4_timing: not enabled (LP2D_Q4_TIMING=1)");
    cudaError_t e = cudaMemcpy(host16, p->q4_timing, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && reset) e = cudaMemset(p->q4_timing, 0, 16 * sizeof(unsigned long long));
    return e == cudaSuccess ? LP2D_OK : cuda_fail(e, "q4_timing");
}

int lp2d_plan_destroy(lp2d_plan* p) {
    if (p == nullptr) return LP2D_OK;
    cudaFree(p->q4_timing);
    cudaFree(p->hp_fac);
    if (p->hp_event) cudaEventDestroy(p->hp_event);
    cudaFree(p->cs_tw[0]); cudaFree(p->cs_tw[1]); cudaFree(p->cs_tw4); cudaFree(p->cs_scratch);
    cudaFree(p->tw);
    cudaFree(p->tabs);
    cudaFree(p->ring);
    cudaFree(p->counters);
    cudaFree(p->p2_tw);
    cudaFree(p->p2_fac);
    if (p->p2_done) cudaEventDestroy(p->p2_done);
    if (p->nstreams > 1) {
        for (int i = 0; i < p->nstreams; ++i) { cudaStreamDestroy(p->aux[i]); cudaEventDestroy(p->ev_join[i]); }
        cudaEventDestroy(p->ev_fork);
    }
    delete p;
    return LP2D_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ typed drivers
namespace {

#define LP2D_CU(call, what)                                         \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);        \
    } while (0)

template <typename T> const SizeOps<T>* ops_of(const lp2d_plan* p);
template <> const SizeOps<float>* ops_of<float>(const lp2d_plan* p) { return p->ops32; }
template <> const SizeOps<double>* ops_of<double>(const lp2d_plan* p) { return p->ops64; }

template <typename T> const XlOps<T>* xl_of(const lp2d_plan* p);
template <> const XlOps<float>* xl_of<float>(const lp2d_plan* p) { return p->xl32; }
template <> const XlOps<double>* xl_of<double>(const lp2d_plan* p) { return p->xl64; }

template <typename T> const Q4Ops<T>* q4_of(const lp2d_plan* p);
template <> const Q4Ops<float>* q4_of<float>(const lp2d_plan* p) { return p->q4_32; }
template <> const Q4Ops<double>* q4_of<double>(const lp2d_plan* p) { return p->q4_64; }

template <typename T> const P2Ops<T>* p2_of(const lp2d_plan* p);
template <> const P2Ops<float>* p2_of<float>(const lp2d_plan* p) { return p->p2_32; }
template <> const P2Ops<double>* p2_of<double>(const lp2d_plan* p) { return p->p2_64; }

// Make the plan's factor plane hold (kind, layout, parameters); it is rebuilt only when they change
// (a z-sweep with a constant step reuses it for every call).
template <typename T>
int p2_ensure_factor(lp2d_plan* p, int kind, int by_column, double p0, double p1, cudaStream_t s) {
    const int full = (kind == P2_FAC_LENS && !p->x2_even) ? 1 : 0;
    if (p->fac_kind == kind && p->fac_by_column == by_column && p->fac_full == full && p->fac_p0 == p0 && p->fac_p1 == p1) return LP2D_OK;
    const P2Ops<T>* ops = p2_of<T>(p);
    const int fl = full ? p->n : ops->fl;
    cudaError_t e;
    if (sizeof(T) == 4)
        e = p2_factor_plane_f32(static_cast<float2*>(p->p2_fac), p->tabs, p->n, fl, kind, by_column, full,
                                kind == P2_FAC_H ? p0 : 0.0, kind == P2_FAC_LENS ? p0 : 0.0, kind == P2_FAC_LENS ? p1 : 0.0, s);
    else
        e = p2_factor_plane_f64(static_cast<double2*>(p->p2_fac), p->tabs, p->n, fl, kind, by_column, full,
                                kind == P2_FAC_H ? p0 : 0.0, kind == P2_FAC_LENS ? p0 : 0.0, kind == P2_FAC_LENS ? p1 : 0.0, s);
    if (e != cudaSuccess) return cuda_fail(e, "factor plane launch");
    p->fac_kind = kind; p->fac_by_column = by_column; p->fac_full = full; p->fac_p0 = p0; p->fac_p1 = p1;
    return LP2D_OK;
}

// Run a chain of `nstages` line passes over `batch` beams in one launch of the fused step kernel.
// fac_kind: P2_FAC_NONE / P2_FAC_H (p0 = z) / P2_FAC_LENS (p0 = k0, p1 = 2f); by_column: layout of the
// factor plane (1 for the fused column stage).  Returns -1 when the kernel is not available.
// stream hand-over of the plan's fused-step state (ring, counters, factor plane)
static int p2_claim_stream(lp2d_plan* p, cudaStream_t s) {
    if (p->p2_used && p->p2_stream != s) {
        if (cudaEventRecord(p->p2_done, p->p2_stream) == cudaSuccess) {
            LP2D_CU(cudaStreamWaitEvent(s, p->p2_done, 0), "fused step: wait for the previous stream");
        } else {
            cudaGetLastError();                     // the previous stream no longer exists: wait for the device
            LP2D_CU(cudaDeviceSynchronize(), "fused step: device synchronise");
        }
    }
    p->p2_stream = s; p->p2_used = true;
    return LP2D_OK;
}

// The persistent form (pipe3.cuh).  Returns -1 when it does not apply (no instance, unaligned user planes).
template <typename T>
int run_q4(lp2d_plan* p, const P2Stage<T>* stages, int nstages, int fac_kind, int by_column, double p0, double p1,
           long long batch, cudaStream_t s) {
    const Q4Ops<T>* ops = q4_of<T>(p);
    if (ops == nullptr) return -1;
    if (stages[nstages - 1].moments != nullptr || stages[nstages - 1].probe_out != nullptr) return -1;   // fused sweep by-products: ticket form
    // bulk copies of the user rows need 16-byte aligned planes
    if (!stages[0].in_ring && (reinterpret_cast<uintptr_t>(stages[0].in) & 15) != 0) return -1;
    const long long plane = (long long)p->n * p->n * (long long)sizeof(cx<T>);
    const int tiles = ops->tiles;
    // items a CTA holds: the one in work plus its landing buffers; the lag between the stages of a beam covers
    // every item that can be held anywhere (with a margin), so that dependencies are complete when they are checked
    const int hints = p->pipe_hints >= 0 ? p->pipe_hints : P2_HINT_KEEP_RING;    // (late draw: slower, r2g_sweep1024.log)
    const long long conc = (long long)p->num_sms * p->q4_ctas * (ops->nbuf + ((hints & P2_HINT_LATE_DRAW) ? 0 : 1));
    const long long reach = (5 * conc) / 4;
    int lag = p->q4_lag > 0 ? p->q4_lag : (int)((reach + (long long)nstages * tiles - 1) / ((long long)nstages * tiles));
    if (lag < 1) lag = 1;
    const int slots = nstages * lag;
    if ((long long)slots * plane > p->ring_bytes) return -1;
    int rc = p2_claim_stream(p, s);
    if (rc) return rc;
    if (fac_kind != P2_FAC_NONE) {
        rc = p2_ensure_factor<T>(p, fac_kind, by_column, p0, p1, s);
        if (rc) return rc;
    }
    const long long groups_cap = (p->counters_cap - 1) / nstages;
    for (long long b0 = 0; b0 < batch; b0 += groups_cap) {
        const long long nb = (batch - b0 < groups_cap) ? batch - b0 : groups_cap;
        P2Args<T> a;
        memset(&a, 0, sizeof a);
        const long long off = b0 * (long long)p->n * p->n;
        for (int i = 0; i < nstages; ++i) {
            a.st[i] = stages[i];
            P2Stage<T>& st = a.st[i];
            if (st.in) st.in += off;
            if (st.out) st.out += off;
            if (st.out_pre) st.out_pre += off;
            if (st.out_mid) st.out_mid += off;
            if (st.inten) st.inten += off;
            if (st.mask) st.mask = static_cast<const char*>(st.mask) + b0 * st.mask_beam_stride *
                                   (st.mask_kind == MASK_CX ? sizeof(cx<T>) : st.mask_kind == MASK_REAL ? sizeof(T) : 1);
        }
        a.ring = static_cast<cx<T>*>(p->ring);
        a.tw = static_cast<const cx<T>*>(p->p2_tw);
        a.fac = static_cast<const cx<T>*>(p->p2_fac);
        a.counters = p->counters + (size_t)p->counter_set * p->counters_cap;
        a.zero_counters = p->counters + (size_t)(p->counter_set ^ 1) * p->counters_cap;
        a.zero_count = p->counters_used[p->counter_set ^ 1];
        a.nbeams = nb;
        a.nstages = nstages;
        a.ngroups = (int)nb;
        a.G = 1; a.lag = lag; a.ring_slots = slots;
        a.fac_map = fac_kind == P2_FAC_LENS ? (p->x2_even ? P2_MAP_CENTRE : P2_MAP_FULL) : P2_MAP_FREQ;
        a.fac_stride = (fac_kind == P2_FAC_LENS && !p->x2_even) ? p->n : ops->fl;
        a.hints = hints;
        a.timing = p->q4_timing;
        a.use_tma = 1;
        const int box_rows = p->n < 256 ? p->n : 256;
        // warp-private final column stage (pipe3.cuh): one warp per column, swizzled tile, results leave by a TMA store.
        // Two-stage steps whose last stage only writes the complex plane; a tile row must be 64 bytes, a line one warp.
        const bool col_private = p->q4_colpriv && nstages == 2 && !a.st[1].mid && !a.st[1].out_t && a.st[1].out != nullptr && a.st[1].inten == nullptr &&
                                 ops->w * (int)sizeof(cx<T>) == 64 && p->n / ops->e == 32 && ops->threads - 64 == 32 * ops->w &&
                                 (reinterpret_cast<uintptr_t>(a.st[1].out) & 15) == 0;
        if (col_private) {
            if (!encode_plane_map(&a.tm_out, a.st[1].out, p->n, nb, (int)sizeof(cx<T>), ops->w, box_rows, true))
                return fail(LP2D_ECUDA, "fused step: cuTensorMapEncodeTiled (output) failed");
            a.use_tma = 2;
        }
        if (!encode_plane_map(&a.tm_ring, p->ring, p->n, (long long)slots, (int)sizeof(cx<T>), ops->w, box_rows, col_private))
            return fail(LP2D_ECUDA, "fused step: cuTensorMapEncodeTiled failed");
        const long long total = ((long long)a.ngroups + (long long)(nstages - 1) * lag) * nstages * tiles;
        if (total > 0x7fffffffll) return fail(LP2D_EINVAL, "fused step: batch too large for one launch");
        a.total_items = (int)total;
        const long long real_items = (long long)a.ngroups * nstages * tiles;
        long long grid = (long long)p->num_sms * p->q4_ctas;
        if (grid > real_items) grid = real_items;
        LP2D_CU(ops->launch(a, (unsigned)grid, s), "fused step launch");
        p->counters_used[p->counter_set ^ 1] = 0;
        p->counters_used[p->counter_set] = 1 + nstages * a.ngroups;
        p->counter_set ^= 1;
    }
    return LP2D_OK;
}

template <typename T>
int run_p2(lp2d_plan* p, const P2Stage<T>* stages, int nstages, int fac_kind, int by_column, double p0, double p1,
           long long batch, cudaStream_t s, int op_bit) {
    const P2Ops<T>* ops = p2_of<T>(p);
    if (ops == nullptr) return -1;
    if (fac_kind == P2_FAC_H && !p->h_even) return -1;      // the folded H plane needs even tables: one kernel per pass
    if (p->q4_ops & op_bit) {
        const int rc = run_q4<T>(p, stages, nstages, fac_kind, by_column, p0, p1, batch, s);
        if (rc != -1) return rc;
    }
    const bool need_factor = fac_kind != P2_FAC_NONE;
    const long long plane = (long long)p->n * p->n * (long long)sizeof(cx<T>);
    const int ctas = need_factor ? p->p2_ctas_factor : p->p2_ctas_plain;
    long long conc = (long long)ctas * p->num_sms;
    const int tiles = ops->tiles;
    // beams per work item: two for planes of up to 512 points (halves the per-item overhead; the ring still
    // fits L2), one for 1024 points (the ring would double to 64 MB)
    long long G = p->pipe_g > 0 ? p->pipe_g : (p->n <= 512 ? 2 : 1);
    if (G > batch) G = batch;
    // Launches that do not fill the GPU (a single field, a handful of small beams): the dependent stages of the fused
    // kernel are then a serial chain of CTA lifetimes, and two plain launches are faster (1024^2 single field 17.5 vs
    // 20.6 us, 512^2 x 4 beams 4.0 vs 6.5 us per beam; from 1024^2 x 2 / 512^2 x 16 on the fused kernel wins again:
    // profiles/r3z_small_batches.log).  Threshold: items per stage against 3/4 of the resident CTA slots
    // (LP2D_FUSED_MIN: percent of the slots, default 75; 0: always fused).  The transposed layout exists only here.
    if (!p->spectrum_t && p->fused_min_pct > 0 &&
        ((batch + G - 1) / G) * (long long)tiles * 100 < conc * (long long)p->fused_min_pct)
        return -1;
    // lag (in groups) between consecutive stages of a group: the items in between cover every resident
    // CTA (with a margin), so that what an item depends on was drawn more than a wave earlier
    const long long reach = (5 * conc) / 4;
    // three-stage chains: one more group of slack (measured at 512: 2.47 -> 2.36 us/beam, profiles/r1d_pipe_sweep512.log)
    int lag = p->pipe_lag > 0 ? p->pipe_lag : (int)((reach + (long long)nstages * tiles - 1) / ((long long)nstages * tiles)) + (nstages == 3 ? 1 : 0);
    if (lag < 1) lag = 1;
    // ring slots (groups): the stage that refills a slot must also be `lag` groups behind the last
    // stage of the group that held it
    int slots = nstages * lag;
    if (p->pipe_slots > 0) slots = p->pipe_slots > (nstages - 1) * lag + 1 ? p->pipe_slots : (nstages - 1) * lag + 1;
    while (G > 1 && slots * G * plane > p->ring_bytes) --G;
    if (slots * G * plane > p->ring_bytes) return -1;
    // launches on one stream are ordered anyway; when the caller switches streams, the new stream waits for
    // everything queued so far on the old one (which includes the previous fused launch of this plan)
    if (p->p2_used && p->p2_stream != s) {
        if (cudaEventRecord(p->p2_done, p->p2_stream) == cudaSuccess) {
            LP2D_CU(cudaStreamWaitEvent(s, p->p2_done, 0), "fused step: wait for the previous stream");
        } else {
            cudaGetLastError();                     // the previous stream no longer exists: wait for the device
            LP2D_CU(cudaDeviceSynchronize(), "fused step: device synchronise");
        }
    }
    p->p2_stream = s; p->p2_used = true;
    if (need_factor) {
        const int rc = p2_ensure_factor<T>(p, fac_kind, by_column, p0, p1, s);
        if (rc) return rc;
    }
    const long long groups_cap = (p->counters_cap - 1) / nstages;
    for (long long b0 = 0; b0 < batch; b0 += groups_cap * G) {
        const long long nb = (batch - b0 < groups_cap * G) ? batch - b0 : groups_cap * G;
        P2Args<T> a;
        memset(&a, 0, sizeof a);
        const long long off = b0 * (long long)p->n * p->n;
        for (int i = 0; i < nstages; ++i) {
            a.st[i] = stages[i];
            P2Stage<T>& st = a.st[i];
            if (st.in) st.in += off;
            if (st.out) st.out += off;
            if (st.out_pre) st.out_pre += off;
            if (st.out_mid) st.out_mid += off;
            if (st.inten) st.inten += off;
            if (st.moments) st.moments += b0 * 5;
            if (st.probe_out) st.probe_out += b0;
            if (st.mask) st.mask = static_cast<const char*>(st.mask) + b0 * st.mask_beam_stride *
                                   (st.mask_kind == MASK_CX ? sizeof(cx<T>) : st.mask_kind == MASK_REAL ? sizeof(T) : 1);
        }
        a.ring = static_cast<cx<T>*>(p->ring);
        a.tw = static_cast<const cx<T>*>(p->p2_tw);
        a.fac = static_cast<const cx<T>*>(p->p2_fac);
        a.counters = p->counters + (size_t)p->counter_set * p->counters_cap;
        a.zero_counters = p->counters + (size_t)(p->counter_set ^ 1) * p->counters_cap;
        a.zero_count = p->counters_used[p->counter_set ^ 1];
        a.nbeams = nb;
        a.nstages = nstages;
        a.ngroups = (int)((nb + G - 1) / G);
        a.G = (int)G; a.lag = lag; a.ring_slots = slots;
        a.fac_map = fac_kind == P2_FAC_LENS ? (p->x2_even ? P2_MAP_CENTRE : P2_MAP_FULL) : P2_MAP_FREQ;
        a.fac_stride = (fac_kind == P2_FAC_LENS && !p->x2_even) ? p->n : ops->fl;
        a.hints = p->pipe_hints >= 0 ? p->pipe_hints : (P2_HINT_KEEP_RING | 1 | (nstages == 2 ? P2_HINT_PREFETCH : 0));
        a.use_tma = 0;
        if (p->use_tma && nstages == 2 && a.st[1].in_ring && !a.st[1].mid && !a.st[1].out_t && a.st[1].moments == nullptr && a.st[1].probe_out == nullptr && a.st[1].out != nullptr &&
            (ops->w * (int)sizeof(cx<T>)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.st[1].out) & 15) == 0) {
            const int box_rows = p->n < 256 ? p->n : 256;
            if (encode_plane_map(&a.tm_ring, p->ring, p->n, (long long)slots * G, (int)sizeof(cx<T>), ops->w, box_rows) &&
                encode_plane_map(&a.tm_out, a.st[1].out, p->n, nb, (int)sizeof(cx<T>), ops->w, box_rows))
                a.use_tma = 1;
        }
        const long long total = ((long long)a.ngroups + (long long)(nstages - 1) * lag) * nstages * tiles;
        if (total > 0x7fffffffll) return fail(LP2D_EINVAL, "fused step: batch too large for one launch");
        a.total_items = (int)total;
        const unsigned grid = (unsigned)total;      // one CTA per item
        LP2D_CU(ops->launch(a, grid, need_factor, s, p->ring, p->pipe_persist ? (size_t)(slots * G * plane) : 0), "fused step launch");
        p->counters_used[p->counter_set ^ 1] = 0;
        p->counters_used[p->counter_set] = 1 + nstages * a.ngroups;
        p->counter_set ^= 1;
    }
    return LP2D_OK;
}

template <typename T>
P2Stage<T> p2_stage(int axis, int inverse, T scale) {
    P2Stage<T> st;
    memset(&st, 0, sizeof st);
    st.axis = axis; st.inverse = inverse; st.scale = scale;
    return st;
}

template <typename T>
LineArgs<T> base_args(const lp2d_plan* p) {
    LineArgs<T> a;
    memset(&a, 0, sizeof a);
    a.tw = static_cast<const cx<T>*>(p->tw);
    a.a_tab = p->tabs;
    a.b_tab = p->tabs + p->n;
    a.x2_tab = p->tabs + 2 * p->n;
    a.scale = (T)1;
    a.beam_stride = (long long)p->n * p->n;
    a.groups = 1;
    return a;
}

// beams per kernel pair so that the intermediate plane(s) stay L2 resident
template <typename T>
long long chunk_beams(const lp2d_plan* p, long long batch) {
    const long long plane = (long long)p->n * p->n * (long long)sizeof(cx<T>);
    long long c = p->chunk_bytes / plane;
    // Planes of 16 MB and more (complex128 from 1024 points, complex64 from 2048): a chunk would be one or two beams,
    // i.e. launches of one or two waves fanned out over streams that then compete for L2; the whole batch per pass is
    // 13-24 % faster there (profiles/r3k_sweep_perpass_sizes.log: 4096^2 complex64 239 -> 181 us per beam)
    if ((c < 4 && !p->chunk_explicit) || p->cs_a != 0) c = batch;      // (the two-step column pass owns ONE scratch: one stream)
    if (c < 1) c = 1;
    // keep the row-count of a launch inside int range
    const long long cap = (long long)(0x7fffffff) / p->n;
    if (c > cap) c = cap;
    return c < batch ? c : batch;
}

// Chunks of a batch are independent: fan them out over the plan's auxiliary streams so the
// tail of one chunk's kernel overlaps the head of the next, then join back on the user stream.
struct Fan {
    const lp2d_plan* p;
    cudaStream_t user;
    int ns, k;
    Fan(const lp2d_plan* plan, cudaStream_t u, long long nchunks) : p(plan), user(u), ns(0), k(0) {
        if (nchunks > 1 && plan->nstreams > 1) ns = (int)(nchunks < plan->nstreams ? nchunks : plan->nstreams);
    }
    cudaError_t fork() {
        if (!ns) return cudaSuccess;
        cudaError_t e = cudaEventRecord(p->ev_fork, user);
        for (int i = 0; i < ns && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(p->aux[i], p->ev_fork, 0);
        return e;
    }
    cudaStream_t next() { return ns ? p->aux[k++ % ns] : user; }
    cudaError_t join() {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < ns && e == cudaSuccess; ++i) {
            e = cudaEventRecord(p->ev_join[i], p->aux[i]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(user, p->ev_join[i], 0);
        }
        return e;
    }
};

template <typename T>
int rows(const lp2d_plan* p, int pre, LineArgs<T> a, long long beams, cudaStream_t s) {
    // the transfer function is the same for every beam: let a CTA reuse it over `groups` slabs
    long long groups = (pre == PRE_H) ? p->h_groups : 1;
    if (groups > beams) groups = beams;
    if (groups < 1) groups = 1;
    const long long gbeams = (beams + groups - 1) / groups;
    a.groups = (int)((beams + gbeams - 1) / gbeams);
    a.group_beams = gbeams;
    a.nbeams = beams;
    a.nlines = (int)(gbeams * p->n);
    LP2D_CU(ops_of<T>(p)->row(pre, a, s), "row pass launch");
    return LP2D_OK;
}
static int claim_side_buffers(lp2d_plan* p, cudaStream_t s) {
    if (p->side_used && p->side_stream != s) {
        if (cudaStreamSynchronize(p->side_stream) != cudaSuccess) {
            cudaGetLastError();                         // the previous stream no longer exists: wait for the device
            LP2D_CU(cudaDeviceSynchronize(), "plan buffers: device synchronise");
        }
    }
    p->side_stream = s; p->side_used = true;
    return LP2D_OK;
}

template <typename T> const ColStepOps<T>* cs_of(const lp2d_plan* p, int i);
template <> const ColStepOps<float>* cs_of<float>(const lp2d_plan* p, int i) { return p->cs32[i]; }
template <> const ColStepOps<double>* cs_of<double>(const lp2d_plan* p, int i) { return p->cs64[i]; }

// The column pass in two steps (colstep.cuh): in -> scratch (transforms of length cs_a over rows cs_b apart, twiddle),
// scratch -> out (transforms of length cs_b over consecutive rows, result rows cs_a apart).  in may be out.
template <typename T>
int cols_two_step(lp2d_plan* p, const LineArgs<T>& a, long long beams, cudaStream_t s) {
    const long long n = p->n, plane = n * n;
    { const int rc = claim_side_buffers(p, s); if (rc) return rc; }
    // One beam at a time through a scratch of one plane (default), or one PANEL of columns at a time
    // (LP2D_COLSTEP_PANEL_MB > 0): the scratch of a panel (n rows x wp columns) is then written by step 1 and read by
    // step 2 while it is still in L2, so the pass moves 2 P B through HBM like the one-kernel pass -- but in launches of
    // a few microseconds each, which measured slower at every size (profiles/r3o_colstep_panels.log).
    const long long panel_bytes = (long long)p->cs_panel_mb << 20;
    long long wp = panel_bytes > 0 ? panel_bytes / (n * (long long)sizeof(cx<T>)) : n;
    const int wc = cs_of<T>(p, 0)->wc;
    if (wp > n) wp = n;
    if (wp < wc) wp = wc;
    while (wp & (wp - 1)) wp &= wp - 1;                      // power of two: divides n, multiple of the tile width
    const size_t need = (size_t)n * wp * sizeof(cx<T>);
    if (p->cs_scratch_bytes < need) {
        if (p->cs_scratch != nullptr) { LP2D_CU(cudaStreamSynchronize(s), "column scratch: synchronise"); cudaFree(p->cs_scratch); p->cs_scratch = nullptr; p->cs_scratch_bytes = 0; }
        LP2D_CU(cudaMalloc(&p->cs_scratch, need), "cudaMalloc(column scratch)");
        p->cs_scratch_bytes = need;
    }
    const bool inv = (a.flags & F_INVERSE) != 0;
    const int A = p->cs_a, B = p->cs_b;
    for (long long b0 = 0; b0 < beams; ++b0) {
        for (long long c0 = 0; c0 < n; c0 += wp) {
            for (int step = 0; step < 2; ++step) {
                const ColStepOps<T>* ops = cs_of<T>(p, step);
                ColStepArgs<T> c;
                memset(&c, 0, sizeof c);
                c.tw = static_cast<const cx<T>*>(p->cs_tw[step]);
                c.tiles = (int)(wp / ops->wc);
                c.scale = (T)1;
                if (step == 0) {
                    c.in = a.in + b0 * a.beam_stride + c0; c.in_plane = a.beam_stride;
                    c.out = static_cast<cx<T>*>(p->cs_scratch); c.out_plane = n * wp;
                    c.tw4 = static_cast<const cx<T>*>(p->cs_tw4);
                    c.U = B;                                       // u = j2
                    c.in_u = n; c.in_pos = (long long)B * n;       // row j1 * B + j2 of the grid
                    c.out_u = wp; c.out_pos = (long long)B * wp;   // (k1, j2) -> row k1 * B + j2 of the panel
                    c.swap_in = inv ? 1 : 0;
                } else {
                    c.in = static_cast<const cx<T>*>(p->cs_scratch); c.in_plane = n * wp;
                    c.out = a.out ? a.out + b0 * a.beam_stride + c0 : nullptr; c.out_plane = a.beam_stride;
                    c.inten = a.inten ? a.inten + b0 * a.beam_stride + c0 : nullptr;
                    c.U = A;                                       // u = k1
                    c.in_u = (long long)B * wp; c.in_pos = wp;     // row k1 * B + j2 of the panel
                    c.out_u = n; c.out_pos = (long long)A * n;     // k = k1 + A * k2
                    c.swap_out = inv ? 1 : 0;
                    c.scale = a.scale;
                    c.stream_out = (a.flags & F_OUT_STREAM) ? 1 : 0;
                }
                c.nitems = (long long)c.U * c.tiles;
                long long grid = p->cs_grid_mult > 0 ? (long long)p->cs_sms * p->cs_ctas[step] * p->cs_grid_mult : c.nitems;
                if (grid > c.nitems) grid = c.nitems;
                LP2D_CU(ops->launch(c, (unsigned)grid, s), "two-step column pass launch");
            }
        }
    }
    return LP2D_OK;
}

template <typename T>
int cols(const lp2d_plan* p, int mid, LineArgs<T> a, long long beams, cudaStream_t s) {
    if (mid == MID_NONE && p->cs_a != 0 && a.out_pre == nullptr && a.mask_kind == MASK_NONE)
        return cols_two_step<T>(const_cast<lp2d_plan*>(p), a, beams, s);
    a.groups = 1;
    a.group_beams = beams;
    a.nbeams = beams;
    a.nlines = (int)(beams * (p->n / ops_of<T>(p)->w_col));
    LP2D_CU(ops_of<T>(p)->col(mid, a, s), "column pass launch");
    return LP2D_OK;
}

// The H plane of a repeated step on the one-kernel-per-pass path: the second consecutive call with the same z builds it
// (FP64, reference operation order, half the samples by the kx fold), later calls read it.  *use = the plane holds H(z).
template <typename T>
int ensure_h_plane(lp2d_plan* p, double z, cudaStream_t s, bool* use) {
    *use = false;
    if (p->hp_mode == 0 || !p->h_even) return LP2D_OK;
    if (p->hp_valid && p->hp_z == z) {
        // readers on other streams than the one that built the plane wait for the build (device-side), nothing else
        if (s != p->hp_stream) {
            LP2D_CU(cudaStreamWaitEvent(s, p->hp_event, 0), "H plane: wait for the build");
            p->hp_multi = true;
        }
        *use = true;
    } else if (p->hp_mode == 2 || (p->hp_seen && p->hp_seen_z == z)) {
        const int fl = p->n / 2 + 2;
        if (p->hp_fac == nullptr) {
            if (cudaMalloc(&p->hp_fac, (size_t)p->n * fl * sizeof(cx<T>)) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->hp_event, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError(); cudaFree(p->hp_fac); p->hp_fac = nullptr;
            }
        } else if (p->hp_valid) {
            // rebuilt for another z: everything that still reads the old plane must be done
            if (p->hp_multi) { LP2D_CU(cudaDeviceSynchronize(), "H plane: device synchronise"); }
            else if (s != p->hp_stream && cudaStreamSynchronize(p->hp_stream) != cudaSuccess) { cudaGetLastError(); LP2D_CU(cudaDeviceSynchronize(), "H plane: device synchronise"); }
        }
        if (p->hp_fac != nullptr) {
            cudaError_t e = sizeof(T) == 4
                ? p2_factor_plane_f32(static_cast<float2*>(p->hp_fac), p->tabs, p->n, fl, P2_FAC_H, 0, 0, z, 0.0, 0.0, s)
                : p2_factor_plane_f64(static_cast<double2*>(p->hp_fac), p->tabs, p->n, fl, P2_FAC_H, 0, 0, z, 0.0, 0.0, s);
            if (e != cudaSuccess) return cuda_fail(e, "H plane launch");
            LP2D_CU(cudaEventRecord(p->hp_event, s), "H plane: event");
            p->hp_valid = true; p->hp_z = z; p->hp_stream = s; p->hp_multi = false; *use = true;
        }
    }
    p->hp_seen = true; p->hp_seen_z = z;
    return LP2D_OK;
}

// data *= mask (optional), state stored to `data`, other = (i)fft2(data)
template <typename T>
int mul_fft2(lp2d_plan* p, const void* in, void* state_out, void* other, const void* mask, int mask_kind,
             long long mask_stride, long long batch, int inverse, int pre, double k0, double two_f, cudaStream_t s) {
    const T sc = inverse ? (T)(1.0 / (double)p->n) : (T)1;
    {
        P2Stage<T> st[2] = {p2_stage<T>(AX_ROW, inverse, sc), p2_stage<T>(AX_COL, inverse, sc)};
        st[0].in = static_cast<const cx<T>*>(in);
        st[0].out_pre = static_cast<cx<T>*>(state_out);
        st[0].out_ring = 1;
        st[0].pre = (pre == PRE_LENS) ? P2_FAC_LENS : P2_FAC_NONE;
        if (mask != nullptr) { st[0].mask = mask; st[0].mask_kind = mask_kind; st[0].mask_beam_stride = mask_stride; }
        st[1].in_ring = 1;
        st[1].out = static_cast<cx<T>*>(other);
        st[1].out_t = p->spectrum_t;
        const int rc = (p->spectrum_t || (p->fused_ops & (state_out != nullptr ? 2 : 1)))
                           ? run_p2<T>(p, st, 2, st[0].pre, 0, k0, two_f, batch, s, state_out != nullptr ? 2 : 1) : -1;
        if (rc >= 0) return rc;
        if (p->spectrum_t) return fail(LP2D_EUNSUPPORTED, "transposed spectrum layout: the fused step kernel does not apply to this call");
    }
    const long long plane = (long long)p->n * p->n;
    const long long cb = chunk_beams<T>(p, batch);
    const int H = p->hints;
    Fan fan(p, s, (batch + cb - 1) / cb);
    LP2D_CU(fan.fork(), "stream fork");
    for (long long b0 = 0; b0 < batch; b0 += cb) {
        const long long nb = (batch - b0 < cb) ? batch - b0 : cb;
        s = fan.next();
        LineArgs<T> a = base_args<T>(p);
        a.in = static_cast<const cx<T>*>(in) + b0 * plane;
        a.out = static_cast<cx<T>*>(other) + b0 * plane;
        a.out_pre = state_out ? static_cast<cx<T>*>(state_out) + b0 * plane : nullptr;
        a.flags = (inverse ? F_INVERSE : 0) | (H ? (F_IN_STREAM | F_STATE_STREAM) : 0);
        a.scale = sc;
        a.lens_k0 = k0; a.lens_2f = two_f;
        if (mask != nullptr) {
            a.mask = static_cast<const char*>(mask) +
                     b0 * mask_stride * (mask_kind == MASK_CX ? sizeof(cx<T>) : mask_kind == MASK_REAL ? sizeof(T) : 1);
            a.mask_kind = mask_kind;
            a.mask_beam_stride = mask_stride;
        }
        int rc = rows<T>(p, pre, a, nb, s);
        if (rc) return rc;
        LineArgs<T> c = base_args<T>(p);
        c.in = a.out; c.out = a.out;
        c.flags = (inverse ? F_INVERSE : 0) | (H ? F_OUT_STREAM : 0); c.scale = sc;
        rc = cols<T>(p, MID_NONE, c, nb, s);
        if (rc) return rc;
    }
    LP2D_CU(fan.join(), "stream join");
    return LP2D_OK;
}

template <typename T>
int propagate_state(lp2d_plan* p, void* spectrum, void* field, void* intensity, long long batch, double z,
                    cudaStream_t s, double* moments = nullptr, double dL = 0.0, bool* moments_done = nullptr,
                    void* probe_out = nullptr, long long probe_index = 0) {
    const T sc = (T)(1.0 / (double)p->n);
    {
        // rows: spectrum *= H (state stored back), inverse row transform -> ring; columns: inverse -> field
        P2Stage<T> st[2] = {p2_stage<T>(AX_ROW, 1, sc), p2_stage<T>(AX_COL, 1, sc)};
        st[0].in = static_cast<const cx<T>*>(spectrum);
        st[0].out_pre = static_cast<cx<T>*>(spectrum);
        st[0].pre = P2_FAC_H;
        st[0].out_ring = 1;
        st[1].in_ring = 1;
        st[1].out = static_cast<cx<T>*>(field);
        st[1].inten = static_cast<T*>(intensity);
        // transposed spectrum: a memory row is a grid COLUMN of the spectrum (factor rows by column), and the column
        // stage's transposed store lands the field in its natural orientation
        st[1].out_t = p->spectrum_t;
        st[1].moments = moments; st[1].dL = dL;                 // (zeroed by the caller; inverse stage only; the per-pass path ignores them)
        st[1].probe_out = static_cast<T*>(probe_out); st[1].probe_index = probe_index;
        const int rc = (p->spectrum_t || p->fused_prop) ? run_p2<T>(p, st, 2, P2_FAC_H, p->spectrum_t, z, 0.0, batch, s, 2) : -1;
        if (rc >= 0) {
            if (moments_done != nullptr) *moments_done = (rc == 0);   // the fused kernel delivered moments and probe
            return rc;
        }
        if (p->spectrum_t) return fail(LP2D_EUNSUPPORTED, "transposed spectrum layout: the fused step kernel does not apply to this call");
    }
    const long long plane = (long long)p->n * p->n;
    const long long cb = chunk_beams<T>(p, batch);
    const int H = p->hints;
    // The H plane of a repeated step: the second consecutive call with the same z builds it (FP64, reference operation
    // order, half the samples by the kx fold), later calls read it -- the per-sample evaluation costs a single-field row
    // pass 17-25 % of the whole step (profiles/r3s_h_cost.log).  Needs even tables (the fold).
    bool use_plane = false;
    { const int rc = ensure_h_plane<T>(p, z, s, &use_plane); if (rc) return rc; }
    Fan fan(p, s, (batch + cb - 1) / cb);
    LP2D_CU(fan.fork(), "stream fork");
    for (long long b0 = 0; b0 < batch; b0 += cb) {
        const long long nb = (batch - b0 < cb) ? batch - b0 : cb;
        s = fan.next();
        // rows: spectrum *= H (stored back), inverse row transform -> field buffer
        LineArgs<T> a = base_args<T>(p);
        a.in = static_cast<const cx<T>*>(spectrum) + b0 * plane;
        a.out_pre = static_cast<cx<T>*>(spectrum) + b0 * plane;
        a.out = static_cast<cx<T>*>(field) + b0 * plane;
        a.flags = F_INVERSE | (H ? (F_IN_STREAM | F_STATE_STREAM) : 0); a.scale = sc; a.z = z;
        a.fac = static_cast<const cx<T>*>(p->hp_fac); a.fac_stride = p->n / 2 + 2;
        int rc = rows<T>(p, use_plane ? PRE_HPLANE : PRE_H, a, nb, s);
        if (rc) return rc;
        // columns: inverse transform in place, optional |E|^2
        LineArgs<T> c = base_args<T>(p);
        c.in = a.out; c.out = a.out;
        c.inten = intensity ? static_cast<T*>(intensity) + b0 * plane : nullptr;
        c.flags = F_INVERSE | (H ? F_OUT_STREAM : 0); c.scale = sc;
        rc = cols<T>(p, MID_NONE, c, nb, s);
        if (rc) return rc;
    }
    LP2D_CU(fan.join(), "stream join");
    return LP2D_OK;
}

template <typename T>
int propagate_fresh(lp2d_plan* p, const void* fin, void* fout, void* spec_out, void* intensity, void* work,
                    long long work_bytes, long long batch, double z, cudaStream_t s) {
    if (p->spectrum_t && spec_out != nullptr) {
        // transposed spectrum wanted as a by-product: fft2 (-> transposed spectrum), then the stateful step
        if (fout == nullptr) return fail(LP2D_EUNSUPPORTED, "propagate_field: transposed spectrum_out needs field_out");
        int rc = mul_fft2<T>(p, fin, nullptr, spec_out, nullptr, 0, 0, batch, 0, PRE_NONE, 0, 0, s);
        if (rc) return rc;
        return propagate_state<T>(p, spec_out, fout, intensity, batch, z, s);
    }
    if (p->cs_a != 0 && spec_out != nullptr && fout != nullptr) {
        // grids whose column pass runs in two steps: the fused column stage (FFT, H, inverse FFT) would need the
        // one-kernel column pass with its 16-32-byte pieces (8192^2 complex128: 42 ms); with the spectrum wanted
        // anyway, fft2 + the stateful step are four streaming passes instead
        int rc = mul_fft2<T>(p, fin, nullptr, spec_out, nullptr, 0, 0, batch, 0, PRE_NONE, 0, 0, s);
        if (rc) return rc;
        return propagate_state<T>(p, spec_out, fout, intensity, batch, z, s);
    }
    {
        // forward rows -> ring; columns: FFT, H, inverse FFT in the ring; inverse rows -> field / |E|^2
        const T inv_n = (T)(1.0 / (double)p->n);
        P2Stage<T> st[3] = {p2_stage<T>(AX_ROW, 0, (T)1), p2_stage<T>(AX_COL, 0, inv_n), p2_stage<T>(AX_ROW, 1, inv_n)};
        st[0].in = static_cast<const cx<T>*>(fin);
        st[0].out_ring = 1;
        st[1].in_ring = 1; st[1].out_ring = 1; st[1].mid = 1;
        st[1].out_mid = static_cast<cx<T>*>(spec_out);
        st[2].in_ring = 1;
        st[2].out = static_cast<cx<T>*>(fout);
        st[2].inten = static_cast<T*>(intensity);
        const int rc = (p->fused_ops & 4) ? run_p2<T>(p, st, 3, P2_FAC_H, 1, z, 0.0, batch, s, 4) : -1;
        if (rc >= 0) return rc;
    }
    const long long plane = (long long)p->n * p->n;
    long long cb = chunk_beams<T>(p, batch);
    const bool via_work = (fout == nullptr);
    if (via_work) {
        const long long fit = work_bytes / (plane * (long long)sizeof(cx<T>));
        if (work == nullptr || fit < 1) return fail(LP2D_EINVAL, "propagate_field: field_out is NULL and work holds less than one plane");
        if (cb > fit) cb = fit;
    }
    const T sc = (T)(1.0 / (double)p->n);
    const int H = p->hints;
    // H plane of a repeated z for the fused column stage: complex64 only (2048^2: 59.2 -> 56.9 us per beam); the complex128
    // column kernel is bound by its L1 wavefronts (67 % busy) and the plane reads cost more than the FP64 evaluation
    // that otherwise runs in an idle pipe (1024^2: 24.2 -> 28.3)
    bool use_plane = false;
    if (sizeof(T) == 4) { const int rc = ensure_h_plane<T>(p, z, s, &use_plane); if (rc) return rc; }
    else { p->hp_seen = true; p->hp_seen_z = z; }
    // the shared work buffer serialises the chunks; with per-chunk output buffers they fan out
    Fan fan(p, s, via_work ? 1 : (batch + cb - 1) / cb);
    LP2D_CU(fan.fork(), "stream fork");
    for (long long b0 = 0; b0 < batch; b0 += cb) {
        const long long nb = (batch - b0 < cb) ? batch - b0 : cb;
        s = fan.next();
        cx<T>* buf = via_work ? static_cast<cx<T>*>(work) : static_cast<cx<T>*>(fout) + b0 * plane;
        // 1. forward row transform: field -> buf
        LineArgs<T> a = base_args<T>(p);
        a.in = static_cast<const cx<T>*>(fin) + b0 * plane;
        a.out = buf;
        a.flags = H ? F_IN_STREAM : 0;
        int rc = rows<T>(p, PRE_NONE, a, nb, s);
        if (rc) return rc;
        // 2. fused column pass in place: FFT_y, H, inverse FFT_y (scaled 1/N)
        LineArgs<T> c = base_args<T>(p);
        c.in = buf; c.out = buf; c.z = z; c.scale = sc;
        c.flags = H ? F_STATE_STREAM : 0;
        c.out_mid = spec_out ? static_cast<cx<T>*>(spec_out) + b0 * plane : nullptr;
        c.fac = static_cast<const cx<T>*>(p->hp_fac); c.fac_stride = p->n / 2 + 2;
        rc = cols<T>(p, use_plane ? MID_HPLANE : MID_H, c, nb, s);
        if (rc) return rc;
        // 3. inverse row transform (scaled 1/N): buf -> field_out and/or |E|^2
        LineArgs<T> r = base_args<T>(p);
        r.in = buf;
        r.out = via_work ? nullptr : buf;
        r.flags = F_INVERSE | (H ? F_OUT_STREAM : 0); r.scale = sc;
        r.inten = intensity ? static_cast<T*>(intensity) + b0 * plane : nullptr;
        rc = rows<T>(p, PRE_NONE, r, nb, s);
        if (rc) return rc;
    }
    LP2D_CU(fan.join(), "stream join");
    return LP2D_OK;
}

}  // namespace

extern "C" {

int lp2d_fft2(lp2d_plan* p, const void* in, void* out, long long batch, int inverse, void* stream) {
    LP2D_CHECK_2D(p);
    if (in == nullptr || out == nullptr || batch < 0) return fail(LP2D_EINVAL, "fft2: NULL pointer or negative batch");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return p->dtype == LP2D_C64
               ? mul_fft2<float>(p, in, nullptr, out, nullptr, 0, 0, batch, inverse, PRE_NONE, 0, 0, s)
               : mul_fft2<double>(p, in, nullptr, out, nullptr, 0, 0, batch, inverse, PRE_NONE, 0, 0, s);
}

int lp2d_propagate(lp2d_plan* p, void* spectrum, void* field, void* intensity, long long batch, double z,
                   void* stream) {
    LP2D_CHECK_2D(p);
    if (spectrum == nullptr || field == nullptr || batch < 0) return fail(LP2D_EINVAL, "propagate: NULL pointer or negative batch");
    if (spectrum == field) return fail(LP2D_EINVAL, "propagate: spectrum and field must be distinct buffers");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return p->dtype == LP2D_C64 ? propagate_state<float>(p, spectrum, field, intensity, batch, z, s)
                                : propagate_state<double>(p, spectrum, field, intensity, batch, z, s);
}

int lp2d_propagate_sweep(lp2d_plan* p, void* spectrum, void* field, long long batch, int nsteps, double dz, long long probe_index,
                         void* probe_out, double dL, double* moments_out, void* stream) {
    LP2D_CHECK_2D(p);
    if (spectrum == nullptr || field == nullptr || batch < 0 || nsteps < 0) return fail(LP2D_EINVAL, "propagate_sweep: NULL pointer or negative count");
    if (spectrum == field) return fail(LP2D_EINVAL, "propagate_sweep: spectrum and field must be distinct buffers");
    const long long plane = (long long)p->n * p->n;
    if (probe_out != nullptr && (probe_index < 0 || probe_index >= plane)) return fail(LP2D_EINVAL, "propagate_sweep: probe_index outside the plane");
    if (batch == 0 || nsteps == 0 || dz == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t rsz = p->dtype == LP2D_C64 ? sizeof(float) : sizeof(double);
    if (moments_out != nullptr)                                 // the fused step accumulates into it (atomics)
        LP2D_CU(cudaMemsetAsync(moments_out, 0, (size_t)nsteps * batch * 5 * sizeof(double), s), "propagate_sweep: memset");
    for (int k = 0; k < nsteps; ++k) {
        bool fused_moments = false;
        double* mk_ = moments_out ? moments_out + (size_t)k * batch * 5 : nullptr;
        void* pk_ = probe_out ? static_cast<char*>(probe_out) + (size_t)k * batch * rsz : nullptr;
        const int rc = p->dtype == LP2D_C64 ? propagate_state<float>(p, spectrum, field, nullptr, batch, dz, s, mk_, dL, &fused_moments, pk_, probe_index)
                                            : propagate_state<double>(p, spectrum, field, nullptr, batch, dz, s, mk_, dL, &fused_moments, pk_, probe_index);
        if (rc) return rc;
        if (probe_out != nullptr && !fused_moments) {
            char* dst = static_cast<char*>(probe_out) + (size_t)k * batch * rsz;
            const unsigned blocks = (unsigned)((batch + 127) / 128);
            if (p->dtype == LP2D_C64)
                probe_kernel<float><<<blocks, 128, 0, s>>>(static_cast<const float2*>(field), reinterpret_cast<float*>(dst), plane, probe_index, batch);
            else
                probe_kernel<double><<<blocks, 128, 0, s>>>(static_cast<const double2*>(field), reinterpret_cast<double*>(dst), plane, probe_index, batch);
            ++g_launches;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "probe launch");
        }
        if (moments_out != nullptr && !fused_moments) {
            const size_t csz = p->dtype == LP2D_C64 ? sizeof(float2) : sizeof(double2);
            for (long long b = 0; b < batch; ++b) {
                const int rc2 = lp2d_moments(p->dtype, static_cast<const char*>(field) + (size_t)b * plane * csz, p->n, dL, 0.0, 0.0,
                                             moments_out + ((size_t)k * batch + b) * 5, stream);
                if (rc2) return rc2;
            }
        }
    }
    return LP2D_OK;
}

int lp2d_propagate_field(lp2d_plan* p, const void* field_in, void* field_out, void* spectrum_out, void* intensity,
                         void* work, long long work_bytes, long long batch, double z, void* stream) {
    LP2D_CHECK_2D(p);
    if (field_in == nullptr || batch < 0) return fail(LP2D_EINVAL, "propagate_field: NULL input or negative batch");
    if (field_out == nullptr && intensity == nullptr && spectrum_out == nullptr)
        return fail(LP2D_EINVAL, "propagate_field: no output requested");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return p->dtype == LP2D_C64
               ? propagate_fresh<float>(p, field_in, field_out, spectrum_out, intensity, work, work_bytes, batch, z, s)
               : propagate_fresh<double>(p, field_in, field_out, spectrum_out, intensity, work, work_bytes, batch, z, s);
}

int lp2d_lens(lp2d_plan* p, void* field, void* spectrum, long long batch, double k0, double f, void* stream) {
    LP2D_CHECK_2D(p);
    if (field == nullptr || spectrum == nullptr || batch < 0) return fail(LP2D_EINVAL, "lens: NULL pointer or negative batch");
    if (field == spectrum) return fail(LP2D_EINVAL, "lens: field and spectrum must be distinct buffers");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return p->dtype == LP2D_C64
               ? mul_fft2<float>(p, field, field, spectrum, nullptr, 0, 0, batch, 0, PRE_LENS, k0, 2.0 * f, s)
               : mul_fft2<double>(p, field, field, spectrum, nullptr, 0, 0, batch, 0, PRE_LENS, k0, 2.0 * f, s);
}

int lp2d_mul_fft2(lp2d_plan* p, void* data, void* other, const void* mask, int mask_kind, long long mask_beam_stride,
                  long long batch, int inverse, void* stream) {
    LP2D_CHECK_2D(p);
    if (data == nullptr || other == nullptr || batch < 0) return fail(LP2D_EINVAL, "mul_fft2: NULL pointer or negative batch");
    if (data == other) return fail(LP2D_EINVAL, "mul_fft2: data and other must be distinct buffers");
    if (mask != nullptr && (mask_kind < LP2D_MASK_COMPLEX || mask_kind > LP2D_MASK_U8))
        return fail(LP2D_EINVAL, "mul_fft2: unknown mask_kind");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    void* state = mask ? data : nullptr;
    return p->dtype == LP2D_C64
               ? mul_fft2<float>(p, data, state, other, mask, mask_kind, mask_beam_stride, batch, inverse, PRE_NONE, 0, 0, s)
               : mul_fft2<double>(p, data, state, other, mask, mask_kind, mask_beam_stride, batch, inverse, PRE_NONE, 0, 0, s);
}

int lp2d_pass(lp2d_plan* p, int axis, int op, const void* in, void* out, void* out_state, void* intensity,
              long long batch, int inverse, double z, void* stream) {
    LP2D_CHECK_2D(p);
    if (in == nullptr || batch < 0 || (axis != LP2D_AXIS_ROWS && axis != LP2D_AXIS_COLS))
        return fail(LP2D_EINVAL, "pass: bad arguments");
    if (op < LP2D_OP_FFT || op > LP2D_OP_FFT_H_IFFT) return fail(LP2D_EINVAL, "pass: unknown op");
    if (axis == LP2D_AXIS_ROWS && op == LP2D_OP_FFT_H_IFFT) return fail(LP2D_EUNSUPPORTED, "pass: the fused FFT-H-IFFT pass exists for columns only");
    if (axis == LP2D_AXIS_COLS && op == LP2D_OP_H_FFT) return fail(LP2D_EUNSUPPORTED, "pass: the H-then-transform pass exists for rows only");
    if (batch == 0) return LP2D_OK;
    if (batch * (long long)p->n > 0x7fffffffll) return fail(LP2D_EINVAL, "pass: batch too large for one launch");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double sc = (inverse || op == LP2D_OP_FFT_H_IFFT) ? 1.0 / (double)p->n : 1.0;
#define LP2D_PASS_BODY(T)                                                                        \
    {                                                                                            \
        LineArgs<T> a = base_args<T>(p);                                                         \
        a.in = static_cast<const cx<T>*>(in);                                                    \
        a.out = static_cast<cx<T>*>(out);                                                        \
        if (op == LP2D_OP_FFT_H_IFFT) a.out_mid = static_cast<cx<T>*>(out_state);                \
        else a.out_pre = static_cast<cx<T>*>(out_state);                                         \
        a.inten = static_cast<T*>(intensity);                                                    \
        a.flags = inverse ? F_INVERSE : 0;                                                       \
        a.scale = (T)sc;                                                                         \
        a.z = z;                                                                                 \
        return axis == LP2D_AXIS_ROWS ? rows<T>(p, op == LP2D_OP_H_FFT ? PRE_H : PRE_NONE, a, batch, s) \
                                      : cols<T>(p, op == LP2D_OP_FFT_H_IFFT ? MID_H : MID_NONE, a, batch, s); \
    }
    if (p->dtype == LP2D_C64) LP2D_PASS_BODY(float) else LP2D_PASS_BODY(double)
#undef LP2D_PASS_BODY
}

static int pass_lines_impl(lp2d_plan* p, int op, const void* in, int tile, void* out, void* out_state, void* intensity, int line0,
                           int nlines, int inverse, double z, void* stream) {
    LP2D_CHECK_PLAN(p);
    if (in == nullptr || nlines < 0 || line0 < 0 || line0 + nlines > p->n) return fail(LP2D_EINVAL, "pass_lines: bad arguments");
    if (op != LP2D_OP_FFT && op != LP2D_OP_H_FFT) return fail(LP2D_EINVAL, "pass_lines: op must be LP2D_OP_FFT or LP2D_OP_H_FFT");
    int seg_shift = 0;
    if (tile != 0) {
        if (tile < 1 || (tile & (tile - 1)) != 0 || p->n % tile != 0)
            return fail(LP2D_EINVAL, "pass_lines_tiles: tile must be a power of two that divides npoints");
        if (tile != p->n) {                            // one tile of n samples = the contiguous slab
            while ((1 << seg_shift) < tile) ++seg_shift;
            if (in == out || in == out_state) return fail(LP2D_EINVAL, "pass_lines_tiles: the tiled input cannot be transformed in place");
        }
    }
    if (nlines == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double sc = inverse ? 1.0 / (double)p->n : 1.0;
    const long long back = (long long)line0 * p->n;      // the kernel indexes from line 0 of the plane
#define LP2D_LINES_BODY(T)                                                                        \
    {                                                                                             \
        LineArgs<T> a = base_args<T>(p);                                                          \
        a.in = seg_shift ? static_cast<const cx<T>*>(in) : static_cast<const cx<T>*>(in) - back;  \
        a.out = out ? static_cast<cx<T>*>(out) - back : nullptr;                                  \
        a.out_pre = out_state ? static_cast<cx<T>*>(out_state) - back : nullptr;                  \
        a.inten = intensity ? static_cast<T*>(intensity) - back : nullptr;                        \
        a.flags = inverse ? F_INVERSE : 0;                                                        \
        a.scale = (T)sc;                                                                          \
        a.z = z;                                                                                  \
        a.groups = 1; a.group_beams = 1; a.nbeams = 1;                                            \
        a.line0 = line0; a.nlines = nlines;                                                       \
        a.seg_shift = seg_shift; a.seg_rows = nlines;                                             \
        if (ops_of<T>(p) == nullptr) {                                                            \
            LP2D_CU(xl_of<T>(p)->launch(op == LP2D_OP_H_FFT ? PRE_H : PRE_NONE, a, s), "cluster line pass launch"); \
            return LP2D_OK;                                                                       \
        }                                                                                         \
        LP2D_CU(ops_of<T>(p)->row(op == LP2D_OP_H_FFT ? PRE_H : PRE_NONE, a, s), "slab row pass launch"); \
        return LP2D_OK;                                                                           \
    }
    if (p->dtype == LP2D_C64) LP2D_LINES_BODY(float) else LP2D_LINES_BODY(double)
#undef LP2D_LINES_BODY
}

int lp2d_pass_lines(lp2d_plan* p, int op, const void* in, void* out, void* out_state, void* intensity, int line0,
                    int nlines, int inverse, double z, void* stream) {
    return pass_lines_impl(p, op, in, 0, out, out_state, intensity, line0, nlines, inverse, z, stream);
}

int lp2d_pass_lines_tiles(lp2d_plan* p, int op, const void* in_tiles, int tile, void* out, void* out_state, void* intensity,
                          int line0, int nlines, int inverse, double z, void* stream) {
    if (tile <= 0) return fail(LP2D_EINVAL, "pass_lines_tiles: tile must be positive");
    return pass_lines_impl(p, op, in_tiles, tile, out, out_state, intensity, line0, nlines, inverse, z, stream);
}

int lp2d_intensity(lp2d_plan* p, const void* field, void* out, long long batch, void* stream) {
    LP2D_CHECK_PLAN(p);
    if (field == nullptr || out == nullptr || batch < 0) return fail(LP2D_EINVAL, "intensity: NULL pointer or negative batch");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long n = batch * (long long)p->n * p->n;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (p->dtype == LP2D_C64)
        intensity_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const float2*>(field), static_cast<float*>(out), n);
    else
        intensity_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const double2*>(field), static_cast<double*>(out), n);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "intensity launch");
    return LP2D_OK;
}

}  // extern "C"
