// Initial-field and mask generators on the device: the eleven profiles of
// lightprop2d/profiles.py:35-136 evaluated per sample from the 1-D coordinate tables, so that ensembles never touch
// host memory.  Deterministic profiles use the reference's FP64 expressions (same operation order); the random ones
// draw from the counter-based generator of lp2d_random_fields (Philox4x32-10 keyed by seed, beam, sample) -- numpy's
// Mersenne-Twister stream cannot be reproduced on the device, seeded parity tests use the host generators.
#include <string>

#include "../../include/lp2d.h"
#include "dispatch.h"

namespace lp2d {
int set_error(int code, const std::string& msg);   // lp2d_api.cu

__device__ __forceinline__ void gen_philox(unsigned int (&c)[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// uniform [0, 1) number of cell `idx` of beam `beam`: the same words, in the same order, as random_fields_kernel
// (lp2d_aux.cu) draws for sample idx, so both entry points generate the same ensemble
template <typename T>
__device__ __forceinline__ double gen_uniform(unsigned long long seed, unsigned long long beam, long long idx) {
    const long long q = idx >> 2;
    const int j = (int)(idx & 3);
    unsigned int c[4] = {(unsigned int)q, (unsigned int)((unsigned long long)q >> 32), (unsigned int)beam, (unsigned int)(beam >> 32)};
    if (sizeof(T) == 8) {
        unsigned int d[4] = {c[0], c[1] | 0x80000000u, c[2], c[3]};
        gen_philox(c, (unsigned int)seed, (unsigned int)(seed >> 32));
        gen_philox(d, (unsigned int)seed, (unsigned int)(seed >> 32));
        const unsigned int w[8] = {c[0], c[1], c[2], c[3], d[0], d[1], d[2], d[3]};
        const unsigned long long bits = ((unsigned long long)w[2 * j] << 21) ^ (unsigned long long)(w[2 * j + 1] >> 11);
        return (double)(bits & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
    }
    gen_philox(c, (unsigned int)seed, (unsigned int)(seed >> 32));
    return (double)((float)(c[j] >> 8) * (1.0f / 16777216.0f));
}

struct GenArgs {
    const double* x;       // device: X (n values, already shifted by the caller where the reference shifts in place)
    const double* y;       // device: Y (n values)
    double p[6];           // profile parameters (see lp2d.h)
    int kind, n, bin;
    long long batch;
    unsigned long long seed, first_beam;
};

// the reference's round_hole: gaussian_beam(x - x0, y - y0, 1, R) >= 1 / exp(0.5)   (profiles.py:64-81)
// (thr = 1 / exp(0.5) as the HOST evaluates it, p[5])
__device__ __forceinline__ bool gen_round_hole(double x, double y, double R, double x0, double y0, double thr) {
    const double xs = x - x0, ys = y - y0;
    return 1 * exp(-(xs * xs + ys * ys) / (R * R)) >= thr;
}
__device__ __forceinline__ bool gen_rect(double x, double y, double dx, double dy, double x0, double y0) {
    return (fabs(x - x0) < (dx / 2)) & (fabs(y - y0) < (dy / 2));
}

template <typename T, bool MASK>
__global__ void generate_kernel(void* __restrict__ out_, GenArgs a) {
    const long long plane = (long long)a.n * a.n;
    const long long total = plane * a.batch;
    const int nc = a.n / a.bin;                            // coarse cells per axis of the binned random profiles
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, s = i - b * plane;
        const int r = (int)(s / a.n), c = (int)(s - (long long)r * a.n);
        const double x = a.x[c], y = a.y[r];
        double re = 0.0, im = 0.0;
        bool m = false;
        const long long cell = (a.bin == 1) ? s : (long long)(r / a.bin) * nc + c / a.bin;
        const unsigned long long beam = a.first_beam + (unsigned long long)b;
        switch (a.kind) {
            case LP2D_GEN_PLANE_WAVE: re = 1.0; m = true; break;
            case LP2D_GEN_RANDOM_WAVE: re = gen_uniform<T>(a.seed, beam, cell); break;
            case LP2D_GEN_RANDOM_WAVE_BIN: re = gen_uniform<T>(a.seed, beam, cell) < 0.5 ? 0.0 : 1.0; m = re != 0.0; break;
            case LP2D_GEN_GAUSSIAN_BEAM: re = a.p[0] * exp(-(x * x + y * y) / (a.p[1] * a.p[1])); break;
            case LP2D_GEN_ROUND_HOLE: m = gen_round_hole(x, y, a.p[0], a.p[1], a.p[2], a.p[5]); re = m ? 1.0 : 0.0; break;
            case LP2D_GEN_RANDOM_ROUND_HOLE:
                if (gen_round_hole(x, y, a.p[0], a.p[1], a.p[2], a.p[5])) re = gen_uniform<T>(a.seed, beam, cell);
                break;
            case LP2D_GEN_RANDOM_ROUND_HOLE_BIN:
                if (gen_round_hole(x, y, a.p[0], a.p[1], a.p[2], a.p[5])) re = gen_uniform<T>(a.seed, beam, cell) < 0.5 ? 0.0 : 1.0;
                m = re != 0.0;
                break;
            case LP2D_GEN_RANDOM_ROUND_HOLE_PHASE:
                if (gen_round_hole(x, y, a.p[0], a.p[1], a.p[2], a.p[5])) {
                    if (sizeof(T) == 8) sincospi(2.0 * gen_uniform<T>(a.seed, beam, cell), &im, &re);
                    else { float sn, cs; sincospif(2.0f * (float)gen_uniform<T>(a.seed, beam, cell), &sn, &cs); re = cs; im = sn; }
                }
                break;
            case LP2D_GEN_RECTANGLE_HOLE: m = gen_rect(x, y, a.p[0], a.p[1], a.p[2], a.p[3]); re = m ? 1.0 : 0.0; break;
            case LP2D_GEN_SQUARE_SLITS: {
                const double dl = a.p[1] / 2;               // square_hole(d, x0 - dl, y0) | square_hole(d, x0 + dl, y0)
                m = gen_rect(x, y, a.p[0], a.p[0], a.p[2] - dl, a.p[3]) | gen_rect(x, y, a.p[0], a.p[0], a.p[2] + dl, a.p[3]);
                re = m ? 1.0 : 0.0;
                break;
            }
            default: break;
        }
        if (MASK) static_cast<unsigned char*>(out_)[i] = m ? 1 : 0;
        else static_cast<cx<T>*>(out_)[i] = mk<T>((T)re, (T)im);
    }
}

}  // namespace lp2d

using namespace lp2d;

extern "C" int lp2d_generate(int out_kind, void* out, int n, long long batch, const double* dev_x, const double* dev_y, int kind,
                             const double* host_params6, int binning_order, unsigned long long seed,
                             unsigned long long first_beam, void* stream) {
    if (out == nullptr || dev_x == nullptr || dev_y == nullptr || n <= 0 || batch < 0 || kind < LP2D_GEN_PLANE_WAVE ||
        kind > LP2D_GEN_SQUARE_SLITS || (out_kind != LP2D_C64 && out_kind != LP2D_C128 && out_kind != LP2D_GEN_OUT_MASK))
        return set_error(LP2D_EINVAL, "generate: bad arguments");
    if (binning_order < 1 || n % binning_order != 0) return set_error(LP2D_EINVAL, "generate: binning_order must divide npoints");
    const bool boolean = kind == LP2D_GEN_PLANE_WAVE || kind == LP2D_GEN_ROUND_HOLE || kind == LP2D_GEN_RECTANGLE_HOLE ||
                         kind == LP2D_GEN_SQUARE_SLITS || kind == LP2D_GEN_RANDOM_WAVE_BIN || kind == LP2D_GEN_RANDOM_ROUND_HOLE_BIN;
    if (out_kind == LP2D_GEN_OUT_MASK && !boolean) return set_error(LP2D_EINVAL, "generate: this profile is not a 0/1 mask");
    if (batch == 0) return LP2D_OK;
    GenArgs a;
    a.x = dev_x; a.y = dev_y;
    for (int i = 0; i < 6; ++i) a.p[i] = host_params6 ? host_params6[i] : 0.0;
    a.kind = kind; a.n = n; a.bin = binning_order; a.batch = batch; a.seed = seed; a.first_beam = first_beam;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long total = (long long)n * n * batch;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (out_kind == LP2D_GEN_OUT_MASK) generate_kernel<float, true><<<(unsigned)blocks, 256, 0, s>>>(out, a);
    else if (out_kind == LP2D_C64) generate_kernel<float, false><<<(unsigned)blocks, 256, 0, s>>>(out, a);
    else generate_kernel<double, false><<<(unsigned)blocks, 256, 0, s>>>(out, a);
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("generate launch: ") + cudaGetErrorString(e));
    return LP2D_OK;
}
