// Kernels for the grids the fused / per-pass power-of-two kernel sets do not cover:
//   * pointwise factors (transfer function, lens phase, mask multiply, |E|^2) on a slab of rows of an n x n plane,
//     ANY n, with the reference's FP64 operation order (kernels.cuh: transfer_factor / lens_factor);
//   * the chirp steps of Bluestein's algorithm, which turns a DFT of any length n into a cyclic convolution of a
//     power-of-two length m >= 2n - 1 (run by the power-of-two fft2 kernels): the reference accepts every npoints
//     (beam2d.py:299-308) and expand() produces e.g. 82 points (beam2d.py:385-389).
//
//   X[k] = sum_j x[j] e^{-2 pi i jk/n},  jk = (j^2 + k^2 - (k-j)^2) / 2
//        = w[k] * sum_j (x[j] w[j]) * conj(w)[k - j],          w[j] = e^{-pi i j^2 / n}
// in two dimensions with the separable chirp w[r] w[c].  j^2 is reduced mod 2n in integers, the chirp is evaluated in
// FP64 (sincospi) also for complex64 and rounded once.
#include <string>

#include "../../include/lp2d.h"
#include "dispatch.h"

namespace lp2d {
int set_error(int code, const std::string& msg);   // lp2d_api.cu

// e^{-+ pi i k^2 / n}: sign = -1 forward chirp w[k], +1 its conjugate
__device__ __forceinline__ double2 chirp(long long k, int n, int sign) {
    const long long q = (k * k) % (2ll * n);          // k^2 mod 2n: the phase is pi * q / n
    double s, c;
    sincospi((double)q / (double)n, &s, &c);
    return make_double2(c, sign < 0 ? -s : s);
}
__device__ __forceinline__ double2 dmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// data[b][r][c] *= factor(row0 + r, c)
template <typename T>
__global__ void pointwise_kernel(cx<T>* __restrict__ data, long long batch, int rows, int n, int row0, int kind,
                                 const double* __restrict__ a_tab, const double* __restrict__ b_tab,
                                 const double* __restrict__ x2_tab, double p0, double p1) {
    const long long plane = (long long)rows * n;
    const long long total = batch * plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long q = i % plane;
        const int r = (int)(q / n), c = (int)(q - (long long)r * n);
        const cx<T> f = (kind == LP2D_PW_H) ? transfer_factor<T>(a_tab[c], b_tab[row0 + r], p0)
                                            : lens_factor<T>(x2_tab[c], x2_tab[row0 + r], p0, p1);
        data[i] = cmul(data[i], f);
    }
}

template <typename T>
__global__ void mul_mask_kernel(cx<T>* __restrict__ data, const void* __restrict__ mask, int kind, long long plane,
                                long long mask_beam_stride, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, q = i - b * plane;
        const long long mi = b * mask_beam_stride + q;
        cx<T> v = data[i];
        if (kind == LP2D_MASK_COMPLEX) v = cmul(v, static_cast<const cx<T>*>(mask)[mi]);
        else {
            const T f = (kind == LP2D_MASK_REAL) ? static_cast<const T*>(mask)[mi] : (static_cast<const unsigned char*>(mask)[mi] ? (T)1 : (T)0);
            v.x *= f; v.y *= f;
        }
        data[i] = v;
    }
}

template <typename T>
__global__ void intensity_n_kernel(const cx<T>* __restrict__ f, T* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const cx<T> v = f[i];
        out[i] = v.x * v.x + v.y * v.y;
    }
}

// out[b] (m x m) = in[b] (n x n) * w[r] w[c] in the top-left corner, zero elsewhere
template <typename T>
__global__ void bluestein_pre_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n, int m, long long batch, int sign) {
    const long long plane = (long long)m * m;
    const long long total = batch * plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, q = i - b * plane;
        const int r = (int)(q / m), c = (int)(q - (long long)r * m);
        cx<T> v = mk<T>((T)0, (T)0);
        if (r < n && c < n) {
            const cx<T> x = in[(b * n + r) * (long long)n + c];
            const double2 w = dmul(chirp(r, n, sign), chirp(c, n, sign));
            const double2 y = dmul(make_double2((double)x.x, (double)x.y), w);
            v = mk<T>((T)y.x, (T)y.y);
        }
        out[i] = v;
    }
}

// out[b] (n x n) = in[b][0..n)[0..n) * w[r] w[c] * scale
template <typename T>
__global__ void bluestein_post_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n, int m, long long batch, int sign,
                                      double scale) {
    const long long plane = (long long)n * n;
    const long long total = batch * plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, q = i - b * plane;
        const int r = (int)(q / n), c = (int)(q - (long long)r * n);
        const cx<T> x = in[(b * m + r) * (long long)m + c];
        const double2 w = dmul(chirp(r, n, sign), chirp(c, n, sign));
        const double2 y = dmul(make_double2((double)x.x, (double)x.y), w);
        out[i] = mk<T>((T)(y.x * scale), (T)(y.y * scale));
    }
}

// the convolution kernel in the sample domain, wrapped to length m on both axes: f[r][c] = conj(w)[dr] conj(w)[dc],
// dr = r for r < n, r - m for r > m - n, (zero in between)
template <typename T>
__global__ void bluestein_filter_kernel(cx<T>* __restrict__ out, int n, int m, int sign) {
    const long long total = (long long)m * m;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / m), c = (int)(i - (long long)r * m);
        const int dr = r < n ? r : (r > m - n ? m - r : -1);
        const int dc = c < n ? c : (c > m - n ? m - c : -1);
        cx<T> v = mk<T>((T)0, (T)0);
        if (dr >= 0 && dc >= 0) {
            const double2 w = dmul(chirp(dr, n, -sign), chirp(dc, n, -sign));
            v = mk<T>((T)w.x, (T)w.y);
        }
        out[i] = v;
    }
}

static unsigned grid_for(long long total) {
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}
}  // namespace lp2d

using namespace lp2d;

#define LP2D_LAUNCH_CHECK(what)                                                                        \
    do {                                                                                               \
        ++g_launches;                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                          \
        if (e__ != cudaSuccess) return set_error(LP2D_ECUDA, std::string(what) + ": " + cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

int lp2d_pointwise(int dtype, void* data, long long batch, int rows, int n, int row0, int kind, const double* dev_tabs,
                   double p0, double p1, void* stream) {
    if (data == nullptr || dev_tabs == nullptr || batch < 0 || rows < 0 || n <= 0 || row0 < 0 || row0 + rows > n)
        return set_error(LP2D_EINVAL, "pointwise: bad arguments");
    if (kind != LP2D_PW_H && kind != LP2D_PW_LENS) return set_error(LP2D_EINVAL, "pointwise: kind must be LP2D_PW_H or LP2D_PW_LENS");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "pointwise: bad dtype");
    const long long total = batch * rows * (long long)n;
    if (total == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64)
        pointwise_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<float2*>(data), batch, rows, n, row0, kind, dev_tabs, dev_tabs + n,
                                                                 dev_tabs + 2 * n, p0, p1);
    else
        pointwise_kernel<double><<<grid_for(total), 256, 0, s>>>(static_cast<double2*>(data), batch, rows, n, row0, kind, dev_tabs, dev_tabs + n,
                                                                  dev_tabs + 2 * n, p0, p1);
    LP2D_LAUNCH_CHECK("pointwise launch");
    return LP2D_OK;
}

int lp2d_mul_mask(int dtype, void* data, const void* mask, int mask_kind, long long plane, long long mask_beam_stride, long long batch,
                  void* stream) {
    if (data == nullptr || mask == nullptr || plane < 0 || batch < 0) return set_error(LP2D_EINVAL, "mul_mask: bad arguments");
    if (mask_kind < LP2D_MASK_COMPLEX || mask_kind > LP2D_MASK_U8) return set_error(LP2D_EINVAL, "mul_mask: unknown mask_kind");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "mul_mask: bad dtype");
    const long long total = plane * batch;
    if (total == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64)
        mul_mask_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<float2*>(data), mask, mask_kind, plane, mask_beam_stride, total);
    else
        mul_mask_kernel<double><<<grid_for(total), 256, 0, s>>>(static_cast<double2*>(data), mask, mask_kind, plane, mask_beam_stride, total);
    LP2D_LAUNCH_CHECK("mul_mask launch");
    return LP2D_OK;
}

int lp2d_intensity_n(int dtype, const void* field, void* out, long long count, void* stream) {
    if (field == nullptr || out == nullptr || count < 0) return set_error(LP2D_EINVAL, "intensity_n: bad arguments");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "intensity_n: bad dtype");
    if (count == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64)
        intensity_n_kernel<float><<<grid_for(count), 256, 0, s>>>(static_cast<const float2*>(field), static_cast<float*>(out), count);
    else
        intensity_n_kernel<double><<<grid_for(count), 256, 0, s>>>(static_cast<const double2*>(field), static_cast<double*>(out), count);
    LP2D_LAUNCH_CHECK("intensity_n launch");
    return LP2D_OK;
}

int lp2d_bluestein_size(int n) {
    if (n < 1) return 0;
    int m = 16;
    while (m < 2 * n - 1) m <<= 1;
    return m;
}

int lp2d_bluestein_pre(int dtype, const void* in, void* out, int n, int m, long long batch, int inverse, void* stream) {
    if (in == nullptr || out == nullptr || n < 1 || m < 2 * n - 1 || batch < 0) return set_error(LP2D_EINVAL, "bluestein_pre: bad arguments");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "bluestein_pre: bad dtype");
    const long long total = batch * (long long)m * m;
    if (total == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sign = inverse ? +1 : -1;
    if (dtype == LP2D_C64)
        bluestein_pre_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<const float2*>(in), static_cast<float2*>(out), n, m, batch, sign);
    else
        bluestein_pre_kernel<double><<<grid_for(total), 256, 0, s>>>(static_cast<const double2*>(in), static_cast<double2*>(out), n, m, batch, sign);
    LP2D_LAUNCH_CHECK("bluestein_pre launch");
    return LP2D_OK;
}

int lp2d_bluestein_post(int dtype, const void* in, void* out, int n, int m, long long batch, int inverse, void* stream) {
    if (in == nullptr || out == nullptr || n < 1 || m < 2 * n - 1 || batch < 0) return set_error(LP2D_EINVAL, "bluestein_post: bad arguments");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "bluestein_post: bad dtype");
    const long long total = batch * (long long)n * n;
    if (total == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sign = inverse ? +1 : -1;
    const double scale = inverse ? 1.0 / ((double)n * (double)n) : 1.0;        // numpy ifft2 convention
    if (dtype == LP2D_C64)
        bluestein_post_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<const float2*>(in), static_cast<float2*>(out), n, m, batch, sign, scale);
    else
        bluestein_post_kernel<double><<<grid_for(total), 256, 0, s>>>(static_cast<const double2*>(in), static_cast<double2*>(out), n, m, batch, sign, scale);
    LP2D_LAUNCH_CHECK("bluestein_post launch");
    return LP2D_OK;
}

int lp2d_bluestein_filter(int dtype, void* out, int n, int m, int inverse, void* stream) {
    if (out == nullptr || n < 1 || m < 2 * n - 1) return set_error(LP2D_EINVAL, "bluestein_filter: bad arguments");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "bluestein_filter: bad dtype");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sign = inverse ? +1 : -1;
    const long long total = (long long)m * m;
    if (dtype == LP2D_C64)
        bluestein_filter_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<float2*>(out), n, m, sign);
    else
        bluestein_filter_kernel<double><<<grid_for(total), 256, 0, s>>>(static_cast<double2*>(out), n, m, sign);
    LP2D_LAUNCH_CHECK("bluestein_filter launch");
    return LP2D_OK;
}

}  // extern "C"
