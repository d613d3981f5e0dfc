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

// ---------------------------------------------------------------------------------------------------------------
// Small dense solve and bilinear resampling (the "next" rows of SURVEY.md section 8f).
namespace lp2d {

// A X = B for one M x M complex128 system with nrhs right-hand sides, LU with partial pivoting, ONE CTA (the systems
// of the mode decomposition are a few hundred unknowns: beam2d.py:601-603, 649-650).  A and B are overwritten
// (B by the solution).  info[0] = 0, or k + 1 when the k-th pivot is exactly zero.
__global__ void __launch_bounds__(1024) lu_solve_kernel(double2* __restrict__ A, double2* __restrict__ B, int M, int nrhs, int* info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ double2 s_inv;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) info[0] = 0;
    for (int k = 0; k < M; ++k) {
        // pivot search in column k
        double best = -1.0; int bi = k;
        for (int r = k + tid; r < M; r += nt) {
            const double2 v = A[(long long)r * M + k];
            const double a = fabs(v.x) + fabs(v.y);              // LAPACK's cabs1
            if (a > best) { best = a; bi = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = s_val[0]; int i = s_idx[0];
            for (int w = 1; w < (nt + 31) / 32; ++w)
                if (s_val[w] > b || (s_val[w] == b && s_idx[w] < i)) { b = s_val[w]; i = s_idx[w]; }
            s_piv = i;
            if (b == 0.0 && info[0] == 0) info[0] = k + 1;
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k) {                                               // swap rows k and p (columns k.., and B)
            for (int c = k + tid; c < M; c += nt) {
                const double2 t = A[(long long)k * M + c]; A[(long long)k * M + c] = A[(long long)p * M + c]; A[(long long)p * M + c] = t;
            }
            for (int c = tid; c < nrhs; c += nt) {
                const double2 t = B[(long long)k * nrhs + c]; B[(long long)k * nrhs + c] = B[(long long)p * nrhs + c]; B[(long long)p * nrhs + c] = t;
            }
        }
        __syncthreads();
        if (tid == 0) {
            const double2 d = A[(long long)k * M + k];
            const double den = d.x * d.x + d.y * d.y;
            s_inv = den > 0.0 ? make_double2(d.x / den, -d.y / den) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const double2 inv = s_inv;
        // multipliers, then the rank-1 update of the trailing block and of B
        for (int r = k + 1 + tid; r < M; r += nt) A[(long long)r * M + k] = dmul(A[(long long)r * M + k], inv);
        __syncthreads();
        const int rem = M - k - 1;
        for (long long e = tid; e < (long long)rem * (rem + nrhs); e += nt) {
            const int r = k + 1 + (int)(e / (rem + nrhs));
            const int c = (int)(e % (rem + nrhs));
            const double2 l = A[(long long)r * M + k];
            if (c < rem) {
                const double2 u = A[(long long)k * M + k + 1 + c];
                double2& t = A[(long long)r * M + k + 1 + c];
                const double2 pr = dmul(l, u);
                t.x -= pr.x; t.y -= pr.y;
            } else {
                const double2 u = B[(long long)k * nrhs + (c - rem)];
                double2& t = B[(long long)r * nrhs + (c - rem)];
                const double2 pr = dmul(l, u);
                t.x -= pr.x; t.y -= pr.y;
            }
        }
        __syncthreads();
    }
    // back substitution: U X = B
    for (int k = M - 1; k >= 0; --k) {
        if (tid == 0) {
            const double2 d = A[(long long)k * M + k];
            const double den = d.x * d.x + d.y * d.y;
            s_inv = den > 0.0 ? make_double2(d.x / den, -d.y / den) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const double2 inv = s_inv;
        for (int c = tid; c < nrhs; c += nt) B[(long long)k * nrhs + c] = dmul(B[(long long)k * nrhs + c], inv);
        __syncthreads();
        for (long long e = tid; e < (long long)k * nrhs; e += nt) {
            const int r = (int)(e / nrhs), c = (int)(e % nrhs);
            const double2 pr = dmul(A[(long long)r * M + k], B[(long long)k * nrhs + c]);
            double2& t = B[(long long)r * nrhs + c];
            t.x -= pr.x; t.y -= pr.y;
        }
        __syncthreads();
    }
}

// out[b][r][c] = sum_{i, j in {0, 1}} in[b][iy[r] + i][ix[c] + j] * wy[i][r] * wx[j][c]: the tensor-product linear
// spline FITPACK evaluates for the reference's interp2d (beam2d.py:427-434, 527-535), weights and knot intervals from
// the host (they depend on 2 n numbers only), summed in FITPACK's order ((c * hx) * hy, i outer, j inner) in FP64.
template <typename T>
__global__ void resample_kernel(const cx<T>* __restrict__ in, long long in_plane, int ld_in, cx<T>* __restrict__ out, int n_out,
                                const int* __restrict__ ix, const double* __restrict__ wx, const int* __restrict__ iy,
                                const double* __restrict__ wy, long long batch) {
    const long long plane = (long long)n_out * n_out;
    const long long total = batch * plane;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / plane, q = e - b * plane;
        const int r = (int)(q / n_out), c = (int)(q - (long long)r * n_out);
        const cx<T>* src = in + b * in_plane + (long long)iy[r] * ld_in + ix[c];
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {                  // x interval (FITPACK's first index = our columns)
#pragma unroll
            for (int j = 0; j < 2; ++j) {              // y interval
                const cx<T> v = src[(long long)j * ld_in + i];
                const double hx = wx[i * n_out + c], hy = wy[j * n_out + r];
                sx += ((double)v.x * hx) * hy;
                sy += ((double)v.y * hx) * hy;
            }
        }
        out[e] = mk<T>((T)sx, (T)sy);
    }
}
}  // namespace lp2d

extern "C" {

int lp2d_solve(void* A, void* B, int m, int nrhs, int* dev_info, void* stream) {
    if (A == nullptr || B == nullptr || dev_info == nullptr || m < 1 || nrhs < 1) return set_error(LP2D_EINVAL, "solve: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    lu_solve_kernel<<<1, 1024, 0, s>>>(static_cast<double2*>(A), static_cast<double2*>(B), m, nrhs, dev_info);
    LP2D_LAUNCH_CHECK("solve launch");
    return LP2D_OK;
}

int lp2d_resample(int dtype, const void* in, long long in_plane, int ld_in, void* out, int n_out, const int* dev_ix,
                  const double* dev_wx, const int* dev_iy, const double* dev_wy, long long batch, void* stream) {
    if (in == nullptr || out == nullptr || dev_ix == nullptr || dev_wx == nullptr || dev_iy == nullptr || dev_wy == nullptr || n_out < 1 ||
        ld_in < 2 || batch < 0)
        return set_error(LP2D_EINVAL, "resample: bad arguments");
    if (dtype != LP2D_C64 && dtype != LP2D_C128) return set_error(LP2D_EINVAL, "resample: bad dtype");
    const long long total = batch * (long long)n_out * n_out;
    if (total == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64)
        resample_kernel<float><<<grid_for(total), 256, 0, s>>>(static_cast<const float2*>(in), in_plane, ld_in, static_cast<float2*>(out), n_out,
                                                                dev_ix, dev_wx, dev_iy, dev_wy, batch);
    else
        resample_kernel<double><<<grid_for(total), 256, 0, s>>>(static_cast<const double2*>(in), in_plane, ld_in, static_cast<double2*>(out), n_out,
                                                                 dev_ix, dev_wx, dev_iy, dev_wy, batch);
    LP2D_LAUNCH_CHECK("resample launch");
    return LP2D_OK;
}

}  // extern "C"
