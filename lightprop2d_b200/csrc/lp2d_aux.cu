// Pointwise / resampling / reduction kernels around the transform passes:
// expand (edge-replicating pad), coarse (block mean), phase, intensity moments,
// and the SIMT form of the mode projection / synthesis.
#include <cstdlib>
#include <string>

#include "../../include/lp2d.h"
#include "dispatch.h"

namespace lp2d {
int set_error(int code, const std::string& msg);   // lp2d_api.cu
int modes_project_tc(const void* field, const void* modes, void* coeffs, int n, int nb, int nmodes, long long batch,
                     cudaStream_t s);              // lp2d_modes_tc.cu
int modes_synthesize_tc(const void* coeffs, const void* modes, void* field, int n, int nb, int nmodes, long long batch,
                        cudaStream_t s);           // lp2d_modes_tc.cu

// out[b][r][c] = in[b][clamp(r-pad)][clamp(c-pad)]      (beam2d.py:392-401)
template <typename T>
__global__ void expand_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n_in, int n_out,
                              long long total) {
    const int pad = (n_out - n_in) / 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % n_out);
        const long long rb = i / n_out;
        const int r = (int)(rb % n_out);
        const long long b = rb / n_out;
        int rs = r - pad; rs = rs < 0 ? 0 : (rs >= n_in ? n_in - 1 : rs);
        int cs = c - pad; cs = cs < 0 ? 0 : (cs >= n_in ? n_in - 1 : cs);
        out[i] = in[(b * n_in + rs) * (long long)n_in + cs];
    }
}

// Transposing tile pack of the slab exchange: out[c][i][r] = in[r][c*tile + i].  32 x 32 samples per
// CTA through padded shared memory, coalesced on both sides.
template <typename T>
__global__ void pack_transpose_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int rows, int ld, int tile) {
    __shared__ cx<T> s[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;          // column / row origin of this 32 x 32 block
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int r = r0 + k, c = c0 + threadIdx.x;
        if (r < rows && c < ld) s[k][threadIdx.x] = in[(long long)r * ld + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int c = c0 + k, r = r0 + threadIdx.x;                // output row = input column
        if (r < rows && c < ld) {
            const int t = c / tile, i = c - t * tile;
            out[((long long)t * tile + i) * rows + r] = s[threadIdx.x][k];
        }
    }
}

// out[b][c][r] = in[b][r][c], n x n planes: 32 x 32 samples per CTA through padded shared memory (conversion between
// the natural and the transposed spectrum layout; not on the step path)
template <typename T>
__global__ void transpose_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n, long long batch) {
    __shared__ cx<T> s[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
        const cx<T>* src = in + b * (long long)n * n;
        cx<T>* dst = out + b * (long long)n * n;
        for (int k = threadIdx.y; k < 32; k += blockDim.y) {
            const int r = r0 + k, c = c0 + threadIdx.x;
            if (r < n && c < n) s[k][threadIdx.x] = src[(long long)r * n + c];
        }
        __syncthreads();
        for (int k = threadIdx.y; k < 32; k += blockDim.y) {
            const int c = c0 + k, r = r0 + threadIdx.x;
            if (r < n && c < n) dst[(long long)c * n + r] = s[threadIdx.x][k];
        }
        __syncthreads();
    }
}

// Counter-based RNG (Philox4x32-10, Salmon et al. 2011) for ensembles generated on the device: the four
// words of block `ctr` of beam `beam` depend only on (seed, beam, ctr), so a beam is the same whatever the
// batch / shard it is generated in.
__device__ __forceinline__ void philox4x32_10(unsigned int (&c)[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// random_round_hole (kind 0: amplitude uniform in [0,1) inside the aperture) / random_round_hole_phase
// (kind 1: exp(2 pi i u) inside the aperture), zero outside (profiles.py:94-113); the aperture mask is the
// host-built round_hole array (exactly the reference's comparison), one byte per grid point.
template <typename T>
__global__ void random_fields_kernel(cx<T>* __restrict__ out, const unsigned char* __restrict__ mask, long long plane,
                                     long long batch, unsigned long long first_beam, unsigned long long seed, int kind) {
    const long long quads = (plane + 3) / 4;
    const long long total = quads * batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / quads, q = i - b * quads;
        const unsigned long long beam = first_beam + (unsigned long long)b;
        unsigned int c[4] = {(unsigned int)q, (unsigned int)((unsigned long long)q >> 32), (unsigned int)beam, (unsigned int)(beam >> 32)};
        if (sizeof(T) == 8) {
            // two 32-bit words per sample: two Philox blocks per quad
            unsigned int d[4] = {c[0], c[1] | 0x80000000u, c[2], c[3]};
            philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
            philox4x32_10(d, (unsigned int)seed, (unsigned int)(seed >> 32));
            const unsigned int w[8] = {c[0], c[1], c[2], c[3], d[0], d[1], d[2], d[3]};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long idx = q * 4 + j;
                if (idx >= plane) break;
                const unsigned long long bits = ((unsigned long long)w[2 * j] << 21) ^ (unsigned long long)(w[2 * j + 1] >> 11);   // 53 bits
                const double u = (double)(bits & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
                cx<T> v = mk<T>((T)0, (T)0);
                if (mask[idx]) {
                    if (kind == 0) v = mk<T>((T)u, (T)0);
                    else { double sn, cs; sincospi(2.0 * u, &sn, &cs); v = mk<T>((T)cs, (T)sn); }
                }
                out[b * plane + idx] = v;
            }
        } else {
            philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long idx = q * 4 + j;
                if (idx >= plane) break;
                const float u = (float)(c[j] >> 8) * (1.0f / 16777216.0f);                // 24 bits, [0, 1)
                cx<T> v = mk<T>((T)0, (T)0);
                if (mask[idx]) {
                    if (kind == 0) v = mk<T>((T)u, (T)0);
                    else { float sn, cs; sincospif(2.0f * u, &sn, &cs); v = mk<T>((T)cs, (T)sn); }
                }
                out[b * plane + idx] = v;
            }
        }
    }
}

// out[b][r][c] = mean of the m x m block                  (beam2d.py:456-459)
template <typename T>
__global__ void coarse_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, int n_in, int m,
                              long long total) {
    const int n_out = n_in / m;
    const T inv = (T)1 / (T)(m * m);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % n_out);
        const long long rb = i / n_out;
        const int r = (int)(rb % n_out);
        const long long b = rb / n_out;
        const cx<T>* src = in + (b * n_in + (long long)r * m) * n_in + (long long)c * m;
        T sx = 0, sy = 0;
        for (int dr = 0; dr < m; ++dr)
            for (int dc = 0; dc < m; ++dc) {
                const cx<T> v = src[(long long)dr * n_in + dc];
                sx += v.x; sy += v.y;
            }
        out[i] = mk<T>(sx * inv, sy * inv);
    }
}

template <typename T>
__global__ void phase_kernel(const cx<T>* __restrict__ f, T* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const cx<T> v = f[i];
        out[i] = atan2(v.y, v.x);
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[0..4] += [sum I, sum x I, sum y I, sum (x-xc)^2 I, sum (y-yc)^2 I],  x = (i - n/2) dL
// (the mgrid[-n:n] * dL grids of beam2d.py:680-684, 696-702), I in the field's real type.
template <typename T>
__global__ void moments_kernel(const cx<T>* __restrict__ f, int n, double dL, double xc, double yc,
                               double* __restrict__ out) {
    double acc[5] = {0, 0, 0, 0, 0};
    const long long total = (long long)n * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int h = n / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % n), r = (int)(i / n);
        const cx<T> v = f[i];
        const double I = (double)(T)(v.x * v.x + v.y * v.y);
        const double x = (double)(c - h) * dL, y = (double)(r - h) * dL;
        acc[0] += I;
        acc[1] += x * I;
        acc[2] += y * I;
        acc[3] += (x - xc) * (x - xc) * I;
        acc[4] += (y - yc) * (y - yc) * I;
    }
    __shared__ double sm[5][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double s = warp_sum(acc[k]);
        if (lane == 0) sm[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += sm[threadIdx.x][q];
        atomicAdd(&out[threadIdx.x], s);
    }
}

// SIMT mode projection: coeffs[b][k] = sum_p M[k][p] * field_b[crop(p)]  (no conjugate, beam2d.py:592-600).
// One CTA per (mode, tile of BT beams): a mode sample is loaded ONCE per BT beams (the mode matrix is the big
// operand: nmodes x nb^2), the BT field samples of the same grid point are coalesced loads of BT planes.
// FP64 accumulation for both field types.
template <typename T, int BT>
__global__ void __launch_bounds__(256) modes_project_kernel(const cx<T>* __restrict__ field, const cx<T>* __restrict__ modes,
                                                            cx<T>* __restrict__ coeffs, int n, int nb, int nmodes, long long batch) {
    const int k = blockIdx.x;
    const long long b0 = (long long)blockIdx.y * BT;
    const int dn = (n - nb) / 2;
    const long long plane = (long long)n * n;
    const cx<T>* mrow = modes + (long long)k * nb * nb;
    const cx<T>* f = field + b0 * plane;
    const int nbt = (batch - b0 < BT) ? (int)(batch - b0) : BT;
    double ax[BT], ay[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) ax[j] = ay[j] = 0.0;
    for (int p = threadIdx.x; p < nb * nb; p += blockDim.x) {
        const int r = p / nb, c = p % nb;
        const cx<T> mv = __ldcs(&mrow[p]);
        const long long off = (long long)(r + dn) * n + (c + dn);
#pragma unroll
        for (int j = 0; j < BT; ++j) {
            if (j < nbt) {
                const cx<T> fv = f[j * plane + off];
                ax[j] += (double)mv.x * fv.x - (double)mv.y * fv.y;
                ay[j] += (double)mv.x * fv.y + (double)mv.y * fv.x;
            }
        }
    }
    __shared__ double sm[2][BT][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < BT; ++j) {
        const double sx = warp_sum(ax[j]), sy = warp_sum(ay[j]);
        if (lane == 0) { sm[0][j][w] = sx; sm[1][j][w] = sy; }
    }
    __syncthreads();
    if (threadIdx.x < nbt) {
        double sx = 0, sy = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { sx += sm[0][threadIdx.x][q]; sy += sm[1][threadIdx.x][q]; }
        coeffs[(b0 + threadIdx.x) * nmodes + k] = mk<T>((T)sx, (T)sy);
    }
}

// Tiled complex128 projection (the tensor cores have no FP64 path worth a 3-term split here): a register-tiled GEMM
// C[b][m] += sum_k F[b][crop(k)] * M[m][k] over a slice of k.  CTA = 4 warps: a warp owns 32 consecutive modes x MT
// (mode tm + 32 i) and BT beams, so a k step costs MT conflict-free + BT broadcast shared-memory loads for MT*BT
// complex FMAs (4 DFMA each) -- the FP64 pipe, not memory, bounds it (8*Bt*M*K flop against (M + Bt)*K*16 bytes:
// 7.5 flop/byte at 16 beams).  The mode matrix is read ONCE per 16 beams (the one-mode-per-CTA kernel above re-reads
// every field sample once per mode through L2: 5.6 ms at 256 modes x 512^2 x 16 beams, profiles/r3h_*).  k is split
// over gridDim.z CTAs; partial sums are added with FP64 atomics into the zeroed coefficient array.
template <int MT, int BT, int KC>
__global__ void __launch_bounds__(128) modes_project_f64_kernel(const double2* __restrict__ field, const double2* __restrict__ modes,
                                                                double2* __restrict__ coeffs, int n, int nb, int nmodes, long long batch,
                                                                int kslice) {
    constexpr int MTILE = 32 * MT, BTILE = 4 * BT;
    __shared__ double2 Ms[KC][MTILE + 1];
    __shared__ double2 Fs[KC][BTILE];
    const int tid = threadIdx.x, tm = tid & 31, tb = tid >> 5;
    const int m0 = blockIdx.x * MTILE;
    const long long b0 = (long long)blockIdx.y * BTILE;
    const long long kk = (long long)nb * nb;
    const long long k_begin = (long long)blockIdx.z * kslice;
    const long long k_end = (k_begin + kslice < kk) ? k_begin + kslice : kk;
    const int dn = (n - nb) / 2;
    const long long plane = (long long)n * n;
    double2 acc[MT][BT];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < BT; ++j) acc[i][j] = make_double2(0.0, 0.0);
    for (long long k0 = k_begin; k0 < k_end; k0 += KC) {                // (KC divides nb and kslice: a chunk lies in one crop row)
        __syncthreads();
        // modes tile: MTILE rows of KC consecutive samples, stored transposed
        for (int q = tid; q < MTILE * KC; q += 128) {
            const int m = q / KC, kq = q % KC;
            Ms[kq][m] = (m0 + m < nmodes) ? __ldcs(&modes[(long long)(m0 + m) * kk + k0 + kq]) : make_double2(0.0, 0.0);
        }
        const int r = (int)(k0 / nb), c = (int)(k0 % nb);
        const long long off = (long long)(r + dn) * n + (c + dn);
        for (int q = tid; q < BTILE * KC; q += 128) {
            const int b = q / KC, kq = q % KC;
            Fs[kq][b] = (b0 + b < batch) ? field[(b0 + b) * plane + off + kq] : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll 4
        for (int kq = 0; kq < KC; ++kq) {
            double2 mv[MT], fv[BT];
#pragma unroll
            for (int i = 0; i < MT; ++i) mv[i] = Ms[kq][tm + 32 * i];
#pragma unroll
            for (int j = 0; j < BT; ++j) fv[j] = Fs[kq][tb * BT + j];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < BT; ++j) {
                    acc[i][j].x += mv[i].x * fv[j].x - mv[i].y * fv[j].y;
                    acc[i][j].y += mv[i].x * fv[j].y + mv[i].y * fv[j].x;
                }
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int m = m0 + tm + 32 * i;
#pragma unroll
        for (int j = 0; j < BT; ++j) {
            const long long b = b0 + tb * BT + j;
            if (m < nmodes && b < batch) {
                double* dst = reinterpret_cast<double*>(&coeffs[b * nmodes + m]);
                atomicAdd(dst, acc[i][j].x);
                atomicAdd(dst + 1, acc[i][j].y);
            }
        }
    }
}

// Flat form for a strip of the crop (a slab of a grid decomposed over several GPUs): coeffs[k] = sum_{p < K}
// modes[k * ldm + p] * f[p]; the partial sums of the ranks are added by the caller (one all-reduce).
template <typename T>
__global__ void __launch_bounds__(256) modes_project_flat_kernel(const cx<T>* __restrict__ f, const cx<T>* __restrict__ modes,
                                                                 cx<T>* __restrict__ coeffs, long long K, long long ldm) {
    const int k = blockIdx.x;
    const cx<T>* mrow = modes + (long long)k * ldm;
    double ax = 0, ay = 0;
    for (long long p = threadIdx.x; p < K; p += blockDim.x) {
        const cx<T> mv = __ldcs(&mrow[p]);
        const cx<T> fv = f[p];
        ax += (double)mv.x * fv.x - (double)mv.y * fv.y;
        ay += (double)mv.x * fv.y + (double)mv.y * fv.x;
    }
    __shared__ double sm[2][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    ax = warp_sum(ax); ay = warp_sum(ay);
    if (lane == 0) { sm[0][w] = ax; sm[1][w] = ay; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sx = 0, sy = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { sx += sm[0][q]; sy += sm[1][q]; }
        coeffs[k] = mk<T>((T)sx, (T)sy);
    }
}

// SIMT mode synthesis: field_b[r][c] = sum_k coeffs[b][k] * M[k][crop^-1(r, c)], zero outside the
// centred nb x nb window (beam2d.py:552-562, 664-666).  The mode matrix is the big operand (nmodes x nb^2):
// a thread owns one grid point and BT beams, so a mode sample is loaded once per BT beams; the coefficients
// of the beam tile are staged in shared memory MC modes at a time and read as warp broadcasts.  Accumulation
// in the field's real type (at most a few hundred terms per sum).
template <typename T, int BT, int MC>
__global__ void __launch_bounds__(256) modes_synthesize_kernel(const cx<T>* __restrict__ coeffs, const cx<T>* __restrict__ modes,
                                                               cx<T>* __restrict__ field, int n, int nb, int nmodes, long long batch) {
    __shared__ cx<T> sc[MC][BT];
    const long long b0 = (long long)blockIdx.y * BT;
    const int dn = (n - nb) / 2;
    const int nt = n - dn;
    const long long total = (long long)n * n;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(i % n), r = (int)(i / n);
    const bool inside = i < total && r >= dn && r < nt && c >= dn && c < nt;
    const long long p = inside ? (long long)(r - dn) * nb + (c - dn) : 0;
    const long long kk = (long long)nb * nb;
    cx<T> acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = mk<T>((T)0, (T)0);
    for (int k0 = 0; k0 < nmodes; k0 += MC) {
        __syncthreads();
        for (int q = threadIdx.x; q < MC * BT; q += blockDim.x) {
            const int k = q / BT, j = q % BT;
            sc[k][j] = (k0 + k < nmodes && b0 + j < batch) ? coeffs[(b0 + j) * nmodes + k0 + k] : mk<T>((T)0, (T)0);
        }
        __syncthreads();
        if (inside) {
            const int kn = (nmodes - k0 < MC) ? nmodes - k0 : MC;
            for (int k = 0; k < kn; ++k) {
                const cx<T> mv = __ldg(&modes[(long long)(k0 + k) * kk + p]);
#pragma unroll
                for (int j = 0; j < BT; ++j) {
                    const cx<T> cv = sc[k][j];
                    acc[j].x += mv.x * cv.x - mv.y * cv.y;
                    acc[j].y += mv.x * cv.y + mv.y * cv.x;
                }
            }
        }
    }
    if (i < total) {
#pragma unroll
        for (int j = 0; j < BT; ++j)
            if (b0 + j < batch) __stcs(&field[(b0 + j) * total + i], acc[j]);
    }
}

static unsigned grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = 148ll * 32;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

static int done(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return LP2D_OK;
}
}  // namespace lp2d

using namespace lp2d;

#define LP2D_DTYPE_OK(d) ((d) == LP2D_C64 || (d) == LP2D_C128)

extern "C" {

int lp2d_expand(int dtype, const void* in, void* out, int n_in, int n_out, long long batch, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || in == nullptr || out == nullptr || n_in <= 0 || n_out < n_in || ((n_out - n_in) & 1) || batch < 0)
        return set_error(LP2D_EINVAL, "expand: bad arguments (n_out - n_in must be even and >= 0)");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long total = batch * (long long)n_out * n_out;
    if (dtype == LP2D_C64)
        expand_kernel<float><<<grid_for(total, 256), 256, 0, s>>>((const float2*)in, (float2*)out, n_in, n_out, total);
    else
        expand_kernel<double><<<grid_for(total, 256), 256, 0, s>>>((const double2*)in, (double2*)out, n_in, n_out, total);
    return done("expand launch");
}

int lp2d_pack_transpose(int dtype, const void* in, void* out, int rows, int ld, int tile, void* stream) {
    if (in == nullptr || out == nullptr || rows <= 0 || ld <= 0 || tile <= 0 || ld % tile != 0 || in == out)
        return set_error(LP2D_EINVAL, "pack_transpose: bad arguments (tile must divide ld, buffers distinct)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)((ld + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    if (dtype == LP2D_C64)
        pack_transpose_kernel<float><<<grid, block, 0, s>>>(static_cast<const float2*>(in), static_cast<float2*>(out), rows, ld, tile);
    else if (dtype == LP2D_C128)
        pack_transpose_kernel<double><<<grid, block, 0, s>>>(static_cast<const double2*>(in), static_cast<double2*>(out), rows, ld, tile);
    else
        return set_error(LP2D_EINVAL, "pack_transpose: unknown dtype");
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("pack_transpose launch: ") + cudaGetErrorString(e));
    return LP2D_OK;
}

int lp2d_transpose(int dtype, const void* in, void* out, int n, long long batch, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || in == nullptr || out == nullptr || n <= 0 || batch < 0 || in == out)
        return set_error(LP2D_EINVAL, "transpose: bad arguments (buffers must be distinct)");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32), (unsigned)(batch < 1024 ? batch : 1024)), block(32, 8);
    if (dtype == LP2D_C64) transpose_kernel<float><<<grid, block, 0, s>>>(static_cast<const float2*>(in), static_cast<float2*>(out), n, batch);
    else transpose_kernel<double><<<grid, block, 0, s>>>(static_cast<const double2*>(in), static_cast<double2*>(out), n, batch);
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("transpose launch: ") + cudaGetErrorString(e));
    return LP2D_OK;
}

int lp2d_random_fields(int dtype, void* out, const void* mask_u8, int n, long long batch, unsigned long long first_beam,
                       unsigned long long seed, int kind, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || out == nullptr || mask_u8 == nullptr || n <= 0 || batch < 0 || (kind != 0 && kind != 1))
        return set_error(LP2D_EINVAL, "random_fields: bad arguments");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long plane = (long long)n * n;
    const unsigned grid = grid_for(((plane + 3) / 4) * batch, 256);
    if (dtype == LP2D_C64)
        random_fields_kernel<float><<<grid, 256, 0, s>>>(static_cast<float2*>(out), static_cast<const unsigned char*>(mask_u8), plane, batch, first_beam, seed, kind);
    else
        random_fields_kernel<double><<<grid, 256, 0, s>>>(static_cast<double2*>(out), static_cast<const unsigned char*>(mask_u8), plane, batch, first_beam, seed, kind);
    return done("random_fields launch");
}

int lp2d_coarse(int dtype, const void* in, void* out, int n_in, int m, long long batch, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || in == nullptr || out == nullptr || n_in <= 0 || m <= 0 || n_in % m != 0 || batch < 0)
        return set_error(LP2D_EINVAL, "coarse: bad arguments (mean_order must divide npoints)");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long total = batch * (long long)(n_in / m) * (n_in / m);
    if (dtype == LP2D_C64)
        coarse_kernel<float><<<grid_for(total, 256), 256, 0, s>>>((const float2*)in, (float2*)out, n_in, m, total);
    else
        coarse_kernel<double><<<grid_for(total, 256), 256, 0, s>>>((const double2*)in, (double2*)out, n_in, m, total);
    return done("coarse launch");
}

int lp2d_phase(int dtype, const void* field, void* out, long long count, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || field == nullptr || out == nullptr || count < 0)
        return set_error(LP2D_EINVAL, "phase: bad arguments");
    if (count == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64)
        phase_kernel<float><<<grid_for(count, 256), 256, 0, s>>>((const float2*)field, (float*)out, count);
    else
        phase_kernel<double><<<grid_for(count, 256), 256, 0, s>>>((const double2*)field, (double*)out, count);
    return done("phase launch");
}

int lp2d_moments(int dtype, const void* field, int n, double dL, double xc, double yc, double* out5, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || field == nullptr || out5 == nullptr || n <= 0)
        return set_error(LP2D_EINVAL, "moments: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out5, 0, 5 * sizeof(double), s);
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("moments memset: ") + cudaGetErrorString(e));
    const unsigned grid = grid_for((long long)n * n, 256 * 8);
    if (dtype == LP2D_C64)
        moments_kernel<float><<<grid, 256, 0, s>>>((const float2*)field, n, dL, xc, yc, out5);
    else
        moments_kernel<double><<<grid, 256, 0, s>>>((const double2*)field, n, dL, xc, yc, out5);
    return done("moments launch");
}

int lp2d_modes_project(int dtype, const void* field, const void* modes, void* coeffs, int n, int nb, int nmodes,
                       long long batch, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || field == nullptr || modes == nullptr || coeffs == nullptr || n <= 0 || nb <= 0 || nb > n ||
        nmodes <= 0 || batch < 0 || batch > 65535)
        return set_error(LP2D_EINVAL, "modes_project: bad arguments");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64) {
        // tensor-core form (tcgen05, 3xTF32) for complex64 when the crop rows are a multiple of the K block
        const char* env = getenv("LP2D_MODES_TC");
        if (env == nullptr || atoi(env) != 0) {
            const int rc = modes_project_tc(field, modes, coeffs, n, nb, nmodes, batch, s);
            if (rc != LP2D_EUNSUPPORTED) return rc;
        }
    }
    if (dtype == LP2D_C128 && nb % 16 == 0) {
        // register-tiled FP64 GEMM, k split so that the grid fills the GPU a few times over; up to 4 beams: one beam
        // per warp (the call is a matrix-vector product bound by the read of the mode matrix)
        constexpr int MT = 4, KC = 16;
        const int btt = batch <= 4 ? 1 : 4;
        const long long kk = (long long)nb * nb;
        const unsigned gx = (unsigned)((nmodes + 32 * MT - 1) / (32 * MT)), gy = (unsigned)((batch + 4 * btt - 1) / (4 * btt));
        long long splits = (148ll * 4 + (long long)gx * gy - 1) / ((long long)gx * gy);
        long long kslice = ((kk + splits - 1) / splits + KC - 1) / KC * KC;
        if (kslice < 4 * KC) kslice = 4 * KC;
        splits = (kk + kslice - 1) / kslice;
        cudaError_t e = cudaMemsetAsync(coeffs, 0, (size_t)batch * nmodes * sizeof(double2), s);
        if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_project memset: ") + cudaGetErrorString(e));
        const dim3 grid(gx, gy, (unsigned)splits);
        if (btt == 1)
            modes_project_f64_kernel<MT, 1, KC><<<grid, 128, 0, s>>>((const double2*)field, (const double2*)modes, (double2*)coeffs, n, nb, nmodes, batch, (int)kslice);
        else
            modes_project_f64_kernel<MT, 4, KC><<<grid, 128, 0, s>>>((const double2*)field, (const double2*)modes, (double2*)coeffs, n, nb, nmodes, batch, (int)kslice);
        return done("modes_project (tiled FP64) launch");
    }
    constexpr int BT = 8;
    dim3 grid((unsigned)nmodes, (unsigned)((batch + BT - 1) / BT));
    if (dtype == LP2D_C64)
        modes_project_kernel<float, BT><<<grid, 256, 0, s>>>((const float2*)field, (const float2*)modes, (float2*)coeffs, n, nb, nmodes, batch);
    else
        modes_project_kernel<double, BT><<<grid, 256, 0, s>>>((const double2*)field, (const double2*)modes, (double2*)coeffs, n, nb, nmodes, batch);
    return done("modes_project launch");
}

int lp2d_modes_project_flat(int dtype, const void* flat_field, const void* modes, void* coeffs, long long count, long long ldm,
                            int nmodes, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || flat_field == nullptr || modes == nullptr || coeffs == nullptr || count < 0 || ldm < count || nmodes <= 0)
        return set_error(LP2D_EINVAL, "modes_project_flat: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64)
        modes_project_flat_kernel<float><<<(unsigned)nmodes, 256, 0, s>>>((const float2*)flat_field, (const float2*)modes, (float2*)coeffs, count, ldm);
    else
        modes_project_flat_kernel<double><<<(unsigned)nmodes, 256, 0, s>>>((const double2*)flat_field, (const double2*)modes, (double2*)coeffs, count, ldm);
    return done("modes_project_flat launch");
}

int lp2d_modes_synthesize(int dtype, const void* coeffs, const void* modes, void* field, int n, int nb, int nmodes,
                          long long batch, void* stream) {
    if (!LP2D_DTYPE_OK(dtype) || field == nullptr || modes == nullptr || coeffs == nullptr || n <= 0 || nb <= 0 || nb > n ||
        nmodes <= 0 || batch < 0 || batch > 65535)
        return set_error(LP2D_EINVAL, "modes_synthesize: bad arguments");
    if (batch == 0) return LP2D_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == LP2D_C64) {
        // tensor-core form (tcgen05, 3xTF32) when the crop is made of whole 32-point runs
        const char* env = getenv("LP2D_MODES_TC");
        if (env == nullptr || atoi(env) != 0) {
            const int rc = modes_synthesize_tc(coeffs, modes, field, n, nb, nmodes, batch, s);
            if (rc != LP2D_EUNSUPPORTED) return rc;
        }
    }
    const unsigned gx = (unsigned)(((long long)n * n + 255) / 256);
    if (dtype == LP2D_C64) {
        constexpr int BT = 8;
        dim3 grid(gx, (unsigned)((batch + BT - 1) / BT));
        modes_synthesize_kernel<float, BT, 128><<<grid, 256, 0, s>>>((const float2*)coeffs, (const float2*)modes, (float2*)field, n, nb, nmodes, batch);
    } else if (batch > 4) {
        // 16 beams per thread: the mode matrix (the big operand) is read once per 16 beams (4: 4.3 GB of DRAM reads
        // for 1.07 GB of modes at 256 modes x 512^2 x 16 beams, profiles/r3h_*)
        constexpr int BT = 16;
        dim3 grid(gx, (unsigned)((batch + BT - 1) / BT));
        modes_synthesize_kernel<double, BT, 64><<<grid, 256, 0, s>>>((const double2*)coeffs, (const double2*)modes, (double2*)field, n, nb, nmodes, batch);
    } else {
        constexpr int BT = 4;
        dim3 grid(gx, (unsigned)((batch + BT - 1) / BT));
        modes_synthesize_kernel<double, BT, 128><<<grid, 256, 0, s>>>((const double2*)coeffs, (const double2*)modes, (double2*)field, n, nb, nmodes, batch);
    }
    return done("modes_synthesize launch");
}

}  // extern "C"
