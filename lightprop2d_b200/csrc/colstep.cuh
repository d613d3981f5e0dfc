// Long columns in two steps ("four-step" split of the column transform), for the grids whose column tile in the
// one-kernel column pass is only 16-32 bytes wide (a whole column must sit in one CTA there: 4096 complex128 samples
// are 64 KB of registers plus a 64 KB exchange buffer, so only two columns fit, and every grid row is touched in
// 32-byte pieces: 0.25-0.4 of the HBM peak on the pass's own bytes, profiles/r1d_ncu_full_rowcol_4096c128_summary.txt).
//
// A column of N = A * B samples, j = j1 * B + j2, k = k1 + A * k2:
//     X[k1 + A k2] = sum_j2 W_B^(j2 k2) * [ W_N^(j2 k1) * sum_j1 x[j1 B + j2] W_A^(j1 k1) ]
//   step 1: for every (j2, column): A-point transform over j1 (rows B apart), times the twiddle W_N^(j2 k1);
//           result (k1, j2) written to row k1 * B + j2 of the scratch plane;
//   step 2: for every (k1, column): B-point transform over j2 (consecutive rows of the scratch), result k2 written to
//           row k1 + A * k2 of the destination (out of place: the caller's in == out is fine, the scratch is the
//           intermediate), with the scale, the inverse's re/im swap and the optional |E|^2.
// Both steps are short-line kernels (A, B = 32 ... 128: a handful of registers per thread, a few KB of exchange per
// CTA, many resident CTAs) whose CTAs work on WC ADJACENT columns, so every access is a 256-byte run of a grid row.
// They move 2 * P * B each (4 * P * B for the column pass instead of 2), at streaming efficiency.
#pragma once
#include <vector>
#include "kernels.cuh"

namespace lp2d {

template <typename T>
struct ColStepArgs {
    const cx<T>* in;
    cx<T>* out;
    T* inten;                 // |out|^2 (step 2) or nullptr
    const cx<T>* tw;          // twiddles of LinePlan<L, E>
    const cx<T>* tw4;         // step 1: exp(-2 pi i q / N), q < N; step 2: nullptr
    long long in_plane, out_plane;                  // samples between beams
    long long in_u, in_pos, out_u, out_pos;         // sample strides of the outer index u and of the line position
    int U;                    // outer count: lines per column and beam
    int tiles;                // column tiles per grid row (N / WC)
    long long nitems;         // beams * U * tiles
    T scale;
    int swap_in, swap_out;    // inverse transform: re/im swapped at the first load / at the last store
    int stream_out;
};

template <typename T, int L, int E, int WC>
struct CSCfg {
    using LP = LinePlan<L, E>;
    static constexpr int TPL = LP::TPL;
    static constexpr int THREADS = WC * TPL;
    static constexpr int S128 = 128 / (int)sizeof(cx<T>);
    static constexpr int RES = (WC >= S128) ? 1 : S128 / WC;       // consecutive lines land in different banks
    static constexpr int LS = LP::LINE_SLOTS + ((RES - LP::LINE_SLOTS % S128) + S128) % S128;
    static constexpr size_t SMEM = (size_t)WC * LS * sizeof(cx<T>);
    static_assert(THREADS >= 32 && THREADS <= 1024 && LP::npass() >= 2, "shape");
};

template <typename T, int L, int E, int WC>
__global__ void __launch_bounds__(CSCfg<T, L, E, WC>::THREADS) colstep_kernel(const ColStepArgs<T> p) {
    using K = CSCfg<T, L, E, WC>;
    constexpr int TPL = K::TPL;
    extern __shared__ __align__(16) unsigned char cs_smem[];
    cx<T>* smem = reinterpret_cast<cx<T>*>(cs_smem);
    const int tid = threadIdx.x;
    const int l = tid % WC, t = tid / WC;
    cx<T>* sline = smem + l * K::LS;
    for (long long item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int ct = (int)(item % p.tiles);
        const long long r = item / p.tiles;
        const int u = (int)(r % p.U);
        const long long beam = r / p.U;
        const long long col = (long long)ct * WC + l;
        const cx<T>* src = p.in + beam * p.in_plane + (long long)u * p.in_u + col;
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = __ldcs(src + (long long)(t + TPL * m) * p.in_pos);
        if (p.swap_in) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
        }
        line_fft<T, L, E, SYNC_CTA>(v, sline, p.tw, t, l);
        if (p.tw4 != nullptr) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], __ldg(p.tw4 + (long long)u * (t + TPL * m)));
        }
        if (p.swap_out) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
        }
        if (p.scale != (T)1) {
#pragma unroll
            for (int m = 0; m < E; ++m) { v[m].x *= p.scale; v[m].y *= p.scale; }
        }
        const long long obase = beam * p.out_plane + (long long)u * p.out_u + col;
        if (p.out != nullptr) {
#pragma unroll
            for (int m = 0; m < E; ++m) st_cx(p.out + obase + (long long)(t + TPL * m) * p.out_pos, v[m], p.stream_out != 0);
        }
        if (p.inten != nullptr) {
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(p.inten + obase + (long long)(t + TPL * m) * p.out_pos, v[m].x * v[m].x + v[m].y * v[m].y);
        }
        // (line_fft synchronises the CTA before it scatters: the next item may start right away)
    }
}

// host side: one entry per (L) and type
template <typename T>
struct ColStepOps {
    int l, e, wc, tw_size;
    cudaError_t (*launch)(const ColStepArgs<T>& a, unsigned grid, cudaStream_t s);
    cudaError_t (*configure)(int* ctas_per_sm);
    void (*twiddles)(std::vector<double>& re_im);
};
const ColStepOps<float>* find_colstep_f32(int l, int variant);
const ColStepOps<double>* find_colstep_f64(int l, int variant);

}  // namespace lp2d
