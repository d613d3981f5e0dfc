// Host-side entry of one (T, N, E, W) instance of the fused step kernel (pipe2.cuh, t3_kernel).
#pragma once
#include <cmath>
#include <cuda_runtime.h>
#include <vector>
#include "pipe3.cuh"

namespace lp2d {

template <typename T>
struct P2Ops {
    int n, e, w;
    int tiles;          // items per (stage, group)
    int threads;
    int fl;             // factor-plane row stride (samples)
    int tw_size;        // twiddle table entries of LinePlan<n, e>
    size_t smem_x;      // exchange area
    size_t smem_f;      // unused (the factor rows are read from the L2-resident plane)
    // win_ptr / win_bytes: optional L2 persisting access window (the ring); win_bytes = 0: none
    cudaError_t (*launch)(const P2Args<T>& a, unsigned grid, bool need_factor, cudaStream_t s, void* win_ptr,
                          size_t win_bytes);
    // opt in to large dynamic shared memory; resident CTAs per SM with / without the factor rows
    cudaError_t (*configure)(int* ctas_per_sm_factor, int* ctas_per_sm_plain);
    void (*twiddles)(std::vector<double>& re_im);
};

// persistent form (pipe3.cuh, q4_kernel): same arguments, one CTA per SM
template <typename T>
struct Q4Ops {
    int n, e, w;
    int tiles, threads, fl, tw_size;
    size_t smem;
    int nbuf;           // landing buffers per CTA (items a CTA holds besides the one it works on)
    cudaError_t (*launch)(const P2Args<T>& a, unsigned grid, cudaStream_t s);
    cudaError_t (*configure)(int* ctas_per_sm);
    void (*twiddles)(std::vector<double>& re_im);
};
const Q4Ops<float>* find_q4_f32(int n, int variant);
const Q4Ops<double>* find_q4_f64(int n, int variant);

const P2Ops<float>* find_p2_f32(int n, int variant);
const P2Ops<double>* find_p2_f64(int n, int variant);
cudaError_t p2_factor_plane_f32(float2* fac, const double* tabs, int n, int fl, int kind, int by_column, int full, double z,
                                double k0, double two_f, cudaStream_t s);
cudaError_t p2_factor_plane_f64(double2* fac, const double* tabs, int n, int fl, int kind, int by_column, int full, double z,
                                double k0, double two_f, cudaStream_t s);

extern long long g_launches;

// exp(-2 pi i k r / (Ns R)) for every pass after the first of the schedule LP, FP64 re/im interleaved
template <typename LP>
void p2_twiddles(std::vector<double>& w) {
    w.assign(2 * (size_t)LP::tw_size(), 0.0);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int p = 1; p < LP::npass(); ++p) {
        const int R = LP::radix(p), Ns = LP::ns(p), off = LP::tw_offset(p);
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                const long double ang = -two_pi * (long double)((long long)k * r) / (long double)((long long)Ns * R);
                w[2 * (size_t)(off + (r - 1) * Ns + k)] = (double)cosl(ang);
                w[2 * (size_t)(off + (r - 1) * Ns + k) + 1] = (double)sinl(ang);
            }
    }
}

template <typename T, int N, int E, int W, int MINB, int HOOKS = 0>
struct T3Inst {
    using K = T3Cfg<T, N, E, W>;
    using LP = LinePlan<N, E>;
    static cudaError_t launch(const P2Args<T>& a, unsigned grid, bool, cudaStream_t s, void*, size_t) {
        t3_kernel<T, N, E, W, MINB, HOOKS><<<grid, K::THREADS, K::SMEM_X, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t configure(int* with_factor, int* plain) {
        auto kern = t3_kernel<T, N, E, W, MINB, HOOKS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_X);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(with_factor, kern, K::THREADS, K::SMEM_X);
        *plain = *with_factor;
        return e;
    }
    static void twiddles(std::vector<double>& w) { p2_twiddles<LP>(w); }
    static constexpr P2Ops<T> entry() {
        return P2Ops<T>{N, E, W, K::TILES, K::THREADS, K::FL, LP::tw_size(), K::SMEM_X, 0, &launch, &configure, &twiddles};
    }
};

template <typename T, int N, int E, int W, int NBUF, int HOOKS = 2, int MINB = 1>
struct Q4Inst {
    using K = Q4Cfg<T, N, E, W, NBUF, (MINB >= 3 ? 1 : 2)>;
    using LP = LinePlan<N, E>;
    static cudaError_t launch(const P2Args<T>& a, unsigned grid, cudaStream_t s) {
        q4_kernel<T, N, E, W, NBUF, HOOKS, MINB><<<grid, K::THREADS, K::SMEM, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t configure(int* ctas) {
        auto kern = q4_kernel<T, N, E, W, NBUF, HOOKS, MINB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, kern, K::THREADS, K::SMEM);
    }
    static void twiddles(std::vector<double>& w) { p2_twiddles<LP>(w); }
    static constexpr Q4Ops<T> entry() {
        return Q4Ops<T>{N, E, W, K::TILES, K::THREADS, K::FL, LP::tw_size(), K::SMEM, NBUF, &launch, &configure, &twiddles};
    }
};

}  // namespace lp2d
