// Long lines: ONE 1-D transform spread over a thread-block cluster.
//
// A line of 16384 / 32768 complex64 samples (128 / 256 KB) no longer fits the registers and the
// shared memory of one CTA next to its Stockham exchange buffer.  Here a line belongs to a
// cluster of CL CTAs: TPL = N / E threads in total, THREADS = TPL / CL per CTA, thread t (global
// within the cluster) owns the samples t + TPL * m exactly as in fft_core.cuh, so loads and stores
// stay HBM <-> registers, coalesced.  The exchange buffer of N samples is spread over the CTAs so
// that a thread GATHERS only from its own CTA (plain ld.shared); the scatter stores go to whichever
// CTA will gather the slot, through distributed shared memory (mapa + st.shared::cluster), and the
// two barriers of a pass are cluster barriers.  Same pass schedule, same twiddle tables, same arithmetic as the
// single-CTA line kernel (line_pass), so results agree with it to rounding.
//
// Used by lp2d_pass_lines (the row pass of the slab-decomposed grid, BASELINE.json config 4:
// 32768^2 complex64 over 8 GPUs) for the sizes the single-CTA kernels do not cover.
#pragma once
#include <vector>
#include <cmath>
#include "dispatch.h"

namespace lp2d {

template <typename T> struct ClMem;
template <> struct ClMem<float> {
    __device__ __forceinline__ static float2 ld(unsigned a) {
        float2 v;
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
        return v;
    }
    __device__ __forceinline__ static void st(unsigned a, float2 v) {
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
    }
};
template <> struct ClMem<double> {
    __device__ __forceinline__ static double2 ld(unsigned a) {
        double2 v;
        asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
        return v;
    }
    __device__ __forceinline__ static void st(unsigned a, double2 v) {
        asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
    }
};

template <int CL>
__device__ __forceinline__ void xl_sync() {
    if (CL == 1) {
        __syncthreads();
    } else {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}
// Rendezvous without memory ordering: before the first scatter of a line there is nothing to publish (the barrier
// only makes sure every CTA of the cluster is running), and a release here would wait for the state stores to HBM.
template <int CL>
__device__ __forceinline__ void xl_rendezvous() {
    if (CL == 1) {
        __syncthreads();
    } else {
        asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    }
}

template <typename T, int N, int E, int CL>
struct XlCfg {
    using LP = LinePlan<N, E>;
    static constexpr int TPL = LP::TPL;
    static constexpr int THREADS = TPL / CL;
    static constexpr int CH = N / CL;                            // exchange slots held by one CTA (= E * THREADS)
    static constexpr size_t SMEM = (size_t)CH * sizeof(cx<T>);
    // resident CTAs per SM: registers (2 E sizeof(T)/4 for the data, as many again for the butterflies), shared memory, threads
    static constexpr int REGS = (E * (int)sizeof(T) >= 128) ? 128 : 64;
    static constexpr int MINB_R = 65536 / (THREADS * REGS), MINB_S = (int)((227 * 1024) / (SMEM + 1024)), MINB_T = 2048 / THREADS;
    static constexpr int MINB0 = MINB_R < MINB_S ? (MINB_R < MINB_T ? MINB_R : MINB_T) : (MINB_S < MINB_T ? MINB_S : MINB_T);
    static constexpr int MINB = MINB0 < 1 ? 1 : MINB0;
    static_assert(TPL % CL == 0 && (CL & (CL - 1)) == 0, "cluster split");
    static_assert(THREADS <= 1024 && THREADS % 32 == 0, "block size");
    static_assert(SMEM <= 227 * 1024, "exchange piece does not fit an SM");
    static_assert(CL >= 1 && CL <= 8, "portable cluster size");
};

// Where sample position g of the line lives between two passes.  The owner is the CTA whose thread
// GATHERS it (thread t reads t + TPL * m, so owner = (g mod TPL) / THREADS): every gather is a local,
// conflict-free ld.shared of consecutive slots, and only the scatter stores -- fire and forget -- cross
// SMs.  Within the owner the slot is (g / TPL) * THREADS + g mod THREADS, XOR-swizzled with the next five
// bits so that the stride-R scatter of the first pass spreads over the banks (a warp's gather reads 32
// consecutive slots with those bits constant: the swizzle permutes them among themselves).
__device__ __forceinline__ int xl_swz(int off) { return off ^ ((off >> 5) & 31); }

template <typename T, int N, int E, int CL>
__device__ __forceinline__ unsigned xl_scatter_addr(const unsigned (&sb)[CL], int g) {
    constexpr int TPL = N / E, TH = TPL / CL;
    const int gm = g & (TPL - 1);
    const int owner = gm / TH;
    const int off = (g / TPL) * TH + (gm & (TH - 1));
    unsigned b = sb[0];
#pragma unroll
    for (int r = 1; r < CL; ++r)
        if (owner == r) b = sb[r];
    return b + (unsigned)xl_swz(off) * (unsigned)sizeof(cx<T>);
}

// One Stockham pass (see line_pass in fft_core.cuh) with the exchange in distributed shared memory.
template <typename T, int N, int E, int CL, int P>
__device__ __forceinline__ void xl_pass(cx<T> (&v)[E], const unsigned (&sb)[CL], const cx<T>* local, const cx<T>* __restrict__ tw, int t) {
    using LP = LinePlan<N, E>;
    using K = XlCfg<T, N, E, CL>;
    constexpr int R = LP::radix(P);
    constexpr int Ns = LP::ns(P);
    constexpr int Q = E / R;
    constexpr int TPL = LP::TPL;
    constexpr bool last = (Ns * R == N);
    constexpr int twoff = LP::tw_offset(P);

    if (P > 0) {
        xl_sync<CL>();                       // the scatter of the previous pass is visible cluster-wide
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = local[xl_swz(m * K::THREADS + (int)threadIdx.x)];
    }
    if (!last) {
        if (P == 0) xl_rendezvous<CL>();     // every CTA of the cluster runs
        else xl_sync<CL>();                  // everyone has read its inputs before anyone overwrites the exchange buffer
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int j = t + TPL * q;
        cx<T> u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = v[q + r * Q];
        if (Ns > 1) {
            const int k = j & (Ns - 1);
            if constexpr (R > 16) {
                // radix 32: load and apply the twiddles eight at a time (register pressure, see line_pass)
#pragma unroll
                for (int r0 = 1; r0 < R; r0 += 8) {
                    cx<T> w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (r0 + i < R) w[i] = __ldg(&tw[twoff + (r0 + i - 1) * Ns + k]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (r0 + i < R) u[r0 + i] = cmul(u[r0 + i], w[i]);
                    asm volatile("" ::: "memory");
                }
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) u[r] = cmul(u[r], __ldg(&tw[twoff + (r - 1) * Ns + k]));
            }
        }
        dft_reg<T, R>(u);
        if (last) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[q + r * Q] = u[r];
        } else {
            const int base = (j / Ns) * (Ns * R) + (j & (Ns - 1));
#pragma unroll
            for (int r = 0; r < R; ++r) ClMem<T>::st(xl_scatter_addr<T, N, E, CL>(sb, base + r * Ns), u[r]);
        }
    }
}

template <typename T, int N, int E, int CL, int P = 0>
__device__ __forceinline__ void xl_fft(cx<T> (&v)[E], const unsigned (&sb)[CL], const cx<T>* local, const cx<T>* tw, int t) {
    xl_pass<T, N, E, CL, P>(v, sb, local, tw, t);
    if constexpr (P + 1 < LinePlan<N, E>::npass()) xl_fft<T, N, E, CL, P + 1>(v, sb, local, tw, t);
}

// Row pass over lines p.line0 .. p.line0 + p.nlines - 1 (one line per cluster):
//   load -> [transfer function H, computed per sample in FP64] -> [store state] -> (i)FFT -> scale
//        -> [store] [store |.|^2]
template <typename T, int N, int E, int CL, int PRE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(XlCfg<T, N, E, CL>::THREADS, XlCfg<T, N, E, CL>::MINB)
xl_line_kernel(const LineArgs<T> p) {
    using K = XlCfg<T, N, E, CL>;
    constexpr int TPL = K::TPL;
    extern __shared__ __align__(16) unsigned char xl_smem[];

    unsigned rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    unsigned sb[CL];
    {
        const unsigned local = (unsigned)__cvta_generic_to_shared(xl_smem);
#pragma unroll
        for (int r = 0; r < CL; ++r) {
            if (CL == 1) sb[r] = local;
            else asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(sb[r]) : "r"(local), "r"(r));
        }
    }
    const int t = (int)rank * K::THREADS + (int)threadIdx.x;
    const long long line = (long long)(blockIdx.x / CL) + p.line0;
    const long long beam = line / N;
    const int fixed = (int)(line % N);
    const long long base = beam * p.beam_stride + (long long)fixed * N;
    const bool inv = (p.flags & F_INVERSE) != 0;

    cx<T> v[E];
    if (p.seg_shift) {
        // input in all-to-all tiles (see LineArgs::seg_shift)
        const int lrel = (int)(blockIdx.x / CL);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int pos = t + TPL * m;
            v[m] = __ldcs(&p.in[(((long long)(pos >> p.seg_shift) * p.seg_rows + lrel) << p.seg_shift) + (pos & ((1 << p.seg_shift) - 1))]);
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = __ldcs(&p.in[base + t + TPL * m]);
    }

    if constexpr (PRE == PRE_H) {
        const double b_j = __ldg(&p.b_tab[fixed]);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cmul(v[m], transfer_factor<T>(__ldg(&p.a_tab[t + TPL * m]), b_j, p.z));
    }
    if (p.out_pre != nullptr) {
#pragma unroll
        for (int m = 0; m < E; ++m) __stcs(&p.out_pre[base + t + TPL * m], v[m]);
    }
    if (inv) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
    }
    xl_fft<T, N, E, CL>(v, sb, reinterpret_cast<const cx<T>*>(xl_smem), p.tw, t);
    if (inv) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
    }
    const T sc = p.scale;
    if (sc != (T)1) {
#pragma unroll
        for (int m = 0; m < E; ++m) { v[m].x *= sc; v[m].y *= sc; }
    }
    if (p.out != nullptr) {
#pragma unroll
        for (int m = 0; m < E; ++m) __stcs(&p.out[base + t + TPL * m], v[m]);
    }
    if (p.inten != nullptr) {
#pragma unroll
        for (int m = 0; m < E; ++m) __stcs(&p.inten[base + t + TPL * m], v[m].x * v[m].x + v[m].y * v[m].y);
    }
    // no barrier at the end: gathers are local, and the last scatter into this CTA's shared memory was
    // published by the barrier before the final gather (a release here would hold the SM until the stores drain)
}

template <typename T>
struct XlOps {
    int n, e, cl, tw_size, variant;
    cudaError_t (*launch)(int pre, const LineArgs<T>& a, cudaStream_t s);   // a.nlines lines from a.line0
    cudaError_t (*configure)();
    void (*twiddles)(std::vector<double>& re_im);
};

template <typename T, int N, int E, int CL>
struct Xl {
    using LP = LinePlan<N, E>;
    using K = XlCfg<T, N, E, CL>;
    template <int PRE>
    static cudaError_t go(const LineArgs<T>& a, cudaStream_t s) {
        xl_line_kernel<T, N, E, CL, PRE><<<(unsigned)a.nlines * CL, K::THREADS, K::SMEM, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t launch(int pre, const LineArgs<T>& a, cudaStream_t s) {
        return pre == PRE_H ? go<PRE_H>(a, s) : pre == PRE_NONE ? go<PRE_NONE>(a, s) : cudaErrorInvalidValue;
    }
    static cudaError_t configure() {
        cudaError_t e = cudaFuncSetAttribute(xl_line_kernel<T, N, E, CL, PRE_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(xl_line_kernel<T, N, E, CL, PRE_H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
    }
    // same table layout as Ops::twiddles (sizes.cuh)
    static void twiddles(std::vector<double>& w) {
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
    static constexpr XlOps<T> entry(int variant = 0) { return XlOps<T>{N, E, CL, LP::tw_size(), variant, &launch, &configure, &twiddles}; }
};

const XlOps<float>* find_xl_f32(int n, int variant);
const XlOps<double>* find_xl_f64(int n, int variant);

}  // namespace lp2d
