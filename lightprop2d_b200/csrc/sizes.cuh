// Builds the SizeOps entry of one (T, N, E, W_ROW, W_COL) configuration.
#pragma once
#include <cmath>
#include "dispatch.h"

namespace lp2d {

template <typename T, int N, int E, int WR, int WC>
struct Ops {
    using LP = LinePlan<N, E>;
    using KC = KCfg<T, N, E, AX_COL, WC>;

    template <int PRE>
    static cudaError_t launch_row(const LineArgs<T>& a, cudaStream_t s) {
        using KR = KCfg<T, N, E, AX_ROW, WR, PRE>;
        const unsigned grid = (unsigned)((a.nlines + WR - 1) / WR);
        line_kernel<T, N, E, AX_ROW, WR, PRE, MID_NONE><<<grid, KR::THREADS, KR::SMEM, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    template <int MID>
    static cudaError_t launch_col(const LineArgs<T>& a, cudaStream_t s) {
        line_kernel<T, N, E, AX_COL, WC, PRE_NONE, MID><<<(unsigned)a.nlines, KC::THREADS, KC::SMEM, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t row(int pre, const LineArgs<T>& a, cudaStream_t s) {
        switch (pre) {
            case PRE_NONE: return launch_row<PRE_NONE>(a, s);
            case PRE_LENS: return launch_row<PRE_LENS>(a, s);
            case PRE_H:    return launch_row<PRE_H>(a, s);
            case PRE_HPLANE: return launch_row<PRE_HPLANE>(a, s);
        }
        return cudaErrorInvalidValue;
    }
    static cudaError_t col(int mid, const LineArgs<T>& a, cudaStream_t s) {
        switch (mid) {
            case MID_NONE: return launch_col<MID_NONE>(a, s);
            case MID_H:    return launch_col<MID_H>(a, s);
            case MID_HPLANE: return launch_col<MID_HPLANE>(a, s);
        }
        return cudaErrorInvalidValue;
    }
    template <typename F>
    static cudaError_t optin(F f, size_t smem) {
        if (smem <= 48 * 1024) return cudaSuccess;
        return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    static cudaError_t configure() {
        cudaError_t e;
        if ((e = optin(line_kernel<T, N, E, AX_ROW, WR, PRE_NONE, MID_NONE>, KCfg<T, N, E, AX_ROW, WR, PRE_NONE>::SMEM)) != cudaSuccess) return e;
        if ((e = optin(line_kernel<T, N, E, AX_ROW, WR, PRE_LENS, MID_NONE>, KCfg<T, N, E, AX_ROW, WR, PRE_LENS>::SMEM)) != cudaSuccess) return e;
        if ((e = optin(line_kernel<T, N, E, AX_ROW, WR, PRE_H, MID_NONE>, KCfg<T, N, E, AX_ROW, WR, PRE_H>::SMEM)) != cudaSuccess) return e;
        if ((e = optin(line_kernel<T, N, E, AX_ROW, WR, PRE_HPLANE, MID_NONE>, KCfg<T, N, E, AX_ROW, WR, PRE_HPLANE>::SMEM)) != cudaSuccess) return e;
        if ((e = optin(line_kernel<T, N, E, AX_COL, WC, PRE_NONE, MID_NONE>, KC::SMEM)) != cudaSuccess) return e;
        if ((e = optin(line_kernel<T, N, E, AX_COL, WC, PRE_NONE, MID_H>, KC::SMEM)) != cudaSuccess) return e;
        if ((e = optin(line_kernel<T, N, E, AX_COL, WC, PRE_NONE, MID_HPLANE>, KC::SMEM)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    // exp(-2 pi i k r / (Ns R)) for every pass after the first, FP64 (rounded to T by the plan)
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
    static constexpr SizeOps<T> entry(int variant = 0) {
        return SizeOps<T>{N, E, LP::tw_size(), WR, WC, variant, &row, &col, &configure, &twiddles};
    }
};

}  // namespace lp2d
