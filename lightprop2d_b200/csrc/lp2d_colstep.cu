// Instances of the two-step long-column kernels (colstep.cuh): short lines of 32 / 64 / 128 points, column tiles of
// 256 bytes per grid row.
#include "colstep.cuh"

namespace lp2d {
extern long long g_launches;

template <typename T, int L, int E, int WC>
struct CSInst {
    using K = CSCfg<T, L, E, WC>;
    static cudaError_t launch(const ColStepArgs<T>& a, unsigned grid, cudaStream_t s) {
        colstep_kernel<T, L, E, WC><<<grid, K::THREADS, K::SMEM, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t configure(int* ctas) {
        auto kern = colstep_kernel<T, L, E, WC>;
        if (K::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, kern, K::THREADS, K::SMEM);
    }
    static void twiddles(std::vector<double>& w) {
        using LP = typename K::LP;
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
    static constexpr ColStepOps<T> entry() { return ColStepOps<T>{L, E, WC, K::LP::tw_size(), &launch, &configure, &twiddles}; }
};

// (measured, profiles/r3y_colstep_variants.log: 512-byte tiles and E = 16 / 8 swapped are 3-20 % slower than these; the
// two steps together stream 4 P B at 0.91 of the HBM peak at 4096^2 complex128)
static const ColStepOps<float> kCS32[] = {
    CSInst<float, 32, 8, 32>::entry(),
    CSInst<float, 64, 8, 32>::entry(),
    CSInst<float, 128, 16, 32>::entry(),
};
static const ColStepOps<double> kCS64[] = {
    CSInst<double, 32, 8, 16>::entry(),
    CSInst<double, 64, 8, 16>::entry(),
    CSInst<double, 128, 16, 16>::entry(),
};
const ColStepOps<float>* find_colstep_f32(int l, int variant) {
    const ColStepOps<float>* first = nullptr;
    int seen = 0;
    for (const auto& o : kCS32) {
        if (o.l != l) continue;
        if (seen++ == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
const ColStepOps<double>* find_colstep_f64(int l, int variant) {
    const ColStepOps<double>* first = nullptr;
    int seen = 0;
    for (const auto& o : kCS64) {
        if (o.l != l) continue;
        if (seen++ == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
}  // namespace lp2d
