// complex128 instances of the fused step kernel: (N, E, lines per CTA, min CTAs/SM)
#include "pipe2_ops.cuh"

namespace lp2d {
static const P2Ops<double> kP2F64[] = {
    // (1024 points, E = 16: 252 registers, one CTA per SM: 24 us per beam against 20 for the per-pass kernels,
    // profiles/r2b_sweep_1024c128_2048c64.log -- not built)
    T3Inst<double, 512, 16, 4, 3>::entry(),
    T3Inst<double, 512, 16, 8, 2>::entry(),
    T3Inst<double, 256, 16, 8, 3>::entry(),
    T3Inst<double, 256, 8, 8, 4>::entry(),
};
static const Q4Ops<double> kQ4F64[] = {
    Q4Inst<double, 512, 16, 8, 2>::entry(),
    Q4Inst<double, 256, 16, 16, 2>::entry(),
};
const Q4Ops<double>* find_q4_f64(int n, int variant) {
    (void)variant;
    for (const auto& o : kQ4F64)
        if (o.n == n) return &o;
    return nullptr;
}
const P2Ops<double>* find_p2_f64(int n, int variant) {
    const P2Ops<double>* first = nullptr;
    int seen = 0;
    for (const auto& o : kP2F64) {
        if (o.n != n) continue;
        if (seen++ == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
cudaError_t p2_factor_plane_f64(double2* fac, const double* tabs, int n, int fl, int kind, int by_column, int full, double z,
                                double k0, double two_f, cudaStream_t s) {
    factor_plane_kernel<double><<<148 * 4, 256, 0, s>>>(fac, tabs, tabs + n, tabs + 2 * n, n, fl, kind, by_column, full, z, k0, two_f);
    ++g_launches;
    return cudaGetLastError();
}
}  // namespace lp2d
