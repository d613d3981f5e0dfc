// complex64 instances of the fused step kernel: (N, E, lines per CTA, min CTAs/SM)
#define LP2D_PACKED_F32 1
#ifndef LP2D_TWO_LEVEL_TWIDDLES
#define LP2D_TWO_LEVEL_TWIDDLES 1
#endif
#include "pipe2_ops.cuh"

namespace lp2d {
static const P2Ops<float> kP2F32[] = {
    // 2048 points: a line is two warps (E = 32), 4 lines per CTA; a plane is 32 MB and the ring is 2 planes: the
    // steps with a state store gain 14 % (43.4 vs 50.3 us per beam), plain fft2 and the fresh chain do not
    // (profiles/r2b_sweep_1024c128_2048c64.log)
    T3Inst<float, 2048, 32, 4, 2>::entry(),
    T3Inst<float, 2048, 32, 4, 2, 3>::entry(),
    // the first entry of each size is the default (measured best, profiles/r1c_size_sweep.log)
    T3Inst<float, 1024, 32, 8, 2, 3>::entry(),
    T3Inst<float, 1024, 32, 8, 2, 0>::entry(),
    T3Inst<float, 1024, 32, 4, 3>::entry(),
#if defined(LP2D_ABLATION) && LP2D_ABLATION
    // data-movement skeletons (HOOKS bit 2: no transform, bit 3: no factor): wrong results, timing experiments only
    // (profiles/r3b_ablate.log: 5.1-5.3 us per beam without any arithmetic against 6.5 with it)
    T3Inst<float, 1024, 32, 8, 2, 4>::entry(),
    T3Inst<float, 1024, 32, 8, 2, 8>::entry(),
    T3Inst<float, 1024, 32, 8, 2, 12>::entry(),
#endif
    T3Inst<float, 512, 32, 8, 3>::entry(),
    T3Inst<float, 512, 32, 16, 2>::entry(),
    T3Inst<float, 256, 16, 16, 4>::entry(),
    T3Inst<float, 256, 16, 8, 8>::entry(),
};
static const Q4Ops<float> kQ4F32[] = {
    Q4Inst<float, 1024, 32, 8, 2, 2>::entry(),
    Q4Inst<float, 1024, 32, 8, 2, 3>::entry(),
    Q4Inst<float, 1024, 32, 8, 2, 0>::entry(),
    // half tiles, two CTAs per SM: their load/store and floating-point phases interleave
    Q4Inst<float, 1024, 32, 4, 2, 2, 2>::entry(),
    Q4Inst<float, 1024, 32, 4, 2, 3, 2>::entry(),
    Q4Inst<float, 1024, 32, 4, 1, 2, 3>::entry(),
#if defined(LP2D_ABLATION) && LP2D_ABLATION
    Q4Inst<float, 1024, 32, 8, 2, 4>::entry(),
    Q4Inst<float, 1024, 32, 8, 2, 12>::entry(),
    Q4Inst<float, 1024, 32, 4, 2, 12, 2>::entry(),
    Q4Inst<float, 1024, 32, 8, 2, 2 + 16>::entry(),        // transform, no result stores
    Q4Inst<float, 1024, 32, 8, 2, 2 + 32>::entry(),        // transform, no state stores
    Q4Inst<float, 1024, 32, 8, 2, 2 + 48>::entry(),        // transform, no stores at all
#endif
    Q4Inst<float, 512, 32, 16, 2>::entry(),
    Q4Inst<float, 512, 32, 16, 1>::entry(),
    Q4Inst<float, 256, 16, 16, 2>::entry(),
};
const Q4Ops<float>* find_q4_f32(int n, int variant) {
    const Q4Ops<float>* first = nullptr;
    int seen = 0;
    for (const auto& o : kQ4F32) {
        if (o.n != n) continue;
        if (seen++ == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
const P2Ops<float>* find_p2_f32(int n, int variant) {
    const P2Ops<float>* first = nullptr;
    int seen = 0;
    for (const auto& o : kP2F32) {
        if (o.n != n) continue;
        if (seen++ == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
cudaError_t p2_factor_plane_f32(float2* fac, const double* tabs, int n, int fl, int kind, int by_column, int full, double z,
                                double k0, double two_f, cudaStream_t s) {
    factor_plane_kernel<float><<<148 * 4, 256, 0, s>>>(fac, tabs, tabs + n, tabs + 2 * n, n, fl, kind, by_column, full, z, k0, two_f);
    ++g_launches;
    return cudaGetLastError();
}
}  // namespace lp2d
