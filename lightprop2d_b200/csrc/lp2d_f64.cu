// complex128 kernel set: (N, E, rows per CTA, columns per CTA)
#include "sizes.cuh"

namespace lp2d {
static const SizeOps<double> kOpsF64[] = {
    Ops<double, 16, 8, 64, 8>::entry(),
    Ops<double, 32, 8, 32, 8>::entry(),
    Ops<double, 64, 8, 16, 8>::entry(),
    Ops<double, 128, 8, 8, 8>::entry(),
    Ops<double, 256, 8, 8, 8>::entry(),
    Ops<double, 512, 8, 4, 4>::entry(),
    // 1024: E = 16 (two exchanges instead of three): fresh chain +16 %, stateful +3 % (profiles/r1c_large_n_variants.log)
    // (one row per CTA with whole-batch passes: 15.62 vs 15.91 us per beam stateful, 12.81 vs 13.09 fft2; 128-byte column
    // tiles (W = 8, one 512-thread CTA per SM): 17.5)
    Ops<double, 1024, 16, 1, 4>::entry(),
    Ops<double, 1024, 8, 2, 4>::entry(1),
    Ops<double, 1024, 8, 2, 8>::entry(2),
    Ops<double, 1024, 16, 2, 4>::entry(3),
    // E = 16 (two exchanges) with whole-batch passes: fft2 58.9 vs 65.7, stateful 73.2 vs 77.6, chain 113.7 vs 127.6 us per
    // beam against the round-1 default E = 8 (profiles/r3w_variants_whole_batch.log)
    Ops<double, 2048, 16, 1, 2>::entry(),
    Ops<double, 2048, 8, 1, 4>::entry(1),
    Ops<double, 2048, 16, 2, 4>::entry(2),
    Ops<double, 2048, 8, 1, 2>::entry(3),
    Ops<double, 4096, 8, 1, 2>::entry(),
    Ops<double, 4096, 8, 1, 1>::entry(1),
    // 8192 points (tests/gaussbeam_performance.py:15 runs it): a row is 128 KB, one CTA of 512 threads holds it (E = 16,
    // three exchanges through 135 KB of shared memory); the column pass runs in two steps (colstep.cuh, 64 x 128), the
    // one-column-per-CTA kernel below only serves the fused column stage of the fresh chain
    Ops<double, 8192, 16, 1, 1>::entry(),
};
const SizeOps<double>* find_ops_f64(int n, int variant) {
    const SizeOps<double>* first = nullptr;
    for (const auto& o : kOpsF64) {
        if (o.n != n) continue;
        if (o.variant == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
}  // namespace lp2d
