// complex64 kernel set: (N, E, rows per CTA, columns per CTA)
#include "sizes.cuh"

namespace lp2d {
static const SizeOps<float> kOpsF32[] = {
    Ops<float, 16, 16, 128, 16>::entry(),
    Ops<float, 32, 16, 64, 16>::entry(),
    Ops<float, 64, 16, 32, 16>::entry(),
    Ops<float, 128, 16, 16, 16>::entry(),
    Ops<float, 256, 16, 16, 16>::entry(),
    Ops<float, 512, 16, 8, 8>::entry(),
    Ops<float, 1024, 16, 4, 8>::entry(),
    Ops<float, 1024, 16, 4, 4>::entry(1),
    Ops<float, 1024, 16, 4, 16>::entry(2),
    Ops<float, 1024, 32, 8, 8>::entry(3),
    Ops<float, 1024, 32, 8, 4>::entry(4),
    Ops<float, 1024, 32, 8, 16>::entry(5),
    Ops<float, 2048, 16, 2, 4>::entry(),
    Ops<float, 2048, 16, 2, 8>::entry(1),
    // large grids: the widest column tile that fits 1024 threads (measured: profiles/r1c_large_n_variants.log)
    Ops<float, 4096, 16, 1, 4>::entry(),
    Ops<float, 4096, 16, 1, 2>::entry(1),
    Ops<float, 8192, 16, 1, 2>::entry(),
    Ops<float, 8192, 16, 1, 1>::entry(1),
};
const SizeOps<float>* find_ops_f32(int n, int variant) {
    const SizeOps<float>* first = nullptr;
    for (const auto& o : kOpsF32) {
        if (o.n != n) continue;
        if (o.variant == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
}  // namespace lp2d
