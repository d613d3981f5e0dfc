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
    Ops<float, 2048, 16, 2, 4>::entry(),
    Ops<float, 4096, 16, 1, 2>::entry(),
    Ops<float, 8192, 16, 1, 1>::entry(),
};
const SizeOps<float>* find_ops_f32(int n) {
    for (const auto& o : kOpsF32) if (o.n == n) return &o;
    return nullptr;
}
}  // namespace lp2d
