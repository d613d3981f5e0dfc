// Long-line kernel set (cluster / distributed shared memory): (N, E, CTAs per line)
#include "xl_line.cuh"

namespace lp2d {
static const XlOps<float> kXlF32[] = {
    Xl<float, 16384, 16, 1>::entry(),
    Xl<float, 32768, 16, 2>::entry(),
};
static const XlOps<double> kXlF64[] = {
    Xl<double, 8192, 8, 1>::entry(),
    Xl<double, 16384, 8, 2>::entry(),
    Xl<double, 32768, 8, 4>::entry(),
};
const XlOps<float>* find_xl_f32(int n) {
    for (const auto& o : kXlF32) if (o.n == n) return &o;
    return nullptr;
}
const XlOps<double>* find_xl_f64(int n) {
    for (const auto& o : kXlF64) if (o.n == n) return &o;
    return nullptr;
}
}  // namespace lp2d
