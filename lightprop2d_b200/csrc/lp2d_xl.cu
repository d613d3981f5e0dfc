// Long-line kernel set (cluster / distributed shared memory): (N, E, CTAs per line); LP2D_XL_VARIANT picks
// an alternative split of the same size (tuning)
#include "xl_line.cuh"

namespace lp2d {
static const XlOps<float> kXlF32[] = {
    Xl<float, 16384, 16, 1>::entry(),
    Xl<float, 16384, 16, 2>::entry(1),
    Xl<float, 16384, 32, 1>::entry(2),
    Xl<float, 16384, 32, 2>::entry(3),
    Xl<float, 16384, 32, 4>::entry(4),
    // measured (profiles/r1d_xl_variants*.log): E = 32 (two exchanges instead of three) wins at 32768; more CTAs
    // per line lose (a larger share of the exchange crosses SMs); packed f32x2 arithmetic and two-level
    // twiddles change nothing (r1d_xl_packed_tw2.log): the phases of the one resident line are serialised
    Xl<float, 32768, 32, 2>::entry(),
    Xl<float, 32768, 16, 4>::entry(1),
    Xl<float, 32768, 16, 2>::entry(2),
    Xl<float, 32768, 32, 4>::entry(3),
    Xl<float, 32768, 32, 8>::entry(4),
};
static const XlOps<double> kXlF64[] = {
    Xl<double, 8192, 8, 1>::entry(),
    Xl<double, 8192, 8, 2>::entry(1),
    Xl<double, 8192, 16, 1>::entry(2),
    Xl<double, 8192, 16, 2>::entry(3),
    Xl<double, 16384, 16, 2>::entry(),
    Xl<double, 16384, 8, 4>::entry(1),
    Xl<double, 16384, 8, 2>::entry(2),
    Xl<double, 16384, 16, 4>::entry(3),
    Xl<double, 32768, 8, 4>::entry(),
    Xl<double, 32768, 16, 4>::entry(2),
    Xl<double, 32768, 16, 8>::entry(3),
};
template <typename T, size_t K>
static const XlOps<T>* pick(const XlOps<T> (&tab)[K], int n, int variant) {
    const XlOps<T>* first = nullptr;
    for (const auto& o : tab) {
        if (o.n != n) continue;
        if (o.variant == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
const XlOps<float>* find_xl_f32(int n, int variant) { return pick(kXlF32, n, variant); }
const XlOps<double>* find_xl_f64(int n, int variant) { return pick(kXlF64, n, variant); }
}  // namespace lp2d
