// complex128 instances of the persistent pipeline kernel: (N, E, lines per CTA, min CTAs/SM)
#include "pipe_ops.cuh"

namespace lp2d {
static const PipeOps<double> kPipeF64[] = {
    PipeInst<double, 256, 8, 8, 3>::entry(),
    PipeInst<double, 512, 8, 4, 3>::entry(),
    PipeInst<double, 1024, 8, 2, 3>::entry(),
};
const PipeOps<double>* find_pipe_f64(int n, int variant) {
    (void)variant;
    for (const auto& o : kPipeF64)
        if (o.n == n) return &o;
    return nullptr;
}
}  // namespace lp2d
