// complex64 instances of the persistent pipeline kernel: (N, E, lines per CTA, min CTAs/SM)
#define LP2D_PACKED_F32 1
#include "pipe_ops.cuh"

namespace lp2d {
static const PipeOps<float> kPipeF32[] = {
    PipeInst<float, 256, 16, 16, 3>::entry(),
    PipeInst<float, 512, 16, 8, 3>::entry(),
    PipeInst<float, 1024, 16, 4, 3>::entry(),
    PipeInst<float, 2048, 16, 4, 1>::entry(),
    // variant 1: register cap of 2 CTAs per SM
    PipeInst<float, 256, 16, 16, 2>::entry(),
    PipeInst<float, 512, 16, 8, 2>::entry(),
    PipeInst<float, 1024, 16, 4, 2>::entry(),
    // variant 2 (1024 only, needs the E = 32 twiddle table: LP2D_VARIANT=3): one warp per row
    PipeInst<float, 1024, 32, 4, 3>::entry(),
    PipeInst<float, 1024, 32, 4, 4>::entry(),
};
const PipeOps<float>* find_pipe_f32(int n, int variant) {
    const PipeOps<float>* first = nullptr;
    int seen = 0;
    for (const auto& o : kPipeF32) {
        if (o.n != n) continue;
        if (seen++ == variant) return &o;
        if (first == nullptr) first = &o;
    }
    return first;
}
}  // namespace lp2d
