// complex64 instances of the ticket form of the fused step kernel with scalar FP32 complex
// arithmetic (no packed f32x2: the pair-aligned register allocation spills at 64 registers)
#include "pipe2_ops.cuh"

namespace lp2d {
static const P2Ops<float> kT3F32[] = {
    T3Inst<float, 1024, 16, 4, 4>::entry(),
    T3Inst<float, 1024, 16, 8, 2>::entry(),
    T3Inst<float, 1024, 16, 2, 8>::entry(),
    T5Inst<float, 1024, 16, 8, 2>::entry(),
    T5Inst<float, 1024, 16, 4, 4>::entry(),
};
const P2Ops<float>* find_t3u_f32(int n, int variant) {
    int seen = 0;
    for (const auto& o : kT3F32) {
        if (o.n != n) continue;
        if (seen++ == variant) return &o;
    }
    return nullptr;
}
}  // namespace lp2d
