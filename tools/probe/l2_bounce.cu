// What the memory system allows for the DATA MOVEMENT of the fused step, with no arithmetic at all: per 64 KB of input
// (8 grid rows of a 1024^2 complex64 plane) read it from HBM, write a state copy (HBM) and an intermediate copy to an
// L2-resident ring, read a (different) 64 KB tile of the ring and write it to the output plane (HBM) -- the ring tile
// and the output either as contiguous rows or as the column tile of the real kernel (1024 pieces of PIECE bytes, 8 KB
// apart).  Reports microseconds per 8 MB plane, to be compared with the step kernel's 6.5 us per beam.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_bounce l2_bounce.cu && ./l2_bounce
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 1024;                    // plane = N x N float2
constexpr size_t PLANE = (size_t)N * N;    // samples

// STATE: state store; RP / OP: pattern of the ring read / of the output store: 0 = contiguous rows, else the column tile of
// the real kernel with pieces of RP / OP bytes (8 KB apart); RP < 0: no ring at all
template <int STATE, int RP, int OP>
__global__ void __launch_bounds__(256) bounce_kernel(const float2* __restrict__ in, float2* __restrict__ state, float2* __restrict__ out,
                                                     float2* __restrict__ ring, int ring_planes, int nplanes, int lag) {
    constexpr int RT = N / 8;
    const int tid = threadIdx.x;
    const long long total = (long long)nplanes * RT;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int plane = (int)(u / RT), tile = (int)(u % RT);
        const size_t o = (size_t)tile * 8 * N + (tid >> 5) * N + (tid & 31);
        const float2* src = in + plane * PLANE + o;
        float2 v[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) v[m] = __ldcs(src + 32 * m);
        if (STATE) {
#pragma unroll
            for (int m = 0; m < 32; ++m) __stcs(state + plane * PLANE + o + 32 * m, v[m]);
        }
        int p2 = plane;
        if (RP >= 0) {
            float2* r = ring + (size_t)(plane % ring_planes) * PLANE + o;
#pragma unroll
            for (int m = 0; m < 32; ++m) __stcg(r + 32 * m, v[m]);
            p2 = plane >= lag ? plane - lag : plane;
            const float2* rr = ring + (size_t)(p2 % ring_planes) * PLANE;
            if (RP == 0) {
#pragma unroll
                for (int m = 0; m < 32; ++m) v[m] = __ldcg(rr + o + 32 * m);
            } else {
                constexpr int CW = (RP > 0 ? RP : 8) / 8, TPT = 256 / CW, UPT = CW >= 8 ? CW / 8 : 1;   // units per tile
                const int l = tid % CW, t = tid / CW;
                const size_t col = (size_t)(tile / UPT) * CW % N + l + (CW < 8 ? 0 : 0);
                const size_t rbase = (size_t)(tile % UPT) * (N / UPT);
                if (CW >= 8) {
#pragma unroll
                    for (int m = 0; m < 32; ++m) v[m] = __ldcg(rr + (rbase + t + (size_t)TPT * m) * N + col);
                } else {
#pragma unroll
                    for (int m = 0; m < 32; ++m) v[m] = __ldcg(rr + ((size_t)(t + TPT * m) % N) * N + (size_t)tile * 8 % N + ((t + TPT * m) / N) * CW + l);
                }
            }
        }
        float2* oo = out + p2 * PLANE;
        if (OP == 0) {
#pragma unroll
            for (int m = 0; m < 32; ++m) __stcs(oo + o + 32 * m, v[m]);
        } else {
            constexpr int CW = (OP > 0 ? OP : 8) / 8, TPT = 256 / CW, UPT = CW >= 8 ? CW / 8 : 1;
            const int l = tid % CW, t = tid / CW;
            const size_t col = (size_t)(tile / UPT) * CW % N + l;
            const size_t rbase = (size_t)(tile % UPT) * (N / UPT);
            if (CW >= 8) {
#pragma unroll
                for (int m = 0; m < 32; ++m) __stcs(oo + (rbase + t + (size_t)TPT * m) * N + col, v[m]);
            } else {
#pragma unroll
                for (int m = 0; m < 32; ++m) __stcs(oo + ((size_t)(t + TPT * m) % N) * N + (size_t)tile * 8 % N + ((t + TPT * m) / N) * CW + l, v[m]);
            }
        }
    }
}

template <typename F>
static float time_ms(F f, int iters = 10) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main() {
    const int nplanes = 256, ring_planes = 6, lag = 3;
    float2 *in, *st, *out, *ring;
    cudaMalloc(&in, nplanes * PLANE * 8); cudaMalloc(&st, nplanes * PLANE * 8); cudaMalloc(&out, nplanes * PLANE * 8);
    cudaMalloc(&ring, ring_planes * PLANE * 8);
    cudaMemset(in, 0, nplanes * PLANE * 8); cudaMemset(ring, 0, ring_planes * PLANE * 8);
    for (int ctas = 2; ctas <= 8; ctas *= 4) {
        const int grid = 148 * ctas;
        printf("grid = 148 x %d CTAs of 256 threads (64 KB per unit in registers)\n", ctas);
#define RUN(STATE, RP, OP, what)                                                                                      \
    {                                                                                                                  \
        float ms = time_ms([&] { bounce_kernel<STATE, RP, OP><<<grid, 256>>>(in, st, out, ring, ring_planes, nplanes, lag); });   \
        printf("  %-74s %6.2f us / plane\n", what, ms * 1e3 / nplanes);                                               \
    }
        RUN(0, -1, 0, "read + out rows (2 PB HBM)");
        RUN(1, -1, 0, "read + state + out rows (3 PB HBM)");
        RUN(1, -1, 64, "read + state + out 64 B pieces (3 PB HBM, no ring)");
        RUN(1, -1, 128, "read + state + out 128 B pieces (3 PB HBM, no ring)");
        RUN(1, -1, 256, "read + state + out 256 B pieces (3 PB HBM, no ring)");
        RUN(1, 0, 0, "read, state, ring rows, out rows (3 PB HBM + 2 PB L2)");
        RUN(1, 64, 0, "read, state, ring read 64 B pieces, out rows");
        RUN(1, 0, 64, "read, state, ring rows, out 64 B pieces");
        RUN(1, 64, 64, "read, state, ring read 64 B pieces, out 64 B pieces");
        RUN(1, 0, 128, "read, state, ring rows, out 128 B pieces");
        RUN(1, 128, 128, "read, state, ring read 128 B pieces, out 128 B pieces");
        RUN(1, 0, 256, "read, state, ring rows, out 256 B pieces");
        RUN(1, 32, 32, "read, state, ring read 32 B pieces, out 32 B pieces");
    }
    return 0;
}
