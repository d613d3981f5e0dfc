// DRAM throughput of streaming kernels with different read : write mixes (bytes moved / time), to put the stateful
// step's 1 : 2 mix (read spectrum, write spectrum, write field) next to the 1 : 1 copy that MEASURED_PEAKS.json reports.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rw_mix rw_mix.cu && ./rw_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NW>
__global__ void mix_kernel(const float4* __restrict__ in, float4* __restrict__ o0, float4* __restrict__ o1, float4* __restrict__ o2, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 v = __ldcs(in + i);
        if (NW >= 1) __stcs(o0 + i, v);
        if (NW >= 2) { v.x += 1.f; __stcs(o1 + i, v); }
        if (NW >= 3) { v.y += 1.f; __stcs(o2 + i, v); }
    }
}
__global__ void read_kernel(const float4* __restrict__ in, float* out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const float4 v = __ldcs(in + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}
__global__ void write_kernel(float4* __restrict__ o0, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(o0 + i, make_float4(1.f, 2.f, 3.f, 4.f));
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
    const size_t bytes = 2ull << 30, n = bytes / sizeof(float4);
    float4 *in, *o0, *o1, *o2; float* r;
    cudaMalloc(&in, bytes); cudaMalloc(&o0, bytes); cudaMalloc(&o1, bytes); cudaMalloc(&o2, bytes); cudaMalloc(&r, 4);
    cudaMemset(in, 0, bytes);
    const int grid = 148 * 16, block = 256;
    float ms;
    ms = time_ms([&] { read_kernel<<<grid, block>>>(in, r, n); });
    printf("read only      : %7.1f GB/s\n", bytes / ms / 1e6);
    ms = time_ms([&] { write_kernel<<<grid, block>>>(o0, n); });
    printf("write only     : %7.1f GB/s\n", bytes / ms / 1e6);
    ms = time_ms([&] { mix_kernel<1><<<grid, block>>>(in, o0, o1, o2, n); });
    printf("1 read : 1 write: %7.1f GB/s\n", 2.0 * bytes / ms / 1e6);
    ms = time_ms([&] { mix_kernel<2><<<grid, block>>>(in, o0, o1, o2, n); });
    printf("1 read : 2 write: %7.1f GB/s\n", 3.0 * bytes / ms / 1e6);
    ms = time_ms([&] { mix_kernel<3><<<grid, block>>>(in, o0, o1, o2, n); });
    printf("1 read : 3 write: %7.1f GB/s\n", 4.0 * bytes / ms / 1e6);
    ms = time_ms([&] { cudaMemcpyAsync(o0, in, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D  : %7.1f GB/s\n", 2.0 * bytes / ms / 1e6);
    return 0;
}
