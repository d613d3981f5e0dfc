// which L2 cache-hint forms execute on this GPU
#include <cstdio>
#include <cuda_runtime.h>
template <int K> __device__ unsigned long long pol() {
    unsigned long long p;
    if (K == 0) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    if (K == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    if (K == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    if (K == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <int K> __global__ void k_cp(const float4* in, float4* out) {
    __shared__ float4 s[128];
    unsigned long long p = pol<K>();
    unsigned d = (unsigned)__cvta_generic_to_shared(&s[threadIdx.x]);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(in + threadIdx.x), "l"(p) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    out[threadIdx.x] = s[threadIdx.x];
}
template <int K, int CG> __global__ void k_st(const float2* in, float2* out) {
    unsigned long long p = pol<K>();
    float2 v = in[threadIdx.x];
    if (CG) asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(out + threadIdx.x), "f"(v.x), "f"(v.y), "l"(p) : "memory");
    else asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(out + threadIdx.x), "f"(v.x), "f"(v.y), "l"(p) : "memory");
}
#define RUN(name, ...) do { __VA_ARGS__; cudaError_t e = cudaDeviceSynchronize(); printf("%-28s %s\n", name, cudaGetErrorString(e)); if (e != cudaSuccess) return 1; } while (0)
int main() {
    float4 *a, *b; cudaMalloc(&a, 4096); cudaMalloc(&b, 4096); cudaMemset(a, 0, 4096);
    RUN("cp evict_first", (k_cp<1><<<1, 128>>>(a, b)));
    RUN("cp evict_last", (k_cp<2><<<1, 128>>>(a, b)));
    RUN("cp evict_unchanged", (k_cp<3><<<1, 128>>>(a, b)));
    RUN("st evict_first", (k_st<1, 0><<<1, 128>>>((float2*)a, (float2*)b)));
    RUN("st evict_last", (k_st<2, 0><<<1, 128>>>((float2*)a, (float2*)b)));
    RUN("st.cg evict_last", (k_st<2, 1><<<1, 128>>>((float2*)a, (float2*)b)));
    RUN("st.cg evict_first", (k_st<1, 1><<<1, 128>>>((float2*)a, (float2*)b)));
    RUN("st evict_normal", (k_st<0, 0><<<1, 128>>>((float2*)a, (float2*)b)));
    RUN("cp evict_normal", (k_cp<0><<<1, 128>>>(a, b)));
    return 0;
}
