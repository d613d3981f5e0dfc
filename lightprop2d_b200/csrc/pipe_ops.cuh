// Host-side entry of one (T, N, E, LP) instance of the persistent pipeline kernel.
#pragma once
#include <cuda_runtime.h>
#include "pipeline.cuh"

namespace lp2d {

template <typename T>
struct PipeOps {
    int n, e, w;
    int lines;          // lines (rows / adjacent columns) of one work item
    int tiles;          // items per (stage, group)
    int threads;
    size_t smem_x;      // exchange area + prefetched tile
    size_t smem_f;      // per-tile factor table (H / lens), only for steps that need it
    cudaError_t (*launch)(const PipeArgs<T>& a, unsigned grid, bool need_factor, cudaStream_t s);
    // opt in to large dynamic shared memory; resident CTAs per SM with / without the factor table
    cudaError_t (*configure)(int* ctas_per_sm_factor, int* ctas_per_sm_plain);
    // ticket form (one CTA per item)
    cudaError_t (*launch_ticket)(const PipeArgs<T>& a, unsigned grid, bool need_factor, cudaStream_t s);
    cudaError_t (*configure_ticket)(int* ctas_per_sm_factor, int* ctas_per_sm_plain);
};

const PipeOps<float>* find_pipe_f32(int n, int variant);
const PipeOps<double>* find_pipe_f64(int n, int variant);

extern long long g_launches;

template <typename T, int N, int E, int W, int MINB>
struct PipeInst {
    using K = PCfg<T, N, E, W>;
    static cudaError_t launch(const PipeArgs<T>& a, unsigned grid, bool need_factor, cudaStream_t s) {
        const size_t smem = K::SMEM_X + K::SMEM_P + (need_factor ? K::SMEM_F : 0);
        pipe_kernel<T, N, E, W, MINB><<<grid, K::THREADS, smem, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t configure(int* with_factor, int* plain) {
        auto kern = pipe_kernel<T, N, E, W, MINB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(K::SMEM_X + K::SMEM_P + K::SMEM_F));
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(with_factor, kern, K::THREADS, K::SMEM_X + K::SMEM_P + K::SMEM_F);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(plain, kern, K::THREADS, K::SMEM_X + K::SMEM_P);
    }
    static cudaError_t launch_ticket(const PipeArgs<T>& a, unsigned grid, bool need_factor, cudaStream_t s) {
        const size_t smem = K::SMEM_X + (need_factor ? K::SMEM_F : 0);
        ticket_kernel<T, N, E, W, MINB><<<grid, K::THREADS, smem, s>>>(a);
        ++g_launches;
        return cudaGetLastError();
    }
    static cudaError_t configure_ticket(int* with_factor, int* plain) {
        auto kern = ticket_kernel<T, N, E, W, MINB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(K::SMEM_X + K::SMEM_F));
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(with_factor, kern, K::THREADS, K::SMEM_X + K::SMEM_F);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(plain, kern, K::THREADS, K::SMEM_X);
    }
    static constexpr PipeOps<T> entry() {
        return PipeOps<T>{N, E, W, K::LINES, K::TILES, K::THREADS, K::SMEM_X + K::SMEM_P, K::SMEM_F, &launch, &configure,
                          &launch_ticket, &configure_ticket};
    }
};

}  // namespace lp2d
