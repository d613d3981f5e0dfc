// Host-side dispatch from a runtime grid size to the compiled (T, N) kernel set.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "kernels.cuh"

namespace lp2d {

template <typename T>
struct SizeOps {
    int n;          // grid points per axis
    int e;          // samples per thread
    int tw_size;    // twiddle table entries
    int w_row;      // rows per CTA
    int w_col;      // columns per CTA
    int variant;    // alternative tiling of the same size (tuning)
    // pre: PRE_NONE / PRE_LENS / PRE_H ; nunits = beams*N rows
    cudaError_t (*row)(int pre, const LineArgs<T>& a, cudaStream_t s);
    // mid: MID_NONE / MID_H ; a.nlines = beams*N/w_col tiles
    cudaError_t (*col)(int mid, const LineArgs<T>& a, cudaStream_t s);
    cudaError_t (*configure)();                         // opt in to > 48 KB dynamic shared memory
    void (*twiddles)(std::vector<double>& re_im);       // interleaved re, im in FP64
};

const SizeOps<float>* find_ops_f32(int n, int variant);
const SizeOps<double>* find_ops_f64(int n, int variant);

extern long long g_launches;   // kernels launched by this library

}  // namespace lp2d
