// In-register radix butterflies and the per-line Stockham pass schedule.
//
// A "line" is one 1-D transform of length N (a grid row or a grid column).  It is
// processed by TPL = N / E threads; every thread keeps E complex samples in
// registers at line positions  t + TPL*m  (m = 0..E-1).  That ownership pattern is
// (a) what a coalesced global load/store wants (consecutive threads touch
// consecutive samples) and (b) invariant under the autosort passes below, so the
// first load and the last store go straight between HBM and registers and only
// the passes in between exchange data through shared memory.
//
// Pass p uses radix R_p (R_p <= E, so each thread runs E/R_p butterflies) with
// Ns_p = prod_{q<p} R_q:
//     j      = t + TPL*q                      butterfly index, q < E/R_p
//     in[r]  = x[j + r*N/R_p]                 == register slot q + r*(E/R_p)
//     in[r] *= exp(-2*pi*i * (j mod Ns) * r / (Ns*R_p))          (table lookup)
//     out    = DFT_R(in)
//     y[(j/Ns)*Ns*R + (j mod Ns) + r*Ns] = out[r]
// After the last pass (Ns*R == N) the destination index is again j + r*N/R, i.e.
// the thread's own register slots: natural-order output, no bit reversal pass.
//
// The inverse transform is the same code with re/im swapped on the way in and on
// the way out (ifft(x) = swap(fft(swap(x))) / N).
#pragma once
#include <cuda_runtime.h>

#ifndef LP2D_TWO_LEVEL_TWIDDLES
#define LP2D_TWO_LEVEL_TWIDDLES 0
#endif

namespace lp2d {

template <typename T> struct Cx2;
template <> struct Cx2<float>  { using type = float2; };
template <> struct Cx2<double> { using type = double2; };
template <typename T> using cx = typename Cx2<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
template <typename C> __device__ __forceinline__ C cswap(C a) { C r; r.x = a.y; r.y = a.x; return r; }
template <typename C> __device__ __forceinline__ C cconj(C a) { C r; r.x = a.x; r.y = -a.y; return r; }

// complex64 arithmetic in the packed f32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2):
// a complex sample is one 64-bit register pair, a complex add is ONE instruction, a complex
// multiply is two.  The half swap, the per-half negation and the scalar broadcast written
// below with make_float2 are operand modifiers of those instructions (R.F32x2.LO_HI, .NP,
// R.F32), not extra moves.  Enabled per translation unit with LP2D_PACKED_F32 (the fused step
// kernel, which is issue bound); the one-kernel-per-pass line kernels are latency bound and
// measured ~5 % slower with the pair-aligned register allocation, so they keep scalar FP32.
#if defined(LP2D_PACKED_F32) && defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) { return __fadd2_rn(a, b); }
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
template <> __device__ __forceinline__ float2 cmul<float2>(float2 a, float2 b) {
    // (ax bx - ay by, ax by + ay bx) = bx * (ax, ay) + by * (-ay, ax): the halves of b enter as
    // broadcast scalars, the swap + half negation of a is an operand modifier of FFMA2
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
#endif
// a * s for a real s
__device__ __forceinline__ float2 cscale(float2 a, float s) {
#if defined(LP2D_PACKED_F32) && defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fmul2_rn(a, make_float2(s, s));
#else
    return make_float2(a.x * s, a.y * s);
#endif
}
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// (a.x * s.x, a.y * s.y)
__device__ __forceinline__ float2 cmul_each(float2 a, float2 s) {
#if defined(LP2D_PACKED_F32) && defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    return __fmul2_rn(a, s);
#else
    return make_float2(a.x * s.x, a.y * s.y);
#endif
}
__device__ __forceinline__ double2 cmul_each(double2 a, double2 s) { return make_double2(a.x * s.x, a.y * s.y); }

// cos(2*pi*k/32), k = 0..8
__host__ __device__ constexpr double kCos32(int k) {
    return k == 0 ? 1.0
         : k == 1 ? 0.98078528040323044912618223613424
         : k == 2 ? 0.92387953251128675612818318939679
         : k == 3 ? 0.83146961230254523707878837761791
         : k == 4 ? 0.70710678118654752440084436210485
         : k == 5 ? 0.55557023301960222474283081394853
         : k == 6 ? 0.38268343236508977172845998403040
         : k == 7 ? 0.19509032201612826784828486847702
         : 0.0;
}

// d * exp(-2*pi*i * k32 / 32) for 0 <= k32 < 16 (k32 is a compile-time value after unrolling)
template <typename T>
__device__ __forceinline__ cx<T> mul_w32(cx<T> d, int k32) {
    if (k32 == 0) return d;
    if (k32 == 8) return mk<T>(d.y, -d.x);                       // * (-i)
    if (k32 == 4) {                                              // * (1-i)/sqrt2 : (d + (-i) d) h
        const T h = (T)kCos32(4);
        return cscale(cadd(d, mk<T>(d.y, -d.x)), h);
    }
    if (k32 == 12) {                                             // * (-1-i)/sqrt2 : ((-i) d - d) h
        const T h = (T)kCos32(4);
        return cscale(csub(mk<T>(d.y, -d.x), d), h);
    }
    // general: w = c - i s  with c = cos(2 pi k/32), s = sin(2 pi k/32):  d w = c d + s ((-i) d)
    const T c = (T)(k32 <= 8 ? kCos32(k32) : -kCos32(16 - k32));
    const T s = (T)(k32 <= 8 ? kCos32(8 - k32) : kCos32(k32 - 8));
    return cmul(mk<T>(c, -s), d);
}

__host__ __device__ constexpr int ilog2(int v) { int l = 0; while (v > 1) { v >>= 1; ++l; } return l; }
__host__ __device__ constexpr int bitrev(int k, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((k >> i) & 1) << (bits - 1 - i);
    return r;
}

// R-point forward DFT on registers, natural order in and out (R = 2..32, power of two).
// Radix-2 decimation in frequency; the twiddles and the final bit reversal are
// compile-time after unrolling (register renaming only).
// Hooks: `pre(i, x[i])` is called when input i is first touched (its value is still the input), `post(k, y[k])`
// when output k (natural order) has its final value.  They let the caller spread the global stores of the state
// before the transform / of the result after it between the butterflies instead of issuing 32 of them in a row
// (a store burst fills the load/store queue and stalls the warp; profiles/r2e_*).
struct NoHook {
    template <typename C> __device__ __forceinline__ void operator()(int, const C&) const {}
};

template <typename T, int R, typename PRE = NoHook, typename POST = NoHook>
__device__ __forceinline__ void dft_reg(cx<T> (&x)[R], PRE pre = PRE(), POST post = POST()) {
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix");
    if (R == 1) { pre(0, x[0]); post(0, x[0]); return; }
#pragma unroll
    for (int s = R / 2; s >= 1; s >>= 1) {
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const int pos = k % s;
            const int i0 = (k / s) * 2 * s + pos;
            const int i1 = i0 + s;
            const cx<T> a = x[i0], b = x[i1];
            if (s == R / 2) { pre(i0, a); pre(i1, b); }
            x[i0] = cadd(a, b);
            x[i1] = mul_w32<T>(csub(a, b), pos * (16 / s));
            if (s == 1) { post(bitrev(i0, ilog2(R)), x[i0]); post(bitrev(i1, ilog2(R)), x[i1]); }
        }
    }
    cx<T> y[R];
#pragma unroll
    for (int k = 0; k < R; ++k) y[k] = x[bitrev(k, ilog2(R))];
#pragma unroll
    for (int k = 0; k < R; ++k) x[k] = y[k];
}

// ------------------------------------------------------------------ pass schedule
template <int N, int E>
struct LinePlan {
    static_assert((N & (N - 1)) == 0 && (E & (E - 1)) == 0 && E <= N && E <= 32 && E >= 2, "sizes");
    static constexpr int TPL = N / E;                       // threads per line
    __host__ __device__ static constexpr int npass() {
        int n = 0, rem = N;
        while (rem > 1) { rem /= (E < rem ? E : rem); ++n; }
        return n;
    }
    __host__ __device__ static constexpr int radix(int p) {
        int rem = N, r = 1;
        for (int i = 0; i <= p; ++i) { r = (E < rem ? E : rem); rem /= r; }
        return r;
    }
    __host__ __device__ static constexpr int ns(int p) {    // product of the radices before pass p
        int v = 1;
        for (int i = 0; i < p; ++i) v *= radix(i);
        return v;
    }
    // twiddle table: for pass p >= 1, entry (r-1)*Ns + k = exp(-2 pi i k r / (Ns R)), r = 1..R-1, k < Ns
    __host__ __device__ static constexpr int tw_offset(int p) {
        int off = 0;
        for (int i = 1; i < p; ++i) off += (radix(i) - 1) * ns(i);
        return off;
    }
    __host__ __device__ static constexpr int tw_size() { return tw_offset(npass()); }
    // shared-memory padding: one extra slot every PADSTEP samples so that the
    // stride-R scatter of pass 0 (Ns = 1) spreads over the banks
    static constexpr int PADSTEP = (radix(0) < 4 ? 4 : radix(0));
    __host__ __device__ static constexpr int pad(int i) { return i + i / PADSTEP; }
    static constexpr int LINE_SLOTS = N + N / PADSTEP;      // padded samples per line
};

// How the TPL threads of one line synchronise around a shared-memory exchange.
// SYNC_CONSUMERS + n: the n transform threads of a warp-specialised CTA (pipe3.cuh) without its helper warps
enum { SYNC_WARP = 0, SYNC_NAMED = 1, SYNC_CTA = 2, SYNC_CONSUMERS = 4096 };

template <int MODE, int TPL>
__device__ __forceinline__ void line_sync(int line_in_cta) {
    if (MODE == SYNC_WARP) {
        __syncwarp();
    } else if (MODE == SYNC_NAMED) {
        asm volatile("bar.sync %0, %1;" ::"r"(line_in_cta + 1), "r"(TPL) : "memory");
    } else if (MODE >= SYNC_CONSUMERS) {
        asm volatile("bar.sync 15, %0;" ::"n"(MODE - SYNC_CONSUMERS) : "memory");
    } else {
        __syncthreads();
    }
}

// One pass of the schedule on the thread's registers v[E].
// TWS: the twiddle table is in shared memory (plain loads) instead of global memory (read-only path)
template <bool TWS, typename C>
__device__ __forceinline__ C tw_load(const C* p) {
    if constexpr (TWS) return *p;
    else return __ldg(p);
}

// pre(m, v[m]) in the first pass / post(m, v[m]) in the last pass: see dft_reg (m = register slot, line position
// t + TPL*m)
template <typename T, int N, int E, int P, int SYNC, bool TWS = false, typename PRE = NoHook, typename POST = NoHook>
__device__ __forceinline__ void line_pass(cx<T> (&v)[E], cx<T>* __restrict__ sline,
                                          const cx<T>* __restrict__ tw, int t, int line_in_cta, PRE pre = PRE(), POST post = POST()) {
    using LP = LinePlan<N, E>;
    constexpr int R = LP::radix(P);
    constexpr int Ns = LP::ns(P);
    constexpr int Q = E / R;
    constexpr int TPL = LP::TPL;
    constexpr bool last = (Ns * R == N);
    constexpr int twoff = LP::tw_offset(P);

    if (P > 0) {
        // registers of the previous pass were scattered to smem; previous readers are done
        line_sync<SYNC, TPL>(line_in_cta);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = sline[LP::pad(t + TPL * m)];
    }
    if (!last) {
        // everyone has read its inputs before anyone overwrites the exchange buffer
        line_sync<SYNC, TPL>(line_in_cta);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int j = t + TPL * q;
        cx<T> u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = v[q + r * Q];
        if (Ns > 1) {
            const int k = j & (Ns - 1);
            if constexpr (R > 16 && sizeof(T) == 4 && LP2D_TWO_LEVEL_TWIDDLES) {
                // radix 32, FP32: w^(k r) = w^(4a k) * w^(b k) with r = 4a + b: 10 table loads (b = 1..3,
                // 4a = 4..28) instead of 31, 52 complex multiplies instead of 31 -- the L1 data pipe, not
                // the FP pipe, is the scarce resource of the fused kernel
                cx<T> wb[3], wa[7];
#pragma unroll
                for (int b = 1; b < 4; ++b) wb[b - 1] = tw_load<TWS>(&tw[twoff + (b - 1) * Ns + k]);
#pragma unroll
                for (int a = 1; a < 8; ++a) wa[a - 1] = tw_load<TWS>(&tw[twoff + (4 * a - 1) * Ns + k]);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    if ((r >> 2) > 0) u[r] = cmul(u[r], wa[(r >> 2) - 1]);
                    if ((r & 3) > 0) u[r] = cmul(u[r], wb[(r & 3) - 1]);
                }
            } else if constexpr (R > 16) {
                // radix 32: 31 twiddle loads hoisted above their uses would need 62 registers on top of
                // the 64 of the data (spills at 128 registers): load and apply them eight at a time
#pragma unroll
                for (int r0 = 1; r0 < R; r0 += 8) {
                    cx<T> w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (r0 + i < R) w[i] = tw_load<TWS>(&tw[twoff + (r0 + i - 1) * Ns + k]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (r0 + i < R) u[r0 + i] = cmul(u[r0 + i], w[i]);
                    asm volatile("" ::: "memory");
                }
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) u[r] = cmul(u[r], tw_load<TWS>(&tw[twoff + (r - 1) * Ns + k]));
            }
        }
        if constexpr (P == 0 && last) {
            dft_reg<T, R>(u, [&](int r, const cx<T>& val) { pre(q + r * Q, val); }, [&](int r, const cx<T>& val) { post(q + r * Q, val); });
        } else if constexpr (P == 0) {
            dft_reg<T, R>(u, [&](int r, const cx<T>& val) { pre(q + r * Q, val); }, NoHook());
        } else if constexpr (last) {
            dft_reg<T, R>(u, NoHook(), [&](int r, const cx<T>& val) { post(q + r * Q, val); });
        } else {
            dft_reg<T, R>(u);
        }
        if (last) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[q + r * Q] = u[r];
        } else {
            const int base = (j / Ns) * (Ns * R) + (j & (Ns - 1));
            // pad(base + r*Ns) is affine in r when a step of Ns never carries into the padding
            // differently: Ns a multiple of PADSTEP, or the whole butterfly inside one PADSTEP block
            constexpr bool affine = (Ns % LP::PADSTEP == 0) || (LP::PADSTEP % (Ns * R) == 0);
            if constexpr (affine) {
                constexpr int RS = (Ns % LP::PADSTEP == 0) ? Ns + Ns / LP::PADSTEP : Ns;
                cx<T>* d = sline + LP::pad(base);
#pragma unroll
                for (int r = 0; r < R; ++r) d[r * RS] = u[r];
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) sline[LP::pad(base + r * Ns)] = u[r];
            }
        }
    }
}

// Forward FFT of the line held in v (natural-strided ownership in, same out).
// Exchange-buffer contract: every scatter is preceded by a sync of the line's
// threads, so back-to-back calls on the same sline (next line of a loop, or the
// inverse transform that follows the transfer function) need no extra barrier.
template <typename T, int N, int E, int SYNC, int P = 0, bool TWS = false, typename PRE = NoHook, typename POST = NoHook>
__device__ __forceinline__ void line_fft(cx<T> (&v)[E], cx<T>* sline, const cx<T>* tw, int t, int line_in_cta, PRE pre = PRE(),
                                         POST post = POST()) {
    line_pass<T, N, E, P, SYNC, TWS, PRE, POST>(v, sline, tw, t, line_in_cta, pre, post);
    if constexpr (P + 1 < LinePlan<N, E>::npass())
        line_fft<T, N, E, SYNC, P + 1, TWS, PRE, POST>(v, sline, tw, t, line_in_cta, pre, post);
}

// line_fft for a transform whose threads change lines at the first exchange: pass 0 runs with the ownership of the
// load (sline0, t0, l0), the passes after it with (sline1, t1, l1).  Column tiles are read with consecutive threads on
// consecutive columns of a grid row (coalesced pieces of a row); when the result is stored TRANSPOSED, a warp should own
// one line instead, so that its stores are whole 256-byte runs of the destination row.  SYNC must cover every thread
// of the tile (SYNC_CTA / SYNC_CONSUMERS).
template <typename T, int N, int E, int SYNC, bool TWS = false, typename POST = NoHook>
__device__ __forceinline__ void line_fft_remap(cx<T> (&v)[E], cx<T>* sline0, int t0, int l0, cx<T>* sline1, int t1, int l1,
                                               const cx<T>* tw, POST post = POST()) {
    static_assert(LinePlan<N, E>::npass() >= 2, "needs an exchange");
    line_pass<T, N, E, 0, SYNC, TWS, NoHook, POST>(v, sline0, tw, t0, l0, NoHook(), post);
    line_fft<T, N, E, SYNC, 1, TWS, NoHook, POST>(v, sline1, tw, t1, l1, NoHook(), post);
}

}  // namespace lp2d
