// The line kernel: one CTA transforms W lines (rows or a tile of W adjacent
// columns) of one beam, with the pointwise stages of the propagation path fused
// around the transform:
//
//   load -> [mask multiply] -> [lens phase | transfer function H]  -> [store state]
//        -> FFT (forward or inverse)  or  FFT -> H -> inverse FFT (fused column pass)
//        -> scale -> [store field/spectrum] [store |E|^2]
//
// Rows are contiguous: consecutive threads load consecutive samples.  Columns are
// processed W at a time so that every global access of a warp is made of W*sizeof
// contiguous bytes per grid row (64 B for W=8 complex64).
//
// Reference semantics reproduced here (lightprop2d/beam2d.py, /root/reference):
//   * H = exp(i*2*pi*(trunc(Kz*z) - Kz*z)), Kz = nan_to_num(sqrt(k0^2 - Kx^2 - Ky^2)),
//     phase formed in FP64 with the reference's operation order (beam2d.py:307-308,
//     485-489); Kz^2 comes from two host-built FP64 tables a[i] = k0^2 - Kx[i]^2 and
//     b[j] = Ky[j]^2 so the subtraction order matches numpy's broadcasting bit for bit.
//   * lens phase (X^2 + Y^2) * k0 / (2 f) in FP64, k0 = 1/wl (beam2d.py:508-510).
//   * inverse transform scaled by 1/N per axis (numpy ifft2 convention).
#pragma once
#include "fft_core.cuh"

namespace lp2d {

enum { AX_ROW = 0, AX_COL = 1 };
// PRE_HPLANE: the transfer function is read from a plane the host has built once for this z (factor_plane_kernel, folded by
// the kx -> -kx symmetry) instead of being evaluated in FP64 per sample: a z-sweep with a constant step, single field,
// spends a quarter of its row pass on sqrt / sincospi otherwise (profiles/r3s_h_cost.log)
enum { PRE_NONE = 0, PRE_LENS = 1, PRE_H = 2, PRE_HPLANE = 3 };
enum { MID_NONE = 0, MID_H = 1, MID_HPLANE = 2 };     // MID_HPLANE: as MID_H with H read from the plane of PRE_HPLANE
enum { F_INVERSE = 1,
       F_IN_STREAM = 2,     // input is read once: evict-first loads
       F_OUT_STREAM = 4,    // output is not re-read soon: evict-first stores
       F_STATE_STREAM = 8   // same for out_pre / out_mid
};

// cache-policy aware accessors: streaming data must not push the inter-pass
// intermediate out of L2
template <typename C> __device__ __forceinline__ C ld_cx(const C* p, bool stream) { return stream ? __ldcs(p) : *p; }
template <typename C> __device__ __forceinline__ void st_cx(C* p, C v, bool stream) { if (stream) __stcs(p, v); else *p = v; }
enum { MASK_NONE = 0, MASK_CX = 1, MASK_REAL = 2, MASK_U8 = 3 };

template <typename T>
struct LineArgs {
    const cx<T>* in;        // input plane(s)
    cx<T>* out;             // transformed output (may alias in), or nullptr
    cx<T>* out_pre;         // state after the pointwise pre-stage (spectrum*H, field*lens, field*mask), or nullptr
    cx<T>* out_mid;         // fused column pass only: the 2-D spectrum after H (reference's spectrum state), or nullptr
    T* inten;               // |out|^2, or nullptr
    const void* mask;       // optional multiplier applied at load
    const cx<T>* tw;        // twiddle table of the (N, E) plan
    const double* a_tab;    // k0^2 - Kx[i]^2
    const double* b_tab;    // Ky[j]^2
    const double* x2_tab;   // X[i]^2
    const cx<T>* fac;       // PRE_HPLANE: H plane [row j][idx = min(i, N - i)], row stride fac_stride
    int fac_stride;
    double z;               // propagation distance
    double lens_k0;         // 1 / wl
    double lens_2f;         // 2 * f
    T scale;                // applied to the output of the transform
    long long beam_stride;  // samples between consecutive beams (N*N)
    long long mask_beam_stride;  // 0: one mask for every beam
    int line0;              // rows only: global index of the first line of the launch (slab passes), else 0
    int seg_shift;          // slab row passes only: != 0 -> `in` is the receive buffer of the tile exchange, n >> seg_shift
    int seg_rows;           //   tiles of (seg_rows lines x (1 << seg_shift) samples), first line of the launch first
    int nlines;             // lines (rows) or column tiles in ONE group
    int groups;             // beam groups a CTA walks over (>= 1)
    long long group_beams;  // beams per group
    long long nbeams;       // beams in the launch (the last group may be partial)
    int mask_kind;
    int flags;
};

// ---------------------------------------------------------------- pointwise factors
template <typename T> struct Trig;
template <> struct Trig<float> {
    // cycles in (-1, 1) -> (cos, sin) of 2*pi*cycles
    __device__ __forceinline__ static float2 unit_from_cycles(double cyc) {
        float s, c;
        sincospif(2.0f * (float)cyc, &s, &c);
        return make_float2(c, s);
    }
    __device__ __forceinline__ static float2 unit_from_radians(double ph) {
        // range-reduce in FP64, evaluate in FP32
        double r = ph * 0.15915494309189533576888376337251;   // 1/(2 pi)
        r -= rint(r);
        return unit_from_cycles(r);
    }
};
template <> struct Trig<double> {
    __device__ __forceinline__ static double2 unit_from_cycles(double cyc) {
        double s, c;
        sincospi(2.0 * cyc, &s, &c);
        return make_double2(c, s);
    }
    __device__ __forceinline__ static double2 unit_from_radians(double ph) {
        double s, c;
        sincos(ph, &s, &c);
        return make_double2(c, s);
    }
};

// exp(i*2*pi*(trunc(Kz z) - Kz z)); a_i - b_j is Kz^2 exactly as the reference forms it
template <typename T>
__device__ __forceinline__ cx<T> transfer_factor(double a_i, double b_j, double z) {
    const double s = __dsub_rn(a_i, b_j);
    const double kz = (s > 0.0) ? __dsqrt_rn(s) : 0.0;      // NaN (evanescent) -> 0 -> H = 1
    const double d = __dmul_rn(kz, z);
    const double fr = __dsub_rn(trunc(d), d);                // exact; in (-1, 1)
    return Trig<T>::unit_from_cycles(fr);
}

template <typename T>
__device__ __forceinline__ cx<T> lens_factor(double x2_i, double y2_j, double k0, double two_f) {
    const double ph = __ddiv_rn(__dmul_rn(__dadd_rn(x2_i, y2_j), k0), two_f);
    return Trig<T>::unit_from_radians(ph);
}

// ------------------------------------------------------------------ configuration
template <typename T, int N, int E, int AXIS, int W, int PRE = PRE_NONE>
struct KCfg {
    using LP = LinePlan<N, E>;
    static constexpr int TPL = LP::TPL;
    static constexpr int THREADS = W * TPL;
    static constexpr int SYNC = (AXIS == AX_COL) ? (THREADS <= 32 ? SYNC_WARP : SYNC_CTA)
                              : (TPL <= 32 ? SYNC_WARP : (W == 1 ? SYNC_CTA : SYNC_NAMED));
    static constexpr int S128 = 128 / (int)sizeof(cx<T>);          // samples per 128 B
    static constexpr int RES = (AXIS == AX_COL) ? (W >= S128 ? 1 : S128 / W) : (TPL < S128 ? TPL : 0);
    static constexpr int LS = LP::LINE_SLOTS + ((RES - LP::LINE_SLOTS % S128) + S128) % S128;
    static constexpr size_t SMEM_X = (LP::npass() > 1) ? (size_t)W * LS * sizeof(cx<T>) : 0;   // exchange
    // PRE_H keeps the transfer function of the CTA's W lines in shared memory so that it is
    // computed once and reused for every beam group the CTA walks over
    // (8192 complex128 points: exchange + H would not fit 227 KB; H is then evaluated where it is used, for every beam)
    static constexpr bool H_IN_SMEM = (PRE == PRE_H) && (SMEM_X + (size_t)W * N * sizeof(cx<T>) <= 226 * 1024);
    static constexpr size_t SMEM_H = H_IN_SMEM ? (size_t)W * N * sizeof(cx<T>) : 0;
    static constexpr size_t SMEM = SMEM_X + SMEM_H;
    // resident CTAs per SM the register allocation should allow: data registers are
    // 2*E*sizeof(T)/4, the butterflies need about as many again
    static constexpr int REGS_EST = 4 * E * (int)sizeof(T) / 4;
    static constexpr int MINB_R = 65536 / (THREADS * REGS_EST) < 1 ? 1 : 65536 / (THREADS * REGS_EST);
    static constexpr int MINB_S = SMEM == 0 ? 32 : (int)((227 * 1024) / (SMEM + 1024)) < 1 ? 1 : (int)((227 * 1024) / (SMEM + 1024));
    static constexpr int MINB_T = 2048 / THREADS < 1 ? 1 : 2048 / THREADS;
    static constexpr int MINB = MINB_R < MINB_S ? (MINB_R < MINB_T ? MINB_R : MINB_T) : (MINB_S < MINB_T ? MINB_S : MINB_T);
    static_assert(THREADS <= 1024, "block too large");
    static_assert(SYNC != SYNC_NAMED || (W <= 15 && TPL % 32 == 0), "named barriers");
};

// One CTA: W lines (rows, or a tile of W adjacent columns).  The batch is seen as p.groups
// slabs of (p.nlines lines) each; the CTA handles its lines in every slab, so per-line
// constants (the transfer function) are built once and reused across beams.
template <typename T, int N, int E, int AXIS, int W, int PRE, int MID>
__global__ void __launch_bounds__(KCfg<T, N, E, AXIS, W, PRE>::THREADS, KCfg<T, N, E, AXIS, W, PRE>::MINB)
line_kernel(const LineArgs<T> p) {
    using K = KCfg<T, N, E, AXIS, W, PRE>;
    constexpr int TPL = K::TPL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);

    const int tid = threadIdx.x;
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;     // line within the CTA
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;     // thread within the line
    cx<T>* sline = smem + l * K::LS;

    // which line of which beam; "fixed" is the grid index that is constant along the line
    long long base, mbase, beam;
    int fixed;
    int lrel = 0;                                             // line relative to the first line of the launch
    bool in_range;
    if constexpr (AXIS == AX_ROW) {
        lrel = (int)blockIdx.x * W + l;
        const long long line = (long long)blockIdx.x * W + l + p.line0;
        in_range = line < (long long)p.line0 + p.nlines;
        beam = line / N;
        fixed = (int)(line % N);                              // row j
        base = beam * p.beam_stride + (long long)fixed * N;
        mbase = beam * p.mask_beam_stride + (long long)fixed * N;
    } else {
        constexpr int TILES = N / W;
        const long long tile = blockIdx.x;
        in_range = tile < p.nlines;
        beam = tile / TILES;
        fixed = (int)(tile % TILES) * W + l;                  // column i
        base = beam * p.beam_stride + fixed;
        mbase = beam * p.mask_beam_stride + fixed;
    }
    constexpr long long STRIDE = (AXIS == AX_ROW) ? 1 : N;
    const bool in_stream = (p.flags & F_IN_STREAM) != 0;
    const bool out_stream = (p.flags & F_OUT_STREAM) != 0;
    const bool state_stream = (p.flags & F_STATE_STREAM) != 0;
    const bool inv = (p.flags & F_INVERSE) != 0;

    cx<T>* sH = nullptr;
    if constexpr (PRE == PRE_H && K::H_IN_SMEM) {
        // H for this thread's samples, kept at [line][pos]: written and read by the same thread
        sH = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X) + l * N;
        const double fx = (AXIS == AX_ROW) ? __ldg(&p.b_tab[fixed]) : __ldg(&p.a_tab[fixed]);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int pos = t + TPL * m;
            const double a_i = (AXIS == AX_ROW) ? __ldg(&p.a_tab[pos]) : fx;
            const double b_j = (AXIS == AX_ROW) ? fx : __ldg(&p.b_tab[pos]);
            sH[pos] = transfer_factor<T>(a_i, b_j, p.z);
        }
    }

    for (int g = 0; g < p.groups; ++g) {
        const bool active = in_range && (beam + (long long)g * p.group_beams < p.nbeams);
        const long long gb = base + (long long)g * p.group_beams * p.beam_stride;

        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int pos = t + TPL * m;
            if (AXIS == AX_ROW && p.seg_shift) {
                // input in all-to-all tiles: tile (pos >> seg_shift) holds seg_rows lines of (1 << seg_shift) samples
                const long long si = ((((long long)(pos >> p.seg_shift) * p.seg_rows + lrel) << p.seg_shift) + (pos & ((1 << p.seg_shift) - 1)));
                v[m] = active ? ld_cx(&p.in[si], in_stream) : mk<T>(0, 0);
            } else {
                v[m] = active ? ld_cx(&p.in[gb + pos * STRIDE], in_stream) : mk<T>(0, 0);
            }
        }

        if (p.mask_kind != MASK_NONE && active) {
            const long long gm = mbase + (long long)g * p.group_beams * p.mask_beam_stride;
            if (p.mask_kind == MASK_CX) {
                const cx<T>* mk_ = static_cast<const cx<T>*>(p.mask);
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], mk_[gm + (t + TPL * m) * STRIDE]);
            } else if (p.mask_kind == MASK_REAL) {
                const T* mk_ = static_cast<const T*>(p.mask);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const T f = mk_[gm + (t + TPL * m) * STRIDE];
                    v[m].x *= f; v[m].y *= f;
                }
            } else {
                const unsigned char* mk_ = static_cast<const unsigned char*>(p.mask);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const T f = mk_[gm + (t + TPL * m) * STRIDE] ? (T)1 : (T)0;
                    v[m].x *= f; v[m].y *= f;
                }
            }
        }

        if constexpr (PRE == PRE_LENS) {
            const double f2 = __ldg(&p.x2_tab[fixed]);
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const double g2 = __ldg(&p.x2_tab[t + TPL * m]);
                // (X^2 + Y^2): X indexes columns, Y rows; the sum is commutative
                v[m] = cmul(v[m], lens_factor<T>(AXIS == AX_ROW ? g2 : f2, AXIS == AX_ROW ? f2 : g2, p.lens_k0, p.lens_2f));
            }
        }
        if constexpr (PRE == PRE_H && K::H_IN_SMEM) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], sH[t + TPL * m]);
        }
        if constexpr (PRE == PRE_H && !K::H_IN_SMEM) {
            const double fx = (AXIS == AX_ROW) ? __ldg(&p.b_tab[fixed]) : __ldg(&p.a_tab[fixed]);
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int pos = t + TPL * m;
                const double a_i = (AXIS == AX_ROW) ? __ldg(&p.a_tab[pos]) : fx;
                const double b_j = (AXIS == AX_ROW) ? fx : __ldg(&p.b_tab[pos]);
                v[m] = cmul(v[m], transfer_factor<T>(a_i, b_j, p.z));
            }
        }
        if constexpr (PRE == PRE_HPLANE && AXIS == AX_ROW) {
            const cx<T>* frow = p.fac + (long long)fixed * p.fac_stride;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int pos = t + TPL * m;
                v[m] = cmul(v[m], __ldg(frow + (pos <= N / 2 ? pos : N - pos)));
            }
        }
        if (p.out_pre != nullptr && active) {
#pragma unroll
            for (int m = 0; m < E; ++m) st_cx(&p.out_pre[gb + (t + TPL * m) * STRIDE], v[m], state_stream);
        }

        if constexpr (MID == MID_H || MID == MID_HPLANE) {
            // column pass of the fresh chain: FFT_y, transfer function, inverse FFT_y
            line_fft<T, N, E, K::SYNC>(v, sline, p.tw, t, l);
            if constexpr (MID == MID_H) {
                const double a_i = __ldg(&p.a_tab[fixed]);        // AXIS == AX_COL
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const cx<T> h = transfer_factor<T>(a_i, __ldg(&p.b_tab[t + TPL * m]), p.z);
                    v[m] = cmul(v[m], h);
                }
            } else {
                // plane [row j][min(i, N - i)]: the W threads of a grid row read W consecutive (or mirrored) entries
                const int fi = fixed <= N / 2 ? fixed : N - fixed;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], __ldg(p.fac + (long long)(t + TPL * m) * p.fac_stride + fi));
            }
            if (p.out_mid != nullptr && active) {
#pragma unroll
                for (int m = 0; m < E; ++m) st_cx(&p.out_mid[gb + (t + TPL * m) * STRIDE], v[m], state_stream);
            }
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
            line_fft<T, N, E, K::SYNC>(v, sline, p.tw, t, l);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
        } else {
            if (inv) {
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
            }
            line_fft<T, N, E, K::SYNC>(v, sline, p.tw, t, l);
            if (inv) {
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
            }
        }

        if (active) {
            const T sc = p.scale;
            if (sc != (T)1) {
#pragma unroll
                for (int m = 0; m < E; ++m) { v[m].x *= sc; v[m].y *= sc; }
            }
            if (p.out != nullptr) {
#pragma unroll
                for (int m = 0; m < E; ++m) st_cx(&p.out[gb + (t + TPL * m) * STRIDE], v[m], out_stream);
            }
            if (p.inten != nullptr) {
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(&p.inten[gb + (t + TPL * m) * STRIDE], v[m].x * v[m].x + v[m].y * v[m].y);
            }
        }
    }
}

}  // namespace lp2d
