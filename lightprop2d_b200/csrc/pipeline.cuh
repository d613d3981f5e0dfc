// The persistent pipeline kernel: a whole 2-D step (fft2, lens + fft2, mask + fft2,
// the stateful propagate, the fresh field -> field chain) in ONE launch.
//
// A step is a chain of up to three line passes ("stages": rows, columns, rows).  The
// batch is cut into groups of G beams; a work item is (stage, group, tile) where a
// tile is LINES adjacent lines (rows, or adjacent columns) and the item transforms
// that tile in every beam of the group, so that per-tile constants -- the transfer
// function H(z) or the lens phase, both formed in FP64 -- are built once per item
// into shared memory and reused for the G beams.
//
// Items are numbered so that stage s of group g comes `lag` phases after stage s-1 of
// the same group,
//     phase p :  stage0(group p), stage1(group p - lag), stage2(group p - 2 lag)
// and CTAs take them in that order from one atomic counter.  An item waits (acquire
// load on a per-(stage, group) completion counter) until the previous stage of its
// group is complete; because everything it can wait for has a smaller item number,
// it was taken by a CTA that is already running: no deadlock and no co-residency
// requirement.  With lag * nstages * tiles >= the number of resident CTAs the wait is
// (almost) never taken.
//
// Between two stages the data lives in a small ring of planes owned by the plan
// (ring_slots groups).  The ring is written and re-read within a few microseconds and
// its lines are reused for every group, so it stays in the 126 MB L2: per beam the
// step reads its input and writes its outputs from/to HBM once (the algorithmic
// bytes) and the row <-> column transposition is L2 traffic only.  Ring accesses use
// ld/st.global.cg (L1 is not coherent across SMs); user buffers are streamed
// (ld/st.global.cs) so they do not push the ring out.
//
// Arithmetic: fft_core.cuh (packed f32x2 complex arithmetic for complex64, FP64 for complex128).
// Reference semantics: see kernels.cuh (transfer_factor / lens_factor are shared).
#pragma once
#include "kernels.cuh"

namespace lp2d {

enum { FAC_NONE = 0, FAC_H = 1, FAC_LENS = 2 };
enum { PIPE_MAX_STAGES = 3 };

template <typename T>
struct StageArgs {
    const cx<T>* in;     // user input planes (ignored when in_ring)
    cx<T>* out;          // user output planes (ignored when out_ring), may be nullptr
    cx<T>* out_pre;      // user: state after the pointwise pre-stage, or nullptr
    cx<T>* out_mid;      // user: 2-D spectrum after H in the fused column stage, or nullptr
    T* inten;            // user: |out|^2, or nullptr
    const void* mask;    // user: multiplier applied at load, or nullptr
    long long mask_beam_stride;
    T scale;
    int axis;            // AX_ROW / AX_COL
    int pre;             // FAC_* applied before the transform
    int mid;             // 1: forward transform, H, inverse transform (scaled) in one item
    int inverse;         // inverse transform (ignored when mid)
    int in_ring, out_ring;
    int mask_kind;
};

template <typename T>
struct PipeArgs {
    StageArgs<T> st[PIPE_MAX_STAGES];
    cx<T>* ring;              // ring_slots * G planes
    const cx<T>* tw;
    const double* a_tab;      // k0^2 - Kx[i]^2
    const double* b_tab;      // Ky[j]^2
    const double* x2_tab;     // X[i]^2
    unsigned int* counters;   // [0] next item; [1 + s*ngroups + g] completed items of (s, g)
    double z, lens_k0, lens_2f;
    long long nbeams;
    int nstages, ngroups, G, lag, ring_slots, total_items;
};

template <typename T, int N, int E, int W>
struct PCfg {
    using P = LinePlan<N, E>;
    static constexpr int TPL = P::TPL;
    static constexpr int THREADS = TPL * W;
    static constexpr int LINES = W;                // lines (rows / adjacent columns) of one item
    static constexpr int TILES = N / W;            // items per (stage, group)
    // exchange-buffer line stride in samples: the W threads that share a line position (column
    // mapping) must fall into different banks (128 B of shared memory per wavefront)
    static constexpr int S128 = 128 / (int)sizeof(cx<T>);
    static constexpr int WANT = (W >= S128) ? 1 : S128 / W;
    static constexpr int LS = P::LINE_SLOTS + ((WANT - P::LINE_SLOTS % S128) + S128) % S128;
    static constexpr size_t SMEM_X = (P::npass() > 1) ? (size_t)W * LS * sizeof(cx<T>) : 0;   // exchange
    static constexpr size_t SMEM_P = (size_t)W * N * sizeof(cx<T>);                           // prefetched tile
    static constexpr size_t SMEM_F = (size_t)W * N * sizeof(cx<T>);                           // factor table
    static constexpr int CHUNKS = (int)(SMEM_P / 16);          // 16-byte pieces of one tile
    static constexpr int CPT = CHUNKS / THREADS;               // ... per thread
    static constexpr int CPR = (int)(W * sizeof(cx<T>) / 16);  // ... per grid row of a column tile
    // row stages: a line belongs to TPL consecutive threads, which synchronise among themselves
    // (one warp: __syncwarp, several warps: a named barrier per line); column stages spread a line
    // over all warps of the CTA
    static constexpr int SYNC_ROW = (TPL <= 32) ? SYNC_WARP : ((TPL % 32 == 0 && W <= 15) ? SYNC_NAMED : SYNC_CTA);
    static_assert(THREADS <= 1024 && THREADS >= 32, "block size");
    static_assert(N % W == 0, "tile");
    static_assert(CHUNKS % THREADS == 0 && CPR >= 1 && SMEM_X % 16 == 0, "copy geometry");
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* ctr) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* ctr) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr));
    return v;
}
// completion signal: release at gpu scope (waits for the stores of this CTA that happened before
// the preceding barrier); no acquire side, hence no L1 invalidation
__device__ __forceinline__ void signal_release(unsigned int* ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ cx<T> mask_value(const void* mask, int kind, long long idx) {
    if (kind == MASK_CX) return static_cast<const cx<T>*>(mask)[idx];
    if (kind == MASK_REAL) return mk<T>(static_cast<const T*>(mask)[idx], (T)0);
    return mk<T>(static_cast<const unsigned char*>(mask)[idx] ? (T)1 : (T)0, (T)0);
}

struct PipeItem {
    int s, g, tile;
};

// item -> (stage, group, tile); nstages is 2 or 3 (no runtime division)
template <typename T, int TILES>
__device__ __forceinline__ bool pipe_decode(const PipeArgs<T>& p, int item, PipeItem& it) {
    const int q = item / TILES;
    it.tile = item - q * TILES;
    int phase;
    if (p.nstages == 2) { phase = q >> 1; it.s = q & 1; }
    else { phase = q / 3; it.s = q - 3 * phase; }
    it.g = phase - it.s * p.lag;
    return it.g >= 0 && it.g < p.ngroups;
}

// thread 0: take items until a valid one (or the end) comes up; `first` is an index already taken
template <typename T, int TILES>
__device__ __forceinline__ int pipe_next_valid(const PipeArgs<T>& p, int first) {
    int item = first;
    PipeItem it;
    while (item < p.total_items && !pipe_decode<T, TILES>(p, item, it)) item = (int)atomicAdd(&p.counters[0], 1u);
    return item < p.total_items ? item : -1;
}

// thread 0: what `item` waits for -- the previous stage of its group, and, for a stage that starts
// to fill a ring slot, the last stage of the group that held the slot before
template <typename T, int TILES>
__device__ __forceinline__ bool pipe_ready(const PipeArgs<T>& p, int item) {
    PipeItem it;
    pipe_decode<T, TILES>(p, item, it);
    const unsigned int* done = p.counters + 1;
    const StageArgs<T>& st = p.st[it.s];
    if (it.s > 0 && ld_relaxed(&done[(it.s - 1) * p.ngroups + it.g]) < (unsigned)TILES) return false;
    if (st.out_ring && !st.in_ring && it.g >= p.ring_slots &&
        ld_relaxed(&done[(p.nstages - 1) * p.ngroups + (it.g - p.ring_slots)]) < (unsigned)TILES) return false;
    return true;
}

// thread 0: start the (relaxed, non-blocking) reads of the two counters `item` depends on; the
// values are looked at one tile later, so their latency is never exposed
template <typename T, int TILES>
__device__ __forceinline__ void pipe_poll_issue(const PipeArgs<T>& p, int item, unsigned int& pa, unsigned int& pb) {
    pa = pb = (unsigned)TILES;
    if (item < 0) return;
    PipeItem it;
    pipe_decode<T, TILES>(p, item, it);
    const unsigned int* done = p.counters + 1;
    const StageArgs<T>& st = p.st[it.s];
    if (it.s > 0) pa = ld_relaxed(&done[(it.s - 1) * p.ngroups + it.g]);
    if (st.out_ring && !st.in_ring && it.g >= p.ring_slots)
        pb = ld_relaxed(&done[(p.nstages - 1) * p.ngroups + (it.g - p.ring_slots)]);
}

// origin (bytes) of tile `tile` of beam bi of group g in the input of stage st; ring_slots is a power of two
template <typename T, int N, int W>
__device__ __forceinline__ const char* pipe_src(const PipeArgs<T>& p, const StageArgs<T>& st, int g, int tile, int bi) {
    constexpr long long PLANE = (long long)N * N;
    const cx<T>* plane = st.in_ring ? p.ring + ((long long)(g & (p.ring_slots - 1)) * p.G + bi) * PLANE
                                    : st.in + ((long long)g * p.G + bi) * PLANE;
    return reinterpret_cast<const char*>(plane + (st.axis == AX_ROW ? (long long)tile * W * N : (long long)tile * W));
}

// all threads: start the asynchronous copy of one tile into the prefetch buffer.
// rows: W consecutive rows are one contiguous block, buffer layout [line][pos];
// columns: W adjacent columns, CPR 16-byte pieces per grid row, buffer layout [pos][line]
template <typename T, int N, int E, int W, int AXIS>
__device__ __forceinline__ void pipe_copy(cx<T>* pbuf, const char* src, int tid) {
    using K = PCfg<T, N, E, W>;
    char* dst = reinterpret_cast<char*>(pbuf) + (size_t)tid * 16;
    if constexpr (AXIS == AX_ROW) {
        src += (size_t)tid * 16;
#pragma unroll
        for (int k = 0; k < K::CPT; ++k) cp_async16(dst + (size_t)K::THREADS * 16 * k, src + (size_t)K::THREADS * 16 * k);
    } else {
        static_assert(K::THREADS % K::CPR == 0, "column copy");
        src += (size_t)(tid / K::CPR) * N * sizeof(cx<T>) + (size_t)(tid % K::CPR) * 16;
#pragma unroll
        for (int k = 0; k < K::CPT; ++k)
            cp_async16(dst + (size_t)K::THREADS * 16 * k, src + (size_t)(K::THREADS / K::CPR) * k * N * sizeof(cx<T>));
    }
    cp_async_commit();
}

struct PipeCtl {            // shared memory mailbox written by thread 0
    const char* src;        // origin of beam 0 of the next item
    int item;               // next item of this CTA, or -1
    int ready;              // its dependencies were satisfied when looked at
    int axis;               // its axis
    int pad;
};

struct PipeT0 {             // registers of thread 0: one step of look-ahead
    int i1;                 // next item
    unsigned int f2;        // fetch behind it (taken at the start of the current item)
    unsigned int pa, pb;    // polled dependency counters of i1
};

// thread 0, one tile before the end of the current item: describe the next item in the mailbox (the
// other threads read it after the next-but-one barrier) and start polling its dependencies
template <typename T, int N, int E, int W>
__device__ __forceinline__ void pipe_prepare_next(const PipeArgs<T>& p, PipeT0& t0, PipeCtl* ctl) {
    using K = PCfg<T, N, E, W>;
    const int ni = t0.i1;
    ctl->item = ni;
    if (ni >= 0) {
        PipeItem nit;
        pipe_decode<T, K::TILES>(p, ni, nit);
        const StageArgs<T>& nst = p.st[nit.s];
        ctl->src = pipe_src<T, N, W>(p, nst, nit.g, nit.tile, 0);
        ctl->axis = nst.axis;
    }
    pipe_poll_issue<T, K::TILES>(p, ni, t0.pa, t0.pb);
}

// One tile of one beam: registers v (already loaded) -> pointwise stages -> transform(s) -> stores.
// INV: inverse transform, computed as the forward transform with the output index reversed,
// out[(N - k) mod N] = FFT(x)[k] (times the scale): no conjugation or swap instructions at all.
template <typename T, int N, int E, int W, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void pipe_tile(const PipeArgs<T>& p, const StageArgs<T>& st, cx<T> (&v)[E], long long b,
                                          long long ring_plane, int tile, cx<T>* sline, const cx<T>* fline, int l, int t) {
    using K = PCfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;      // along the line
    constexpr bool REV = MID || INV;
    const int fixed = tile * W + l;                              // grid index constant along the line
    const long long off0 = (AXIS == AX_ROW) ? (long long)fixed * N : fixed;
    // output offset of register m: position t + TPL*m, or its mirror (N - pos) mod N (only pos 0 wraps)
    const long long o_first = (REV ? (long long)((N - t) & (N - 1)) : (long long)t) * PSTRIDE;
    const long long o_base = (REV ? (long long)(N - t) : (long long)t) * PSTRIDE;
    constexpr long long o_step = (REV ? -(long long)TPL : (long long)TPL) * PSTRIDE;

    if constexpr (!IN_RING) {
        if (st.mask_kind != MASK_NONE) {
            const long long mb = b * st.mask_beam_stride + off0;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], mask_value<T>(st.mask, st.mask_kind, mb + (t + TPL * m) * PSTRIDE));
        }
        if (st.pre != FAC_NONE) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[t + TPL * m]);
        }
        if (st.out_pre != nullptr) {
            cx<T>* dst = st.out_pre + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
        }
    }

    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CTA;
    line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
    if constexpr (MID) {
        // fused column stage of the fresh chain: FFT_y, H, inverse FFT_y
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[t + TPL * m]);
        if (st.out_mid != nullptr) {
            cx<T>* dst = st.out_mid + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
        }
        line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
    }
    if (st.scale != (T)1) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cscale(v[m], st.scale);
    }
    if constexpr (OUT_RING) {
        // coherent at L2: the ring is read by other SMs
        cx<T>* dst = p.ring + ring_plane * PLANE + off0;
        __stcg(dst + o_first, v[0]);
        dst += o_base;
#pragma unroll
        for (int m = 1; m < E; ++m) __stcg(dst + m * o_step, v[m]);
    } else {
        if (st.out != nullptr) {
            cx<T>* dst = st.out + b * PLANE + off0;
            __stcs(dst + o_first, v[0]);
            dst += o_base;
#pragma unroll
            for (int m = 1; m < E; ++m) __stcs(dst + m * o_step, v[m]);
        }
        if (st.inten != nullptr) {
            T* dst = st.inten + b * PLANE + off0;
            __stcs(dst + o_first, v[0].x * v[0].x + v[0].y * v[0].y);
            dst += o_base;
#pragma unroll
            for (int m = 1; m < E; ++m) __stcs(dst + m * o_step, v[m].x * v[m].x + v[m].y * v[m].y);
        }
    }
}

// One work item of a statically known stage shape (S = index into p.st, so every stage field is a
// constant-bank operand): all beams of group it.g for tile it.tile.  Per tile: [wait for the
// prefetched tile] [registers <- shared] [start the copy of the NEXT tile] [pointwise stages,
// transform, stores].  Returns with ctl (shared) describing the next item of this CTA.
template <typename T, int N, int E, int W, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void pipe_run_item(const PipeArgs<T>& p, const PipeItem it, int cur_src_is_set, cx<T>* xbuf,
                                              cx<T>* pbuf, cx<T>* fbuf, PipeCtl* ctl, PipeT0& t0, unsigned int*& pend_done,
                                              PipeCtl& next, int tid) {
    using K = PCfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr size_t PLANE_BYTES = (size_t)N * N * sizeof(cx<T>);
    const StageArgs<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;          // line within the tile
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;          // thread within the line
    cx<T>* sline = xbuf + l * K::LS;
    cx<T>* fline = fbuf + l * N;
    const long long b0 = (long long)it.g * p.G;
    const long long slot0 = (long long)(it.g & (p.ring_slots - 1)) * p.G;
    const long long left = p.nbeams - b0;
    const int nb = left < p.G ? (int)left : p.G;
    const char* src0 = pipe_src<T, N, W>(p, st, it.g, it.tile, 0);
    if (tid == 0) {
        t0.f2 = atomicAdd(&p.counters[0], 1u);                     // looked at when this item is finished
        if (nb == 1) pipe_prepare_next<T, N, E, W>(p, t0, ctl);
    }

    if constexpr (MID || !IN_RING) {
        if (MID || st.pre != FAC_NONE) {
            // pointwise factor of this tile, FP64 phase (kernels.cuh), kept at [line][pos]: written and
            // read by the same thread
            const bool lens = !MID && (st.pre == FAC_LENS);
            const int fixed = it.tile * W + l;
            const double fx = __ldg(&(lens ? p.x2_tab : (AXIS == AX_ROW ? p.b_tab : p.a_tab))[fixed]);
            const double* rtab = lens ? p.x2_tab : (AXIS == AX_ROW ? p.a_tab : p.b_tab);
#pragma unroll 2
            for (int m = 0; m < E; ++m) {
                const int pos = t + TPL * m;
                const double rv = __ldg(&rtab[pos]);
                // lens: (X^2 + Y^2), the sum is commutative; H: a[i] - b[j] with i the column, j the row
                fline[pos] = lens ? lens_factor<T>(rv, fx, p.lens_k0, p.lens_2f)
                                  : transfer_factor<T>(AXIS == AX_ROW ? rv : fx, AXIS == AX_ROW ? fx : rv, p.z);
            }
        }
    }

    const cx<T>* pl = (AXIS == AX_ROW) ? pbuf + l * N + t : pbuf + t * W + l;
    for (int bi = 0; bi < nb; ++bi) {
        cp_async_wait_all();
        __syncthreads();                                   // the tile has landed, for every thread
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = pl[(AXIS == AX_ROW) ? TPL * m : TPL * m * W];
        const bool last = (bi == nb - 1);
        if (last && tid == 0)
            ctl->ready = (t0.i1 >= 0 && t0.pa >= (unsigned)K::TILES && t0.pb >= (unsigned)K::TILES) ? 1 : 0;
        __syncthreads();                                   // prefetch buffer consumed; ctl visible
        if (pend_done != nullptr) {
            // every thread is past its stores of the previous item (two barriers ago); the fence comes
            // before this thread has new copies in flight, so it only waits for stores issued long ago
            if (tid == 0) signal_release(pend_done);
            pend_done = nullptr;
        }
        if (!last) {
            pipe_copy<T, N, E, W, AXIS>(pbuf, src0 + (size_t)(bi + 1) * PLANE_BYTES, tid);
        } else {
            // the mailbox is only rewritten after every thread has gone through the barriers of this tile
            next = *ctl;
            if (next.item >= 0 && next.ready) {
                if (next.axis == AX_ROW) pipe_copy<T, N, E, W, AX_ROW>(pbuf, next.src, tid);
                else pipe_copy<T, N, E, W, AX_COL>(pbuf, next.src, tid);
            }
        }
        if (tid == 0 && bi == nb - 2) pipe_prepare_next<T, N, E, W>(p, t0, ctl);
        pipe_tile<T, N, E, W, AXIS, IN_RING, OUT_RING, MID, INV>(p, st, v, b0 + bi, slot0 + bi, it.tile, sline, fline, l, t);
    }
    pend_done = p.counters + 1 + S * p.ngroups + it.g;
}

// The CTA loop.  The copy of the next tile (cp.async, no registers) runs under the arithmetic of
// the current one.  All queue bookkeeping is done by thread 0 one step ahead and without waiting:
// the item after the next one is taken (atomic) at the start of the current item, the dependency
// counters of the next item are read (relaxed) one tile before they are needed, and the completion
// of an item is signalled during the first tile of the next one.  An item whose dependency is not
// yet satisfied is only waited for after the current item has been completed AND signalled (no CTA
// ever blocks while it holds unfinished or unsignalled work, so the in-order argument above still
// gives deadlock freedom).
template <typename T, int N, int E, int W, int MINB>
__global__ void __launch_bounds__(PCfg<T, N, E, W>::THREADS, MINB) pipe_kernel(const __grid_constant__ PipeArgs<T> p) {
    using K = PCfg<T, N, E, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ PipeCtl s_ctl;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* pbuf = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X);
    cx<T>* fbuf = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X + K::SMEM_P);
    const int tid = threadIdx.x;

    PipeT0 t0;
    t0.i1 = -1; t0.f2 = 0; t0.pa = 0; t0.pb = 0;
    if (tid == 0) {
        const int first = pipe_next_valid<T, K::TILES>(p, (int)atomicAdd(&p.counters[0], 1u));
        s_ctl.item = first;
        if (first >= 0) {
            while (!pipe_ready<T, K::TILES>(p, first)) __nanosleep(64);
            t0.i1 = pipe_next_valid<T, K::TILES>(p, (int)atomicAdd(&p.counters[0], 1u));
            PipeItem it;
            pipe_decode<T, K::TILES>(p, first, it);
            s_ctl.src = pipe_src<T, N, W>(p, p.st[it.s], it.g, it.tile, 0);
            s_ctl.axis = p.st[it.s].axis;
        }
    }
    __syncthreads();
    int cur = s_ctl.item;
    if (cur < 0) return;
    if (s_ctl.axis == AX_ROW) pipe_copy<T, N, E, W, AX_ROW>(pbuf, s_ctl.src, tid);
    else pipe_copy<T, N, E, W, AX_COL>(pbuf, s_ctl.src, tid);
    unsigned int* pend_done = nullptr;                     // completion signal still to be sent (uniform)

    for (;;) {
        PipeItem it;
        PipeCtl next;
        pipe_decode<T, K::TILES>(p, cur, it);
        // the stage shapes a step is made of (2 stages: rows user->ring, columns ring->user;
        // 3 stages: rows user->ring, fused columns ring->ring, inverse rows ring->user)
        if (it.s == 0) {
            if (p.st[0].inverse) pipe_run_item<T, N, E, W, 0, AX_ROW, false, true, false, true>(p, it, 0, xbuf, pbuf, fbuf, &s_ctl, t0, pend_done, next, tid);
            else pipe_run_item<T, N, E, W, 0, AX_ROW, false, true, false, false>(p, it, 0, xbuf, pbuf, fbuf, &s_ctl, t0, pend_done, next, tid);
        } else if (it.s == 1) {
            if (p.st[1].mid) pipe_run_item<T, N, E, W, 1, AX_COL, true, true, true, false>(p, it, 0, xbuf, pbuf, fbuf, &s_ctl, t0, pend_done, next, tid);
            else if (p.st[1].inverse) pipe_run_item<T, N, E, W, 1, AX_COL, true, false, false, true>(p, it, 0, xbuf, pbuf, fbuf, &s_ctl, t0, pend_done, next, tid);
            else pipe_run_item<T, N, E, W, 1, AX_COL, true, false, false, false>(p, it, 0, xbuf, pbuf, fbuf, &s_ctl, t0, pend_done, next, tid);
        } else {
            pipe_run_item<T, N, E, W, 2, AX_ROW, true, false, false, true>(p, it, 0, xbuf, pbuf, fbuf, &s_ctl, t0, pend_done, next, tid);
        }
        const int nxt = next.item;
        const int nxt_ready = next.ready;
        if (nxt < 0) break;
        if (tid == 0) t0.i1 = pipe_next_valid<T, K::TILES>(p, (int)t0.f2);
        if (!nxt_ready) {
            __syncthreads();                               // all stores of the item are issued
            if (tid == 0) {
                signal_release(pend_done);
                while (!pipe_ready<T, K::TILES>(p, nxt)) __nanosleep(64);
            }
            pend_done = nullptr;
            __syncthreads();
            if (next.axis == AX_ROW) pipe_copy<T, N, E, W, AX_ROW>(pbuf, next.src, tid);
            else pipe_copy<T, N, E, W, AX_COL>(pbuf, next.src, tid);
        }
        cur = nxt;
    }
    __syncthreads();                                       // all stores of the last item are issued
    if (tid == 0 && pend_done != nullptr) signal_release(pend_done);
}

// --------------------------------------------------------------------------------------------
// Ticket form of the same pipeline: one CTA per work item instead of a persistent loop.  A CTA
// takes its item number from the atomic counter when it starts (so the numbering follows the
// order in which CTAs really start, whatever the hardware dispatch order is), waits until the
// dependencies of the item are complete, transforms the tile in the G beams of the group with
// plain register loads, signals and exits.  Latencies (ticket, dependency poll, completion fence)
// are hidden by the other resident CTAs exactly as in the one-kernel-per-pass path, while the
// intermediate still lives in the L2-resident ring.
template <typename T, int N, int E, int W, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void ticket_item(const PipeArgs<T>& p, const PipeItem it, cx<T>* xbuf, cx<T>* fbuf, int tid) {
    using K = PCfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;
    const StageArgs<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;          // line within the tile
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;          // thread within the line
    cx<T>* sline = xbuf + l * K::LS;
    cx<T>* fline = fbuf + l * N;
    const long long b0 = (long long)it.g * p.G;
    const long long slot0 = (long long)(it.g & (p.ring_slots - 1)) * p.G;
    const long long left = p.nbeams - b0;
    const int nb = left < p.G ? (int)left : p.G;
    const int fixed = it.tile * W + l;
    const long long off0 = ((AXIS == AX_ROW) ? (long long)fixed * N : fixed) + t * PSTRIDE;

    if constexpr (MID || !IN_RING) {
        if (MID || st.pre != FAC_NONE) {
            const bool lens = !MID && (st.pre == FAC_LENS);
            const double fx = __ldg(&(lens ? p.x2_tab : (AXIS == AX_ROW ? p.b_tab : p.a_tab))[fixed]);
            const double* rtab = lens ? p.x2_tab : (AXIS == AX_ROW ? p.a_tab : p.b_tab);
#pragma unroll 2
            for (int m = 0; m < E; ++m) {
                const int pos = t + TPL * m;
                const double rv = __ldg(&rtab[pos]);
                fline[pos] = lens ? lens_factor<T>(rv, fx, p.lens_k0, p.lens_2f)
                                  : transfer_factor<T>(AXIS == AX_ROW ? rv : fx, AXIS == AX_ROW ? fx : rv, p.z);
            }
        }
    }
    for (int bi = 0; bi < nb; ++bi) {
        const cx<T>* src = (IN_RING ? p.ring + (slot0 + bi) * PLANE : st.in + (b0 + bi) * PLANE) + off0;
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = IN_RING ? __ldcg(src + (TPL * m) * PSTRIDE) : __ldcs(src + (TPL * m) * PSTRIDE);
        pipe_tile<T, N, E, W, AXIS, IN_RING, OUT_RING, MID, INV>(p, st, v, b0 + bi, slot0 + bi, it.tile, sline, fline, l, t);
    }
}

template <typename T, int N, int E, int W, int MINB>
__global__ void __launch_bounds__(PCfg<T, N, E, W>::THREADS, MINB) ticket_kernel(const __grid_constant__ PipeArgs<T> p) {
    using K = PCfg<T, N, E, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_item;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* fbuf = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X);
    const int tid = threadIdx.x;
    unsigned int* done = p.counters + 1;

    if (tid == 0) {
        const int item = (int)atomicAdd(&p.counters[0], 1u);
        PipeItem it;
        if (item < p.total_items && pipe_decode<T, K::TILES>(p, item, it)) {
            while (!pipe_ready<T, K::TILES>(p, item)) __nanosleep(100);
            s_item = item;
        } else {
            s_item = -1;
        }
    }
    __syncthreads();
    const int item = s_item;
    if (item < 0) return;
    PipeItem it;
    pipe_decode<T, K::TILES>(p, item, it);
    if (it.s == 0) {
        if (p.st[0].inverse) ticket_item<T, N, E, W, 0, AX_ROW, false, true, false, true>(p, it, xbuf, fbuf, tid);
        else ticket_item<T, N, E, W, 0, AX_ROW, false, true, false, false>(p, it, xbuf, fbuf, tid);
    } else if (it.s == 1) {
        if (p.st[1].mid) ticket_item<T, N, E, W, 1, AX_COL, true, true, true, false>(p, it, xbuf, fbuf, tid);
        else if (p.st[1].inverse) ticket_item<T, N, E, W, 1, AX_COL, true, false, false, true>(p, it, xbuf, fbuf, tid);
        else ticket_item<T, N, E, W, 1, AX_COL, true, false, false, false>(p, it, xbuf, fbuf, tid);
    } else {
        ticket_item<T, N, E, W, 2, AX_ROW, true, false, false, true>(p, it, xbuf, fbuf, tid);
    }
    __syncthreads();                                       // all stores of the item are issued
    if (tid == 0) signal_release(&done[it.s * p.ngroups + it.g]);
}

}  // namespace lp2d
