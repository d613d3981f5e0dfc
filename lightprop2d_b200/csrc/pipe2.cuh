// The fused step kernel ("P2"): a whole 2-D step -- fft2, lens + fft2, mask + fft2, the stateful
// propagate, the fresh field -> field chain -- in ONE persistent launch whose row <-> column
// intermediate lives in an L2-resident ring, so that HBM sees only the algorithmic bytes.
//
// Shape of the kernel (what the ncu captures of the first pipeline, profiles/r1b_*, asked for):
//   * one line = TPL threads with E = N/TPL samples each, E as large as the register file allows
//     (1024-point complex64: E = 32, one warp per line, radix-32 x radix-32: ONE shared-memory
//     exchange per transform, 29 issued instructions per sample per pass with packed f32x2
//     arithmetic); small CTAs of W lines (128 threads), 4 resident per SM at <= 128 registers;
//   * no global load ever feeds a register directly: tiles arrive by cp.async in the CTA's
//     exchange buffer (which is free between the last gather of one transform and the first
//     scatter of the next), issued one tile ahead, across work items too;
//   * row stages are warp-independent (a line is synchronised by __syncwarp / a named barrier of
//     its own threads); only the column stages and the item hand-over use CTA barriers;
//   * the pointwise factor (transfer function H(z) or lens phase) is NOT rebuilt per CTA in FP64
//     (25 % of the instructions of the first pipeline): a factor plane is produced once per
//     (plan, z) by factor_plane_kernel with the reference's FP64 operation order, stays L2
//     resident (4 MB for 1024^2 complex64, folded by the kx -> -kx symmetry), and a CTA copies
//     the W rows it needs next to its tile.
//
// Scheduling is the in-order ticket scheme of pipeline.cuh: items (stage, group, tile) are numbered
// so that stage s of a group comes `lag` groups after stage s-1, CTAs draw numbers from one atomic
// counter and an item waits on the per-(stage, group) completion counter of what it depends on;
// everything it can wait for has a smaller number and was therefore drawn by a running CTA.
#pragma once
#include "kernels.cuh"

namespace lp2d {

enum { P2_FAC_NONE = 0, P2_FAC_H = 1, P2_FAC_LENS = 2 };
enum { P2_MAX_STAGES = 3 };
// how a line position maps to the folded factor row: H is even in kx (index min(pos, N - pos)),
// the lens phase is even in x around the grid centre (index |pos - N/2|)
enum { P2_MAP_FREQ = 0, P2_MAP_CENTRE = 1 };

template <typename T>
struct P2Stage {
    const cx<T>* in;     // user input planes (ignored when in_ring)
    cx<T>* out;          // user output planes (ignored when out_ring), may be nullptr
    cx<T>* out_pre;      // user: state after the pointwise pre-stage, or nullptr
    cx<T>* out_mid;      // user: 2-D spectrum after H in the fused column stage, or nullptr
    T* inten;            // user: |out|^2, or nullptr
    const void* mask;    // user: multiplier applied at load, or nullptr
    long long mask_beam_stride;
    T scale;
    int axis;            // AX_ROW / AX_COL
    int pre;             // P2_FAC_* applied before the transform (from the factor plane)
    int mid;             // 1: forward transform, factor, inverse transform in one item
    int inverse;         // inverse transform (ignored when mid)
    int in_ring, out_ring;
    int mask_kind;
};

template <typename T>
struct P2Args {
    P2Stage<T> st[P2_MAX_STAGES];
    cx<T>* ring;              // ring_slots * G planes
    const cx<T>* tw;          // twiddles of LinePlan<N, E>
    const cx<T>* fac;         // factor plane [N lines][FL], folded (P2_MAP_*)
    unsigned int* counters;   // [0] next item; [1 + s*ngroups + g] completed items of (s, g)
    long long nbeams;
    int nstages, ngroups, G, lag, ring_slots, total_items, fac_map;
    int hints;                // P2_HINT_*
};

template <typename T, int N, int E, int W>
struct P2Cfg {
    using P = LinePlan<N, E>;
    static constexpr int TPL = P::TPL;
    static constexpr int THREADS = TPL * W;
    static constexpr int TILES = N / W;
    static constexpr int SZ = (int)sizeof(cx<T>);
    static constexpr int S128 = 128 / SZ;                  // samples per shared-memory wavefront
    static constexpr int WANT = (W >= S128) ? 1 : S128 / W;
    // exchange line stride: the W lanes that share a line position in the column mapping fall
    // into different banks
    static constexpr int LS = P::LINE_SLOTS + ((WANT - P::LINE_SLOTS % S128) + S128) % S128;
    static constexpr int FL = N / 2 + WANT;                // folded factor row (N/2 + 1 used)
    static constexpr size_t SMEM_X = (size_t)W * LS * SZ;
    static constexpr size_t SMEM_F = (size_t)W * FL * SZ;
    static constexpr int CPT = E * SZ / 16;                // 16-byte pieces of a tile per thread
    static constexpr int CPR = W * SZ / 16;                // ... per grid row of a column tile
    static constexpr int FCH = FL * SZ / 16;               // ... per factor row
    static constexpr int SYNC_ROW = (TPL <= 32) ? SYNC_WARP : ((TPL % 32 == 0 && W <= 15) ? SYNC_NAMED : SYNC_CTA);
    static_assert(P::npass() >= 2, "P2 needs at least one exchange");
    static_assert(THREADS >= 32 && THREADS <= 1024 && THREADS % 32 == 0, "block size");
    static_assert((LS * SZ) % 16 == 0 && (FL * SZ) % 16 == 0 && (E * SZ) % 16 == 0 && CPR >= 1, "copy geometry");
    static_assert(W * N <= W * LS, "landing area");
    // every warp of a CTA signals the completion of its part of an item
    static constexpr unsigned int FULL = (unsigned int)TILES * (THREADS / 32);
};

__device__ __forceinline__ unsigned int p2_ld_relaxed(const unsigned int* ctr) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr));
    return v;
}
__device__ __forceinline__ unsigned int p2_ld_acquire(const unsigned int* ctr) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    return v;
}
__device__ __forceinline__ void p2_signal(unsigned int* ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void p2_cp16(void* smem_dst, const void* gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// L2 eviction priorities: user planes are streamed (evict first), the ring is kept (evict last)
enum { P2_HINT_STREAM_IN = 1, P2_HINT_KEEP_RING = 2, P2_HINT_STREAM_OUT = 4, P2_HINT_BLOCK_TICKET = 8 };
// KIND is a compile-time value: a policy selected at run time ends up in a misaligned uniform
// register pair (ptxas 12.9) and the copy traps as an illegal instruction.
template <int KIND>     // 0 no hint, 1 evict first, 2 evict last
__device__ __forceinline__ unsigned long long p2_policy() {
    unsigned long long pol = 0;
    if (KIND == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (KIND == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int KIND>
__device__ __forceinline__ void p2_cp16_hint(void* smem_dst, const void* gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    if (KIND == 0) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    } else {
        const unsigned long long pol = p2_policy<KIND>();
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
    }
}
template <int KIND>
__device__ __forceinline__ void p2_st_ring(float2* a, float2 v) {
    if (KIND == 0) { __stcg(a, v); return; }
    const unsigned long long pol = p2_policy<KIND>();
    asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(a), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
template <int KIND>
__device__ __forceinline__ void p2_st_ring(double2* a, double2 v) {
    if (KIND == 0) { __stcg(a, v); return; }
    const unsigned long long pol = p2_policy<KIND>();
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void p2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void p2_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct P2Item {
    int s, g, tile;
};

// item -> (stage, group, tile); nstages is 2 or 3
template <typename T, int TILES>
__device__ __forceinline__ bool p2_decode(const P2Args<T>& p, int item, P2Item& it) {
    const int q = item / TILES;
    it.tile = item - q * TILES;
    int phase;
    if (p.nstages == 2) { phase = q >> 1; it.s = q & 1; }
    else { phase = q / 3; it.s = q - 3 * phase; }
    it.g = phase - it.s * p.lag;
    return it.g >= 0 && it.g < p.ngroups;
}

// thread 0: draw numbers until a valid item (or the end) comes up; `first` is a number already drawn
template <typename T, int TILES>
__device__ __forceinline__ int p2_next_valid(const P2Args<T>& p, int first) {
    int item = first;
    P2Item it;
    while (item < p.total_items && !p2_decode<T, TILES>(p, item, it)) item = (int)atomicAdd(&p.counters[0], 1u);
    return item < p.total_items ? item : -1;
}

// thread 0: the two counters `item` depends on -- the previous stage of its group and, for the stage
// that starts to fill a ring slot, the last stage of the group that held the slot before
template <typename T, int TILES>
__device__ __forceinline__ void p2_poll(const P2Args<T>& p, int item, unsigned int& pa, unsigned int& pb, unsigned int full) {
    pa = pb = full;
    if (item < 0) return;
    P2Item it;
    p2_decode<T, TILES>(p, item, it);
    const unsigned int* done = p.counters + 1;
    const P2Stage<T>& st = p.st[it.s];
    // relaxed, non-blocking reads: everything that is read once the counters are complete (the ring)
    // goes to L2 (cp.async.cg), is requested only after a CTA barrier that follows the test of these
    // values, and was made visible at L2 before the counter was incremented (red.release)
    if (it.s > 0) pa = p2_ld_relaxed(&done[(it.s - 1) * p.ngroups + it.g]);
    if (st.out_ring && !st.in_ring && it.g >= p.ring_slots)
        pb = p2_ld_relaxed(&done[(p.nstages - 1) * p.ngroups + (it.g - p.ring_slots)]);
}
template <typename T, int TILES>
__device__ __forceinline__ bool p2_ready(const P2Args<T>& p, int item, unsigned int full) {
    unsigned int pa, pb;
    p2_poll<T, TILES>(p, item, pa, pb, full);
    return pa >= full && pb >= full;
}

// Wait until the dependencies of `item` are complete.  The wait is bounded: if they do not complete
// within ~2 s (they are a few microseconds away in a healthy run) the kernel traps instead of hanging
// the device.
template <typename T, int TILES>
__device__ __forceinline__ void p2_wait_ready(const P2Args<T>& p, int item, unsigned int full) {
    if (p2_ready<T, TILES>(p, item, full)) return;
    const long long t0 = clock64();
    while (!p2_ready<T, TILES>(p, item, full)) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

// origin of tile `tile` of beam bi of group g in the input of stage st
template <typename T, int N, int W>
__device__ __forceinline__ const cx<T>* p2_src(const P2Args<T>& p, const P2Stage<T>& st, int g, int tile, int bi) {
    constexpr long long PLANE = (long long)N * N;
    const cx<T>* plane = st.in_ring ? p.ring + ((long long)(g % p.ring_slots) * p.G + bi) * PLANE
                                    : st.in + ((long long)g * p.G + bi) * PLANE;
    return plane + (st.axis == AX_ROW ? (long long)tile * W * N : (long long)tile * W);
}

// Start the asynchronous copy of one tile into the exchange buffer.
// rows: every line is copied by its own TPL threads to [line * LS + pos] (so that a line only
// ever needs its own threads' synchronisation); columns: W adjacent columns, CPR 16-byte pieces per
// grid row, copied by all threads to [pos * W + line].
template <typename T, int N, int E, int W, int AXIS, int KIND>
__device__ __forceinline__ void p2_copy_tile_k(cx<T>* xbuf, const cx<T>* src, int tid) {
    using K = P2Cfg<T, N, E, W>;
    if constexpr (AXIS == AX_ROW) {
        const int l = tid / K::TPL, t = tid % K::TPL;
        char* dst = reinterpret_cast<char*>(xbuf + l * K::LS) + (size_t)t * 16;
        const char* s = reinterpret_cast<const char*>(src + (long long)l * N) + (size_t)t * 16;
#pragma unroll
        for (int k = 0; k < K::CPT; ++k) p2_cp16_hint<KIND>(dst + (size_t)K::TPL * 16 * k, s + (size_t)K::TPL * 16 * k);
    } else {
        static_assert(K::THREADS % K::CPR == 0, "column copy");
        char* dst = reinterpret_cast<char*>(xbuf) + (size_t)tid * 16;
        const char* s = reinterpret_cast<const char*>(src) + (size_t)(tid / K::CPR) * N * K::SZ + (size_t)(tid % K::CPR) * 16;
#pragma unroll
        for (int k = 0; k < K::CPT; ++k)
            p2_cp16_hint<KIND>(dst + (size_t)K::THREADS * 16 * k, s + (size_t)(K::THREADS / K::CPR) * k * N * K::SZ);
    }
}

// kind: 0 no hint, 1 evict first (user planes), 2 evict last (ring)
template <typename T, int N, int E, int W, int AXIS>
__device__ __forceinline__ void p2_copy_tile(cx<T>* xbuf, const cx<T>* src, int tid, int kind) {
    if (kind == 1) p2_copy_tile_k<T, N, E, W, AXIS, 1>(xbuf, src, tid);
    else if (kind == 2) p2_copy_tile_k<T, N, E, W, AXIS, 2>(xbuf, src, tid);
    else p2_copy_tile_k<T, N, E, W, AXIS, 0>(xbuf, src, tid);
}

// Copy the W folded factor rows of a tile (lines tile*W .. tile*W + W - 1) into fbuf[line * FL + idx].
template <typename T, int N, int E, int W, int AXIS>
__device__ __forceinline__ void p2_copy_factor(cx<T>* fbuf, const cx<T>* fac, int tile, int tid) {
    using K = P2Cfg<T, N, E, W>;
    const char* s = reinterpret_cast<const char*>(fac + (long long)tile * W * K::FL);
    char* d = reinterpret_cast<char*>(fbuf);
    if constexpr (AXIS == AX_ROW) {
        // by the line's own threads
        const int l = tid / K::TPL, t = tid % K::TPL;
        for (int c = t; c < K::FCH; c += K::TPL) p2_cp16(d + ((size_t)l * K::FCH + c) * 16, s + ((size_t)l * K::FCH + c) * 16);
    } else {
        for (int c = tid; c < W * K::FCH; c += K::THREADS) p2_cp16(d + (size_t)c * 16, s + (size_t)c * 16);
    }
}

template <typename T>
__device__ __forceinline__ int p2_in_policy(const P2Args<T>& p, const P2Stage<T>& st) {
    // cp.async with an L2 cache hint traps as an illegal instruction in this kernel (ptxas 12.9 puts the
    // policy in a misaligned uniform register pair) and the hint is dropped altogether in a minimal
    // probe (tools/probe/hint_probe.cu): no hints on the LDGSTS path
    (void)p; (void)st;
    return 0;
}

// all threads: start the copies the first tile of `item` needs (factor rows and beam 0)
template <typename T, int N, int E, int W>
__device__ __forceinline__ void p2_prefetch_item(const P2Args<T>& p, int item, cx<T>* xbuf, cx<T>* fbuf, int tid) {
    using K = P2Cfg<T, N, E, W>;
    P2Item it;
    p2_decode<T, K::TILES>(p, item, it);
    const P2Stage<T>& st = p.st[it.s];
    const cx<T>* src = p2_src<T, N, W>(p, st, it.g, it.tile, 0);
    const int pol = p2_in_policy(p, st);
    if (st.axis == AX_ROW) {
        if (st.pre != P2_FAC_NONE) p2_copy_factor<T, N, E, W, AX_ROW>(fbuf, p.fac, it.tile, tid);
        p2_copy_tile<T, N, E, W, AX_ROW>(xbuf, src, tid, pol);
    } else {
        if (st.mid) p2_copy_factor<T, N, E, W, AX_COL>(fbuf, p.fac, it.tile, tid);
        p2_copy_tile<T, N, E, W, AX_COL>(xbuf, src, tid, pol);
    }
    p2_commit();
}

// folded factor index of line position pos = t + TPL*m (m is a compile-time value after unrolling)
template <int N, int E>
__device__ __forceinline__ int p2_fac_index(int map, int pos, int m) {
    if (map == P2_MAP_FREQ) return (m < E / 2) ? pos : N - pos;
    return (m < E / 2) ? N / 2 - pos : pos - N / 2;
}

template <typename T>
__device__ __forceinline__ cx<T> p2_mask_value(const void* mask, int kind, long long idx) {
    if (kind == MASK_CX) return static_cast<const cx<T>*>(mask)[idx];
    if (kind == MASK_REAL) return mk<T>(static_cast<const T*>(mask)[idx], (T)0);
    return mk<T>(static_cast<const unsigned char*>(mask)[idx] ? (T)1 : (T)0, (T)0);
}

// One pass of the Stockham schedule (fft_core.cuh) with a hook: `after_gather()` runs once the
// LAST gather of the transform is done, i.e. at the point from which the transform no longer
// touches shared memory.  The hook is responsible for the synchronisation it needs before it
// reuses the exchange buffer.
template <typename T, int N, int E, int SYNC, int P, typename Hook>
__device__ __forceinline__ void p2_pass(cx<T> (&v)[E], cx<T>* __restrict__ sline, const cx<T>* __restrict__ tw, int t,
                                        int line_in_cta, Hook& after_gather) {
    using LP = LinePlan<N, E>;
    constexpr int R = LP::radix(P);
    constexpr int Ns = LP::ns(P);
    constexpr int Q = E / R;
    constexpr int TPL = LP::TPL;
    constexpr bool last = (Ns * R == N);
    constexpr int twoff = LP::tw_offset(P);

    if (P > 0) {
        line_sync<SYNC, TPL>(line_in_cta);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = sline[LP::pad(t + TPL * m)];
    }
    if (last) after_gather();
    else line_sync<SYNC, TPL>(line_in_cta);     // everyone has read its inputs before the scatter
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int j = t + TPL * q;
        cx<T> u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = v[q + r * Q];
        if (Ns > 1) {
            const int k = j & (Ns - 1);
#pragma unroll
            for (int r = 1; r < R; ++r) u[r] = cmul(u[r], __ldg(&tw[twoff + (r - 1) * Ns + k]));
        }
        dft_reg<T, R>(u);
        if (last) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[q + r * Q] = u[r];
        } else {
            const int base = (j / Ns) * (Ns * R) + (j & (Ns - 1));
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
template <typename T, int N, int E, int SYNC, typename Hook, int P = 0>
__device__ __forceinline__ void p2_fft(cx<T> (&v)[E], cx<T>* sline, const cx<T>* tw, int t, int l, Hook& hook) {
    p2_pass<T, N, E, SYNC, P, Hook>(v, sline, tw, t, l, hook);
    if constexpr (P + 1 < LinePlan<N, E>::npass()) p2_fft<T, N, E, SYNC, Hook, P + 1>(v, sline, tw, t, l, hook);
}

struct P2Mail {     // shared: the next item of this CTA, written by thread 0
    int item;       // -1: none
    int ready;      // its dependencies were complete when looked at
};

struct P2Sched {    // registers of thread 0: two items of look-ahead
    int nxt;                // the next item of this CTA (-1: none), drawn one item ago
    unsigned int drawn;     // the number drawn for the item after that (atomic in flight)
};

struct P2NoHook {
    __device__ __forceinline__ void operator()() const {}
};

// One work item of a statically known stage shape: the tile it.tile in every beam of group it.g.
// Returns with `next` = the next item of this CTA; when next.ready its first copies are in flight.
template <typename T, int N, int E, int W, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void p2_run_item(const P2Args<T>& p, const P2Item it, cx<T>* xbuf, cx<T>* fbuf, P2Mail* mail,
                                            P2Mail& next, P2Sched& sc, int tid) {
    using K = P2Cfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;      // along the line
    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CTA;
    constexpr bool REV = MID || INV;        // inverse = forward with mirrored output index
    const P2Stage<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;          // line within the tile
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;          // thread within the line
    cx<T>* sline = xbuf + l * K::LS;
    const cx<T>* fline = fbuf + l * K::FL;
    const long long b0 = (long long)it.g * p.G;
    const long long slot0 = (long long)(it.g % p.ring_slots) * p.G;
    const long long left = p.nbeams - b0;
    const int nb = left < p.G ? (int)left : p.G;
    const cx<T>* src0 = p2_src<T, N, W>(p, st, it.g, it.tile, 0);
    const int fixed = it.tile * W + l;                             // grid index constant along the line
    const long long off0 = (AXIS == AX_ROW) ? (long long)fixed * N : fixed;
    // output offset of register m: position t + TPL*m, or its mirror (N - pos) mod N (only pos 0 wraps)
    const long long o_first = (REV ? (long long)((N - t) & (N - 1)) : (long long)t) * PSTRIDE;
    const long long o_base = (REV ? (long long)(N - t) : (long long)t) * PSTRIDE;
    constexpr long long o_step = (REV ? -(long long)TPL : (long long)TPL) * PSTRIDE;
    const bool use_pre = !IN_RING && !MID && st.pre != P2_FAC_NONE;
    const int pol_in = p2_in_policy(p, st);
    const bool keep_ring = (p.hints & P2_HINT_KEEP_RING) != 0;

    // thread 0: the state of what the next item depends on, requested now and looked at in the hook of
    // the last beam (the latency of these reads is never exposed)
    unsigned int pa = 0, pb = 0;
    if (tid == 0) p2_poll<T, K::TILES>(p, sc.nxt, pa, pb, K::FULL);

    const cx<T>* land = (AXIS == AX_ROW) ? sline + t : xbuf + t * W + l;
    for (int bi = 0; bi < nb; ++bi) {
        const bool last_beam = (bi == nb - 1);
        p2_wait_all();
        line_sync<SYNC, TPL>(l);                               // the tile (and the factor rows) have landed
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = land[(AXIS == AX_ROW) ? TPL * m : TPL * m * W];
        const long long b = b0 + bi;

        if constexpr (!IN_RING) {
            if (st.mask_kind != MASK_NONE) {
                const long long mb = b * st.mask_beam_stride + off0;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], p2_mask_value<T>(st.mask, st.mask_kind, mb + (t + TPL * m) * PSTRIDE));
            }
            if (use_pre) {
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]);
            }
            if (st.out_pre != nullptr) {
                cx<T>* dst = st.out_pre + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
            }
        }

        // after the last gather of the (last) transform the exchange buffer takes the next tile
        auto hook = [&]() {
            if (!last_beam) {
                line_sync<SYNC, TPL>(l);
                p2_copy_tile<T, N, E, W, AXIS>(xbuf, src0 + (long long)(bi + 1) * PLANE, tid, pol_in);
                p2_commit();
            } else {
                if (tid == 0) {
                    bool ready = sc.nxt >= 0 && pa >= K::FULL && pb >= K::FULL;
                    if (sc.nxt >= 0 && !ready) ready = p2_ready<T, K::TILES>(p, sc.nxt, K::FULL);   // look again, now
                    mail->item = sc.nxt;
                    mail->ready = ready ? 1 : 0;
                }
                __syncthreads();                               // buffers free for everyone; mail visible
                next = *mail;
                if (next.item >= 0 && next.ready) p2_prefetch_item<T, N, E, W>(p, next.item, xbuf, fbuf, tid);
            }
        };
        if constexpr (MID) {
            // fused column stage of the fresh chain: FFT_y, H, inverse FFT_y
            P2NoHook none;
            p2_fft<T, N, E, SYNC, P2NoHook>(v, sline, p.tw, t, l, none);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]);
            if (st.out_mid != nullptr) {
                cx<T>* dst = st.out_mid + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
            }
        }
        p2_fft<T, N, E, SYNC, decltype(hook)>(v, sline, p.tw, t, l, hook);

        if (st.scale != (T)1) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cscale(v[m], st.scale);
        }
        if constexpr (OUT_RING) {
            // coherent at L2: the ring is read by other SMs
            cx<T>* dst = p.ring + (slot0 + bi) * PLANE + off0;
            if (keep_ring) {
                p2_st_ring<2>(dst + o_first, v[0]);
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) p2_st_ring<2>(dst + m * o_step, v[m]);
            } else {
                p2_st_ring<0>(dst + o_first, v[0]);
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) p2_st_ring<0>(dst + m * o_step, v[m]);
            }
        } else {
            if (st.out != nullptr) {
                cx<T>* dst = st.out + b * PLANE + off0;
                if constexpr (AXIS == AX_ROW) {
                    __stcs(dst + o_first, v[0]);
                    dst += o_base;
#pragma unroll
                    for (int m = 1; m < E; ++m) __stcs(dst + m * o_step, v[m]);
                } else {
                    // column tiles write W*sizeof(complex) bytes per grid row
                    if (p.hints & P2_HINT_STREAM_OUT) {
                        __stcs(dst + o_first, v[0]);
                        dst += o_base;
#pragma unroll
                        for (int m = 1; m < E; ++m) __stcs(dst + m * o_step, v[m]);
                    } else {
                        dst[o_first] = v[0];
                        dst += o_base;
#pragma unroll
                        for (int m = 1; m < E; ++m) dst[m * o_step] = v[m];
                    }
                }
            }
            if (st.inten != nullptr) {
                T* dst = st.inten + b * PLANE + off0;
                dst[o_first] = v[0].x * v[0].x + v[0].y * v[0].y;
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) dst[m * o_step] = v[m].x * v[m].x + v[m].y * v[m].y;
            }
        }
    }
    // completion: every warp signals once its own stores are issued (release at gpu scope, cumulative
    // over the warp through __syncwarp); no CTA barrier at the end of an item
    __syncwarp();
    if ((tid & 31) == 0) p2_signal(p.counters + 1 + S * p.ngroups + it.g);
    if (tid == 0) {
        // advance the look-ahead: the number drawn an item ago becomes the next item, draw another
        sc.nxt = p2_next_valid<T, K::TILES>(p, (int)sc.drawn);
        sc.drawn = atomicAdd(&p.counters[0], 1u);
    }
}

template <typename T, int N, int E, int W, int MINB>
__global__ void __launch_bounds__(P2Cfg<T, N, E, W>::THREADS, MINB) p2_kernel(const __grid_constant__ P2Args<T> p) {
    using K = P2Cfg<T, N, E, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ P2Mail s_mail;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* fbuf = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X);
    const int tid = threadIdx.x;

    P2Sched sc;
    sc.nxt = -1; sc.drawn = 0;
    if (tid == 0) {
        const int first = p2_next_valid<T, K::TILES>(p, (int)atomicAdd(&p.counters[0], 1u));
        s_mail.item = first;
        s_mail.ready = 1;
        if (first >= 0) {
            sc.nxt = p2_next_valid<T, K::TILES>(p, (int)atomicAdd(&p.counters[0], 1u));
            sc.drawn = atomicAdd(&p.counters[0], 1u);
            p2_wait_ready<T, K::TILES>(p, first, K::FULL);
        }
    }
    __syncthreads();
    int cur = s_mail.item;
    if (cur < 0) return;
    __syncthreads();                                           // s_mail is rewritten by thread 0 later
    p2_prefetch_item<T, N, E, W>(p, cur, xbuf, fbuf, tid);

    for (;;) {
        P2Item it;
        P2Mail next;
        p2_decode<T, K::TILES>(p, cur, it);
        // the stage shapes a step is made of (2 stages: rows user->ring, columns ring->user;
        // 3 stages: rows user->ring, fused columns ring->ring, inverse rows ring->user)
        if (it.s == 0) {
            if (p.st[0].inverse) p2_run_item<T, N, E, W, 0, AX_ROW, false, true, false, true>(p, it, xbuf, fbuf, &s_mail, next, sc, tid);
            else p2_run_item<T, N, E, W, 0, AX_ROW, false, true, false, false>(p, it, xbuf, fbuf, &s_mail, next, sc, tid);
        } else if (it.s == 1) {
            if (p.st[1].mid) p2_run_item<T, N, E, W, 1, AX_COL, true, true, true, false>(p, it, xbuf, fbuf, &s_mail, next, sc, tid);
            else if (p.st[1].inverse) p2_run_item<T, N, E, W, 1, AX_COL, true, false, false, true>(p, it, xbuf, fbuf, &s_mail, next, sc, tid);
            else p2_run_item<T, N, E, W, 1, AX_COL, true, false, false, false>(p, it, xbuf, fbuf, &s_mail, next, sc, tid);
        } else {
            p2_run_item<T, N, E, W, 2, AX_ROW, true, false, false, true>(p, it, xbuf, fbuf, &s_mail, next, sc, tid);
        }
        if (next.item < 0) break;
        if (!next.ready) {
            // the completion of the current item has been signalled: waiting here cannot deadlock
            if (tid == 0) p2_wait_ready<T, K::TILES>(p, next.item, K::FULL);
            __syncthreads();
            p2_prefetch_item<T, N, E, W>(p, next.item, xbuf, fbuf, tid);
        }
        cur = next.item;
    }
}

// --------------------------------------------------------------------------------------------
// Ticket form of the fused step ("T3"): one CTA per work item instead of a persistent loop, with the
// register-resident line kernel body of kernels.cuh (global loads straight into registers, E samples
// per thread, as many resident CTAs as the register file allows).  A CTA draws its item number when
// it starts (so the numbering follows the real start order), waits until what the item depends on is
// complete, transforms the tile in the G beams of the group, signals and exits.  Latencies (ticket,
// dependency poll, global loads) are hidden by the other resident CTAs, exactly as in the
// one-kernel-per-pass path, while the intermediate lives in the L2-resident ring and the pointwise
// factor comes from the L2-resident factor plane (no FP64 in this kernel).
template <typename T, int N, int E, int W>
struct T3Cfg {
    using P = LinePlan<N, E>;
    static constexpr int TPL = P::TPL;
    static constexpr int THREADS = TPL * W;
    static constexpr int TILES = N / W;
    static constexpr int SZ = (int)sizeof(cx<T>);
    static constexpr int S128 = 128 / SZ;
    static constexpr int WANT = (W >= S128) ? 1 : S128 / W;
    static constexpr int LS = P::LINE_SLOTS + ((WANT - P::LINE_SLOTS % S128) + S128) % S128;
    static constexpr size_t SMEM_X = (P::npass() > 1) ? (size_t)W * LS * SZ : 0;
    static constexpr int SYNC_ROW = (TPL <= 32) ? SYNC_WARP : ((TPL % 32 == 0 && W <= 15) ? SYNC_NAMED : SYNC_CTA);
    static constexpr int FL = N / 2 + 2;                   // folded factor row stride (N/2 + 1 used)
    static_assert(THREADS >= 32 && THREADS <= 1024, "block size");
};

template <typename T, int N, int E, int W, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void t3_item(const P2Args<T>& p, const P2Item it, cx<T>* xbuf, int tid) {
    using K = T3Cfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;
    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CTA;
    constexpr bool REV = MID || INV;
    const P2Stage<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;
    cx<T>* sline = xbuf + l * K::LS;
    const long long b0 = (long long)it.g * p.G;
    const long long slot0 = (long long)(it.g % p.ring_slots) * p.G;
    const long long left = p.nbeams - b0;
    const int nb = left < p.G ? (int)left : p.G;
    const int fixed = it.tile * W + l;
    const long long off0 = (AXIS == AX_ROW) ? (long long)fixed * N : fixed;
    const long long o_first = (REV ? (long long)((N - t) & (N - 1)) : (long long)t) * PSTRIDE;
    const long long o_base = (REV ? (long long)(N - t) : (long long)t) * PSTRIDE;
    constexpr long long o_step = (REV ? -(long long)TPL : (long long)TPL) * PSTRIDE;
    const bool use_pre = !IN_RING && !MID && st.pre != P2_FAC_NONE;
    const cx<T>* fline = p.fac + (long long)fixed * K::FL;      // L2-resident factor row of this line
    const bool keep_ring = (p.hints & P2_HINT_KEEP_RING) != 0;

    for (int bi = 0; bi < nb; ++bi) {
        const long long b = b0 + bi;
        const cx<T>* src = (IN_RING ? p.ring + (slot0 + bi) * PLANE : st.in + b * PLANE) + off0 + t * PSTRIDE;
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = IN_RING ? __ldcg(src + (TPL * m) * PSTRIDE) : __ldcs(src + (TPL * m) * PSTRIDE);
        if constexpr (!IN_RING) {
            if (st.mask_kind != MASK_NONE) {
                const long long mb = b * st.mask_beam_stride + off0;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], p2_mask_value<T>(st.mask, st.mask_kind, mb + (t + TPL * m) * PSTRIDE));
            }
            if (use_pre) {
                // eight factors at a time (32 loads in flight would double the register footprint)
#pragma unroll
                for (int m0 = 0; m0 < E; m0 += 8) {
                    cx<T> h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (m0 + i < E) h[i] = __ldcg(&fline[p2_fac_index<N, E>(p.fac_map, t + TPL * (m0 + i), m0 + i)]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (m0 + i < E) v[m0 + i] = cmul(v[m0 + i], h[i]);
                    if (E > 16) asm volatile("" ::: "memory");
                }
            }
            if (st.out_pre != nullptr) {
                cx<T>* dst = st.out_pre + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
            }
        }
        line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
        if constexpr (MID) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], __ldcg(&fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]));
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
            cx<T>* dst = p.ring + (slot0 + bi) * PLANE + off0;
            if (keep_ring) {
                p2_st_ring<2>(dst + o_first, v[0]);
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) p2_st_ring<2>(dst + m * o_step, v[m]);
            } else {
                p2_st_ring<0>(dst + o_first, v[0]);
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) p2_st_ring<0>(dst + m * o_step, v[m]);
            }
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
}

template <typename T, int N, int E, int W, int MINB>
__global__ void __launch_bounds__(T3Cfg<T, N, E, W>::THREADS, MINB) t3_kernel(const __grid_constant__ P2Args<T> p) {
    using K = T3Cfg<T, N, E, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_item;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    const int tid = threadIdx.x;
    // P2_HINT_BLOCK_TICKET: the item number is the block index (CTAs are dispatched in block-index
    // order), every warp waits for the dependencies and signals completion on its own: no ticket
    // round trip and no CTA barrier outside the transforms.  Otherwise thread 0 draws a ticket.
    const bool block_ticket = (p.hints & P2_HINT_BLOCK_TICKET) != 0;
    int item;
    if (block_ticket) {
        item = (int)blockIdx.x;
        P2Item it0;
        if (item >= p.total_items || !p2_decode<T, K::TILES>(p, item, it0)) return;
        if ((tid & 31) == 0) p2_wait_ready<T, K::TILES>(p, item, (unsigned)K::TILES * (K::THREADS / 32));
        __syncwarp();
    } else {
        if (tid == 0) {
            const int drawn = (int)atomicAdd(&p.counters[0], 1u);
            P2Item it0;
            if (drawn < p.total_items && p2_decode<T, K::TILES>(p, drawn, it0)) {
                p2_wait_ready<T, K::TILES>(p, drawn, (unsigned)K::TILES);
                s_item = drawn;
            } else {
                s_item = -1;
            }
        }
        __syncthreads();
        item = s_item;
        if (item < 0) return;
    }
    P2Item it;
    p2_decode<T, K::TILES>(p, item, it);
    if (it.s == 0) {
        if (p.st[0].inverse) t3_item<T, N, E, W, 0, AX_ROW, false, true, false, true>(p, it, xbuf, tid);
        else t3_item<T, N, E, W, 0, AX_ROW, false, true, false, false>(p, it, xbuf, tid);
    } else if (it.s == 1) {
        if (p.st[1].mid) t3_item<T, N, E, W, 1, AX_COL, true, true, true, false>(p, it, xbuf, tid);
        else if (p.st[1].inverse) t3_item<T, N, E, W, 1, AX_COL, true, false, false, true>(p, it, xbuf, tid);
        else t3_item<T, N, E, W, 1, AX_COL, true, false, false, false>(p, it, xbuf, tid);
    } else {
        t3_item<T, N, E, W, 2, AX_ROW, true, false, false, true>(p, it, xbuf, tid);
    }
    if (block_ticket) {
        __syncwarp();                                      // this warp's stores are issued
        if ((tid & 31) == 0) p2_signal(p.counters + 1 + it.s * p.ngroups + it.g);
    } else {
        __syncthreads();                                   // all stores of the item are issued
        if (tid == 0) p2_signal(p.counters + 1 + it.s * p.ngroups + it.g);
    }
}

// --------------------------------------------------------------------------------------------
// Ticket form with staged tiles ("T4"): one CTA per work item, an item = TG consecutive tiles of
// ONE beam.  The first tile is requested as soon as the item is known; every further tile (and its
// factor rows) is copied by cp.async into the exchange buffer while the previous one is being
// transformed (from the moment its last gather is done), so one ticket, one dependency poll and one
// exposed load latency are paid per TG tiles, while the ring still holds one beam per group.
template <typename T, int N, int E, int W, int TG, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void t4_item(const P2Args<T>& p, const P2Item it, cx<T>* xbuf, cx<T>* fbuf, int tid) {
    using K = P2Cfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;
    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CTA;
    constexpr bool REV = MID || INV;
    const P2Stage<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;
    cx<T>* sline = xbuf + l * K::LS;
    const cx<T>* fline = fbuf + l * K::FL;
    const long long b = it.g;                                   // G == 1: group = beam
    const long long slot = it.g % p.ring_slots;
    const cx<T>* in_plane = IN_RING ? p.ring + slot * PLANE : st.in + b * PLANE;
    const long long o_first = (REV ? (long long)((N - t) & (N - 1)) : (long long)t) * PSTRIDE;
    const long long o_base = (REV ? (long long)(N - t) : (long long)t) * PSTRIDE;
    constexpr long long o_step = (REV ? -(long long)TPL : (long long)TPL) * PSTRIDE;
    const bool use_pre = !IN_RING && !MID && st.pre != P2_FAC_NONE;
    const bool use_fac = use_pre || MID;
    const bool keep_ring = (p.hints & P2_HINT_KEEP_RING) != 0;
    const int tile0 = it.tile * TG;

    auto request = [&](int tile) {
        if (use_fac) p2_copy_factor<T, N, E, W, AXIS>(fbuf, p.fac, tile, tid);
        p2_copy_tile_k<T, N, E, W, AXIS, 0>(xbuf, in_plane + (AXIS == AX_ROW ? (long long)tile * W * N : (long long)tile * W), tid);
        p2_commit();
    };
    request(tile0);
    const cx<T>* land = (AXIS == AX_ROW) ? sline + t : xbuf + t * W + l;
    for (int k = 0; k < TG; ++k) {
        const int tile = tile0 + k;
        const int fixed = tile * W + l;
        const long long off0 = (AXIS == AX_ROW) ? (long long)fixed * N : fixed;
        p2_wait_all();
        line_sync<SYNC, TPL>(l);
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = land[(AXIS == AX_ROW) ? TPL * m : TPL * m * W];
        if constexpr (!IN_RING) {
            if (st.mask_kind != MASK_NONE) {
                const long long mb = b * st.mask_beam_stride + off0;
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], p2_mask_value<T>(st.mask, st.mask_kind, mb + (t + TPL * m) * PSTRIDE));
            }
            if (use_pre) {
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]);
            }
            if (st.out_pre != nullptr) {
                cx<T>* dst = st.out_pre + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
            }
        }
        auto hook = [&]() {
            if (k + 1 < TG) {
                line_sync<SYNC, TPL>(l);                        // every gather of this line / tile is done
                request(tile + 1);
            }
        };
        if constexpr (MID) {
            P2NoHook none;
            p2_fft<T, N, E, SYNC, P2NoHook>(v, sline, p.tw, t, l, none);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]);
            if (st.out_mid != nullptr) {
                cx<T>* dst = st.out_mid + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
            }
        }
        p2_fft<T, N, E, SYNC, decltype(hook)>(v, sline, p.tw, t, l, hook);
        if (st.scale != (T)1) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cscale(v[m], st.scale);
        }
        if constexpr (OUT_RING) {
            cx<T>* dst = p.ring + slot * PLANE + off0;
            if (keep_ring) {
                p2_st_ring<2>(dst + o_first, v[0]);
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) p2_st_ring<2>(dst + m * o_step, v[m]);
            } else {
                p2_st_ring<0>(dst + o_first, v[0]);
                dst += o_base;
#pragma unroll
                for (int m = 1; m < E; ++m) p2_st_ring<0>(dst + m * o_step, v[m]);
            }
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
}

template <typename T, int N, int E, int W, int TG, int MINB>
__global__ void __launch_bounds__(P2Cfg<T, N, E, W>::THREADS, MINB) t4_kernel(const __grid_constant__ P2Args<T> p) {
    using K = P2Cfg<T, N, E, W>;
    constexpr int ITEMS = K::TILES / TG;                    // items per (stage, beam)
    static_assert(K::TILES % TG == 0, "tile group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_item;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* fbuf = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X);
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int item = (int)atomicAdd(&p.counters[0], 1u);
        P2Item it;
        if (item < p.total_items && p2_decode<T, ITEMS>(p, item, it)) {
            p2_wait_ready<T, ITEMS>(p, item, (unsigned)ITEMS);
            s_item = item;
        } else {
            s_item = -1;
        }
    }
    __syncthreads();
    const int item = s_item;
    if (item < 0) return;
    P2Item it;
    p2_decode<T, ITEMS>(p, item, it);
    if (it.s == 0) {
        if (p.st[0].inverse) t4_item<T, N, E, W, TG, 0, AX_ROW, false, true, false, true>(p, it, xbuf, fbuf, tid);
        else t4_item<T, N, E, W, TG, 0, AX_ROW, false, true, false, false>(p, it, xbuf, fbuf, tid);
    } else if (it.s == 1) {
        if (p.st[1].mid) t4_item<T, N, E, W, TG, 1, AX_COL, true, true, true, false>(p, it, xbuf, fbuf, tid);
        else if (p.st[1].inverse) t4_item<T, N, E, W, TG, 1, AX_COL, true, false, false, true>(p, it, xbuf, fbuf, tid);
        else t4_item<T, N, E, W, TG, 1, AX_COL, true, false, false, false>(p, it, xbuf, fbuf, tid);
    } else {
        t4_item<T, N, E, W, TG, 2, AX_ROW, true, false, false, true>(p, it, xbuf, fbuf, tid);
    }
    __syncthreads();                                       // all stores of the item are issued
    if (tid == 0) p2_signal(p.counters + 1 + it.s * p.ngroups + it.g);
}

// --------------------------------------------------------------------------------------------
// "T5": the ticket form with the register budget of the one-kernel-per-pass line kernels (E = 16 at
// 1024 points, 64 registers, 2 x 512 threads per SM = 32 warps, scalar FP32 arithmetic).  Differences
// to t3_item that keep it inside 64 registers: the inverse transform swaps re/im (register renaming,
// free in scalar code) instead of mirroring the output index; the folded factor rows of the tile are
// copied to shared memory once per item instead of being loaded next to the data.
template <typename T, int N, int E, int W, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV>
__device__ __forceinline__ void t5_item(const P2Args<T>& p, const P2Item it, cx<T>* xbuf, cx<T>* fbuf, int tid) {
    using K = P2Cfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;
    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CTA;
    const P2Stage<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;
    cx<T>* sline = xbuf + l * K::LS;
    const cx<T>* fline = fbuf + l * K::FL;
    const long long b = it.g;                                   // G == 1: group = beam
    const long long slot = it.g % p.ring_slots;
    const int fixed = it.tile * W + l;
    const long long off = ((AXIS == AX_ROW) ? (long long)fixed * N : fixed) + t * PSTRIDE;
    const bool use_pre = !IN_RING && !MID && st.pre != P2_FAC_NONE;

    if (use_pre || MID) {
        // the W folded factor rows of this tile (contiguous in the plane) -> shared memory
        const float4* src = reinterpret_cast<const float4*>(p.fac + (long long)it.tile * W * K::FL);
        float4* dst = reinterpret_cast<float4*>(fbuf);
        for (int c = tid; c < W * K::FCH; c += K::THREADS) dst[c] = __ldcg(src + c);
        __syncthreads();
    }
    const cx<T>* src = (IN_RING ? p.ring + slot * PLANE : st.in + b * PLANE) + off;
    cx<T> v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = IN_RING ? __ldcg(src + (TPL * m) * PSTRIDE) : __ldcs(src + (TPL * m) * PSTRIDE);
    if constexpr (!IN_RING) {
        if (st.mask_kind != MASK_NONE) {
            const long long mb = b * st.mask_beam_stride + off;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], p2_mask_value<T>(st.mask, st.mask_kind, mb + (TPL * m) * PSTRIDE));
        }
        if (use_pre) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]);
        }
        if (st.out_pre != nullptr) {
            cx<T>* dst = st.out_pre + b * PLANE + off;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
        }
    }
    if constexpr (INV && !MID) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
    }
    line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
    if constexpr (MID) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(cmul(v[m], fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]));
        if (st.out_mid != nullptr) {
            cx<T>* dst = st.out_mid + b * PLANE + off;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, cswap(v[m]));
        }
        line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
    }
    if constexpr (INV || MID) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
    }
    if (st.scale != (T)1) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cscale(v[m], st.scale);
    }
    if constexpr (OUT_RING) {
        cx<T>* dst = p.ring + slot * PLANE + off;
#pragma unroll
        for (int m = 0; m < E; ++m) p2_st_ring<2>(dst + (TPL * m) * PSTRIDE, v[m]);
    } else {
        if (st.out != nullptr) {
            cx<T>* dst = st.out + b * PLANE + off;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
        }
        if (st.inten != nullptr) {
            T* dst = st.inten + b * PLANE + off;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m].x * v[m].x + v[m].y * v[m].y);
        }
    }
}

template <typename T, int N, int E, int W, int MINB>
__global__ void __launch_bounds__(P2Cfg<T, N, E, W>::THREADS, MINB) t5_kernel(const __grid_constant__ P2Args<T> p) {
    using K = P2Cfg<T, N, E, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_item;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* fbuf = reinterpret_cast<cx<T>*>(smem_raw + K::SMEM_X);
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int drawn = (int)atomicAdd(&p.counters[0], 1u);
        P2Item it0;
        if (drawn < p.total_items && p2_decode<T, K::TILES>(p, drawn, it0)) {
            p2_wait_ready<T, K::TILES>(p, drawn, (unsigned)K::TILES);
            s_item = drawn;
        } else {
            s_item = -1;
        }
    }
    __syncthreads();
    const int item = s_item;
    if (item < 0) return;
    P2Item it;
    p2_decode<T, K::TILES>(p, item, it);
    if (it.s == 0) {
        if (p.st[0].inverse) t5_item<T, N, E, W, 0, AX_ROW, false, true, false, true>(p, it, xbuf, fbuf, tid);
        else t5_item<T, N, E, W, 0, AX_ROW, false, true, false, false>(p, it, xbuf, fbuf, tid);
    } else if (it.s == 1) {
        if (p.st[1].mid) t5_item<T, N, E, W, 1, AX_COL, true, true, true, false>(p, it, xbuf, fbuf, tid);
        else if (p.st[1].inverse) t5_item<T, N, E, W, 1, AX_COL, true, false, false, true>(p, it, xbuf, fbuf, tid);
        else t5_item<T, N, E, W, 1, AX_COL, true, false, false, false>(p, it, xbuf, fbuf, tid);
    } else {
        t5_item<T, N, E, W, 2, AX_ROW, true, false, false, true>(p, it, xbuf, fbuf, tid);
    }
    __syncthreads();                                       // all stores of the item are issued
    if (tid == 0) p2_signal(p.counters + 1 + it.s * p.ngroups + it.g);
}

// The factor plane: fac[line][idx], idx = 0 .. N/2 (row stride FL).
//   kind P2_FAC_H, by_column = 0: line = row j,    fac = H(a[idx], b[j])      (a[i] == a[N - i] bit for bit)
//   kind P2_FAC_H, by_column = 1: line = column i, fac = H(a[i], b[idx])
//   kind P2_FAC_LENS:             line = row j,    fac = lens(x2[N/2 - idx], x2[j])   (x2 even around N/2)
// FP64 phase with the reference's operation order (kernels.cuh).
template <typename T>
__global__ void factor_plane_kernel(cx<T>* __restrict__ fac, const double* __restrict__ a_tab, const double* __restrict__ b_tab,
                                    const double* __restrict__ x2_tab, int n, int fl, int kind, int by_column, double z,
                                    double lens_k0, double lens_2f) {
    const int half = n / 2 + 1;
    const long long total = (long long)n * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int line = (int)(i / half);
        const int idx = (int)(i - (long long)line * half);
        cx<T> f;
        if (kind == P2_FAC_H) {
            f = by_column ? transfer_factor<T>(a_tab[line], b_tab[idx], z) : transfer_factor<T>(a_tab[idx], b_tab[line], z);
        } else {
            f = lens_factor<T>(x2_tab[n / 2 - idx], x2_tab[line], lens_k0, lens_2f);
        }
        fac[(long long)line * fl + idx] = f;
    }
}

}  // namespace lp2d
