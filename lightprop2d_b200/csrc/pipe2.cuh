// The fused step kernel: a whole 2-D step -- the stateful propagate, lens + fft2, mask + fft2 (and, as
// options, plain fft2 and the fresh field -> field chain) -- in ONE launch whose row <-> column
// intermediate lives in an L2-resident ring of planes, so that HBM sees only the algorithmic bytes.
//
// Work item = (stage, beam group, tile of W lines): stage 0 = rows (user -> ring, with the pointwise
// factor and the state store), stage 1 = columns (ring -> user, optional |E|^2); three-stage chains add
// a fused column stage (FFT, H, inverse FFT; ring -> ring) in the middle.  One CTA per item
// ("ticket form"): a CTA draws its item number from one atomic counter when it STARTS, so the
// numbering follows the real start order; items are numbered so that stage s of a group comes `lag`
// groups after stage s-1, and an item waits on the per-(stage, group) completion counter of what it
// depends on.  Everything an item can wait for has a smaller number and was therefore drawn by a CTA
// that is already running: no deadlock, no co-residency requirement.  `lag` is sized so that the wait
// is (almost) never taken; waits are bounded (trap after ~2 s) so that a bug cannot hang the device.
//
// Body: the register-resident line transform of fft_core.cuh with E as large as the register file
// allows (1024-point complex64: E = 32, one warp per line, radix-32 x radix-32, ONE shared-memory
// exchange per transform, packed f32x2 complex arithmetic: ~30 issued instructions per sample per pass);
// inverse transforms are forward transforms with a mirrored output index.
//
// The pointwise factor (transfer function H(z) or lens phase) is NOT rebuilt per CTA in FP64:
// factor_plane_kernel writes it once per (plan, parameters) with the reference's FP64 operation order
// into an L2-resident plane folded by the kx -> -kx (resp. x -> -x) symmetry, reused by every beam of
// the batch and by every call of a constant-step z-sweep.
//
// Ring accesses are L2-coherent (ld/st.global.cg, stores with an evict-last hint), user planes are
// streamed (ld/st.global.cs).  What was tried and dropped (persistent CTAs with a look-ahead queue and
// cp.async staging, staged tile groups, an E = 16 body at 64 registers, block-index tickets, an L2
// persisting window): DESIGN.md section 5 and profiles/r1b_*, r1c_*; the code is in the git history.
#pragma once
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched from the driver at run time)
#include <cstdint>
#include "kernels.cuh"

namespace lp2d {

enum { P2_FAC_NONE = 0, P2_FAC_H = 1, P2_FAC_LENS = 2 };
enum { P2_MAX_STAGES = 3 };
// how a line position maps to the folded factor row: H is even in kx (index min(pos, N - pos)),
// the lens phase is even in x around the grid centre (index |pos - N/2|)
// (only when X^2 is bit-for-bit even around N/2; a generator may have shifted X in place, profiles.py:69-70,
// and then the lens plane is stored unfolded: index pos, P2_MAP_FULL)
enum { P2_MAP_FREQ = 0, P2_MAP_CENTRE = 1, P2_MAP_FULL = 2 };

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
    T* probe_out;        // final column stage, or nullptr: probe_out[beam] = |out[probe_index]|^2, the on-axis (or any one) sample
    long long probe_index;   // a z-sweep reads every step (tests/test_analytics.py:43-50); row * N + column of the natural field
    double* moments;     // final column stage, or nullptr: [beam][5] += sum I, sum x I, sum y I, sum x^2 I, sum y^2 I of the
    double dL;           // stage's output (x = (col - N/2) dL, y = (row - N/2) dL; beam2d.py:680-684, 696-702); zeroed by the host
    int out_t;           // final column stage: out / inten are written TRANSPOSED (line c of the tile = row c of the
                         // destination plane): contiguous 8 KB rows instead of 64-byte pieces 8 KB apart
};

template <typename T>
struct P2Args {
    P2Stage<T> st[P2_MAX_STAGES];
    cx<T>* ring;              // ring_slots * G planes
    const cx<T>* tw;          // twiddles of LinePlan<N, E>
    const cx<T>* fac;         // factor plane [N lines][fac_stride], folded or full (P2_MAP_*)
    int fac_stride;           // samples per factor row
    unsigned int* counters;   // [0] next item; [1 + s*ngroups + g] completed items of (s, g); all zero at launch
    unsigned int* zero_counters;   // the counter set the PREVIOUS launch of the plan used: zeroed by this launch (the plan
    int zero_count;                // alternates between two sets, so no CTA has to find out that it is the last one)
    long long nbeams;
    int nstages, ngroups, G, lag, ring_slots, total_items, fac_map;
    int hints;                // P2_HINT_*
    // TMA path of the final column stage of two-stage steps (use_tma != 0): 2-D tensor maps over the ring
    // and over the output planes, seen as [plane * N + row][N * sizeof(complex) / 4 floats]
    int use_tma;
    // tuning aid (persistent form): per item kind (row / column) x phase, summed clock64 ticks of lane 0 of every
    // transform warp + item counts; NULL in production (LP2D_Q4_TIMING=1 allocates it, lp2d_q4_timing reads it)
    unsigned long long* timing;
    alignas(64) CUtensorMap tm_ring;
    alignas(64) CUtensorMap tm_out;
};

// acquire: pairs with the producers' red.release, so that what is read from the ring after the test of the value
// (by any thread of the CTA, after the CTA barrier that follows) is ordered after the producers' stores
__device__ __forceinline__ unsigned int p2_ld_relaxed(const unsigned int* ctr) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    return v;
}
__device__ __forceinline__ void p2_signal(unsigned int* ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
// completion of a FINAL-stage item: the counter only tells the stage that refills the ring slot that the slot has
// been read (the loads completed before the item's stores were issued); nobody inside the launch reads what the item
// stored, so there is nothing to release and the CTA does not wait for its stores to drain
__device__ __forceinline__ void p2_signal_consumed(unsigned int* ctr) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
// L2 eviction priorities: user planes are streamed (evict first), the ring is kept (evict last)
enum { P2_HINT_KEEP_RING = 2,        // bit 0 was the (dropped) hint on the cp.async path
       P2_HINT_PREFETCH = 4,         // speculative L2 prefetch of the user rows of item blockIdx.x while the ticket is drawn
       P2_HINT_LATE_DRAW = 8,
       P2_HINT_RELAXED_POLL = 16 };  // experiment: relaxed instead of acquire loads on the dependency counters (no L1 invalidation)      // persistent form: draw the next ticket when an item is finished, not when its tile is read
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
    // acquire reads: everything that is read once the counters are complete (the ring) is requested only
    // after a CTA barrier that follows the test of these values, and was made visible before the counter
    // was incremented (red.release)
    const bool rlx = (p.hints & P2_HINT_RELAXED_POLL) != 0;
    if (it.s > 0) {
        const unsigned int* c = &done[(it.s - 1) * p.ngroups + it.g];
        if (rlx) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pa) : "l"(c) : "memory");
        else pa = p2_ld_relaxed(c);
    }
    if (st.out_ring && !st.in_ring && it.g >= p.ring_slots) {
        const unsigned int* c = &done[(p.nstages - 1) * p.ngroups + (it.g - p.ring_slots)];
        if (rlx) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pb) : "l"(c) : "memory");
        else pb = p2_ld_relaxed(c);
    }
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

// folded factor index of line position pos = t + TPL*m (m is a compile-time value after unrolling)
template <int N, int E>
__device__ __forceinline__ int p2_fac_index(int map, int pos, int m) {
    if (map == P2_MAP_FREQ) return (m < E / 2) ? pos : N - pos;
    if (map == P2_MAP_FULL) return pos;
    return (m < E / 2) ? N / 2 - pos : pos - N / 2;
}

__device__ __forceinline__ float2 p2_shfl(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ double2 p2_shfl(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// v[m] *= factor(line position t + TPL*m) from the folded factor row of the line, for a ROW item.
// One warp per line (TPL = 32): the fold makes the factor of position pos equal to the one of N - pos (resp. of the
// mirror image about N/2), so a lane needs E/2 values of its own (index t + 32 j) and E/2 values that lane 32 - t holds
// (index 32 (j+1) - t): E/2 loads in ONE batch and E/2 shuffles instead of E loads in four dependent batches (the four
// L2 round trips were a sixth of a CTA's life in the ticket form).  Lane 0 is its own mirror, shifted by one (and needs
// the single unpaired entry N/2).  Unfolded plane / other thread shapes: plain loads, eight at a time.
template <typename T, int N, int E, int GB>
__device__ __forceinline__ void p2_apply_factor_row(cx<T> (&v)[E], const cx<T>* __restrict__ fline, int map, int t) {
    constexpr int TPL = N / E, H = E / 2;
    if (TPL != 32 || sizeof(T) != 4 || map == P2_MAP_FULL) {      // (complex128: four shuffles per value, measured slower)
#pragma unroll
        for (int m0 = 0; m0 < E; m0 += 8) {
            cx<T> h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) if (m0 + i < E) h[i] = __ldcg(&fline[p2_fac_index<N, E>(map, t + TPL * (m0 + i), m0 + i)]);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (m0 + i < E) v[m0 + i] = cmul(v[m0 + i], h[i]);
            if (E > 16) asm volatile("" ::: "memory");
        }
        return;
    }
    cx<T> gx = mk<T>((T)0, (T)0);
    if (t == 0) gx = __ldcg(fline + N / 2);
    const int src = (TPL - t) & (TPL - 1);
    // GB values per batch: all E/2 at once where the registers allow it (persistent form), two batches otherwise
#pragma unroll
    for (int j0 = 0; j0 < H; j0 += GB) {
        cx<T> g[GB + 1];
#pragma unroll
        for (int i = 0; i <= GB; ++i)                                   // (lane 0 needs g[j + 1]: one value of overlap)
            if (i < GB || j0 + GB < H) g[i] = (i < GB || t == 0) ? __ldcg(fline + t + TPL * (j0 + i)) : mk<T>((T)0, (T)0);
#pragma unroll
        for (int i = 0; i < GB; ++i) {
            const int j = j0 + i;
            cx<T> mv = p2_shfl(g[i], src);
            if (t == 0) mv = (j + 1 < H) ? g[i + 1] : gx;
            if (map == P2_MAP_FREQ) {
                v[j] = cmul(v[j], g[i]);                                // position t + 32 j
                v[E - 1 - j] = cmul(v[E - 1 - j], mv);                  // position t + 32 (E-1-j): index N - pos = 32 (j+1) - t
            } else {
                v[H + j] = cmul(v[H + j], g[i]);                        // position N/2 + t + 32 j: index pos - N/2
                v[H - 1 - j] = cmul(v[H - 1 - j], mv);                  // position t + 32 (H-1-j): index N/2 - pos = 32 (j+1) - t
            }
        }
        if (GB < H) asm volatile("" ::: "memory");
    }
}

template <typename T>
__device__ __forceinline__ cx<T> p2_mask_value(const void* mask, int kind, long long idx) {
    if (kind == MASK_CX) return static_cast<const cx<T>*>(mask)[idx];
    if (kind == MASK_REAL) return mk<T>(static_cast<const T*>(mask)[idx], (T)0);
    return mk<T>(static_cast<const unsigned char*>(mask)[idx] ? (T)1 : (T)0, (T)0);
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

template <typename T, int N, int E, int W, int HOOKS, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV, bool OUT_T = false, bool MOM = false>
__device__ __forceinline__ void t3_item(const P2Args<T>& p, const P2Item it, cx<T>* xbuf, int tid, double (*s_mom)[5] = nullptr) {
    using K = T3Cfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;
    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CTA;
    constexpr bool REV = MID || INV;
    static_assert(!OUT_T || (AXIS == AX_COL && IN_RING && !OUT_RING && !MID), "transposed output: final column stage");
    const P2Stage<T>& st = p.st[S];
    const int l = (AXIS == AX_ROW) ? tid / TPL : tid % W;
    const int t = (AXIS == AX_ROW) ? tid % TPL : tid / W;
    // transposed output: after the first exchange a thread works on line lo at positions to + TPL*m (a warp = a line)
    const int lo = OUT_T ? tid / TPL : l;
    const int to = OUT_T ? tid % TPL : t;
    constexpr int OSTRIDE = (AXIS == AX_ROW || OUT_T) ? 1 : N;
    cx<T>* sline = xbuf + l * K::LS;
    const long long b0 = (long long)it.g * p.G;
    const long long slot0 = (long long)(it.g % p.ring_slots) * p.G;
    const long long left = p.nbeams - b0;
    const int nb = left < p.G ? (int)left : p.G;
    const int fixed = it.tile * W + l;
    const long long off0 = (AXIS == AX_ROW) ? (long long)fixed * N : fixed;
    // (offsets inside a plane fit 32 bits: fewer live registers across the transform than 64-bit ones)
    const long long off0o = OUT_T ? (long long)(it.tile * W + lo) * N : off0;
    const int o_first = (REV ? ((N - to) & (N - 1)) : to) * OSTRIDE;
    const int o_base = (REV ? (N - to) : to) * OSTRIDE;
    constexpr int o_step = (REV ? -TPL : TPL) * OSTRIDE;
    const bool use_pre = !IN_RING && !MID && st.pre != P2_FAC_NONE;
    const cx<T>* fline = p.fac + (long long)fixed * p.fac_stride;   // L2-resident factor row of this line

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
            if (use_pre && !(HOOKS & 8)) {
                if constexpr (AXIS == AX_ROW) {
                    p2_apply_factor_row<T, N, E, (E / 2 < 8 ? E / 2 : 8)>(v, fline, p.fac_map, t);
                } else {
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
            }
        }
        // The global stores ride on the butterflies (hooks of dft_reg): the state goes out while the first radix pass
        // consumes it, the result while the last radix pass produces it (a burst of 32 stores per thread fills the
        // load/store queue: lg_throttle was the second stall reason in profiles/r1d_ncu_full_t3_default_summary.txt).
        cx<T>* const pre_dst = (!IN_RING && st.out_pre != nullptr) ? st.out_pre + b * PLANE + off0 + t * PSTRIDE : nullptr;
        auto pre_hook = [&](int m, const cx<T>& val) {
            if (pre_dst != nullptr) __stcs(pre_dst + (TPL * m) * PSTRIDE, val);
        };
        // (destinations are formed where they are used: with the bursts that is after the transform, and nothing but
        // the plane offset has to stay in registers across it)
        auto post_hook = [&](int m, const cx<T>& val0) {
            const cx<T> val = cscale(val0, st.scale);       // (a scale of 1 is exact)
            const int o = (m == 0) ? o_first : o_base + m * o_step;
            if constexpr (OUT_RING) {
                p2_st_ring<2>(p.ring + (slot0 + bi) * PLANE + off0 + o, val);     // L2 evict-last: re-read by the next stage
            } else {
                if (st.out != nullptr) __stcs(st.out + b * PLANE + off0o + o, val);
                if (st.inten != nullptr) __stcs(st.inten + b * PLANE + off0o + o, val.x * val.x + val.y * val.y);
            }
        };
        // HOOKS bit 0: state stores inside the first radix pass, bit 1: result stores inside the last one; otherwise
        // bursts (measured per size, profiles/r2g_sweep*.log: hooks win at 1024 points, bursts at 512)
        if constexpr (MID) {
            line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], __ldcg(&fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]));
            if (st.out_mid != nullptr) {
                cx<T>* dst = st.out_mid + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
                for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
            }
            if constexpr ((HOOKS & 2) != 0) line_fft<T, N, E, SYNC, 0, false>(v, sline, p.tw, t, l, NoHook(), post_hook);
            else line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
        } else if constexpr (OUT_T) {
            cx<T>* slo = xbuf + lo * K::LS;
            if constexpr ((HOOKS & 2) != 0) line_fft_remap<T, N, E, SYNC, false>(v, sline, t, l, slo, to, lo, p.tw, post_hook);
            else line_fft_remap<T, N, E, SYNC, false>(v, sline, t, l, slo, to, lo, p.tw, NoHook());
        } else {
            if constexpr (!IN_RING && (HOOKS & 1) == 0) {
#pragma unroll
                for (int m = 0; m < E; ++m) pre_hook(m, v[m]);
            }
            if constexpr (!IN_RING && (HOOKS & 1) != 0 && (HOOKS & 2) != 0) line_fft<T, N, E, SYNC, 0, false>(v, sline, p.tw, t, l, pre_hook, post_hook);
            else if constexpr (!IN_RING && (HOOKS & 1) != 0) line_fft<T, N, E, SYNC, 0, false>(v, sline, p.tw, t, l, pre_hook, NoHook());
            else if constexpr ((HOOKS & 2) != 0) line_fft<T, N, E, SYNC, 0, false>(v, sline, p.tw, t, l, NoHook(), post_hook);
            else if constexpr ((HOOKS & 4) == 0) line_fft<T, N, E, SYNC>(v, sline, p.tw, t, l);
        }
        if constexpr ((HOOKS & 2) == 0) {
#pragma unroll
            for (int m = 0; m < E; ++m) post_hook(m, v[m]);
        }
        // Moment sums of the intensity, fused into the final stage: a z-sweep reads
        // centroid / D4sigma of every step without another pass over the field.  v[] holds the (unscaled) results.
        if constexpr (MOM && AXIS == AX_COL && !OUT_RING && !MID) {
            // the line's coordinate and the positions of the thread's results in the natural field
            const int lf = it.tile * W + lo;
            if (st.probe_out != nullptr) {
                const int prow = (int)(st.probe_index / N), pcol = (int)(st.probe_index % N);
                const int want_line = OUT_T ? prow : pcol, want_pos = OUT_T ? pcol : prow;
                if (lf == want_line) {
#pragma unroll
                    for (int m = 0; m < E; ++m) {
                        const int pos = REV ? ((N - to - TPL * m) & (N - 1)) : to + TPL * m;
                        if (pos == want_pos) {
                            const cx<T> val = cscale(v[m], st.scale);
                            st.probe_out[b] = val.x * val.x + val.y * val.y;
                        }
                    }
                }
            }
            if (st.moments != nullptr && s_mom != nullptr) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;                        // sums over the thread's positions along the line
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const cx<T> val = cscale(v[m], st.scale);
                    const double I = (double)(T)(val.x * val.x + val.y * val.y);
                    const int pos = REV ? ((N - to - TPL * m) & (N - 1)) : to + TPL * m;
                    const double u = (double)(pos - N / 2) * st.dL;
                    s0 += I; s1 += u * I; s2 += u * u * I;
                }
                const double w = (double)(lf - N / 2) * st.dL;               // coordinate of the line itself
                // natural store: the line is a grid column (x = w), positions are rows (y); transposed: the other way
                double a[5] = {s0, OUT_T ? s1 : w * s0, OUT_T ? w * s0 : s1, OUT_T ? s2 : w * w * s0, OUT_T ? w * w * s0 : s2};
#pragma unroll
                for (int k = 0; k < 5; ++k) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
                }
                __syncthreads();                                             // (scratch of the previous beam is consumed)
                if ((tid & 31) == 0) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) s_mom[tid >> 5][k] = a[k];
                }
                __syncthreads();
                if (tid < 5) {
                    double acc = 0.0;
                    for (int q = 0; q < K::THREADS / 32; ++q) acc += s_mom[q][tid];
                    atomicAdd(st.moments + b * 5 + tid, acc);
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// TMA form of the final column stage (ring -> user planes) of two-stage steps.  A column tile is N
// rows of W * sizeof(complex) bytes: per-thread loads / stores touch one 128-byte line per 64 bytes
// (twice the L1 wavefronts of a row access); here the tile is fetched and written back by the TMA unit
// as boxes of 256 rows (cp.async.bulk.tensor.2d), the threads only see shared memory.
__device__ __forceinline__ uint32_t p2_smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void p2_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p2_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void p2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p2_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = p2_smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();      // never hang the device on a lost copy
    }
}
__device__ __forceinline__ void p2_tma_load_2d(void* smem_dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     p2_smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(p2_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void p2_tma_store_2d(const CUtensorMap* tm, int x, int y, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(tm)),
                 "r"(x), "r"(y), "r"(p2_smem_u32(smem_src))
                 : "memory");
}

template <typename T, int N, int E, int W, bool INV>
__device__ __forceinline__ void t3_col_tma_item(const P2Args<T>& p, const P2Item it, cx<T>* xbuf, uint64_t* bar, uint32_t& parity,
                                                int tid) {
    using K = T3Cfg<T, N, E, W>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr int BOX = (N < 256) ? N : 256;                   // rows per TMA box
    constexpr int FPC = (int)sizeof(cx<T>) / 4;                // floats per complex sample (tensor maps are in floats)
    const P2Stage<T>& st = p.st[1];
    const int l = tid % W, t = tid / W;
    cx<T>* sline = xbuf + l * K::LS;
    const long long b0 = (long long)it.g * p.G;
    const long long slot0 = (long long)(it.g % p.ring_slots) * p.G;
    const long long left = p.nbeams - b0;
    const int nb = left < p.G ? (int)left : p.G;
    const int col0 = it.tile * W;

    for (int bi = 0; bi < nb; ++bi) {
        const long long b = b0 + bi;
        if (tid == 0) {
            p2_mbar_expect_tx(bar, (uint32_t)(N * W * sizeof(cx<T>)));
#pragma unroll
            for (int k = 0; k < N / BOX; ++k)
                p2_tma_load_2d(xbuf + (long long)k * BOX * W, &p.tm_ring, col0 * FPC, (int)((slot0 + bi) * N) + k * BOX, bar);
        }
        p2_mbar_wait(bar, parity);
        parity ^= 1u;
        cx<T> v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = xbuf[(t + TPL * m) * W + l];
        line_fft<T, N, E, SYNC_CTA>(v, sline, p.tw, t, l);     // the first pass synchronises before it scatters
        if (st.scale != (T)1) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cscale(v[m], st.scale);
        }
        __syncthreads();                                       // every thread has done its last gather
        // results -> shared memory in the plane's orientation [row][W]; inverse = mirrored row index
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int pos = t + TPL * m;
            xbuf[(INV ? ((N - pos) & (N - 1)) : pos) * W + l] = v[m];
        }
        if (st.inten != nullptr) {
            T* dst = st.inten + b * PLANE + col0 + l;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int pos = t + TPL * m;
                __stcs(dst + (long long)(INV ? ((N - pos) & (N - 1)) : pos) * N, v[m].x * v[m].x + v[m].y * v[m].y);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < N / BOX; ++k)
                p2_tma_store_2d(&p.tm_out, col0 * FPC, (int)(b * N) + k * BOX, xbuf + (long long)k * BOX * W);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile has been read out of shared memory
        }
        __syncthreads();
    }
    // the completion signal of this (final) stage only says that the ring slot has been consumed; the
    // bulk stores complete before the grid does
}

template <typename T, int N, int E, int W, int MINB, int HOOKS>
__global__ void __launch_bounds__(T3Cfg<T, N, E, W>::THREADS, MINB) t3_kernel(const __grid_constant__ P2Args<T> p) {
    using K = T3Cfg<T, N, E, W>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_item;
    __shared__ uint64_t s_bar;                             // TMA loads of the column stage
    __shared__ double s_mom[K::THREADS / 32][5];           // CTA-level reduction of the fused moment sums
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem_raw);
    const int tid = threadIdx.x;
    uint32_t tma_parity = 0;
    unsigned int drawn = 0u;
    if (tid == 0) {                                        // first thing: the ticket (its latency covers what follows)
        if (p.use_tma) p2_mbar_init(&s_bar, 1);
        drawn = atomicAdd(&p.counters[0], 1u);
    }
    // Tickets are drawn (almost) in block order: block k will most likely get item k, and if not, a CTA that starts at
    // the same time will.  While the ticket and the dependency check are in flight (two L2 round trips), ask L2 for
    // the user rows of a first-stage item, so that the loads of whoever works on it find them there.
    if ((p.hints & P2_HINT_PREFETCH) && tid < W * p.G && (int)blockIdx.x < p.total_items) {
        P2Item gi;
        if (p2_decode<T, K::TILES>(p, (int)blockIdx.x, gi) && gi.s == 0 && !p.st[0].in_ring) {
            const long long b = (long long)gi.g * p.G + tid / W;
            const cx<T>* row = p.st[0].in + b * ((long long)N * N) + (long long)(gi.tile * W + tid % W) * N;
            if (b < p.nbeams && (reinterpret_cast<uintptr_t>(row) & 15) == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"((unsigned)(N * sizeof(cx<T>))) : "memory");
        }
    }
    // zero the counter set of the previous launch (fire and forget; it is next used by the launch after this one)
    for (unsigned int i = blockIdx.x * K::THREADS + tid; i < (unsigned int)p.zero_count; i += gridDim.x * K::THREADS)
        p.zero_counters[i] = 0u;
    // While the ticket atomic of lane 0 is in flight, lane 1 checks the dependencies of the item
    // the CTA will most likely get (item blockIdx.x: tickets are drawn almost in block order, and the dependencies are
    // per (stage, group), which 128 neighbouring items share).  If the ticket turns out to belong to the same (stage,
    // group) and the check found it complete, the second L2 round trip is skipped; otherwise the normal wait runs.
    if (tid < 32) {
        unsigned int spec = 0u;
        int gs = -1, gg = -1;
        if (tid == 1 && (int)blockIdx.x < p.total_items) {
            P2Item gi;
            if (p2_decode<T, K::TILES>(p, (int)blockIdx.x, gi)) {
                gs = gi.s; gg = gi.g;
                spec = p2_ready<T, K::TILES>(p, (int)blockIdx.x, (unsigned)K::TILES) ? 1u : 0u;
            }
        }
        drawn = __shfl_sync(0xffffffffu, drawn, 0);
        spec = __shfl_sync(0xffffffffu, spec, 1);
        gs = __shfl_sync(0xffffffffu, gs, 1);
        gg = __shfl_sync(0xffffffffu, gg, 1);
        if (tid == 0) {
            P2Item it0;
            if ((int)drawn < p.total_items && p2_decode<T, K::TILES>(p, (int)drawn, it0)) {
                if (!(spec != 0u && it0.s == gs && it0.g == gg)) p2_wait_ready<T, K::TILES>(p, (int)drawn, (unsigned)K::TILES);
                s_item = (int)drawn;
            } else {
                s_item = -1;                               // fill / drain item of the schedule: nothing to do
            }
        }
    }
    __syncthreads();
    const int item = s_item;
    if (item >= 0) {
        P2Item it;
        p2_decode<T, K::TILES>(p, item, it);
        if (it.s == 0) {
            if (p.st[0].inverse) t3_item<T, N, E, W, HOOKS, 0, AX_ROW, false, true, false, true>(p, it, xbuf, tid);
            else t3_item<T, N, E, W, HOOKS, 0, AX_ROW, false, true, false, false>(p, it, xbuf, tid);
        } else if (it.s == 1) {
            if (p.st[1].mid) t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, true, true, false>(p, it, xbuf, tid);
            else if ((p.st[1].moments != nullptr || p.st[1].probe_out != nullptr) && p.st[1].out_t) t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, false, false, true, true, true>(p, it, xbuf, tid, s_mom);
            else if (p.st[1].moments != nullptr || p.st[1].probe_out != nullptr) t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, false, false, true, false, true>(p, it, xbuf, tid, s_mom);
            else if (p.st[1].out_t && p.st[1].inverse) t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, false, false, true, true>(p, it, xbuf, tid, s_mom);
            else if (p.st[1].out_t) t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, false, false, false, true>(p, it, xbuf, tid, s_mom);
            else if (p.use_tma && p.st[1].inverse) t3_col_tma_item<T, N, E, W, true>(p, it, xbuf, &s_bar, tma_parity, tid);
            else if (p.use_tma) t3_col_tma_item<T, N, E, W, false>(p, it, xbuf, &s_bar, tma_parity, tid);
            else if (p.st[1].inverse) t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, false, false, true>(p, it, xbuf, tid, s_mom);
            else t3_item<T, N, E, W, HOOKS, 1, AX_COL, true, false, false, false>(p, it, xbuf, tid, s_mom);
        } else {
            t3_item<T, N, E, W, HOOKS, 2, AX_ROW, true, false, false, true>(p, it, xbuf, tid);
        }
        __syncthreads();                                   // all stores of the item are issued
        if (tid == 0) {
            if (it.s == p.nstages - 1) p2_signal_consumed(p.counters + 1 + it.s * p.ngroups + it.g);
            else p2_signal(p.counters + 1 + it.s * p.ngroups + it.g);
        }
    }
    // No exit protocol: the counters this launch dirties are zeroed by the NEXT launch of the plan (see above), so a
    // CTA leaves as soon as its signal is issued (an "am I the last CTA" atomic kept all eight warps at a barrier
    // for a fence + an L2 round trip: 18 % of the warp time in profiles/r1c_ncu_full_t3_default_summary.txt).
}

// The factor plane: fac[line][idx], idx = 0 .. N/2 (row stride FL).
//   kind P2_FAC_H, by_column = 0: line = row j,    fac = H(a[idx], b[j])      (a[i] == a[N - i] bit for bit)
//   kind P2_FAC_H, by_column = 1: line = column i, fac = H(a[i], b[idx])
//   kind P2_FAC_LENS:             line = row j,    fac = lens(x2[N/2 - idx], x2[j])   (x2 even around N/2)
//   kind P2_FAC_LENS, full = 1:   line = row j,    fac = lens(x2[idx], x2[j]), idx = 0 .. N-1 (row stride fl >= N): the
//                                 grid was shifted in place by a generator and X^2 is not even around N/2
// FP64 phase with the reference's operation order (kernels.cuh).
template <typename T>
__global__ void factor_plane_kernel(cx<T>* __restrict__ fac, const double* __restrict__ a_tab, const double* __restrict__ b_tab,
                                    const double* __restrict__ x2_tab, int n, int fl, int kind, int by_column, int full, double z,
                                    double lens_k0, double lens_2f) {
    const int half = full ? n : n / 2 + 1;
    const long long total = (long long)n * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int line = (int)(i / half);
        const int idx = (int)(i - (long long)line * half);
        cx<T> f;
        if (kind == P2_FAC_H) {
            f = by_column ? transfer_factor<T>(a_tab[line], b_tab[idx], z) : transfer_factor<T>(a_tab[idx], b_tab[line], z);
        } else {
            f = lens_factor<T>(x2_tab[full ? idx : n / 2 - idx], x2_tab[line], lens_k0, lens_2f);
        }
        fac[(long long)line * fl + idx] = f;
    }
}

}  // namespace lp2d
