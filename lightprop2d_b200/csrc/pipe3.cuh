// The persistent form of the fused step kernel ("Q4"): same work items, same L2-resident ring, same ticket order and
// dependency counters as t3_kernel (pipe2.cuh), but memory movement is decoupled from the transform warps.
//
// One CTA per SM, alive for the whole launch:
//   * consumer warps (W * N / E threads): run the register-resident line transform of fft_core.cuh on a tile that is
//     ALREADY in shared memory, and store the results straight from registers;
//   * one producer thread: draws the tickets, checks the dependencies (without stalling the consumers), and has the
//     tile of the NEXT item copied into a free landing buffer by the copy engines: row tiles (W consecutive grid rows
//     are W*N*sizeof contiguous bytes) with `cp.async.bulk` (UBLKCP), column tiles (N rows of W*sizeof bytes) with
//     `cp.async.bulk.tensor.2d` (UTMALDG) -- the TMA unit gathers the 64/128-byte pieces, no LSU wavefronts are spent;
//   * one signaller thread: publishes the completion counter of an item (`red.release.gpu`) once every consumer warp
//     has issued its stores, so that no transform warp ever sits in a release fence.
// Landing buffers cycle through full / empty mbarriers (expect_tx + complete_tx); a buffer is released as soon as its
// samples are in registers, i.e. the copy of item k+2 runs while item k is still being transformed and stored.  Hence
// every SM keeps 64-128 KB of loads in flight all the time, whatever the consumers are doing -- the t3 form had two
// CTAs per SM that each went load -> transform -> store in sequence (23 % of the warp slots active).
//
// Row items never need a CTA-wide exchange: one line belongs to one warp (1024 points, E = 32), its exchange region in
// shared memory is private.  Column items exchange across the consumer warps through a named barrier that the
// producer / signaller threads are not part of.
#pragma once
#include "pipe2.cuh"

namespace lp2d {

// Phase clocks of the transform warps (tools/q4_timing.py): compiled in only with -DLP2D_Q4_CLOCKS=1 (build with
// LP2D_BUILD_DEFS="-DLP2D_Q4_CLOCKS=1"); even a never-taken branch on p.timing costs the plain fft2 8 % (4.83 -> 5.3 us).
#ifndef LP2D_Q4_CLOCKS
#define LP2D_Q4_CLOCKS 0
#endif
enum { Q4_NDONE = 4 };       // (named barrier 15 belongs to the consumer threads, fft_core.cuh)

// HW: helper warps: 2 = producer warp + signaller warp; 1 = both in one warp (lanes 0 and 16), which leaves the
// registers of a warp to the transform warps when three CTAs share an SM
template <typename T, int N, int E, int W, int NBUF, int HW = 2>
struct Q4Cfg {
    using P = LinePlan<N, E>;
    static constexpr int TPL = P::TPL;
    static constexpr int NCT = TPL * W;                    // consumer threads
    static constexpr int NCW = NCT / 32;                   // consumer warps
    static constexpr int THREADS = NCT + 32 * HW;          // + producer warp + signaller warp
    static constexpr int SIGNALLER = (HW == 2) ? 32 : 16;  // helper-thread index of the signaller
    static constexpr int TILES = N / W;
    static constexpr int SZ = (int)sizeof(cx<T>);
    static constexpr int S128 = 128 / SZ;
    static constexpr int WANT = (W >= S128) ? 1 : S128 / W;
    static constexpr int LS = P::LINE_SLOTS + ((WANT - P::LINE_SLOTS % S128) + S128) % S128;
    static constexpr int TILE_BYTES = W * N * SZ;
    static constexpr size_t SMEM_X = (size_t)W * LS * SZ;
    static constexpr size_t SMEM_TW = ((size_t)P::tw_size() * SZ + 127) / 128 * 128;   // twiddle table, copied in once
    static constexpr size_t SMEM = (size_t)NBUF * TILE_BYTES + SMEM_X + SMEM_TW;
    static constexpr int SYNC_ROW = (TPL <= 32) ? SYNC_WARP : SYNC_NAMED;
    static constexpr int BOX = (N < 256) ? N : 256;        // rows per TMA box of a column tile
    static constexpr int FL = N / 2 + 2;
    static_assert(NCT % 32 == 0 && (W * SZ) % 16 == 0 && W <= 256, "tile shape");
    static_assert(TPL <= 32 || (TPL % 32 == 0 && W <= 14), "named row barriers");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ void q4_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p2_smem_u32(bar)) : "memory");
}
// wait of a helper thread: it has nothing to do meanwhile and must not take issue slots from the transform warps of its
// scheduler, so it sleeps between the tests
__device__ __forceinline__ void q4_mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    const uint32_t a = p2_smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(128);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
template <int NCT>
__device__ __forceinline__ void q4_consumer_barrier() {
    asm volatile("bar.sync 15, %0;" ::"n"(NCT) : "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier; the source is read once (evict first)
__device__ __forceinline__ void q4_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, bool stream) {
    if (stream) {
        const unsigned long long pol = p2_policy<1>();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                         p2_smem_u32(smem_dst)),
                     "l"(gsrc), "r"(bytes), "r"(p2_smem_u32(bar)), "l"(pol)
                     : "memory");
    } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(p2_smem_u32(smem_dst)),
                     "l"(gsrc), "r"(bytes), "r"(p2_smem_u32(bar))
                     : "memory");
    }
}

// Row item, tile in registers: the warp hands its part of the landing buffer back.  The exchange region of a line is
// private to its warp in row items, but the threads of a COLUMN item's line are spread over all warps: the first row
// item after a column item waits until every warp has left that item's last gather.
template <int NCT>
__device__ __forceinline__ void q4_row_release(uint64_t* empty_bar, int ctid, bool after_column_item) {
    if (after_column_item) q4_consumer_barrier<NCT>();
    else __syncwarp();
    if ((ctid & 31) == 0) q4_mbar_arrive(empty_bar);
}

// Dependency wait of the producer: relaxed polls (an acquire load invalidates the SM's L1 every time), then ONE
// acquire read of the same counters once they are complete.
template <typename T, int TILES>
__device__ __forceinline__ void q4_wait_ready(const P2Args<T>& p, int item, unsigned int full) {
    P2Item it;
    p2_decode<T, TILES>(p, item, it);
    const unsigned int* done = p.counters + 1;
    const P2Stage<T>& st = p.st[it.s];
    const unsigned int* ca = (it.s > 0) ? &done[(it.s - 1) * p.ngroups + it.g] : nullptr;
    const unsigned int* cb = (st.out_ring && !st.in_ring && it.g >= p.ring_slots)
                                 ? &done[(p.nstages - 1) * p.ngroups + (it.g - p.ring_slots)] : nullptr;
    const long long t0 = clock64();
    for (;;) {
        unsigned int va = full, vb = full;
        if (ca) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(va) : "l"(ca) : "memory");
        if (cb) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(vb) : "l"(cb) : "memory");
        if (va >= full && vb >= full) break;
        __nanosleep(100);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
    if (ca) (void)p2_ld_relaxed(ca);                       // (named "relaxed" for history: it is ld.acquire.gpu)
    if (cb) (void)p2_ld_relaxed(cb);
}

// One item on the consumer threads.  `land` holds the tile: rows as [line][N] (dense), columns as [row][W].
template <typename T, int N, int E, int W, int NBUF, int HOOKS, int S, int AXIS, bool IN_RING, bool OUT_RING, bool MID, bool INV, bool OUT_T = false>
__device__ __forceinline__ void q4_item(const P2Args<T>& p, const P2Item it, const cx<T>* __restrict__ land, cx<T>* xbuf,
                                        const cx<T>* __restrict__ stw, uint64_t* empty_bar, int ctid, bool after_column_item,
                                        long long* tmark) {
    using K = Q4Cfg<T, N, E, W, NBUF>;
    constexpr int TPL = K::TPL;
    constexpr long long PLANE = (long long)N * N;
    constexpr long long PSTRIDE = (AXIS == AX_ROW) ? 1 : N;
    constexpr int SYNC = (AXIS == AX_ROW) ? K::SYNC_ROW : SYNC_CONSUMERS + K::NCT;
    constexpr bool REV = MID || INV;
    const P2Stage<T>& st = p.st[S];
    static_assert(!OUT_T || (AXIS == AX_COL && IN_RING && !OUT_RING && !MID), "transposed output: final column stage");
    const int l = (AXIS == AX_ROW) ? ctid / TPL : ctid % W;
    const int t = (AXIS == AX_ROW) ? ctid % TPL : ctid / W;
    // transposed output (see t3_item): after the first exchange a thread works on line lo at positions to + TPL*m
    const int lo = OUT_T ? ctid / TPL : l;
    const int to = OUT_T ? ctid % TPL : t;
    constexpr long long OSTRIDE = (AXIS == AX_ROW || OUT_T) ? 1 : N;
    cx<T>* sline = xbuf + l * K::LS;
    const long long b = it.g;                              // one beam per group
    const long long slot = it.g % p.ring_slots;
    const int fixed = it.tile * W + l;
    const long long off0 = (AXIS == AX_ROW) ? (long long)fixed * N : fixed;
    const long long off0o = OUT_T ? (long long)(it.tile * W + lo) * N : off0;
    const long long o_first = (REV ? (long long)((N - to) & (N - 1)) : (long long)to) * OSTRIDE;
    const long long o_base = (REV ? (long long)(N - to) : (long long)to) * OSTRIDE;
    constexpr long long o_step = (REV ? -(long long)TPL : (long long)TPL) * OSTRIDE;
    const bool use_pre = !IN_RING && !MID && st.pre != P2_FAC_NONE;
    const cx<T>* fline = p.fac + (long long)fixed * p.fac_stride;      // L2-resident factor row of this line
    const bool keep_ring = (p.hints & P2_HINT_KEEP_RING) != 0;

    cx<T> v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = (AXIS == AX_ROW) ? land[l * N + t + TPL * m] : land[(t + TPL * m) * W + l];
    if (AXIS == AX_ROW) q4_row_release<K::NCT>(empty_bar, ctid, after_column_item);
    if (LP2D_Q4_CLOCKS && tmark) tmark[1] = clock64();     // tile in registers
    if constexpr (!IN_RING) {
        if (st.mask_kind != MASK_NONE) {
            const long long mb = b * st.mask_beam_stride + off0;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cmul(v[m], p2_mask_value<T>(st.mask, st.mask_kind, mb + (t + TPL * m) * PSTRIDE));
        }
        if (use_pre && !(HOOKS & 8)) {
            if constexpr (AXIS == AX_ROW) {
                p2_apply_factor_row<T, N, E, E / 2>(v, fline, p.fac_map, t);
            } else {
#pragma unroll
                for (int m = 0; m < E; ++m) v[m] = cmul(v[m], __ldcg(&fline[p2_fac_index<N, E>(p.fac_map, t + TPL * m, m)]));
            }
        }
    }
    if (LP2D_Q4_CLOCKS && tmark) tmark[2] = clock64();     // pointwise factor applied
    // The global stores ride on the butterflies (hooks of dft_reg): the state (spectrum * H, field * lens ...) goes out
    // while the first radix pass consumes it, the result while the last radix pass produces it.
    cx<T>* const pre_dst = (!IN_RING && st.out_pre != nullptr) ? st.out_pre + b * PLANE + off0 + t * PSTRIDE : nullptr;
    auto pre_hook = [&](int m, const cx<T>& val) {
        if ((HOOKS & 32) && !(p.hints & (1 << 20))) return;  // ablation: no state stores
        if (pre_dst != nullptr) __stcs(pre_dst + (TPL * m) * PSTRIDE, val);
    };
    cx<T>* const ring_dst = OUT_RING ? p.ring + slot * PLANE + off0 : nullptr;
    cx<T>* const out_dst = (!OUT_RING && st.out != nullptr) ? st.out + b * PLANE + off0o : nullptr;
    T* const inten_dst = (!OUT_RING && st.inten != nullptr) ? st.inten + b * PLANE + off0o : nullptr;
    const T scale = st.scale;
    auto post_hook = [&](int m, const cx<T>& val0) {
        if ((HOOKS & 16) && !(p.hints & (1 << 20))) return;  // ablation: transform but no result stores (timing only)
        const cx<T> val = cscale(val0, scale);              // (a scale of 1 is exact)
        const long long o = (m == 0) ? o_first : o_base + m * o_step;
        if constexpr (OUT_RING) {
            p2_st_ring<2>(ring_dst + o, val);              // L2 evict-last: the ring is re-read by the next stage
        } else {
            if (out_dst != nullptr) __stcs(out_dst + o, val);
            if (inten_dst != nullptr) __stcs(inten_dst + o, val.x * val.x + val.y * val.y);
        }
    };
    // HOOKS bit 0: state stores inside the first radix pass, bit 1: result stores inside the last one; otherwise bursts
    auto pre_burst = [&]() {
        if (!(HOOKS & 1)) {
#pragma unroll
            for (int m = 0; m < E; ++m) pre_hook(m, v[m]);
        }
    };
    auto post_burst = [&]() {
        if (!(HOOKS & 2)) {
#pragma unroll
            for (int m = 0; m < E; ++m) post_hook(m, v[m]);
        }
    };
    if constexpr (MID) {
        line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l);
        if ((ctid & 31) == 0) q4_mbar_arrive(empty_bar);   // every thread passed the first exchange barrier
#pragma unroll
        for (int m0 = 0; m0 < E; m0 += 8) {
            cx<T> h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) if (m0 + i < E) h[i] = __ldcg(&fline[p2_fac_index<N, E>(p.fac_map, t + TPL * (m0 + i), m0 + i)]);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (m0 + i < E) v[m0 + i] = cmul(v[m0 + i], h[i]);
        }
        if (st.out_mid != nullptr) {
            cx<T>* dst = st.out_mid + b * PLANE + off0 + t * PSTRIDE;
#pragma unroll
            for (int m = 0; m < E; ++m) __stcs(dst + (TPL * m) * PSTRIDE, v[m]);
        }
        if constexpr ((HOOKS & 2) != 0) line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l, NoHook(), post_hook);
        else line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l);
        post_burst();
    } else if constexpr (OUT_T) {
        cx<T>* slo = xbuf + lo * K::LS;
        if constexpr ((HOOKS & 2) != 0) line_fft_remap<T, N, E, SYNC, true>(v, sline, t, l, slo, to, lo, stw, post_hook);
        else line_fft_remap<T, N, E, SYNC, true>(v, sline, t, l, slo, to, lo, stw, NoHook());
        if ((ctid & 31) == 0) q4_mbar_arrive(empty_bar);
        post_burst();
    } else {
        if constexpr (!IN_RING) pre_burst();
        if constexpr (!IN_RING && (HOOKS & 1) != 0 && (HOOKS & 2) != 0) line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l, pre_hook, post_hook);
        else if constexpr (!IN_RING && (HOOKS & 1) != 0) line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l, pre_hook, NoHook());
        else if constexpr ((HOOKS & 2) != 0) line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l, NoHook(), post_hook);
        else if constexpr ((HOOKS & 4) == 0) line_fft<T, N, E, SYNC, 0, true>(v, sline, stw, t, l);
        else if (AXIS == AX_COL) line_sync<SYNC, TPL>(l);
        if (AXIS == AX_COL && (ctid & 31) == 0) q4_mbar_arrive(empty_bar);
        post_burst();
    }
}

// Warp-private form of the FINAL column stage (use_tma == 2): the tile lands with the TMA unit's 64-byte swizzle
// (16-byte chunk index ^= (row >> 1) & 3), which spreads the rows of one column over the banks well enough (two-way
// instead of eight-way conflicts) for ONE WARP to own a column: no exchange between warps, hence no CTA-wide barrier in
// the item, and the warps of the CTA are free to drift apart -- one can be in its load/store phase while another is in
// its floating-point phase.  The results go back to the same (swizzled) slots of the landing buffer and leave through a
// TMA tensor store issued by the signaller thread, which hands the buffer back when the store has read it.
template <typename T, int N, int E, int W, int NBUF, bool INV>
__device__ __forceinline__ void q4_col_private_item(const P2Args<T>& p, unsigned char* land, cx<T>* xbuf, const cx<T>* __restrict__ stw,
                                                    int ctid) {
    using K = Q4Cfg<T, N, E, W, NBUF>;
    constexpr int TPL = K::TPL;
    static_assert(TPL == 32 && W * K::SZ == 64 && K::NCW == W, "one warp per column, 64-byte tile rows");
    const int w = ctid >> 5, t = ctid & 31;
    const int cw = (w * K::SZ) ^ (((t >> 1) & 3) << 4);                    // swizzled byte offset of column w in rows t + 32 m
    cx<T> v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = *reinterpret_cast<const cx<T>*>(land + (t + TPL * m) * 64 + cw);
    line_fft<T, N, E, SYNC_WARP, 0, true>(v, xbuf + w * K::LS, stw, t, w);
    const T scale = p.st[1].scale;
    const int r0 = INV ? ((N - t) & (N - 1)) : t;                          // output row of v[0]; v[m]: r0 -/+ 32 m (mod N)
    const int co = (w * K::SZ) ^ (((r0 >> 1) & 3) << 4);                   // (32 m never changes bits 1..2 of the row)
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int row = INV ? ((N - t - TPL * m) & (N - 1)) : t + TPL * m;
        *reinterpret_cast<cx<T>*>(land + row * 64 + co) = cscale(v[m], scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy writes -> the TMA store
}

template <typename T, int N, int E, int W, int NBUF, int HOOKS, int MINB = 1>
__global__ void __launch_bounds__((Q4Cfg<T, N, E, W, NBUF, (MINB >= 3 ? 1 : 2)>::THREADS), MINB) q4_kernel(const __grid_constant__ P2Args<T> p) {
    using K = Q4Cfg<T, N, E, W, NBUF, (MINB >= 3 ? 1 : 2)>;
    constexpr int Q4_NBUF = NBUF;
    // (no pointer arithmetic through integers on this array: the compiler must keep seeing shared-memory addresses,
    // or every tile access becomes a generic LD / ST that queues with the global accesses)
    extern __shared__ __align__(1024) unsigned char q4_smem_raw[];
    __shared__ uint64_t s_full[Q4_NBUF], s_empty[Q4_NBUF], s_done[Q4_NDONE];
    __shared__ int s_desc[Q4_NBUF], s_done_item[Q4_NDONE];
    __shared__ volatile unsigned int s_signalled;
    unsigned char* smem = q4_smem_raw;
    cx<T>* xbuf = reinterpret_cast<cx<T>*>(smem + (size_t)Q4_NBUF * K::TILE_BYTES);
    cx<T>* stw = reinterpret_cast<cx<T>*>(smem + (size_t)Q4_NBUF * K::TILE_BYTES + K::SMEM_X);
    const int tid = threadIdx.x;
    // the twiddle table lives in shared memory: the producer's acquire loads invalidate L1 all the time
    for (int i = tid; i < K::P::tw_size(); i += K::THREADS) stw[i] = p.tw[i];

    if (tid == 0) {
        for (int i = 0; i < Q4_NBUF; ++i) {
            p2_mbar_init(&s_full[i], 1);
            p2_mbar_init(&s_empty[i], K::NCW);
        }
        for (int i = 0; i < Q4_NDONE; ++i) p2_mbar_init(&s_done[i], K::NCW);
        s_signalled = 0u;
    }
    // zero the counter set of the previous launch (it is next used by the launch after this one)
    for (unsigned int i = blockIdx.x * K::THREADS + tid; i < (unsigned int)p.zero_count; i += gridDim.x * K::THREADS)
        p.zero_counters[i] = 0u;
    __syncthreads();

    if (tid >= K::NCT) {
        const int h = tid - K::NCT;
        if (h == 0) {
            // ------------------------------------------------------------------ producer
            for (unsigned int k = 0;; ++k) {
                int item;
                P2Item it;
                for (;;) {                                  // fill / drain numbers of the schedule carry no work
                    item = (int)atomicAdd(&p.counters[0], 1u);
                    if (item >= p.total_items) { item = -1; break; }
                    if (p2_decode<T, K::TILES>(p, item, it)) break;
                }
                const int buf = (int)(k % Q4_NBUF);
                if (k >= Q4_NBUF) {
                    q4_mbar_wait_idle(&s_empty[buf], ((k / Q4_NBUF) - 1u) & 1u);
                    // "late draw": the next ticket is taken only when item k - NBUF has left the CTA altogether (all its
                    // stores issued), not as soon as its tile is in registers.  The CTA then holds two unfinished items
                    // instead of three, the lag between the stages of a beam can be one group shorter, and the ring
                    // (stages * lag planes) fits the part of L2 that keeps it resident.  Lead time of the copy: one item.
                    if (p.hints & P2_HINT_LATE_DRAW) {
                        const unsigned int j = k - Q4_NBUF;
                        q4_mbar_wait_idle(&s_done[j % Q4_NDONE], (j / Q4_NDONE) & 1u);
                    }
                }
                if (k >= Q4_NDONE) {                        // slot k % NDONE of the completion ring is free again
                    const long long t0 = clock64();
                    while (s_signalled + Q4_NDONE <= k) {
                        __nanosleep(128);
                        if (clock64() - t0 > 4000000000ll) __trap();
                    }
                }
                s_desc[buf] = item;
                s_done_item[k % Q4_NDONE] = item;
                if (item < 0) {
                    q4_mbar_arrive(&s_full[buf]);
                    break;
                }
                q4_wait_ready<T, K::TILES>(p, item, (unsigned)K::TILES);
                const P2Stage<T>& st = p.st[it.s];
                unsigned char* land = smem + (size_t)buf * K::TILE_BYTES;
                // what the copy engines read from the ring was written with ordinary stores by other SMs
                if (st.in_ring) asm volatile("fence.proxy.async;" ::: "memory");
                p2_mbar_expect_tx(&s_full[buf], (uint32_t)K::TILE_BYTES);
                if (st.axis == AX_ROW) {
                    const cx<T>* src = (st.in_ring ? p.ring + (long long)(it.g % p.ring_slots) * N * N : st.in + (long long)it.g * N * N) +
                                       (long long)it.tile * W * N;
                    // W consecutive rows are contiguous; 32 KB pieces keep several requests in flight
                    constexpr int PIECE = K::TILE_BYTES < 32768 ? K::TILE_BYTES : 32768;
#pragma unroll
                    for (int q = 0; q < K::TILE_BYTES / PIECE; ++q)
                        q4_bulk_load(land + q * PIECE, reinterpret_cast<const unsigned char*>(src) + q * PIECE, PIECE, &s_full[buf], !st.in_ring);
                } else {
                    constexpr int FPC = K::SZ / 4;
#pragma unroll
                    for (int q = 0; q < N / K::BOX; ++q)
                        p2_tma_load_2d(land + (size_t)q * K::BOX * W * K::SZ, &p.tm_ring, it.tile * W * FPC,
                                       (int)((it.g % p.ring_slots) * N) + q * K::BOX, &s_full[buf]);
                }
            }
        } else if (h == K::SIGNALLER) {
            // ----------------------------------------------------------------- signaller
            for (unsigned int k = 0;; ++k) {
                q4_mbar_wait_idle(&s_done[k % Q4_NDONE], (k / Q4_NDONE) & 1u);
                const int item = s_done_item[k % Q4_NDONE];
                if (item < 0) break;
                P2Item it;
                p2_decode<T, K::TILES>(p, item, it);
                if (p.use_tma == 2 && it.s == 1) {
                    // warp-private column item: its results are staged in the landing buffer; store the tile, and hand
                    // the buffer back to the producer once the store has read it
                    const int buf = (int)(k % Q4_NBUF);
                    unsigned char* land = smem + (size_t)buf * K::TILE_BYTES;
                    constexpr int FPC = K::SZ / 4;
#pragma unroll
                    for (int q = 0; q < N / K::BOX; ++q)
                        p2_tma_store_2d(&p.tm_out, it.tile * W * FPC, it.g * N + q * K::BOX, land + (size_t)q * K::BOX * W * K::SZ);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(p2_smem_u32(&s_empty[buf])), "r"(K::NCW) : "memory");
                }
                unsigned int* ctr = p.counters + 1 + it.s * p.ngroups + it.g;
                if (it.s == p.nstages - 1) p2_signal_consumed(ctr);
                else p2_signal(ctr);
                s_signalled = k + 1u;
            }
        }
        return;
    }

    // ------------------------------------------------------------------------ consumers
    bool after_col = false;
    for (unsigned int k = 0;; ++k) {
        const int buf = (int)(k % Q4_NBUF);
        long long tmarks[4];
        long long* tm = (LP2D_Q4_CLOCKS && p.timing != nullptr) ? tmarks : nullptr;
        if (LP2D_Q4_CLOCKS && tm) tm[0] = clock64();
        p2_mbar_wait(&s_full[buf], (k / Q4_NBUF) & 1u);
        if (LP2D_Q4_CLOCKS && tm) { tm[3] = clock64(); tm[1] = tm[2] = tm[3]; }
        const int item = s_desc[buf];
        if (item >= 0) {
            P2Item it;
            p2_decode<T, K::TILES>(p, item, it);
            const cx<T>* land = reinterpret_cast<const cx<T>*>(smem + (size_t)buf * K::TILE_BYTES);
            uint64_t* eb = &s_empty[buf];
            if (it.s == 0) {
                if (p.st[0].inverse) q4_item<T, N, E, W, NBUF, HOOKS, 0, AX_ROW, false, true, false, true>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
                else q4_item<T, N, E, W, NBUF, HOOKS, 0, AX_ROW, false, true, false, false>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
            } else if (it.s == 1) {
                if constexpr (K::TPL == 32 && W * K::SZ == 64 && K::NCW == W) {
                    if (p.use_tma == 2) {
                        unsigned char* lb = smem + (size_t)buf * K::TILE_BYTES;
                        if (p.st[1].inverse) q4_col_private_item<T, N, E, W, NBUF, true>(p, lb, xbuf, stw, tid);
                        else q4_col_private_item<T, N, E, W, NBUF, false>(p, lb, xbuf, stw, tid);
                        goto item_done;
                    }
                }
                if (p.st[1].mid) q4_item<T, N, E, W, NBUF, HOOKS, 1, AX_COL, true, true, true, false>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
                else if (p.st[1].out_t && p.st[1].inverse) q4_item<T, N, E, W, NBUF, HOOKS, 1, AX_COL, true, false, false, true, true>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
                else if (p.st[1].out_t) q4_item<T, N, E, W, NBUF, HOOKS, 1, AX_COL, true, false, false, false, true>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
                else if (p.st[1].inverse) q4_item<T, N, E, W, NBUF, HOOKS, 1, AX_COL, true, false, false, true>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
                else q4_item<T, N, E, W, NBUF, HOOKS, 1, AX_COL, true, false, false, false>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
            } else {
                q4_item<T, N, E, W, NBUF, HOOKS, 2, AX_ROW, true, false, false, true>(p, it, land, xbuf, stw, eb, tid, after_col, tm);
            }
        item_done:
            // (warp-private column mode: the exchange regions are private to a warp in row AND column items)
            after_col = (it.s == 1) && p.use_tma != 2;
        }
        __syncwarp();                                      // the warp's stores of the item are issued
        if (LP2D_Q4_CLOCKS && tm && item >= 0 && (tid & 31) == 0) {
            // bins: [kind][0] wait for the tile, [1] tile -> registers, [2] factor, [3] transform + stores, [4] items
            P2Item it2;
            p2_decode<T, K::TILES>(p, item, it2);
            unsigned long long* bin = p.timing + (p.st[it2.s].axis == AX_ROW ? 0 : 8);
            const long long tend = clock64();
            atomicAdd(&bin[0], (unsigned long long)(tm[3] - tm[0]));
            atomicAdd(&bin[1], (unsigned long long)(tm[1] - tm[3]));
            atomicAdd(&bin[2], (unsigned long long)(tm[2] - tm[1]));
            atomicAdd(&bin[3], (unsigned long long)(tend - tm[2]));
            atomicAdd(&bin[4], 1ull);
        }
        if ((tid & 31) == 0) q4_mbar_arrive(&s_done[k % Q4_NDONE]);
        if (item < 0) break;
    }
}

}  // namespace lp2d
