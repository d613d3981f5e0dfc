// Mode projection on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
//   coeffs[b][m] = sum_p modes[m][p] * field_b[crop(p)]        (no conjugate, beam2d.py:592-600)
//
// is a complex GEMM with a tiny output (batch x nmodes) and a huge reduction (p = nb^2 samples of the
// centred crop).  It is mapped on ONE real TF32 GEMM per CTA tile:
//
//   A (UMMA M = 128 rows)   = 128 modes, as stored: (Mr, Mi) interleaved along K (K-major)
//   B (UMMA N = 2 * 64 rows) = for every beam the two rows (Fr, -Fi) and (Fi, Fr), so that
//   D[m][2b]   = sum Mr*Fr - Mi*Fi = Re coeffs[b][m],     D[m][2b+1] = sum Mr*Fi + Mi*Fr = Im coeffs[b][m]
//
// FP32 accuracy out of TF32 products (the parity bound is 1e-5): every operand is split as
// x = hi + lo with hi = x rounded to TF32 and lo = x - hi (exact in FP32), and the three products
// hi*hi + hi*lo + lo*hi are accumulated (3xTF32; lo*lo is below 2^-22 relative).
//
// The reduction is split over the CTAs (split-K): a CTA owns a contiguous range of K blocks and
// accumulates them in TMEM (128 lanes x 128 FP32 columns).  The tensor core truncates when it adds to
// the accumulator, so the error of one accumulation chain grows linearly with its length (measured:
// 2.5e-5 relative after 55 blocks): every FLUSH blocks the chain is read back (tcgen05.ld) and added
// -- round to nearest -- to FP32 registers, and the next MMA starts a fresh chain; at the end the
// register sums are added to the zero-initialised output with FP32 atomics (red.global.add.f32).
//
// Per K block (16 complex samples = 32 TF32 values per row = 4 MMA K-steps): the 512 producer threads copy the
// mode tile and the field tile from HBM (coalesced 16-byte loads, issued one block ahead), split /
// combine them in registers, and write the four operand tiles to one of two shared-memory stages in
// the canonical no-swizzle K-major UMMA layout ([16-byte K chunk][row], chunk stride padded by 16
// bytes against bank conflicts) and hand the stage to a seventeenth warp through a "full" mbarrier; that warp
// issues the 12 tcgen05.mma of the block and commits them to the stage's "empty" mbarrier, which the
// next overwrite of the stage waits on -- the tensor core works on block k while the producers stage
// block k+1, and issuing MMAs never delays the staging.
#include <cstdint>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "../../include/lp2d.h"
#include <cuda_runtime.h>

namespace lp2d {
int set_error(int code, const std::string& msg);   // lp2d_api.cu
extern long long g_launches;

namespace tc {

constexpr int UM = 128;            // modes per tile (UMMA M)
constexpr int BT = 64;             // beams per tile
constexpr int UN = 2 * BT;         // UMMA N
constexpr int KB = 16;             // complex samples per K block
constexpr int CHUNKS = KB / 2;     // 16-byte K chunks per row and block (4 TF32 values = 2 complex samples)
constexpr int THREADS = 512;       // producer / epilogue threads (warps 0..15); warp 16 issues the MMAs
constexpr int CTA_THREADS = THREADS + 32;
constexpr int FLUSH = 8;            // K blocks per TMEM accumulation chain
constexpr int PF = 2;               // K blocks of global loads in flight per thread (register sets)
constexpr int STAGES = 2;           // operand tile sets in shared memory (the stores of block k+1 overlap the MMAs of block k)
constexpr int ACC = UN / (THREADS / 128);   // accumulator columns per thread: warp w owns lanes 32*(w%4).., columns ACC*(w/4)..
constexpr int CH_A = UM * 16 + 16; // byte stride between K chunks of the A tiles (padded)
constexpr int CH_B = UN * 16 + 16;
constexpr int TILE_A = CHUNKS * CH_A;
constexpr int TILE_B = CHUNKS * CH_B;
constexpr int STAGE_BYTES = 2 * TILE_A + 2 * TILE_B;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int A_LD = UM * CHUNKS / THREADS;        // 16-byte loads per thread and block: modes
constexpr int B_LD = BT * CHUNKS / THREADS;        // ... fields
constexpr int TMEM_COLS = 128;                     // FP32 accumulator columns (= UN)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address [0,14), leading byte offset [16,30) = distance between the two 16-byte K chunks of one
// MMA, stride byte offset [32,46) = distance between 8-row groups, version [46,48) = 1, layout type
// [61,64) = 0; all offsets in units of 16 bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = B = TF32 [7,10), [10,13) = 2,
// both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = clock64();
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
        if (!done && clock64() - t0 > 4000000000ll) __trap();      // never hang the device on a lost arrive
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// x = hi + lo, hi = x rounded to nearest TF32 (10 explicit mantissa bits), lo = x - hi exactly
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

struct Args {
    const float2* field;    // batch x n x n
    const float2* modes;    // nmodes x nb*nb
    float* coeffs;          // batch x nmodes complex, zero-initialised; partial sums are added
    int n, nb, nmodes, batch;
    int kblocks;            // nb*nb / KB
    int ksplit;             // CTAs along the reduction
    int mtiles, btiles;
    int nmma;               // debug: 3 = 3xTF32 (product), 1 = hi*hi only (timing experiments)
};

__global__ void __launch_bounds__(CTA_THREADS, 1) modes_project_tc_kernel(const Args p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_full[STAGES];      // producers -> MMA warp: the stage holds block i
    __shared__ uint64_t s_empty[STAGES];     // tcgen05.commit -> producers: the MMAs that read the stage are done
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;

    const int tile = blockIdx.x / p.ksplit, ks = blockIdx.x % p.ksplit;
    const int mt = tile % p.mtiles, bt = tile / p.mtiles;
    const int m0 = mt * UM, b0 = bt * BT;
    const int kb_lo = (int)((long long)p.kblocks * ks / p.ksplit), kb_hi = (int)((long long)p.kblocks * (ks + 1) / p.ksplit);

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    const int dn = (p.n - p.nb) / 2;
    const long long kk = (long long)p.nb * p.nb;
    // per-thread load assignment: piece q = tid + THREADS*i -> (row, chunk) = (q / CHUNKS, q % CHUNKS):
    // 8 consecutive threads read the 128 contiguous bytes of one row of the block.  Everything that does
    // not change from block to block is computed once: the global row origins, the shared-memory
    // offsets and the validity of the rows; a block then costs one pointer offset per load.
    const float4* pa[A_LD];            // mode rows: origin of chunk c of row (m0 + row) at crop sample 0
    uint32_t oa[A_LD];                 // byte offset of (chunk, row) inside an A tile
    bool va[A_LD];
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
        const int q = tid + THREADS * i, row = q / CHUNKS, c = q % CHUNKS;
        va[i] = (m0 + row < p.nmodes);
        pa[i] = reinterpret_cast<const float4*>(p.modes + (long long)(va[i] ? m0 + row : 0) * kk) + c;
        oa[i] = (uint32_t)(c * CH_A + row * 16);
    }
    const float4* pb[B_LD];            // field planes: origin of chunk c of beam (b0 + row) at crop sample (0, 0)
    uint32_t ob[B_LD];                 // byte offset of (chunk, beam row pair) inside a B tile
    bool vb[B_LD];
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
        const int q = tid + THREADS * i, row = q / CHUNKS, c = q % CHUNKS;
        vb[i] = (b0 + row < p.batch);
        pb[i] = reinterpret_cast<const float4*>(p.field + ((long long)(vb[i] ? b0 + row : 0) * p.n + dn) * p.n + dn) + c;
        ob[i] = (uint32_t)(c * CH_B + (2 * row) * 16);
    }
    // position of the next block to LOAD, advanced incrementally (KB divides nb: a block stays in one crop row)
    long long ld_a = (long long)kb_lo * KB / 2;                              // float4 offset along a mode row
    int ld_cy = (int)(((long long)kb_lo * KB) / p.nb), ld_cx = (int)(((long long)kb_lo * KB) % p.nb);
    float4 ra[PF][A_LD], rb[PF][B_LD];
    auto load_block = [&](auto set_c) {
        constexpr int SET = decltype(set_c)::value;
        const long long ld_b = ((long long)ld_cy * p.n + ld_cx) / 2;         // float4 offset inside the field plane
#pragma unroll
        for (int i = 0; i < A_LD; ++i) ra[SET][i] = va[i] ? __ldcs(pa[i] + ld_a) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < B_LD; ++i) rb[SET][i] = vb[i] ? __ldg(pb[i] + ld_b) : make_float4(0.f, 0.f, 0.f, 0.f);
        ld_a += KB / 2;
        ld_cx += KB;
        if (ld_cx >= p.nb) { ld_cx = 0; ++ld_cy; }
    };
    auto store_block = [&](int stage, auto set_c) {
        constexpr int SET = decltype(set_c)::value;
        unsigned char* a_hi = smem + stage * STAGE_BYTES;
        unsigned char* a_lo = a_hi + TILE_A;
        unsigned char* b_hi = a_lo + TILE_A;
        unsigned char* b_lo = b_hi + TILE_B;
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            float4 h, l;
            split_tf32(ra[SET][i].x, h.x, l.x); split_tf32(ra[SET][i].y, h.y, l.y);
            split_tf32(ra[SET][i].z, h.z, l.z); split_tf32(ra[SET][i].w, h.w, l.w);
            *reinterpret_cast<float4*>(a_hi + oa[i]) = h;
            *reinterpret_cast<float4*>(a_lo + oa[i]) = l;
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            // two complex samples (x, y), (z, w) of a beam: row 2b = (Fr, -Fi), row 2b+1 = (Fi, Fr); the
            // split commutes with the sign and the permutation, so every value is split once
            float4 h, l;
            split_tf32(rb[SET][i].x, h.x, l.x); split_tf32(rb[SET][i].y, h.y, l.y);
            split_tf32(rb[SET][i].z, h.z, l.z); split_tf32(rb[SET][i].w, h.w, l.w);
            *reinterpret_cast<float4*>(b_hi + ob[i]) = make_float4(h.x, -h.y, h.z, -h.w);
            *reinterpret_cast<float4*>(b_lo + ob[i]) = make_float4(l.x, -l.y, l.z, -l.w);
            *reinterpret_cast<float4*>(b_hi + ob[i] + 16) = make_float4(h.y, h.x, h.w, h.z);
            *reinterpret_cast<float4*>(b_lo + ob[i] + 16) = make_float4(l.y, l.x, l.w, l.z);
        }
    };

    // this thread's part of the accumulator: TMEM lane 32*(warp%4) + lane-in-warp, columns col0 .. col0+ACC-1
    float acc[ACC];
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = 0.f;
    const int col0 = (warp >> 2) * ACC;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col0;
    auto flush = [&]() {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < ACC; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(tlane + (uint32_t)c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    // Block i (0-based in this CTA's range) lives in stage i % STAGES; it is the (i / STAGES)-th use of the
    // stage, so both of the stage's mbarriers complete phase i / STAGES for it (parity (i / STAGES) & 1).
    const int nblocks = kb_hi - kb_lo;
    if (warp == THREADS / 32) {
        // ===== MMA issuer (one lane) =====
        if ((tid & 31) == 0) {
            for (int i = 0; i < nblocks; ++i) {
                const int stage = i % STAGES;
                const bool fresh = (i % FLUSH) == 0;                 // this block starts a new accumulation chain
                mbar_wait(&s_full[stage], (uint32_t)((i / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES), sa_lo = sa_hi + TILE_A, sb_hi = sa_lo + TILE_A, sb_lo = sb_hi + TILE_B;
#pragma unroll
                for (int s = 0; s < CHUNKS / 2; ++s) {              // one MMA K-step = 8 TF32 values = 2 chunks
                    const uint64_t dah = umma_desc(sa_hi + 2 * s * CH_A, CH_A, 128), dal = umma_desc(sa_lo + 2 * s * CH_A, CH_A, 128);
                    const uint64_t dbh = umma_desc(sb_hi + 2 * s * CH_B, CH_B, 128), dbl = umma_desc(sb_lo + 2 * s * CH_B, CH_B, 128);
                    mma_tf32(tmem, dah, dbh, (!fresh || s > 0) ? 1u : 0u);
                    if (p.nmma == 3) {
                        mma_tf32(tmem, dah, dbl, 1u);
                        mma_tf32(tmem, dal, dbh, 1u);
                    }
                }
                umma_commit(&s_empty[stage]);                       // arrives when every MMA issued so far is complete
            }
        }
    } else {
        // ===== producers: stage the register set SET (loaded PF blocks ago), refill it, hand the stage over =====
        auto step = [&](int i, auto set_c) {
            const int stage = i % STAGES;
            if ((i % FLUSH) == 0 && i > 0) {
                // the chain closed by block i-1 is complete (a commit covers every earlier MMA); the MMA warp
                // cannot start the next chain before this block is handed over below
                mbar_wait(&s_empty[(i - 1) % STAGES], (uint32_t)(((i - 1) / STAGES) & 1));
                flush();
            }
            if (i >= STAGES) mbar_wait(&s_empty[stage], (uint32_t)((i / STAGES - 1) & 1));   // the stage's previous block was multiplied
            store_block(stage, set_c);
            if (i + PF < nblocks) load_block(set_c);                    // in flight for the next PF blocks
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
            asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");      // all producers have written the stage
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[stage])) : "memory");
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        static_assert(PF == 2, "the loop below rotates two register sets");
        if (0 < nblocks) load_block(I0{});
        if (1 < nblocks) load_block(I1{});
        for (int i = 0; i < nblocks; i += 2) {
            step(i, I0{});
            if (i + 1 < nblocks) step(i + 1, I1{});
        }
    }
    if (nblocks > 0 && warp < THREADS / 32) {
        mbar_wait(&s_empty[(nblocks - 1) % STAGES], (uint32_t)(((nblocks - 1) / STAGES) & 1));
        flush();
        // epilogue: lane = mode m0 + 32*(warp%4) + lane-in-warp, column c = (beam, re/im)
        const int m = m0 + (warp & 3) * 32 + (tid & 31);
        if (m < p.nmodes) {
#pragma unroll
            for (int j = 0; j < ACC; ++j) {
                const int col = col0 + j, b = b0 + (col >> 1);
                if (b < p.batch) atomicAdd(p.coeffs + ((long long)b * p.nmodes + m) * 2 + (col & 1), acc[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Mode synthesis on the tensor cores:  field_b[crop(p)] = sum_m coeffs[b][m] * modes[m][p]   (beam2d.py:552-562, 664-666)
//
// The same real TF32 GEMM with the roles changed: UMMA M = 128 grid points p (the big dimension, nb^2 of them), N = 2 * 64
// (beam, re / im), K = 2 * nmodes (the reduction runs over the modes).  Per K block of 16 modes the producers read
// 128 points x 16 modes of the mode matrix (every warp load = 32 consecutive points of one mode: 256 contiguous bytes)
// and 64 beams x 16 coefficients, split them hi / lo and stage them in the SAME canonical K-major tiles as the projection
// kernel: A row p = (Mr, Mi) of the block's modes, B rows (2b, 2b+1) = (Cr, -Ci), (Ci, Cr).  The MMA warp, the
// descriptors, the 3xTF32 products and the chain flush are shared with the projection.  No split-K: a CTA owns a tile of
// 128 points x 64 beams for all modes and stores its result straight from the TMEM read-back (consecutive lanes =
// consecutive points of one beam: coalesced 8-byte stores).  SIMT needs 8 * batch * nmodes * nb^2 real FMAs on the FP32
// pipes (34 GFLOP at config 5: 1.49 ms measured); here the mode matrix streams once at tensor-core speed.
struct SynthArgs {
    const float2* coeffs;   // batch x nmodes
    const float2* modes;    // nmodes x nb*nb
    float2* field;          // batch x n x n (zero-filled by the caller outside the crop)
    int n, nb, nmodes, batch;
    int kblocks;            // ceil(nmodes / KB)
    int ptiles, btiles;
};

__global__ void __launch_bounds__(CTA_THREADS, 1) modes_synth_tc_kernel(const SynthArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_full[STAGES];
    __shared__ uint64_t s_empty[STAGES];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;

    const int pt_tile = blockIdx.x % p.ptiles, bt = blockIdx.x / p.ptiles;
    const long long p0 = (long long)pt_tile * UM;
    const int b0 = bt * BT;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    const long long kk = (long long)p.nb * p.nb;
    // A: thread = (point pt, mode quad g): modes kb*KB + 4g .. 4g+3 of point p0 + pt -> chunks 2g, 2g+1 of row pt
    const int a_pt = tid & (UM - 1), a_g = tid >> 7;
    const float2* a_src = p.modes + p0 + a_pt;
    const uint32_t a_off = (uint32_t)((2 * a_g) * CH_A + a_pt * 16);
    // B: thread = (beam row, chunk c): coefficients kb*KB + 2c, 2c+1 of beam b0 + row
    const int b_row = tid / CHUNKS, b_c = tid % CHUNKS;
    const bool b_ok = (b0 + b_row < p.batch);
    const float2* b_src = p.coeffs + (long long)(b_ok ? b0 + b_row : 0) * p.nmodes + 2 * b_c;
    const uint32_t b_off = (uint32_t)(b_c * CH_B + (2 * b_row) * 16);

    float2 ra[PF][4], rb[PF][2];
    int ld_kb = 0;
    auto load_block = [&](auto set_c) {
        constexpr int SET = decltype(set_c)::value;
        const int m0 = ld_kb * KB;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + 4 * a_g + j;
            ra[SET][j] = (m < p.nmodes) ? __ldcs(a_src + (long long)m * kk) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + 2 * b_c + j;
            rb[SET][j] = (b_ok && m < p.nmodes) ? __ldg(b_src + m0 + j) : make_float2(0.f, 0.f);
        }
        ++ld_kb;
    };
    auto store_block = [&](int stage, auto set_c) {
        constexpr int SET = decltype(set_c)::value;
        unsigned char* a_hi = smem + stage * STAGE_BYTES;
        unsigned char* a_lo = a_hi + TILE_A;
        unsigned char* b_hi = a_lo + TILE_A;
        unsigned char* b_lo = b_hi + TILE_B;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float4 h, l;
            split_tf32(ra[SET][2 * c].x, h.x, l.x); split_tf32(ra[SET][2 * c].y, h.y, l.y);
            split_tf32(ra[SET][2 * c + 1].x, h.z, l.z); split_tf32(ra[SET][2 * c + 1].y, h.w, l.w);
            *reinterpret_cast<float4*>(a_hi + a_off + c * CH_A) = h;
            *reinterpret_cast<float4*>(a_lo + a_off + c * CH_A) = l;
        }
        {
            float4 h, l;
            split_tf32(rb[SET][0].x, h.x, l.x); split_tf32(rb[SET][0].y, h.y, l.y);
            split_tf32(rb[SET][1].x, h.z, l.z); split_tf32(rb[SET][1].y, h.w, l.w);
            *reinterpret_cast<float4*>(b_hi + b_off) = make_float4(h.x, -h.y, h.z, -h.w);
            *reinterpret_cast<float4*>(b_lo + b_off) = make_float4(l.x, -l.y, l.z, -l.w);
            *reinterpret_cast<float4*>(b_hi + b_off + 16) = make_float4(h.y, h.x, h.w, h.z);
            *reinterpret_cast<float4*>(b_lo + b_off + 16) = make_float4(l.y, l.x, l.w, l.z);
        }
    };

    float acc[ACC];
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = 0.f;
    const int col0 = (warp >> 2) * ACC;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col0;
    auto flush = [&]() {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < ACC; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(tlane + (uint32_t)c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    const int nblocks = p.kblocks;
    if (warp == THREADS / 32) {
        if ((tid & 31) == 0) {
            for (int i = 0; i < nblocks; ++i) {
                const int stage = i % STAGES;
                const bool fresh = (i % FLUSH) == 0;
                mbar_wait(&s_full[stage], (uint32_t)((i / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES), sa_lo = sa_hi + TILE_A, sb_hi = sa_lo + TILE_A, sb_lo = sb_hi + TILE_B;
#pragma unroll
                for (int s = 0; s < CHUNKS / 2; ++s) {
                    const uint64_t dah = umma_desc(sa_hi + 2 * s * CH_A, CH_A, 128), dal = umma_desc(sa_lo + 2 * s * CH_A, CH_A, 128);
                    const uint64_t dbh = umma_desc(sb_hi + 2 * s * CH_B, CH_B, 128), dbl = umma_desc(sb_lo + 2 * s * CH_B, CH_B, 128);
                    mma_tf32(tmem, dah, dbh, (!fresh || s > 0) ? 1u : 0u);
                    mma_tf32(tmem, dah, dbl, 1u);
                    mma_tf32(tmem, dal, dbh, 1u);
                }
                umma_commit(&s_empty[stage]);
            }
        }
    } else {
        auto step = [&](int i, auto set_c) {
            const int stage = i % STAGES;
            if ((i % FLUSH) == 0 && i > 0) {
                mbar_wait(&s_empty[(i - 1) % STAGES], (uint32_t)(((i - 1) / STAGES) & 1));
                flush();
            }
            if (i >= STAGES) mbar_wait(&s_empty[stage], (uint32_t)((i / STAGES - 1) & 1));
            store_block(stage, set_c);
            if (i + PF < nblocks) load_block(set_c);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[stage])) : "memory");
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        if (0 < nblocks) load_block(I0{});
        if (1 < nblocks) load_block(I1{});
        for (int i = 0; i < nblocks; i += 2) {
            step(i, I0{});
            if (i + 1 < nblocks) step(i + 1, I1{});
        }
    }
    if (nblocks > 0 && warp < THREADS / 32) {
        mbar_wait(&s_empty[(nblocks - 1) % STAGES], (uint32_t)(((nblocks - 1) / STAGES) & 1));
        flush();
        // epilogue: TMEM lane = point p0 + 32*(warp%4) + lane-in-warp, columns = (beam, re / im)
        const long long pp = p0 + (warp & 3) * 32 + (tid & 31);
        const int cy = (int)(pp / p.nb), cx = (int)(pp % p.nb);
        const int dn = (p.n - p.nb) / 2;
        float2* dst = p.field + (long long)(cy + dn) * p.n + (cx + dn);
        const long long plane = (long long)p.n * p.n;
#pragma unroll
        for (int j = 0; j < ACC; j += 2) {
            const int b = b0 + ((col0 + j) >> 1);
            if (b < p.batch) __stcs(dst + (long long)b * plane, make_float2(acc[j], acc[j + 1]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc

// complex64 only; nb must be a multiple of the K block.  Returns LP2D_EUNSUPPORTED when the shape
// is not covered (the caller then uses the SIMT kernel).
int modes_project_tc(const void* field, const void* modes, void* coeffs, int n, int nb, int nmodes, long long batch,
                     cudaStream_t s) {
    using namespace tc;
    if (nb % KB != 0 || ((n - nb) / 2) % 2 != 0 || n % 2 != 0) return LP2D_EUNSUPPORTED;   // 16-byte aligned rows
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(modes_project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_project_tc setup: ") + cudaGetErrorString(e));
        configured = true;
    }
    Args a;
    a.field = static_cast<const float2*>(field);
    a.modes = static_cast<const float2*>(modes);
    a.coeffs = static_cast<float*>(coeffs);
    a.n = n; a.nb = nb; a.nmodes = nmodes; a.batch = (int)batch;
    a.kblocks = (int)((long long)nb * nb / KB);
    a.mtiles = (nmodes + UM - 1) / UM;
    a.btiles = (int)((batch + BT - 1) / BT);
    const int tiles = a.mtiles * a.btiles;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ksplit = (2 * sms + tiles - 1) / tiles;            // about two waves of CTAs (one CTA per SM at a time)
    if (ksplit > a.kblocks) ksplit = a.kblocks;
    if (ksplit < 1) ksplit = 1;
    a.ksplit = ksplit;
    { const char* dbg = getenv("LP2D_TC_DEBUG_NMMA"); a.nmma = (dbg != nullptr && atoi(dbg) == 1) ? 1 : 3; }
    cudaError_t e = cudaMemsetAsync(coeffs, 0, sizeof(float2) * (size_t)batch * nmodes, s);
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_project_tc memset: ") + cudaGetErrorString(e));
    modes_project_tc_kernel<<<(unsigned)(tiles * ksplit), CTA_THREADS, SMEM_BYTES, s>>>(a);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_project_tc launch: ") + cudaGetErrorString(e));
    return LP2D_OK;
}

// complex64 only; the crop must be made of whole 32-point runs (nb % 32 == 0, nb^2 % 128 == 0).  Returns
// LP2D_EUNSUPPORTED when the shape is not covered (the caller then uses the SIMT kernel).
int modes_synthesize_tc(const void* coeffs, const void* modes, void* field, int n, int nb, int nmodes, long long batch, cudaStream_t s) {
    using namespace tc;
    if (nb % 32 != 0 || ((long long)nb * nb) % UM != 0 || nmodes < 1) return LP2D_EUNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(modes_synth_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_synthesize_tc setup: ") + cudaGetErrorString(e));
        configured = true;
    }
    SynthArgs a;
    a.coeffs = static_cast<const float2*>(coeffs);
    a.modes = static_cast<const float2*>(modes);
    a.field = static_cast<float2*>(field);
    a.n = n; a.nb = nb; a.nmodes = nmodes; a.batch = (int)batch;
    a.kblocks = (nmodes + KB - 1) / KB;
    a.ptiles = (int)(((long long)nb * nb) / UM);
    a.btiles = (int)((batch + BT - 1) / BT);
    cudaError_t e = cudaSuccess;
    if (nb != n) e = cudaMemsetAsync(field, 0, sizeof(float2) * (size_t)batch * n * n, s);     // zero padding around the crop
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_synthesize_tc memset: ") + cudaGetErrorString(e));
    modes_synth_tc_kernel<<<(unsigned)((long long)a.ptiles * a.btiles), CTA_THREADS, SMEM_BYTES, s>>>(a);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LP2D_ECUDA, std::string("modes_synthesize_tc launch: ") + cudaGetErrorString(e));
    return LP2D_OK;
}

}  // namespace lp2d
