// Hot path (a) as dense one-hot contractions on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// The north star allows tensor cores for the PLL gradient only if the dense recast beats the gather path under
// ncu; bench.py / profiles/ carry that comparison (the gather kernels are kept, see plm_gather.cu): forward
// 5.1 ms -> 2.1 + 0.40 ms, backward 15.7 ms -> 2.2 ms at the same parity tolerance (round 2, config 2).
//
// Maths.  With X[n,(j,b)] = [s_nj = b] (one-hot, exact in bf16), the couplings W[(i,a),(j,b)] = J_ij(a,b) and
// the residuals R[n,(i,a)] = r_ni(a):
//     forward   Zt[(i,a), n]     = sum_(j,b) W[(i,a),(j,b)] * X[n,(j,b)]          (logits without h)
//     backward  Gd[(j,b),(i,a)]  = sum_n     X[n,(j,b)]     * R[n,(i,a)]
//               g_J(i<j)[a][b]   = Gd[(j,b),(i,a)] + Gd[(i,a),(j,b)]
// Precision mode 0 (fp32-equivalent, default): the real-valued operand (W or R) is split in two bf16 terms
// (hi = rn(v), lo = rn(v - hi)): 16 mantissa bits, relative error 2^-17 per term; both products accumulate into the
// SAME fp32 TMEM accumulator.  Precision mode 1 ("bf16 tiles", BASELINE configs[4]): hi only, one product per term.
//
// Operands, all K-major (TMA 2-D, SWIZZLE_128B):
//     forward : Wt_hi, Wt_lo [Mp][Kw] bf16 (written by expand_tc every evaluation), X [Xrows][Kw] bf16 (static)
//     backward: Xt [Mp][Kp] bf16 (static), Rt_hi, Rt_lo [Np][Kp] bf16 (written by plm_softmax_kernel)
//
// tc_gemm_persistent_kernel<SPLIT_A, SINGLE>: one persistent CTA per SM, 128 x 192 tiles, K blocks of 64.
//     warp 0             TMA producer: shared-memory ring (4 x 56 KB / 3 x 64 KB / 5 x 40 KB depending on mode),
//                        mbarrier expect_tx, L2 evict_last on the operand every tile re-reads
//     warp 1             MMA issuer: per K block 4 (x 2 in mode 0) tcgen05.mma.cta_group::1.kind::f16 (M128 N192 K16);
//                        tcgen05.commit frees the stage / publishes the accumulator
//                        Both control warps run their loops CONVERGED and issue through elect.sync: as a single
//                        divergent thread the issue block was ~130 SASS instructions per K block (~700 cycles, more
//                        than the 384 MMA cycles of a mode-1 K block); now ~25
//     warp 2             TMEM allocator: 512 columns = two 192-column fp32 accumulators (double buffered)
//     warps 4..11        epilogue (two per TMEM lane quadrant): tcgen05.ld 32x32b.x16; K-chunk sums are promoted
//                        into registers with IEEE round-to-nearest adds (the tensor core's own fp32 accumulation
//                        truncates: measured -2.6e-5 relative bias over 782 K blocks without promotion); no spills
// Tile order: M tiles in groups whose slice of the coupling operand is ~24 MB (L2-resident), M fastest inside a
// group (decode_tile, forward_mgroup) -- DESIGN.md 4b.
// tc_gemm_pair_kernel<SPLIT_A>: the same product on CTA pairs (cta_group::2, 256 x 192 tiles); parity-green, slower,
// opt-in (EVC_TC_PAIR=1) -- DESIGN.md 4b.
// Roofline: tensor pipe.  Measured (round 2, N=50k, L=200, q=21): mode 0 tensor pipe 90 % / 86-88 % active,
// 1.5-1.6 PFLOP/s executed per GEMM = 0.89-0.95 of the cuBLAS bf16 burst rate on the same part; mode 1 70 % / 83 %.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace evc {

constexpr int TC_BM = 128;
constexpr int TC_BN = 192;
constexpr int TC_BK = 64;
constexpr int TC_MAX_STAGES = 8;   // ring depth is chosen at launch from the stage size (hi+lo: 3-4, bf16x1: 5)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;        // 16384
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;        // 24576
constexpr int TC_SMEM_LIMIT = 232448;                // 227 KB opt-in shared memory per CTA
constexpr int TC_SMEM_HEAD = 2048;                   // 1 KB alignment slack + 1 KB of mbarriers / TMEM slot
constexpr int TC_THREADS = 384;   // warps 0-2: TMA / MMA / TMEM alloc, warps 4-11: epilogue (2 per TMEM lane quadrant)
constexpr int TC_PAIR_DEFAULT = 0; // CTA-pair (cta_group::2) GEMM tiles: opt-in via EVC_TC_PAIR=1 until validated on hardware
constexpr int TC_SPLIT_PRODUCER_DEFAULT = 0;   // two TMA producer threads per CTA: opt-in via EVC_SPLIT_PRODUCER=1 until measured
constexpr int TC_K_CHUNK = 32;   // k-blocks (of 64) accumulated in TMEM before promotion to an fp32 add

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    // spin with a cap so that a programming error becomes a trap instead of a hung GPU
    uint32_t done = 0;
    for (uint64_t it = 0; it < (1ull << 31); it++) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// same, with an L2 cache-policy hint (used to keep the operand that every tile re-reads resident in L2)
__device__ __forceinline__ void tma_load_2d_hint(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                                 uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// true on exactly one lane of a fully converged warp (elect.sync): lets ptxas keep the single-thread tcgen05 /
// TMA instructions on the uniform datapath without the per-instruction "loop over active threads" it emits for
// code that is merely divergent (lane == 0)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n.reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// K-major, 128-byte swizzle, densely packed 8-row x 128-byte atoms (SBO = 1024 B), sm_100 descriptor version 1
__device__ __forceinline__ uint64_t make_desc_sw128(const void *smem_ptr)
{
    const uint32_t addr = smem_u32(smem_ptr);
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);       // start address >> 4          bits [0,14)
    d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major) bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset >> 4      bits [32,46)
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                      // layout type SWIZZLE_128B
    return d;
}

// instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M = 128, N = TC_BN
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------
// tcgen05.ld wrappers (32 lanes x 32 bit, N consecutive columns per thread)
// ---------------------------------------------------------------------------------------------------
template <int NCOL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t *v);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<4>(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// Persistent GEMM with a double-buffered TMEM accumulator (2 x 192 columns): the epilogue of work item t
// overlaps the main loop of work item t+1.
//     forward  (SPLIT_A = 1)  Zt[(i,a), n]     = sum_(j,b) (Wt_hi [+ Wt_lo])[(i,a),(j,b)] * X[n,(j,b)]
//     backward (SPLIT_A = 0)  Gd[(j,b),(i,a)]  = sum_n     Xt[(j,b), n] * (Rt_hi [+ Rt_lo])[(i,a), n]
// Stage layout (the optional lo operand is LAST so that the bf16x1 precision mode uses a compact prefix and
// a deeper ring): SPLIT_A = 1: [A_hi 16 KB][B 24 KB][A_lo 16 KB];  SPLIT_A = 0: [A 16 KB][B_hi 24 KB][B_lo 24 KB].
// `single` != 0 (precision mode 1, "bf16 tiles"): the lo operand is neither loaded nor multiplied -- one
// tcgen05.mma per K slice instead of two.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Tile enumeration: M tiles are swept in groups of `mgroup`; inside a group the M index is fastest, then the N
// tile.  For the forward product (SPLIT_A) a group of M tiles whose slice of the coupling operand fits in L2
// (about 24 MB, see plm_tcf_logits) stays resident while every sequence tile passes by; the backward uses one
// group (all CTAs advance along K together and share operand tiles in time).
__device__ __forceinline__ void decode_tile(int tile, int m_tiles, int n_tiles, int mgroup, int &m_tile, int &n_tile)
{
    const int full = (m_tiles / mgroup) * mgroup * n_tiles;
    if (tile < full) {
        const int per = mgroup * n_tiles;
        const int g = tile / per, r = tile - g * per;
        n_tile = r / mgroup;
        m_tile = g * mgroup + (r - n_tile * mgroup);
    } else {
        const int rem = m_tiles % mgroup, r = tile - full;
        n_tile = r / rem;
        m_tile = (m_tiles / mgroup) * mgroup + (r - n_tile * rem);
    }
}

template <int SPLIT_A, int SINGLE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                          const __grid_constant__ CUtensorMap tm2, float *__restrict__ D, int64_t ldd,
                          int m_tiles, int n_tiles, int num_kb, int k_chunk, int mgroup, int n_stages, int split_prod)
{
    constexpr int single = SINGLE;      // precision mode 1 (bf16 tiles) is a separate instantiation: no lo operand at all
    // Work item = (tile, K chunk).  The tensor core's fp32 accumulator truncates instead of rounding to
    // nearest, so a long accumulation chain picks up a systematic bias (measured -2.6e-5 relative over
    // 782 k-blocks); accumulating at most k_chunk k-blocks in TMEM and adding the chunk results in the
    // epilogue (IEEE round-to-nearest) keeps it at the level of a plain fp32 sum.
    constexpr int BYTES0 = TC_A_BYTES;                              // operand 0: A_hi (fwd) / A (bwd), 128 rows
    constexpr int BYTES1 = TC_B_BYTES;                              // operand 1: B (fwd) / B_hi (bwd), 192 rows
    constexpr int BYTES2 = SPLIT_A ? TC_A_BYTES : TC_B_BYTES;       // operand 2: A_lo (fwd) / B_lo (bwd), optional
    constexpr int stage_bytes = BYTES0 + BYTES1 + (SINGLE ? 0 : BYTES2);
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem0 = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                                             ~static_cast<uintptr_t>(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem0);
    uint64_t *empty = full + TC_MAX_STAGES;
    uint64_t *acc_full = empty + TC_MAX_STAGES;  // [2]
    uint64_t *acc_empty = acc_full + 2;          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    unsigned char *smem = smem0 + 1024;          // operand ring, 1024-byte aligned (SWIZZLE_128B atoms)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = m_tiles * n_tiles;
    const int n_chunks = (num_kb + k_chunk - 1) / k_chunk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 8);         // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || (warp == 3 && split_prod >= 1) || (warp == 2 && split_prod >= 2)) {
        // ===== TMA producer(s): whole warp in the loop, one elected lane issues =====
        // split_prod = 1: the loads of a stage are issued from TWO warps (warp 0: operand 0 + the lo operand +
        // expect_tx, warp 3: operand 1); split_prod = 2: three warps (warp 2, idle after the TMEM allocation, takes the
        // lo operand).  SPLIT_A (forward): the coupling matrix (A_hi, A_lo) is re-read by every sequence tile -> evict_last
        const bool ld0 = warp == 0;
        const bool ld1 = split_prod >= 1 ? warp == 3 : true;
        const bool ld2 = !SINGLE && (split_prod >= 2 ? warp == 2 : warp == 0);
        const uint64_t keep = l2_policy_evict_last();
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_tile, n_tile;
            decode_tile(tile, m_tiles, n_tiles, mgroup, m_tile, n_tile);
            for (int kb = 0; kb < num_kb; kb++) {
                mbar_wait_bounded(&empty[s], ph ^ 1u);
                if (elect_one()) {
                    unsigned char *st = smem + s * stage_bytes;
                    if (ld0) mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    if (SPLIT_A) {
                        if (ld0) tma_load_2d_hint(st, &tm0, kb * TC_BK, m_tile * TC_BM, &full[s], keep);
                        if (ld1) tma_load_2d(st + BYTES0, &tm2, kb * TC_BK, n_tile * TC_BN, &full[s]);
                        if (ld2) tma_load_2d_hint(st + BYTES0 + BYTES1, &tm1, kb * TC_BK, m_tile * TC_BM, &full[s], keep);
                    } else {
                        if (ld0) tma_load_2d(st, &tm0, kb * TC_BK, m_tile * TC_BM, &full[s]);
                        if (ld1) tma_load_2d(st + BYTES0, &tm1, kb * TC_BK, n_tile * TC_BN, &full[s]);
                        if (ld2) tma_load_2d(st + BYTES0 + BYTES1, &tm2, kb * TC_BK, n_tile * TC_BN, &full[s]);
                    }
                }
                __syncwarp();
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (all lanes wait on the barriers, all values are warp-uniform);
        //       the tcgen05 instructions are issued by the lane elect.sync picks =====
        constexpr uint32_t idesc = make_idesc_bf16(TC_BM, TC_BN);
        const uint64_t desc0 = make_desc_sw128(smem);                 // stage 0, operand 0; everything else is an offset
        constexpr uint64_t OFF1 = (uint64_t)(BYTES0 >> 4), OFF2 = (uint64_t)((BYTES0 + BYTES1) >> 4);
        int s = 0, wl = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            for (int c = 0; c < n_chunks; c++, wl++) {
                const int acc = wl & 1;
                mbar_wait_bounded(&acc_empty[acc], (uint32_t)(((wl >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_BN);
                const int kb0 = c * k_chunk, kb1 = min(num_kb, kb0 + k_chunk);
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait_bounded(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t d0 = desc0 + (uint64_t)((s * stage_bytes) >> 4);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++) {
                            const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);
                            const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
                            // operand 0 is always the 128-row (A) tile, operand 1 the 192-row (B) tile
                            umma_bf16(tmem_d, d0 + koff, d0 + OFF1 + koff, idesc, first);
                            if (!SINGLE) {
                                if (SPLIT_A) umma_bf16(tmem_d, d0 + OFF2 + koff, d0 + OFF1 + koff, idesc, 1u);   // A_lo * B
                                else umma_bf16(tmem_d, d0 + koff, d0 + OFF2 + koff, idesc, 1u);                   // A * B_lo
                            }
                        }
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) umma_commit(&acc_full[acc]);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int quad = warp & 3;              // TMEM lane quadrant this warp may access
        const int ehalf = (warp - 4) >> 2;      // which half of the tile's columns this warp drains
        constexpr int ECOLS = TC_BN / 2;        // 96 columns per epilogue warp
        int wl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m_tile, n_tile;
            decode_tile(tile, m_tiles, n_tiles, mgroup, m_tile, n_tile);
            const int64_t row = (int64_t)m_tile * TC_BM + quad * 32 + lane;
            float *out = D + row * ldd + (int64_t)n_tile * TC_BN + ehalf * ECOLS;
            // chunk sums live in registers, added 16 columns at a time (IEEE round-to-nearest adds); one
            // streaming store per tile
            float accr[ECOLS];
            for (int c = 0; c < n_chunks; c++, wl++) {
                const int acc = wl & 1;
                mbar_wait_bounded(&acc_full[acc], (uint32_t)((wl >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                       (uint32_t)(acc * TC_BN + ehalf * ECOLS);
#pragma unroll
                for (int cc = 0; cc < ECOLS / 16; cc++) {
                    uint32_t v[16];
                    tmem_ld_cols<16>(taddr + cc * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 16; u++)
                        accr[cc * 16 + u] = (c == 0) ? __uint_as_float(v[u]) : accr[cc * 16 + u] + __uint_as_float(v[u]);
                }
                // accumulator drained: hand it back to the MMA issuer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
#pragma unroll
            for (int u = 0; u < ECOLS; u += 4)
                __stcs(reinterpret_cast<float4 *>(out + u),      // streaming: do not pollute L2
                       make_float4(accr[u], accr[u + 1], accr[u + 2], accr[u + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x 192 tile per cluster of two CTAs (two SMs of one TPC).
// Each CTA stages ITS 128 rows of the A operand(s) and ITS 96-row half of the B operand(s); the leader CTA
// (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes
// each CTA's 128 accumulator lanes.  Operand bytes per CTA per k-block: 44 KB (forward hi+lo), 40 KB (backward
// hi+lo), 28 KB (bf16 tiles) instead of 56 / 64 / 40 KB for the same MMA work -- the 1-CTA kernel is fed at
// 107 B/clk/SM in bf16-tiles mode and reaches only 54 % tensor-pipe activity (profiles/r2_ncu_full_bf16_tiles.csv).
// Protocol (per stage s; all barriers live at the same shared-memory offsets in both CTAs):
//   full[s]       leader only, 1 arrival + 2 x stage bytes: both CTAs' TMA loads complete_tx on the LEADER's barrier
//                 (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the barrier address cleared)
//   empty[s]      each CTA, 1 arrival: multicast tcgen05.commit of the leader's MMA thread
//   acc_full[a]   each CTA, 1 arrival: multicast commit after the last MMA of a K chunk
//   acc_empty[a]  leader only, 16 arrivals: the 8 epilogue warps of BOTH CTAs (remote mbarrier.arrive for the peer)
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t TC_PEER_MASK = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address (pair leader)
constexpr int TC_BN_HALF = TC_BN / 2;                 // 96 rows of B per CTA
constexpr int TC_BH_BYTES = TC_BN_HALF * TC_BK * 2;   // 12288

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                                      uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1),
        "l"(policy)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)      // arrives on `bar` of BOTH CTAs of the pair
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar)    // arrive on the pair leader's copy of `bar`
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & TC_PEER_MASK)
                 : "memory");
}

template <int SPLIT_A>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                    float *__restrict__ D, int64_t ldd, int m_tiles, int n_tiles, int num_kb, int k_chunk, int pgroup,
                    int single, int n_stages)
{
    // operands: forward (SPLIT_A)  tmA0 = Wt_hi, tmA1 = Wt_lo (box 128 rows), tmB0 = X (box 96 rows), tmB1 unused
    //           backward           tmA0 = Xt (box 128), tmA1 unused, tmB0 = Rt_hi, tmB1 = Rt_lo (box 96 rows)
    // stage layout per CTA: [A0 16 KB][B0 12 KB][optional: A1 16 KB (forward) | B1 12 KB (backward)]
    constexpr int BYTES2 = SPLIT_A ? TC_A_BYTES : TC_BH_BYTES;
    const int stage_bytes = TC_A_BYTES + TC_BH_BYTES + (single ? 0 : BYTES2);
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem0 = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                                             ~static_cast<uintptr_t>(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem0);
    uint64_t *empty = full + TC_MAX_STAGES;
    uint64_t *acc_full = empty + TC_MAX_STAGES;  // [2]
    uint64_t *acc_empty = acc_full + 2;          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    unsigned char *smem = smem0 + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_pairs = (m_tiles + 1) >> 1;
    const int total_tiles = m_pairs * n_tiles;
    const int n_chunks = (num_kb + k_chunk - 1) / k_chunk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 16);        // 8 epilogue warps of each CTA of the pair
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 rows of A, own 96-row half of B; bytes land on the leader's barrier;
        //       whole warp in the loop, one elected lane issues =====
        const uint64_t keep = l2_policy_evict_last();
        int s = 0;
        uint32_t ph = 0;
        for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
            int m_pair, n_tile;
            decode_tile(tile, m_pairs, n_tiles, pgroup, m_pair, n_tile);
            const int row_a = min(2 * m_pair + (int)rank, m_tiles - 1) * TC_BM;  // odd tile count: the peer recomputes the last tile, unstored
            const int row_b = n_tile * TC_BN + (int)rank * TC_BN_HALF;
            for (int kb = 0; kb < num_kb; kb++) {
                mbar_wait_bounded(&empty[s], ph ^ 1u);
                if (elect_one()) {
                    unsigned char *st = smem + s * stage_bytes;
                    if (leader) mbar_expect_tx(&full[s], (uint32_t)(2 * stage_bytes));
                    if (SPLIT_A) {
                        tma_load_2d_pair_hint(st, &tmA0, kb * TC_BK, row_a, &full[s], keep);
                        tma_load_2d_pair(st + TC_A_BYTES, &tmB0, kb * TC_BK, row_b, &full[s]);
                        if (!single) tma_load_2d_pair_hint(st + TC_A_BYTES + TC_BH_BYTES, &tmA1, kb * TC_BK, row_a, &full[s], keep);
                    } else {
                        tma_load_2d_pair(st, &tmA0, kb * TC_BK, row_a, &full[s]);
                        tma_load_2d_pair(st + TC_A_BYTES, &tmB0, kb * TC_BK, row_b, &full[s]);
                        if (!single) tma_load_2d_pair(st + TC_A_BYTES + TC_BH_BYTES, &tmB1, kb * TC_BK, row_b, &full[s]);
                    }
                }
                __syncwarp();
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 && leader) {
        // ===== MMA issuer (leader CTA only): M = 256 across the pair; converged warp, elected lane issues =====
        constexpr uint32_t idesc = make_idesc_bf16(2 * TC_BM, TC_BN);
        const uint64_t desc0 = make_desc_sw128(smem);
        constexpr uint64_t OFF1 = (uint64_t)(TC_A_BYTES >> 4), OFF2 = (uint64_t)((TC_A_BYTES + TC_BH_BYTES) >> 4);
        int s = 0, wl = 0;
        uint32_t ph = 0;
        for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
            for (int c = 0; c < n_chunks; c++, wl++) {
                const int acc = wl & 1;
                mbar_wait_bounded(&acc_empty[acc], (uint32_t)(((wl >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_BN);
                const int kb0 = c * k_chunk, kb1 = min(num_kb, kb0 + k_chunk);
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait_bounded(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t d0 = desc0 + (uint64_t)((s * stage_bytes) >> 4);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++) {
                            const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);
                            const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
                            umma_bf16_pair(tmem_d, d0 + koff, d0 + OFF1 + koff, idesc, first);
                            if (!single) {
                                if (SPLIT_A) umma_bf16_pair(tmem_d, d0 + OFF2 + koff, d0 + OFF1 + koff, idesc, 1u);     // A_lo * B
                                else umma_bf16_pair(tmem_d, d0 + koff, d0 + OFF2 + koff, idesc, 1u);                     // A * B_lo
                            }
                        }
                        umma_commit_pair(&empty[s]);
                    }
                    __syncwarp();
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) umma_commit_pair(&acc_full[acc]);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): this CTA's 128 accumulator lanes =====
        const int quad = warp & 3;
        const int ehalf = (warp - 4) >> 2;
        constexpr int ECOLS = TC_BN / 2;
        int wl = 0;
        for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
            int m_pair, n_tile;
            decode_tile(tile, m_pairs, n_tiles, pgroup, m_pair, n_tile);
            const int m_tile = 2 * m_pair + (int)rank;
            const int64_t row = (int64_t)m_tile * TC_BM + quad * 32 + lane;
            float *out = D + row * ldd + (int64_t)n_tile * TC_BN + ehalf * ECOLS;
            float accr[ECOLS];
            for (int c = 0; c < n_chunks; c++, wl++) {
                const int acc = wl & 1;
                mbar_wait_bounded(&acc_full[acc], (uint32_t)((wl >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                       (uint32_t)(acc * TC_BN + ehalf * ECOLS);
#pragma unroll
                for (int cc = 0; cc < ECOLS / 16; cc++) {
                    uint32_t v[16];
                    tmem_ld_cols<16>(taddr + cc * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 16; u++)
                        accr[cc * 16 + u] = (c == 0) ? __uint_as_float(v[u]) : accr[cc * 16 + u] + __uint_as_float(v[u]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&acc_empty[acc]);
            }
            if (m_tile < m_tiles) {
#pragma unroll
                for (int u = 0; u < ECOLS; u += 4)
                    __stcs(reinterpret_cast<float4 *>(out + u), make_float4(accr[u], accr[u + 1], accr[u + 2], accr[u + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the leader's MMAs read the peer's shared memory: leave together
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Fused forward: logits GEMM with the softmax / residual epilogue on the accumulator.
//   D[n, (i,a)] = sum_(j,b) X[n,(j,b)] * (Wp_hi + Wp_lo)[(i,a),(j,b)]        sequences on M (one TMEM lane each)
// An N tile is 8 sites in two halves of 4 x 21 + 4 zero columns (UMMA N = 176); each of the 8 epilogue warps owns 32
// sequences x 4 sites, so a thread sees whole 21-state logit vectors of its sequence: +h, softmax, fx,
// residuals, bf16 hi/lo split written transposed (sequence fastest) straight into the operand of the backward
// GEMM.  The 847 MB logits matrix never exists.  Per-(site, 32-sequence group) partials of g_h / fx keep the
// reduction deterministic.
// ---------------------------------------------------------------------------------------------------
constexpr int TF_BN = 176;                         // 8 sites x 21 states + 8 pad
constexpr int TF_SITES = 8;
constexpr int TF_B_BYTES = TF_BN * TC_BK * 2;      // 22528

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_fwd_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_whi,
                    const __grid_constant__ CUtensorMap tm_wlo, const float *__restrict__ h,
                    const uint32_t *__restrict__ msa4, const float *__restrict__ wts,
                    __nv_bfloat16 *__restrict__ Rt_hi, __nv_bfloat16 *__restrict__ Rt_lo, int64_t Kp,
                    float *__restrict__ gh_part, double *__restrict__ fx_part, PlmGeom g, int m_tiles, int n_tiles,
                    int num_kb, int single, int n_stages)
{
    constexpr int Q = 21;                          // states per site of this instantiation (q = 21 or 20 -> S = 21)
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                                            ~static_cast<uintptr_t>(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + TC_MAX_STAGES;
    uint64_t *acc_full = empty + TC_MAX_STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    smem += 1024;                                  // operand ring: [X 16 KB][W_hi 22 KB][W_lo 22 KB (hi+lo mode only)]
    const int stage_bytes = TC_A_BYTES + TF_B_BYTES + (single ? 0 : TF_B_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = m_tiles * n_tiles;
    const int q = g.q;                             // 21, or 20 with the ignored gap (column 20 of a site is then zero)

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 8);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: tiles enumerated with the site tile fastest => concurrent CTAs share X tiles =====
        const uint64_t keep = l2_policy_evict_last();
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_tile = tile % n_tiles, m_tile = tile / n_tiles;
            for (int kb = 0; kb < num_kb; kb++) {
                mbar_wait_bounded(&empty[s], ph ^ 1u);
                unsigned char *st = smem + s * stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                tma_load_2d(st, &tm_x, kb * TC_BK, m_tile * TC_BM, &full[s]);
                tma_load_2d_hint(st + TC_A_BYTES, &tm_whi, kb * TC_BK, n_tile * TF_BN, &full[s], keep);
                if (!single)
                    tma_load_2d_hint(st + TC_A_BYTES + TF_B_BYTES, &tm_wlo, kb * TC_BK, n_tile * TF_BN, &full[s], keep);
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_bf16(TC_BM, TF_BN);
        int s = 0, tl = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tl++) {
            const int acc = tl & 1;
            mbar_wait_bounded(&acc_empty[acc], (uint32_t)(((tl >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TF_BN);
            for (int kb = 0; kb < num_kb; kb++) {
                mbar_wait_bounded(&full[s], ph);
                tc_fence_after();
                unsigned char *st = smem + s * stage_bytes;
                const uint64_t da = make_desc_sw128(st);
                const uint64_t dh = make_desc_sw128(st + TC_A_BYTES);
                const uint64_t dl = make_desc_sw128(st + TC_A_BYTES + TF_B_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; k++) {
                    const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);
                    umma_bf16(tmem_d, da + koff, dh + koff, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    if (!single) umma_bf16(tmem_d, da + koff, dl + koff, idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(&acc_full[acc]);
        }
    } else if (warp >= 4) {
        // ===== fused epilogue: thread = sequence, 4 sites =====
        const int quad = warp & 3;
        const int ehalf = (warp - 4) >> 2;
        const int ntile_part = m_tiles * 4;           // partial slots per site: (sequence tile, quadrant)
        int tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tl++) {
            const int n_tile = tile % n_tiles, m_tile = tile / n_tiles;
            const int acc = tl & 1;
            const int64_t n = (int64_t)m_tile * TC_BM + quad * 32 + lane;
            const int64_t nc = n < g.N ? n : g.N - 1;
            const float wn = n < g.N ? wts[nc] : 0.f;
            mbar_wait_bounded(&acc_full[acc], (uint32_t)((tl >> 1) & 1));
            tc_fence_after();
            uint32_t v[84];
            const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TF_BN + ehalf * 88);
            tmem_ld_cols<32>(t0, v);
            tmem_ld_cols<32>(t0 + 32, v + 32);
            tmem_ld_cols<16>(t0 + 64, v + 64);
            tmem_ld_cols<4>(t0 + 80, v + 80);
            tmem_ld_wait();
            // logits are in registers: the accumulator can be reused by the MMA issuer right away
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int i = n_tile * TF_SITES + ehalf * 4 + s;
                if (i >= g.L) break;                                   // padding sites of the last tile (uniform)
                const int si = (int)((msa4[(int64_t)(i >> 2) * g.Nld + nc] >> (8 * (i & 3))) & 0xffu);
                const float w = si < q ? wn : 0.f;
                float z[Q];
                float mx = -INFINITY;
#pragma unroll
                for (int a = 0; a < Q; a++) {
                    z[a] = (a < q) ? __uint_as_float(v[s * Q + a]) + h[i * q + a] : -INFINITY;
                    mx = fmaxf(mx, z[a]);
                }
                float zs = 0.f, sum = 0.f;
#pragma unroll
                for (int a = 0; a < Q; a++) {
                    if (a == si) zs = z[a];
                    z[a] = (a < q) ? expf(z[a] - mx) : 0.f;
                    sum += z[a];
                }
                const double fx_local = (w == 0.f) ? 0.0 : -((double)w * (double)(zs - mx - logf(sum)));
                const float inv = w / sum;
                float *ghp = gh_part + ((int64_t)i * ntile_part + m_tile * 4 + quad) * g.S;
#pragma unroll
                for (int a = 0; a < Q; a++) {
                    const float r = z[a] * inv - (a == si ? w : 0.f);
                    if (a < q) {
                        if (n < g.N) {
                            const int64_t off = ((int64_t)i * q + a) * Kp + n;
                            const __nv_bfloat16 hi = __float2bfloat16_rn(r);
                            Rt_hi[off] = hi;
                            if (!single) Rt_lo[off] = __float2bfloat16_rn(r - __bfloat162float(hi));
                        }
                    }
                    const float tot = warp_sum(r);
                    if (lane == 0) ghp[a] = (a < q) ? tot : 0.f;
                }
                const double fw = warp_sum(fx_local);
                if (lane == 0) fx_part[(int64_t)i * ntile_part + m_tile * 4 + quad] = fw;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// expand for the fused forward: rows regrouped as [site tile][8 sites x q states (+ zero pad to 176)]
__global__ void expand_tcf_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ Wp_hi,
                                  __nv_bfloat16 *__restrict__ Wp_lo, int L, int q, int64_t ldw, int single)
{
    const int i = blockIdx.y, j = blockIdx.x;
    if (j <= i) return;
    const float *J = x + (int64_t)L * q + ((int64_t)i * (2 * L - i - 1) / 2 + (j - i - 1)) * q * q;
    // padded row base of a site: tile of 8 sites = two halves of 88 rows (4 sites x 21 states + 4 zero rows)
    const int64_t ri = (int64_t)(i / TF_SITES) * TF_BN + ((i % TF_SITES) / 4) * 88 + (i % 4) * 21;
    const int64_t rj = (int64_t)(j / TF_SITES) * TF_BN + ((j % TF_SITES) / 4) * 88 + (j % 4) * 21;
    for (int e = threadIdx.x; e < q * q; e += blockDim.x) {
        const int a = e / q, b = e - a * q;
        const float v = J[e];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int64_t p1 = (ri + a) * ldw + (j * q + b);        // row (i,a), K index (j,b)
        const int64_t p2 = (rj + b) * ldw + (i * q + a);        // row (j,b), K index (i,a)
        Wp_hi[p1] = hi;
        Wp_hi[p2] = hi;
        if (!single) { Wp_lo[p1] = lo; Wp_lo[p2] = lo; }
    }
}

// expand for the tensor-core forward: Wt[(i,a)][(j,b)] = J_ij(a,b) as bf16 hi + lo, both orientations
__global__ void expand_tc_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ Wt_hi,
                                 __nv_bfloat16 *__restrict__ Wt_lo, int L, int q, int64_t ldw, int single)
{
    const int i = blockIdx.y, j = blockIdx.x;
    if (j <= i) return;
    const float *J = x + (int64_t)L * q + ((int64_t)i * (2 * L - i - 1) / 2 + (j - i - 1)) * q * q;
    for (int e = threadIdx.x; e < q * q; e += blockDim.x) {
        const int a = e / q, b = e - a * q;
        const float v = J[e];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int64_t p1 = (int64_t)(i * q + a) * ldw + (j * q + b);
        const int64_t p2 = (int64_t)(j * q + b) * ldw + (i * q + a);
        Wt_hi[p1] = hi;
        Wt_hi[p2] = hi;
        if (!single) { Wt_lo[p1] = lo; Wt_lo[p2] = lo; }
    }
}

// one-hot operand of the forward product: X[n][(j,b)], K = (j,b) fastest
__global__ void build_x_kernel(const uint32_t *__restrict__ msa4, __nv_bfloat16 *__restrict__ X, int64_t N,
                               int64_t Nld, int L, int q, int64_t ldx)
{
    const int64_t n = blockIdx.x;
    if (n >= N) return;
    for (int e = threadIdx.x; e < L * q; e += blockDim.x) {
        const int j = e / q, b = e - j * q;
        const int code = (int)((msa4[(int64_t)(j >> 2) * Nld + n] >> (8 * (j & 3))) & 0xffu);
        X[n * ldx + e] = __float2bfloat16(code == b ? 1.0f : 0.0f);
    }
}

// softmax + residuals from the logits Zt[(i,a)][n]: thread = two adjacent sequences (8-byte loads of the
// logits, 4-byte bf16x2 stores of the residuals), CTA = 128 threads = 256 sequences of one site.
// ONEHOT = true: the "residual" is w_n [s_ni = a] (no logits read) -- the operand of the weighted pair counts
// f_ij = sum_n w_n [s_ni = a][s_nj = b] computed by the same tensor-core backward product (row a6).
template <int Q, bool ONEHOT>
__global__ void __launch_bounds__(128)
plm_softmax_kernel(const float *__restrict__ Zt, int64_t ldz, const float *__restrict__ h,
                   const uint32_t *__restrict__ msa4, const float *__restrict__ wts,
                   __nv_bfloat16 *__restrict__ Rt_hi, __nv_bfloat16 *__restrict__ Rt_lo, int64_t Kp,
                   float *__restrict__ gh_part, double *__restrict__ fx_part, PlmGeom g, int ntiles)
{
    __shared__ float s_gh[4 * 32];
    __shared__ double s_fx[4];
    const int tile = blockIdx.x, i = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t N = g.N;
    const int64_t n0 = (int64_t)tile * 256 + 2 * tid;
    // clamped, even load position (rows of Zt / msa4 are padded beyond N, see plm_tcf_geometry / plm_pack_msa)
    const int64_t m0 = n0 < N ? n0 : ((N - 1) & ~(int64_t)1);
    const uint2 wi = *reinterpret_cast<const uint2 *>(msa4 + (int64_t)(i >> 2) * g.Nld + m0);
    const int sh = 8 * (i & 3);
    const int si[2] = {(int)((wi.x >> sh) & 0xffu), (int)((wi.y >> sh) & 0xffu)};
    float w[2];
    w[0] = (n0 < N && si[0] < Q) ? wts[n0] : 0.f;
    w[1] = (n0 + 1 < N && si[1] < Q) ? wts[n0 + 1] : 0.f;
    float z[2][Q];
    double fx_local = 0.0;
    if (ONEHOT) {
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int a = 0; a < Q; a++) z[k][a] = (a == si[k]) ? w[k] : 0.f;
    } else {
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int a = 0; a < Q; a++) {
            const float2 v = *reinterpret_cast<const float2 *>(Zt + ((int64_t)i * Q + a) * ldz + m0);
            const float ha = h[i * Q + a];
            z[0][a] = v.x + ha;
            z[1][a] = v.y + ha;
            mx[0] = fmaxf(mx[0], z[0][a]);
            mx[1] = fmaxf(mx[1], z[1][a]);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float zs = 0.f, sum = 0.f;
#pragma unroll
            for (int a = 0; a < Q; a++) {
                if (a == si[k]) zs = z[k][a];
                z[k][a] = expf(z[k][a] - mx[k]);
                sum += z[k][a];
            }
            if (w[k] != 0.f) fx_local -= (double)w[k] * (double)(zs - mx[k] - logf(sum));
            const float inv = w[k] / sum;
#pragma unroll
            for (int a = 0; a < Q; a++) z[k][a] = z[k][a] * inv - (a == si[k] ? w[k] : 0.f);
        }
    }
#pragma unroll
    for (int a = 0; a < Q; a++) {
        if (n0 < N) {
            // n0 is even and Kp is a multiple of 64: the pair (n0, n0 + 1) is 4-byte aligned and inside the row;
            // a sequence beyond N has weight 0, i.e. writes an exact zero into the K padding
            const int64_t off = ((int64_t)i * Q + a) * Kp + n0;
            const __nv_bfloat162 hi = __floats2bfloat162_rn(z[0][a], z[1][a]);
            *reinterpret_cast<__nv_bfloat162 *>(Rt_hi + off) = hi;
            if (Rt_lo != nullptr)
                *reinterpret_cast<__nv_bfloat162 *>(Rt_lo + off) =
                    __floats2bfloat162_rn(z[0][a] - __low2float(hi), z[1][a] - __high2float(hi));
        }
        const float v = warp_sum(z[0][a] + z[1][a]);
        if (lane == 0) s_gh[warp * 32 + a] = v;
    }
    const double fw = warp_sum(fx_local);
    if (lane == 0) s_fx[warp] = fw;
    __syncthreads();
    if (tid < g.S) {
        float tot = 0.f;
        if (tid < Q)
            for (int ww = 0; ww < 4; ww++) tot += s_gh[ww * 32 + tid];
        gh_part[((int64_t)i * ntiles + tile) * g.S + tid] = tot;
    }
    if (tid == 0) fx_part[(int64_t)i * ntiles + tile] = (s_fx[0] + s_fx[1]) + (s_fx[2] + s_fx[3]);
}

// ---- one-hot operand (static per MSA) ----------------------------------------------------------------
__global__ void build_xt_kernel(const uint32_t *__restrict__ msa4, __nv_bfloat16 *__restrict__ Xt, int64_t N,
                                int64_t Nld, int64_t Kp, int L, int q)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (n >= Kp) return;
    int code = 255;
    if (n < N) code = (int)((msa4[(int64_t)(j >> 2) * Nld + n] >> (8 * (j & 3))) & 0xffu);
    for (int b = 0; b < q; b++)
        Xt[((int64_t)j * q + b) * Kp + n] = __float2bfloat16(code == b ? 1.0f : 0.0f);
}

// g_J(i<j)[a][b] = scale * (Gd[(j,b),(i,a)] + Gd[(i,a),(j,b)])
__global__ void finalize_pairs_tc_kernel(const float *__restrict__ Gd, float *__restrict__ gJ, int L, int q,
                                         int Np, float scale)
{
    const int i = blockIdx.y, j = blockIdx.x;
    if (j <= i) return;
    float *out = gJ + ((int64_t)i * (2 * L - i - 1) / 2 + (j - i - 1)) * q * q;
    for (int e = threadIdx.x; e < q * q; e += blockDim.x) {
        const int a = e / q, b = e - a * q;
        const float v1 = Gd[(int64_t)(j * q + b) * Np + (i * q + a)];
        const float v2 = Gd[(int64_t)(i * q + a) * Np + (j * q + b)];
        out[e] = scale * (v1 + v2);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// tuning hooks for parameter sweeps (read once; the defaults below are what the product uses)
static int env_int_once(const char *name, int *cache)
{
    if (*cache == -2) {
        const char *e = getenv(name);
        *cache = e ? atoi(e) : -1;
    }
    return *cache;
}

static int sm_count_current()
{
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!cache[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = n > 0 ? n : 148;
    }
    return cache[dev];
}

// ring depth for a given stage size: as many stages as fit in the 227 KB opt-in shared memory
static int stages_for(int stage_bytes)
{
    return std::max(2, std::min(TC_MAX_STAGES, (TC_SMEM_LIMIT - TC_SMEM_HEAD) / stage_bytes));
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !p)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

static int make_map(CUtensorMap *m, void *base, int64_t rows, int64_t kp, int box_rows)
{
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
    cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return 1; }
    return 0;
}

void plm_tc_geometry(const PlmGeom &g, PlmTcGeom &t)
{
    const int64_t lq = (int64_t)g.L * g.q;
    t.Mp = round_up(lq, TC_BM);
    t.Np = round_up(lq, TC_BN);
    t.Kp = round_up(g.N, TC_BK);
}

int plm_tc_build_xt(const PlmGeom &g, const PlmTcGeom &t, const uint32_t *d_msa4, void *d_xt, cudaStream_t st)
{
    EVC_CUDA(cudaMemsetAsync(d_xt, 0, (size_t)t.Mp * t.Kp * 2, st));
    dim3 grid((unsigned)ceil_div(t.Kp, 256), (unsigned)g.L);
    build_xt_kernel<<<grid, 256, 0, st>>>(d_msa4, reinterpret_cast<__nv_bfloat16 *>(d_xt), g.N, g.Nld, t.Kp, g.L,
                                          g.q);
    EVC_KERNEL_CHECK();
    return 0;
}

int plm_tc_make_maps(const PlmTcGeom &t, void *d_xt, void *d_rt_hi, void *d_rt_lo, void *maps_out)
{
    CUtensorMap *m = reinterpret_cast<CUtensorMap *>(maps_out);
    if (make_map(&m[0], d_xt, t.Mp, t.Kp, TC_BM)) return 1;
    if (make_map(&m[1], d_rt_hi, t.Np, t.Kp, TC_BN)) return 1;
    if (make_map(&m[2], d_rt_lo, t.Np, t.Kp, TC_BN)) return 1;
    // CTA-pair kernel: each CTA loads a 96-row half of the B tile
    if (make_map(&m[3], d_rt_hi, t.Np, t.Kp, TC_BN_HALF)) return 1;
    if (make_map(&m[4], d_rt_lo, t.Np, t.Kp, TC_BN_HALF)) return 1;
    return 0;
}

// 1 = the tensor loads of a stage are issued by two producer threads (warp 0: A, warp 3: B)
static int split_producer()
{
    static int sp_env = -2;
    const int e = env_int_once("EVC_SPLIT_PRODUCER", &sp_env);
    return e >= 0 ? e : TC_SPLIT_PRODUCER_DEFAULT;
}

// 1 = cta_group::2 tiles (256 x 192 per CTA pair), 0 = one CTA per 128 x 192 tile
static int pair_mode()
{
    static int pm_env = -2;
    const int e = env_int_once("EVC_TC_PAIR", &pm_env);
    return e >= 0 ? e : TC_PAIR_DEFAULT;
}

int plm_tc_backward(const PlmGeom &g, const PlmTcGeom &t, const void *maps, float *d_Gd, int single, cudaStream_t st)
{
    const CUtensorMap *m = reinterpret_cast<const CUtensorMap *>(maps);
    const int stage = TC_A_BYTES + TC_B_BYTES + (single ? 0 : TC_B_BYTES);
    const int n_stages = stages_for(stage);
    const size_t smem = (size_t)n_stages * stage + TC_SMEM_HEAD;
    static int kc_env = -2;
    const int kc = env_int_once("EVC_KCHUNK", &kc_env);
    const int m_tiles = (int)(t.Mp / TC_BM), n_tiles = (int)(t.Np / TC_BN);
    if (pair_mode()) {
        const int pstage = TC_A_BYTES + TC_BH_BYTES + (single ? 0 : TC_BH_BYTES);
        const int pstages = stages_for(pstage);
        const size_t psmem = (size_t)pstages * pstage + TC_SMEM_HEAD;
        EVC_CUDA(cudaFuncSetAttribute(tc_gemm_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        const int m_pairs = (m_tiles + 1) / 2;
        const int pairs = std::min(sm_count_current() / 2, m_pairs * n_tiles);
        tc_gemm_pair_kernel<0><<<2 * pairs, TC_THREADS, psmem, st>>>(m[0], m[0], m[3], m[4], d_Gd, t.Np, m_tiles, n_tiles,
                                                                   (int)(t.Kp / TC_BK), kc > 0 ? kc : TC_K_CHUNK, m_pairs,
                                                                   single, pstages);
        EVC_KERNEL_CHECK();
        return 0;
    }
    const int grid = std::min(sm_count_current(), m_tiles * n_tiles);
    if (single) {
        EVC_CUDA(cudaFuncSetAttribute(tc_gemm_persistent_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        tc_gemm_persistent_kernel<0, 1><<<grid, TC_THREADS, smem, st>>>(m[0], m[1], m[2], d_Gd, t.Np, m_tiles, n_tiles,
                                                                       (int)(t.Kp / TC_BK), kc > 0 ? kc : TC_K_CHUNK, m_tiles,
                                                                       n_stages, split_producer());
    } else {
        EVC_CUDA(cudaFuncSetAttribute(tc_gemm_persistent_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        tc_gemm_persistent_kernel<0, 0><<<grid, TC_THREADS, smem, st>>>(m[0], m[1], m[2], d_Gd, t.Np, m_tiles, n_tiles,
                                                                       (int)(t.Kp / TC_BK), kc > 0 ? kc : TC_K_CHUNK, m_tiles,
                                                                       n_stages, split_producer());
    }
    EVC_KERNEL_CHECK();
    return 0;
}

int plm_tc_finalize_pairs(const PlmGeom &g, const PlmTcGeom &t, const float *d_Gd, float *d_gJ, float scale,
                          cudaStream_t st)
{
    dim3 grid((unsigned)g.L, (unsigned)g.L);
    finalize_pairs_tc_kernel<<<grid, 128, 0, st>>>(d_Gd, d_gJ, g.L, g.q, (int)t.Np, scale);
    EVC_KERNEL_CHECK();
    return 0;
}

// ---- tensor-core forward -----------------------------------------------------------------------------
void plm_tcf_geometry(const PlmGeom &g, PlmTcfGeom &t)
{
    const int64_t lq = (int64_t)g.L * g.q;
    t.Mp = round_up(lq, TC_BM);          // rows of Wt / Zt
    t.Kw = round_up(lq, TC_BK);          // K extent (j,b)
    t.Ns = round_up(g.N, TC_BN);         // sequences rounded to the 192-column tile
    t.Xrows = round_up(g.N, 384);        // allocation of X: covers 128- and 192-row tilings
    t.ntiles_s = (int)ceil_div(g.N, 256);
}

int plm_tcf_build_x(const PlmGeom &g, const PlmTcfGeom &t, const uint32_t *d_msa4, void *d_x1h, cudaStream_t st)
{
    EVC_CUDA(cudaMemsetAsync(d_x1h, 0, (size_t)t.Xrows * t.Kw * 2, st));
    build_x_kernel<<<(unsigned)g.N, 256, 0, st>>>(d_msa4, reinterpret_cast<__nv_bfloat16 *>(d_x1h), g.N, g.Nld, g.L,
                                                g.q, t.Kw);
    EVC_KERNEL_CHECK();
    return 0;
}

int plm_tcf_make_maps(const PlmTcfGeom &t, void *d_wt_hi, void *d_wt_lo, void *d_x1h, void *maps_out)
{
    CUtensorMap *m = reinterpret_cast<CUtensorMap *>(maps_out);
    if (make_map(&m[0], d_wt_hi, t.Mp, t.Kw, TC_BM)) return 1;
    if (make_map(&m[1], d_wt_lo, t.Mp, t.Kw, TC_BM)) return 1;
    if (make_map(&m[2], d_x1h, t.Xrows, t.Kw, TC_BN)) return 1;
    if (make_map(&m[3], d_x1h, t.Xrows, t.Kw, TC_BN_HALF)) return 1;      // CTA-pair kernel: 96-row half of the X tile
    return 0;
}

int plm_tcf_expand(const PlmGeom &g, const PlmTcfGeom &t, const float *d_x, void *d_wt_hi, void *d_wt_lo,
                   int single, cudaStream_t st)
{
    dim3 grid((unsigned)g.L, (unsigned)g.L);
    expand_tc_kernel<<<grid, 128, 0, st>>>(d_x, reinterpret_cast<__nv_bfloat16 *>(d_wt_hi),
                                          reinterpret_cast<__nv_bfloat16 *>(d_wt_lo), g.L, g.q, t.Kw, single);
    EVC_KERNEL_CHECK();
    return 0;
}

// M tiles per group of the forward tile order: the group's slice of the coupling operand (hi [+ lo], all of
// K) should stay L2-resident while every sequence tile passes by.  24 MB per group measured best at config 2
// (71 MB operand); for long alignments (L = 500: 441 MB, L = 800: 1.13 GB) the same byte budget
// gives groups of 4 / 2 M tiles instead of the fixed 11 that round 1 used.
static int forward_mgroup(const PlmTcfGeom &t, int single, int m_tiles)
{
    static int mg_env = -2;
    const int e = env_int_once("EVC_MGROUP", &mg_env);
    if (e > 0) return std::min(e, m_tiles);
    static int mb_env = -2;
    const int mb = env_int_once("EVC_MGROUP_MB", &mb_env);
    const double budget = (mb > 0 ? mb : 24) * 1.0e6;
    const double per_tile = (double)TC_BM * (double)t.Kw * 2.0 * (single ? 1.0 : 2.0);
    return std::max(1, std::min(m_tiles, (int)(budget / per_tile)));
}

int plm_tcf_logits(const PlmGeom &g, const PlmTcfGeom &t, const void *maps, float *d_zt, int single, cudaStream_t st)
{
    const CUtensorMap *m = reinterpret_cast<const CUtensorMap *>(maps);
    const int stage = TC_A_BYTES + TC_B_BYTES + (single ? 0 : TC_A_BYTES);
    const int n_stages = stages_for(stage);
    const size_t smem = (size_t)n_stages * stage + TC_SMEM_HEAD;
    const int m_tiles = (int)(t.Mp / TC_BM), n_tiles = (int)(t.Ns / TC_BN);
    const int num_kb = (int)(t.Kw / TC_BK);
    if (pair_mode()) {
        const int pstage = TC_A_BYTES + TC_BH_BYTES + (single ? 0 : TC_A_BYTES);
        const int pstages = stages_for(pstage);
        const size_t psmem = (size_t)pstages * pstage + TC_SMEM_HEAD;
        EVC_CUDA(cudaFuncSetAttribute(tc_gemm_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        const int m_pairs = (m_tiles + 1) / 2;
        const int pairs = std::min(sm_count_current() / 2, m_pairs * n_tiles);
        const int pgroup = std::max(1, std::min(m_pairs, (forward_mgroup(t, single, m_tiles) + 1) / 2));
        tc_gemm_pair_kernel<1><<<2 * pairs, TC_THREADS, psmem, st>>>(m[0], m[1], m[3], m[3], d_zt, t.Ns, m_tiles, n_tiles,
                                                                   num_kb, num_kb <= 128 ? num_kb : TC_K_CHUNK, pgroup,
                                                                   single, pstages);
        EVC_KERNEL_CHECK();
        return 0;
    }
    const int grid = std::min(sm_count_current(), m_tiles * n_tiles);
    const int kchunk = num_kb <= 128 ? num_kb : TC_K_CHUNK;
    const int mgroup = forward_mgroup(t, single, m_tiles);
    if (single) {
        EVC_CUDA(cudaFuncSetAttribute(tc_gemm_persistent_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        tc_gemm_persistent_kernel<1, 1><<<grid, TC_THREADS, smem, st>>>(m[0], m[1], m[2], d_zt, t.Ns, m_tiles, n_tiles, num_kb,
                                                                       kchunk, mgroup, n_stages, split_producer());
    } else {
        EVC_CUDA(cudaFuncSetAttribute(tc_gemm_persistent_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
        tc_gemm_persistent_kernel<1, 0><<<grid, TC_THREADS, smem, st>>>(m[0], m[1], m[2], d_zt, t.Ns, m_tiles, n_tiles, num_kb,
                                                                       kchunk, mgroup, n_stages, split_producer());
    }
    EVC_KERNEL_CHECK();
    return 0;
}

// softmax / residual kernel; d_rt_lo == nullptr in the bf16x1 precision mode (no lo operand is written)
template <bool ONEHOT>
static int launch_softmax(const PlmGeom &g, int ntiles, const float *d_zt, int64_t ldz, const float *d_x,
                          const uint32_t *d_msa4, const float *d_wts, __nv_bfloat16 *hi, __nv_bfloat16 *lo,
                          int64_t Kp, float *d_gh_part, double *d_fx_part, cudaStream_t st)
{
    dim3 grid((unsigned)ntiles, (unsigned)g.L);
    switch (g.q) {
        case 21: plm_softmax_kernel<21, ONEHOT><<<grid, 128, 0, st>>>(d_zt, ldz, d_x, d_msa4, d_wts, hi, lo, Kp, d_gh_part, d_fx_part, g, ntiles); break;
        case 20: plm_softmax_kernel<20, ONEHOT><<<grid, 128, 0, st>>>(d_zt, ldz, d_x, d_msa4, d_wts, hi, lo, Kp, d_gh_part, d_fx_part, g, ntiles); break;
        case 5: plm_softmax_kernel<5, ONEHOT><<<grid, 128, 0, st>>>(d_zt, ldz, d_x, d_msa4, d_wts, hi, lo, Kp, d_gh_part, d_fx_part, g, ntiles); break;
        case 4: plm_softmax_kernel<4, ONEHOT><<<grid, 128, 0, st>>>(d_zt, ldz, d_x, d_msa4, d_wts, hi, lo, Kp, d_gh_part, d_fx_part, g, ntiles); break;
        default: set_error("plm softmax kernel: unsupported q"); return 1;
    }
    EVC_KERNEL_CHECK();
    return 0;
}

int plm_tcf_softmax(const PlmGeom &g, const PlmTcfGeom &t, const float *d_zt, const float *d_x,
                    const uint32_t *d_msa4, const float *d_wts, void *d_rt_hi, void *d_rt_lo, int64_t Kp,
                    float *d_gh_part, double *d_fx_part, cudaStream_t st)
{
    return launch_softmax<false>(g, t.ntiles_s, d_zt, t.Ns, d_x, d_msa4, d_wts,
                                 reinterpret_cast<__nv_bfloat16 *>(d_rt_hi), reinterpret_cast<__nv_bfloat16 *>(d_rt_lo),
                                 Kp, d_gh_part, d_fx_part, st);
}

// a6 on the tensor cores: Rt = w_n [s_ni = a] as bf16 hi + lo (the weights keep 16 mantissa bits), per-tile
// partials of f_i; the caller then runs the backward product and symmetrises with scale 0.5
int plm_tc_onehot_residual(const PlmGeom &g, int ntiles, const uint32_t *d_msa4, const float *d_wts, void *d_rt_hi,
                           void *d_rt_lo, int64_t Kp, float *d_gh_part, double *d_fx_part, cudaStream_t st)
{
    return launch_softmax<true>(g, ntiles, nullptr, 0, nullptr, d_msa4, d_wts,
                                reinterpret_cast<__nv_bfloat16 *>(d_rt_hi), reinterpret_cast<__nv_bfloat16 *>(d_rt_lo),
                                Kp, d_gh_part, d_fx_part, st);
}

// ---- fused tensor-core forward ------------------------------------------------------------------------
void plm_tcff_geometry(const PlmGeom &g, PlmTcffGeom &t)
{
    t.n_tiles = (int)ceil_div(g.L, TF_SITES);
    t.Np = (int64_t)t.n_tiles * TF_BN;
    t.Kw = round_up((int64_t)g.L * g.q, TC_BK);
    t.m_tiles = (int)ceil_div(g.N, TC_BM);
    t.Xrows = round_up(g.N, 384);                  // covers both the 128-row and the 192-row tilings of X
    t.ntile_part = t.m_tiles * 4;
}

bool plm_tcff_supported(const PlmGeom &g) { return g.S == 21; }

int plm_tcff_make_maps(const PlmTcffGeom &t, void *d_x1h, void *d_wp_hi, void *d_wp_lo, void *maps_out)
{
    CUtensorMap *m = reinterpret_cast<CUtensorMap *>(maps_out);
    if (make_map(&m[0], d_x1h, t.Xrows, t.Kw, TC_BM)) return 1;
    if (make_map(&m[1], d_wp_hi, t.Np, t.Kw, TF_BN)) return 1;
    if (make_map(&m[2], d_wp_lo, t.Np, t.Kw, TF_BN)) return 1;
    return 0;
}

int plm_tcff_expand(const PlmGeom &g, const PlmTcffGeom &t, const float *d_x, void *d_wp_hi, void *d_wp_lo,
                    int single, cudaStream_t st)
{
    dim3 grid((unsigned)g.L, (unsigned)g.L);
    expand_tcf_kernel<<<grid, 128, 0, st>>>(d_x, reinterpret_cast<__nv_bfloat16 *>(d_wp_hi),
                                           reinterpret_cast<__nv_bfloat16 *>(d_wp_lo), g.L, g.q, t.Kw, single);
    EVC_KERNEL_CHECK();
    return 0;
}

int plm_tcff_forward(const PlmGeom &g, const PlmTcffGeom &t, const void *maps, const float *d_x,
                     const uint32_t *d_msa4, const float *d_wts, void *d_rt_hi, void *d_rt_lo, int64_t Kp,
                     float *d_gh_part, double *d_fx_part, int single, cudaStream_t st)
{
    const CUtensorMap *m = reinterpret_cast<const CUtensorMap *>(maps);
    const int stage = TC_A_BYTES + TF_B_BYTES + (single ? 0 : TF_B_BYTES);
    const int n_stages = stages_for(stage);
    const size_t smem = (size_t)n_stages * stage + TC_SMEM_HEAD;
    EVC_CUDA(cudaFuncSetAttribute(tc_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    const int grid = std::min(sm_count_current(), t.m_tiles * t.n_tiles);
    tc_fwd_fused_kernel<<<grid, TC_THREADS, smem, st>>>(m[0], m[1], m[2], d_x, d_msa4, d_wts,
                                                        reinterpret_cast<__nv_bfloat16 *>(d_rt_hi),
                                                        reinterpret_cast<__nv_bfloat16 *>(d_rt_lo), Kp, d_gh_part,
                                                        d_fx_part, g, t.m_tiles, t.n_tiles, (int)(t.Kw / TC_BK), single,
                                                        n_stages);
    EVC_KERNEL_CHECK();
    return 0;
}

size_t plm_tc_map_bytes() { return 6 * sizeof(CUtensorMap); }

}  // namespace evc
