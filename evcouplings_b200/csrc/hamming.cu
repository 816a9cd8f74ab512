// Hot path (b): O(N^2 L) pairwise-Hamming sequence reweighting on sm_100a.
//
// Replaces plmc's reweighting pass and the in-tree numba twin
// evcouplings/align/alignment.py:1192-1233 (num_cluster_members): for every
// sequence s, count the sequences t (self included) with at least
// `min_identical` identical positions (gap == gap is an identity).
//
// Data layout: the uint8 code matrix (codes < 32) is transposed once into five
// bit-planes  planes[p][w][n]  (bit k of word w = bit p of the code at site
// 32w+k, n fastest so a warp reads 128 contiguous bytes).  Two sequences agree
// at a site iff all five plane bits agree, so one 32-site word of a pair costs
// 5 LOP3 + 1 POPC + 1 IADD -- ~6.6x fewer instructions than byte compares and
// integer-exact.  Padded sites (32*W - L) are zero in every sequence and are
// accounted for by raising the threshold.
//
// Tiling: 128 x 128 pair tiles over the upper triangle (each unordered pair
// visited once; a tile credits both its rows and its columns), 256 threads,
// 8x8 pair counters per thread in registers, plane words staged in shared
// memory 16 words (512 sites) at a time.  The plane buffer (N * 5 * W * 4 bytes; 40 MB at
// N=200k, L=300) is L2-resident, so the kernel is bound by the integer pipes.
#include "common.cuh"
#include "internal.h"

namespace evc {

constexpr int HP = 5;        // bit planes (codes < 32)
constexpr int HT = 128;      // pair-tile edge
constexpr int HWC = 16;      // words staged per step (80 KB of dynamic shared memory => 2 CTAs / SM)

__global__ void hamming_pack_kernel(const uint8_t *__restrict__ codes, int64_t N, int L, int W,
                                    uint32_t *__restrict__ planes)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (n >= N) return;
    uint32_t pl[HP] = {0, 0, 0, 0, 0};
    const uint8_t *row = codes + n * L;
    const int k0 = w * 32;
#pragma unroll 4
    for (int k = 0; k < 32; k++) {
        const int site = k0 + k;
        const uint32_t c = site < L ? row[site] : 0u;
#pragma unroll
        for (int p = 0; p < HP; p++) pl[p] |= ((c >> p) & 1u) << k;
    }
#pragma unroll
    for (int p = 0; p < HP; p++) planes[((int64_t)p * W + w) * N + n] = pl[p];
}

__device__ __forceinline__ void tile_from_index(int64_t idx, int64_t T, int64_t &R, int64_t &C)
{
    // idx enumerates (R, C >= R) row-major: offset(R) = R*T - R(R-1)/2
    double t = (double)(2 * T + 1);
    int64_t r = (int64_t)floor((t - sqrt(t * t - 8.0 * (double)idx)) * 0.5);
    if (r < 0) r = 0;
    if (r > T - 1) r = T - 1;
    while (r > 0 && r * T - r * (r - 1) / 2 > idx) r--;
    while ((r + 1) * T - (r + 1) * r / 2 <= idx) r++;
    R = r;
    C = r + (idx - (r * T - r * (r - 1) / 2));
}

__global__ void __launch_bounds__(256, 2)
hamming_tile_kernel(const uint32_t *__restrict__ planes, int64_t N, int W, int thr,
                    int64_t tile_begin, int64_t T, int *__restrict__ counts)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t (*s_row)[HP][HT] = reinterpret_cast<uint32_t (*)[HP][HT]>(s_dyn);
    uint32_t (*s_col)[HP][HT] = reinterpret_cast<uint32_t (*)[HP][HT]>(s_dyn + HWC * HP * HT);
    __shared__ int s_rsum[HT];
    __shared__ int s_csum[HT];

    int64_t R, C;
    tile_from_index(tile_begin + blockIdx.x, T, R, C);
    const int64_t row0 = R * HT, col0 = C * HT;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    int cnt[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) cnt[r][c] = 0;
    if (tid < HT) { s_rsum[tid] = 0; s_csum[tid] = 0; }

    // Exact early termination at warp granularity: after word w a pair can still gain at most 32 * (W - 1 - w)
    // identities; a warp (16 x 128 pairs) whose pairs can no longer reach the threshold stops comparing (it only
    // keeps arriving at the staging barriers), and the tile ends when all of its warps are done.
    bool wdead = false;
    for (int w0 = 0; w0 < W; w0 += HWC) {
        const int nw = min(HWC, W - w0);
        if (!__syncthreads_or(!wdead)) break;
        // stage nw words x 5 planes x 128 sequences for both sides
        for (int e = tid; e < HWC * HP * HT; e += 256) {
            const int s = e & (HT - 1);
            const int p = (e / HT) % HP;
            const int ww = e / (HT * HP);
            uint32_t vr = 0, vc = 0;
            if (ww < nw) {
                const int64_t base = ((int64_t)p * W + (w0 + ww)) * N;
                if (row0 + s < N) vr = planes[base + row0 + s];
                if (col0 + s < N) vc = planes[base + col0 + s];
            }
            s_row[ww][p][s] = vr;
            s_col[ww][p][s] = vc;
        }
        __syncthreads();
        for (int ww = 0; ww < nw && !wdead; ww++) {
            uint32_t a[HP][8], b[HP][8];
#pragma unroll
            for (int p = 0; p < HP; p++) {
                const uint4 a0 = *reinterpret_cast<const uint4 *>(&s_row[ww][p][ty * 8]);
                const uint4 a1 = *reinterpret_cast<const uint4 *>(&s_row[ww][p][ty * 8 + 4]);
                const uint4 b0 = *reinterpret_cast<const uint4 *>(&s_col[ww][p][tx * 8]);
                const uint4 b1 = *reinterpret_cast<const uint4 *>(&s_col[ww][p][tx * 8 + 4]);
                a[p][0] = a0.x; a[p][1] = a0.y; a[p][2] = a0.z; a[p][3] = a0.w;
                a[p][4] = a1.x; a[p][5] = a1.y; a[p][6] = a1.z; a[p][7] = a1.w;
                b[p][0] = b0.x; b[p][1] = b0.y; b[p][2] = b0.z; b[p][3] = b0.w;
                b[p][4] = b1.x; b[p][5] = b1.y; b[p][6] = b1.z; b[p][7] = b1.w;
            }
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t d = a[0][r] ^ b[0][c];
#pragma unroll
                    for (int p = 1; p < HP; p++) d |= a[p][r] ^ b[p][c];
                    cnt[r][c] += __popc(~d);
                }
            const int wdone = w0 + ww;
            if (wdone >= 1 && wdone + 1 < W) {
                const int need = thr - 32 * (W - 1 - wdone);
                int alive = 0;
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) alive |= (cnt[r][c] >= need);
                if (!__any_sync(0xffffffffu, alive)) wdead = true;
            }
        }
    }

    // threshold -> neighbour flags; credit rows (always) and columns (off-diagonal tiles)
    const bool diag = (R == C);
    int rs[8], cs[8];
#pragma unroll
    for (int r = 0; r < 8; r++) rs[r] = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) cs[c] = 0;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const bool ok = (row0 + ty * 8 + r < N) && (col0 + tx * 8 + c < N);
            const int f = (ok && cnt[r][c] >= thr) ? 1 : 0;
            rs[r] += f;
            cs[c] += f;
        }
#pragma unroll
    for (int r = 0; r < 8; r++)
        if (rs[r]) atomicAdd(&s_rsum[ty * 8 + r], rs[r]);
    if (!diag) {
#pragma unroll
        for (int c = 0; c < 8; c++)
            if (cs[c]) atomicAdd(&s_csum[tx * 8 + c], cs[c]);
    }
    __syncthreads();
    if (tid < HT) {
        if (row0 + tid < N && s_rsum[tid]) atomicAdd(&counts[row0 + tid], s_rsum[tid]);
    } else if (tid < 2 * HT && !diag) {
        const int c = tid - HT;
        if (col0 + c < N && s_csum[c]) atomicAdd(&counts[col0 + c], s_csum[c]);
    }
}

int64_t hamming_plane_words(int64_t N, int L) { return (int64_t)HP * ceil_div(L, 32) * N; }

int64_t hamming_num_tiles(int64_t N)
{
    const int64_t T = ceil_div(N, HT);
    return T * (T + 1) / 2;
}

int hamming_pack(const uint8_t *d_codes, int64_t N, int L, uint32_t *d_planes, cudaStream_t st)
{
    if (N <= 0 || L <= 0) { set_error("hamming_pack: empty alignment"); return 1; }
    const int W = (int)ceil_div(L, 32);
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)W);
    hamming_pack_kernel<<<grid, 256, 0, st>>>(d_codes, N, L, W, d_planes);
    EVC_KERNEL_CHECK();
    return 0;
}

int hamming_count_tiles(const uint32_t *d_planes, int64_t N, int L, int min_identical,
                        int64_t tile_begin, int64_t tile_end, int *d_counts, cudaStream_t st)
{
    const int W = (int)ceil_div(L, 32);
    const int64_t T = ceil_div(N, HT);
    if (tile_begin < 0 || tile_end > T * (T + 1) / 2 || tile_begin > tile_end) {
        set_error("hamming_count_tiles: tile range out of bounds");
        return 1;
    }
    const size_t smem = (size_t)2 * HWC * HP * HT * sizeof(uint32_t);
    EVC_CUDA(cudaFuncSetAttribute(hamming_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int thr = min_identical + (W * 32 - L);   // padded sites always "agree"
    int64_t done = tile_begin;
    while (done < tile_end) {                        // grid.x limit 2^31-1
        const int64_t nblk = (tile_end - done) < (int64_t)1 << 30 ? (tile_end - done) : (int64_t)1 << 30;
        hamming_tile_kernel<<<(unsigned)nblk, 256, smem, st>>>(d_planes, N, W, thr, done, T, d_counts);
        EVC_KERNEL_CHECK();
        done += nblk;
    }
    return 0;
}

}  // namespace evc
