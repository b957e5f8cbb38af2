// Stable LSD radix sort of (uint64 key, uint32 value) pairs in HBM -- the "sort + unique" of kmer-db's build
// (console_build.cpp:94-103) restated for the whole genome set at once: tuples (canonical k-mer, genome id) are
// written in genome order, so a STABLE sort by k-mer leaves the genome ids of every k-mer ascending and the
// duplicates of one genome adjacent.
//
// Per pass (RB-bit digit): tile histogram -> per-digit scan over tiles -> stable scatter.  The tile is
// 256 threads x 16 items; ranking inside a warp uses __match_any_sync, across warps a shared-memory table.
// HBM traffic per pass and item: 8 B (histogram read) + 12 B (read) + 12 B (write).
#pragma once
#include "dev_util.cuh"

namespace rsort {

constexpr int BLOCK = 256;
constexpr int ITEMS = 16;
constexpr int TILE = BLOCK * ITEMS;     // 4096
constexpr int WARPS = BLOCK / 32;

template <int RB>
__global__ void __launch_bounds__(BLOCK) hist_kernel(const uint64_t *__restrict__ keys, uint32_t n_tiles, int shift,
                                                     uint32_t *__restrict__ tile_hist /* [digit][tile] */)
{
    constexpr int D = 1 << RB;
    __shared__ uint32_t h[D];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int d = threadIdx.x; d < D; d += BLOCK) h[d] = 0;
        __syncthreads();
        const uint64_t *k = keys + (uint64_t)tile * TILE;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            uint32_t d = (uint32_t)(k[r * BLOCK + threadIdx.x] >> shift) & (D - 1);
            atomicAdd(&h[d], 1u);
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += BLOCK) tile_hist[(uint64_t)d * n_tiles + tile] = h[d];
        __syncthreads();
    }
}

// exclusive scan of each digit row over tiles (one block per digit), row totals to digit_total
static __global__ void __launch_bounds__(BLOCK) scan_rows_kernel(uint32_t *__restrict__ tile_hist, uint32_t n_tiles,
                                                          uint32_t *__restrict__ digit_total)
{
    __shared__ uint32_t warp_sum[WARPS];
    __shared__ uint32_t carry_s;
    uint32_t *row = tile_hist + (uint64_t)blockIdx.x * n_tiles;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_tiles; base += BLOCK) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_tiles ? row[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t pre = carry_s;
        for (int j = 0; j < w; ++j) pre += warp_sum[j];
        if (i < n_tiles) row[i] = pre + x - v;
        __syncthreads();
        if (threadIdx.x == BLOCK - 1) carry_s = pre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_s;
}

template <int RB>
__global__ void scan_digits_kernel(uint32_t *__restrict__ digit_total)
{
    constexpr int D = 1 << RB;
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int d = 0; d < D; ++d) { uint32_t t = digit_total[d]; digit_total[d] = run; run += t; }
    }
}

template <int RB>
__global__ void __launch_bounds__(BLOCK) scatter_kernel(const uint64_t *__restrict__ keys_in,
                                                        const uint32_t *__restrict__ vals_in,
                                                        uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                        uint32_t n_tiles, int shift,
                                                        const uint32_t *__restrict__ tile_hist,
                                                        const uint32_t *__restrict__ digit_base)
{
    constexpr int D = 1 << RB;
    __shared__ uint32_t wh[WARPS][D];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int d = threadIdx.x; d < WARPS * D; d += BLOCK) (&wh[0][0])[d] = 0;
        __syncthreads();
        const uint64_t base = (uint64_t)tile * TILE + (uint64_t)w * (32 * ITEMS);
        uint64_t key[ITEMS];
        uint32_t rank[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) key[r] = keys_in[base + r * 32 + lane];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            uint32_t d = (uint32_t)(key[r] >> shift) & (D - 1);
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = wh[w][d]; wh[w][d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = old + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += BLOCK) {
            uint32_t run = digit_base[d] + tile_hist[(uint64_t)d * n_tiles + tile];
#pragma unroll
            for (int j = 0; j < WARPS; ++j) { uint32_t t = wh[j][d]; wh[j][d] = run; run += t; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            uint32_t d = (uint32_t)(key[r] >> shift) & (D - 1);
            uint32_t o = wh[w][d] + rank[r];
            keys_out[o] = key[r];
            vals_out[o] = vals_in[base + r * 32 + lane];
        }
        __syncthreads();
    }
}

struct Workspace {
    DevBuf<uint32_t> tile_hist, digit_total;
};

// Sorts n_padded (multiple of TILE, < 2^32) items on key bits [0, n_bits).  Returns true when the result ended
// up in (keys_b, vals_b), false when it is in (keys_a, vals_a).
template <int RB = 8>
bool sort_kv(vb_ctx *ctx, uint64_t *keys_a, uint32_t *vals_a, uint64_t *keys_b, uint32_t *vals_b, uint64_t n_padded,
             int n_bits, Workspace &ws)
{
    constexpr int D = 1 << RB;
    if (n_padded == 0) return false;
    if (n_padded % TILE) throw vb_error(VB_ERR_INTERNAL, "sort_kv: size not padded to the tile");
    if (n_padded >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "sort_kv: more than 2^32 tuples in one sort");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    uint32_t n_tiles = (uint32_t)(n_padded / TILE);
    if (ws.tile_hist.n < (size_t)D * n_tiles) ws.tile_hist.alloc((size_t)D * n_tiles);
    if (ws.digit_total.n < (size_t)D) ws.digit_total.alloc(D);
    int grid = (int)std::min<uint32_t>(n_tiles, 148 * 8);
    bool in_b = false;
    for (int shift = 0; shift < n_bits; shift += RB) {
        const uint64_t *ki = in_b ? keys_b : keys_a;
        const uint32_t *vi = in_b ? vals_b : vals_a;
        uint64_t *ko = in_b ? keys_a : keys_b;
        uint32_t *vo = in_b ? vals_a : vals_b;
        hist_kernel<RB><<<grid, BLOCK, 0, st>>>(ki, n_tiles, shift, ws.tile_hist.p);
        VB_LAUNCH_CHECK(ctx);
        scan_rows_kernel<<<D, BLOCK, 0, st>>>(ws.tile_hist.p, n_tiles, ws.digit_total.p);
        VB_LAUNCH_CHECK(ctx);
        scan_digits_kernel<RB><<<1, 32, 0, st>>>(ws.digit_total.p);
        VB_LAUNCH_CHECK(ctx);
        scatter_kernel<RB><<<grid, BLOCK, 0, st>>>(ki, vi, ko, vo, n_tiles, shift, ws.tile_hist.p, ws.digit_total.p);
        VB_LAUNCH_CHECK(ctx);
        in_b = !in_b;
    }
    return in_b;
}

}  // namespace rsort
