// Device-wide exclusive prefix sum over uint32 (counting sort of the cell grid, radix-sort digit
// offsets, voxel compaction).  Three-phase reduce / scan / down-sweep; the middle phase recurses
// on the per-block totals, so any n up to 2^32-1 elements works with O(log_{4096} n) levels.
// The reference does this on the device too, with its own three kernels
// (I/algorithms/common/prefix_sum.hpp:30-166); this is an independent shared-memory design.
#pragma once

#include "spx_common.cuh"

namespace spx {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// block-level exclusive scan of one value per thread; returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp /*[32]*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        uint32_t w = lane < nw ? smem_warp[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;
    }
    __syncthreads();
    const uint32_t off = smem_warp[warp];
    if (total) *total = smem_warp[32];
    return off + inc - v;
}

// phase 1: per-tile totals
static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_totals_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                        uint32_t* __restrict__ totals) {
    __shared__ uint32_t sw[33];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const size_t j = base + i;
        if (j < n) s += in[j];
    }
    uint32_t total;
    block_exclusive_scan(s, sw, &total);
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

// phase 3 (also the whole job when n fits one tile): scan inside the tile + tile offset.
// RAW_TOTALS: `tile_offsets` holds the un-scanned per-tile totals and every block sums its
// predecessors itself (a few KB of L2 reads) — saves the middle launch for up to SCAN_TILE tiles.
template <bool RAW_TOTALS>
static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_apply_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                       const uint32_t* __restrict__ tile_offsets,
                                                                       uint32_t* __restrict__ out,
                                                                       uint32_t* __restrict__ grand_total) {
    __shared__ uint32_t sw[33];
    __shared__ uint32_t tile_off;
    uint32_t my_off = 0;
    if (RAW_TOTALS) {
        uint32_t part = 0;
        for (uint32_t t = threadIdx.x; t < blockIdx.x; t += SCAN_THREADS) part += tile_offsets[t];
        uint32_t tot;
        block_exclusive_scan(part, sw, &tot);
        if (threadIdx.x == 0) tile_off = tot;
        __syncthreads();
        my_off = tile_off;
        __syncthreads();  // sw is reused below
    } else if (tile_offsets) {
        my_off = tile_offsets[blockIdx.x];
    }
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const size_t j = base + i;
        v[i] = j < n ? in[j] : 0u;
        s += v[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, sw, &total) + my_off;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const size_t j = base + i;
        if (j < n) out[j] = run;
        run += v[i];
    }
    if (grand_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *grand_total = my_off + total;
}

inline size_t scan_scratch_elems(size_t n) {
    size_t total = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += align_up(n, 64);
    }
    return total + 64;
}

// out may alias in.  scratch: scan_scratch_elems(n) uint32.  grand_total (device, nullable) gets sum(in).
inline void exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch,
                               uint32_t* grand_total) {
    if (n == 0) {
        if (grand_total) SPX_CUDA(cudaMemsetAsync(grand_total, 0, sizeof(uint32_t), st));
        return;
    }
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_tile_apply_kernel<false><<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, out, grand_total);
        SPX_LAUNCH_CHECK();
        return;
    }
    uint32_t* totals = scratch;
    scan_tile_totals_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, totals);
    SPX_LAUNCH_CHECK();
    if (tiles <= (size_t)SCAN_TILE) {  // two launches: each apply block sums its predecessors' totals itself
        scan_tile_apply_kernel<true><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, totals, out, grand_total);
        SPX_LAUNCH_CHECK();
        return;
    }
    exclusive_scan_u32(st, totals, totals, tiles, scratch + align_up(tiles, 64), nullptr);
    scan_tile_apply_kernel<false><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, totals, out, grand_total);
    SPX_LAUNCH_CHECK();
}

}  // namespace spx
