// KNN on the device: the brute-force tile scan (I/algorithms/knn/bruteforce.hpp:24-96) and the
// GPU-resident exact index that stands in for knn::KDTree (kdtree.hpp:142-562) behind the same
// KNNBase contract (knn.hpp:14-61).
#include <cmath>
#include <cstring>

#include "spx_grid.cuh"
#include "spx_scan.cuh"

using namespace spx;

namespace {

// ------------------------------------------------------------------ brute force
constexpr int BF_THREADS = 256;
constexpr int BF_TILE = 2048;  // float4 targets staged per step (32 KB of shared memory)

// One thread owns QPT queries and streams every target tile out of shared memory (one broadcast
// LDS.128 per target feeds QPT distance evaluations).  Candidate lists live in the thread's rows of
// the output arrays; only the k-th best is cached in registers, so the steady-state inner loop is
// 3 sub + 1 mul + 2 fma + 1 compare per (query, target) pair.  Targets are scanned in index order
// with a strict '<', which yields exactly the (dist, index) order of bruteforce.hpp:71-83.
template <int QPT>
__global__ void __launch_bounds__(BF_THREADS) knn_bruteforce_kernel(const float4* __restrict__ queries, uint32_t nq,
                                                                    const float4* __restrict__ targets, uint32_t nt,
                                                                    int k, Xform T, int has_T,
                                                                    int32_t* __restrict__ idx,
                                                                    float* __restrict__ dist) {
    __shared__ float4 tile[BF_TILE];
    const uint32_t first = (blockIdx.x * BF_THREADS + threadIdx.x) * QPT;

    float qx[QPT], qy[QPT], qz[QPT], wd[QPT];
    float* drow[QPT];
    int32_t* irow[QPT];
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        const uint32_t qi = first + u;
        const bool live = qi < nq;
        float4 q = live ? __ldg(queries + qi) : make_float4(0.f, 0.f, 0.f, 1.f);
        if (has_T) q = transform_point(T, q);
        qx[u] = q.x; qy[u] = q.y; qz[u] = q.z;
        drow[u] = dist + (size_t)(live ? qi : 0) * k;
        irow[u] = idx + (size_t)(live ? qi : 0) * k;
        if (live) {
            for (int j = 0; j < k; ++j) { drow[u][j] = FLT_MAX; irow[u][j] = -1; }
            wd[u] = FLT_MAX;
        } else {
            wd[u] = -1.0f;  // nothing compares below it: the slot never inserts
        }
    }

    for (uint32_t base = 0; base < nt; base += BF_TILE) {
        __syncthreads();
#pragma unroll
        for (int t = threadIdx.x; t < BF_TILE; t += BF_THREADS) {
            const uint32_t j = base + t;
            // sentinel beyond the end: squares overflow to +inf, never below any k-th best
            tile[t] = j < nt ? __ldg(targets + j) : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 1.0f);
        }
        __syncthreads();
        const int lim = min((uint32_t)BF_TILE, nt - base);
        const int lim4 = (lim + 3) & ~3;  // sentinels cover the padding
#pragma unroll 4
        for (int t = 0; t < lim4; ++t) {
            const float4 p = tile[t];
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                const float ds = dist_sq(qx[u], qy[u], qz[u], p.x, p.y, p.z);
                if (ds < wd[u]) {
                    float* d = drow[u];
                    int32_t* id = irow[u];
                    int pos = k - 1;
                    while (pos > 0 && ds < d[pos - 1]) {
                        d[pos] = d[pos - 1];
                        id[pos] = id[pos - 1];
                        --pos;
                    }
                    d[pos] = ds;
                    id[pos] = (int)(base + t);
                    wd[u] = d[k - 1];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ index build
__device__ __forceinline__ int float_ordered(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
inline float float_from_ordered(int o) {
    const int i = o ^ ((o >> 31) & 0x7fffffff);
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}

struct BBoxAcc {
    int mn[3];
    int mx[3];
    uint32_t finite;
    uint32_t pad;
};

__global__ void bbox_kernel(const float4* __restrict__ pts, uint32_t n, BBoxAcc* acc) {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    uint32_t cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(pts + i);
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            mn[0] = min(mn[0], float_ordered(p.x)); mx[0] = max(mx[0], float_ordered(p.x));
            mn[1] = min(mn[1], float_ordered(p.y)); mx[1] = max(mx[1], float_ordered(p.y));
            mn[2] = min(mn[2], float_ordered(p.z)); mx[2] = max(mx[2], float_ordered(p.z));
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&acc->mn[a], mn[a]);
            atomicMax(&acc->mx[a], mx[a]);
        }
        atomicAdd(&acc->finite, cnt);
    }
}

struct GridGeom {
    float ox, oy, oz, inv;
    int dx, dy, dz;
};

__device__ __forceinline__ uint32_t cell_of(const GridGeom& g, const float4 p) {
    const int cx = grid_coord(p.x, g.ox, g.inv, g.dx);
    const int cy = grid_coord(p.y, g.oy, g.inv, g.dy);
    const int cz = grid_coord(p.z, g.oz, g.inv, g.dz);
    return ((uint32_t)cz * (uint32_t)g.dy + (uint32_t)cy) * (uint32_t)g.dx + (uint32_t)cx;
}

__global__ void cell_count_kernel(const float4* __restrict__ pts, uint32_t n, GridGeom g, uint32_t* __restrict__ cell_id,
                                  uint32_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(pts + i);
    uint32_t c = 0xffffffffu;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        c = cell_of(g, p);
        atomicAdd(counts + c, 1u);
    }
    cell_id[i] = c;
}

__global__ void occupied_kernel(const uint32_t* __restrict__ counts, size_t ncells, unsigned long long* occupied) {
    unsigned long long c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (size_t)gridDim.x * blockDim.x)
        c += counts[i] != 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(occupied, c);
}

// Stable scatter: a point's slot inside its cell is its rank among the cell's points in ORIGINAL
// index order, so the sorted copy (and therefore the visiting order) is deterministic run to run.
// Rank = number of earlier points of the same cell; computed with one atomic per point on a
// per-cell cursor would be order-dependent, so instead each cell's points are written in any order
// and then ordered by original index in a second pass (cells are tiny).
__global__ void cell_scatter_kernel(const float4* __restrict__ pts, uint32_t n, const uint32_t* __restrict__ cell_id,
                                    const uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                                    float4* __restrict__ sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_id[i];
    if (c == 0xffffffffu) return;
    const uint32_t pos = start[c] + atomicAdd(cursor + c, 1u);
    const float4 p = __ldg(pts + i);
    sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
}

// insertion sort of each cell's slice by original index (one thread per cell)
__global__ void cell_order_kernel(const uint32_t* __restrict__ start, size_t ncells, float4* __restrict__ sorted) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t s = start[c], e = start[c + 1];
    if (e - s > 64) return;  // pathological cells stay in arrival order (results are order-independent)
    for (uint32_t a = s + 1; a < e; ++a) {
        const float4 v = sorted[a];
        const int key = __float_as_int(v.w);
        uint32_t b = a;
        while (b > s && __float_as_int(sorted[b - 1].w) > key) {
            sorted[b] = sorted[b - 1];
            --b;
        }
        sorted[b] = v;
    }
}

// ------------------------------------------------------------------ index search
constexpr int GRID_THREADS = 128;

// k == 1: the candidate is a register pair.  k > 1: each thread's sorted candidate list lives in
// shared memory, element j of thread t at [j * GRID_THREADS + t] (conflict-free), and the block
// writes its [queries][k] result tile out cooperatively so the global stores coalesce.
template <bool K1>
__global__ void __launch_bounds__(GRID_THREADS) grid_knn_kernel(const GridLevels g, const float4* __restrict__ queries,
                                                                 uint32_t nq, int k, Xform T, int has_T,
                                                                 int32_t* __restrict__ idx, float* __restrict__ dist) {
    extern __shared__ float smem_lists[];
    const uint32_t q0 = blockIdx.x * GRID_THREADS;
    const uint32_t qi = q0 + threadIdx.x;
    const float INF = __int_as_float(0x7f800000);
    float4 q = make_float4(0.f, 0.f, 0.f, 1.f);
    if (qi < nq) {
        q = __ldg(queries + qi);
        if (has_T) q = transform_point(T, q);
    }
    const bool searchable = qi < nq && isfinite(q.x) && isfinite(q.y) && isfinite(q.z) && g.lv[0].n > 0;
    if (K1) {
        Best1 best;
        best.init();
        if (searchable) grid_search_levels(g, q.x, q.y, q.z, best, INF);
        if (qi < nq) {
            idx[qi] = best.i;
            dist[qi] = best.d;
        }
    } else {
        BestK best;
        best.d = smem_lists + threadIdx.x;
        best.i = reinterpret_cast<int*>(smem_lists + (size_t)k * GRID_THREADS) + threadIdx.x;
        best.k = k;
        best.stride = GRID_THREADS;
        best.init();
        if (searchable) grid_search_levels(g, q.x, q.y, q.z, best, INF);
        __syncthreads();
        const float* sd = smem_lists;
        const int* si = reinterpret_cast<const int*>(smem_lists + (size_t)k * GRID_THREADS);
        const uint32_t live = min((uint32_t)GRID_THREADS, nq - q0);
        for (uint32_t e = threadIdx.x; e < live * (uint32_t)k; e += GRID_THREADS) {
            const uint32_t ql = e / (uint32_t)k, j = e - ql * (uint32_t)k;
            dist[(size_t)q0 * k + e] = sd[j * GRID_THREADS + ql];
            idx[(size_t)q0 * k + e] = si[j * GRID_THREADS + ql];
        }
    }
}

// work counters of the k = 1 search (tuning aid): stats[q] = {segments, candidates, shells, last level}
__global__ void __launch_bounds__(GRID_THREADS) grid_nn_stats_kernel(const GridLevels g, const float4* __restrict__ queries,
                                                                      uint32_t nq, Xform T, int has_T, float max_radius,
                                                                      uint4* __restrict__ stats) {
    const uint32_t qi = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (qi >= nq) return;
    float4 q = __ldg(queries + qi);
    if (has_T) q = transform_point(T, q);
    Best1 best;
    best.init();
    CountStats cs;
    if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z) && g.lv[0].n > 0)
        grid_search_levels(g, q.x, q.y, q.z, best, max_radius, &cs);
    stats[qi] = make_uint4(cs.segs, cs.cands, cs.shells, cs.last_level);
}

void launch_bruteforce(spx_queue_t q, const float4* queries, uint32_t nq, const float4* targets, uint32_t nt, int k,
                       const Xform& T, int has_T, int32_t* idx, float* dist) {
    if (nq == 0) return;
    // 2 queries per thread halves the shared-memory traffic per pair; small batches keep 1 so
    // the grid still covers the SMs
    const bool two = nq >= (uint32_t)q->sm_count * BF_THREADS * 4;
    const unsigned blocks = (unsigned)div_up(nq, (size_t)BF_THREADS * (two ? 2 : 1));
    if (two)
        knn_bruteforce_kernel<2><<<blocks, BF_THREADS, 0, q->stream>>>(queries, nq, targets, nt, k, T, has_T, idx, dist);
    else
        knn_bruteforce_kernel<1><<<blocks, BF_THREADS, 0, q->stream>>>(queries, nq, targets, nt, k, T, has_T, idx, dist);
    SPX_LAUNCH_CHECK();
}

constexpr size_t MAX_CELLS = (size_t)1 << 24;

}  // namespace

extern "C" {

int spx_knn_bruteforce(spx_queue_t q, const float* queries, size_t nq, const float* targets, size_t nt, int k,
                       const float* T_host, int32_t* idx, float* dist) {
    return guard([&] {
        SPX_REQUIRE(q, "[knn_search_bruteforce] null queue");
        SPX_REQUIRE(k >= 1 && k <= 128, "[knn_search_bruteforce] `k` must be in [1, 128]");
        SPX_REQUIRE(nq < (1ull << 31) && nt < (1ull << 31), "[knn_search_bruteforce] too many points");
        if (nq == 0) return;
        SPX_REQUIRE(queries && idx && dist && (targets || nt == 0), "[knn_search_bruteforce] null pointer");
        DeviceGuard g(q->device);
        const Xform T = T_host ? xform_from_colmajor(T_host) : xform_identity();
        launch_bruteforce(q, reinterpret_cast<const float4*>(queries), (uint32_t)nq,
                          reinterpret_cast<const float4*>(targets), (uint32_t)nt, k, T, T_host != nullptr, idx, dist);
    });
}

int spx_index_build(spx_queue_t q, const float* targets, size_t nt, float cell_size, spx_index_t* out) {
    return guard([&] {
        SPX_REQUIRE(q && out, "[KDTree::build] null argument");
        SPX_REQUIRE(nt < (1ull << 31), "[KDTree::build] too many points");
        SPX_REQUIRE(targets || nt == 0, "[KDTree::build] null points");
        DeviceGuard dg(q->device);
        auto* ix = new spx_index_s();
        ix->q = q;
        ix->n_total = nt;
        *out = ix;
        GridLevels& L = ix->levels;
        L = GridLevels{};
        L.n_levels = 1;
        L.lv[0].dx = L.lv[0].dy = L.lv[0].dz = 1;
        L.lv[0].cell = 1.0f;
        L.lv[0].inv = 1.0f;
        if (nt == 0) return;  // empty tree: every search returns -1 / FLT_MAX (kdtree.hpp:296-300)
        const float4* pts = reinterpret_cast<const float4*>(targets);
        const uint32_t n = (uint32_t)nt;
        cudaStream_t st = q->stream;

        q->arena_reset();
        q->arena_reserve(sizeof(BBoxAcc) + 256 + (size_t)n * 4 + 2 * (MAX_CELLS + 64) * 4 +
                         scan_scratch_elems(MAX_CELLS + 1) * 4 + 4096);
        BBoxAcc* acc = q->take<BBoxAcc>(1);
        unsigned long long* occ_dev = q->take<unsigned long long>(1);
        uint32_t* cell_id = q->take<uint32_t>(n);
        uint32_t* counts = q->take<uint32_t>(MAX_CELLS + 64);
        uint32_t* scan_tmp = q->take<uint32_t>(scan_scratch_elems(MAX_CELLS + 1));

        BBoxAcc* hacc = static_cast<BBoxAcc*>(q->pinned_get(sizeof(BBoxAcc) + 64));
        for (int a = 0; a < 3; ++a) {
            hacc->mn[a] = INT_MAX;
            hacc->mx[a] = INT_MIN;
        }
        hacc->finite = 0;
        hacc->pad = 0;
        SPX_CUDA(cudaMemcpyAsync(acc, hacc, sizeof(BBoxAcc), cudaMemcpyHostToDevice, st));
        bbox_kernel<<<std::min(div_up(n, 256), q->sm_count * 8), 256, 0, st>>>(pts, n, acc);
        SPX_LAUNCH_CHECK();
        SPX_CUDA(cudaMemcpyAsync(hacc, acc, sizeof(BBoxAcc), cudaMemcpyDeviceToHost, st));
        q->sync();
        const BBoxAcc bb = *hacc;
        ix->n = bb.finite;
        L.lv[0].n = bb.finite;
        if (bb.finite == 0) return;

        float lo[3], ext[3];
        float max_ext = 0.0f, max_abs = 0.0f;
        for (int a = 0; a < 3; ++a) {
            lo[a] = float_from_ordered(bb.mn[a]);
            const float hi = float_from_ordered(bb.mx[a]);
            ext[a] = hi - lo[a];
            max_ext = std::max(max_ext, ext[a]);
            max_abs = std::max(max_abs, std::max(std::fabs(lo[a]), std::fabs(hi)));
        }
        if (!(max_ext > 0.0f)) max_ext = 1.0f;

        auto dims_for = [&](float cell, int dims[3]) {
            double nc = 1.0;
            for (int a = 0; a < 3; ++a) {
                const double d = std::floor((double)ext[a] / (double)cell) + 1.0;
                dims[a] = (int)std::min(d, 2.0e9);
                nc *= d;
            }
            return nc;
        };
        auto count_level = [&](float cell, const int dims[3], size_t ncells) {
            GridGeom geom{lo[0], lo[1], lo[2], 1.0f / cell, dims[0], dims[1], dims[2]};
            SPX_CUDA(cudaMemsetAsync(counts, 0, (ncells + 1) * 4, st));
            cell_count_kernel<<<div_up(n, 256), 256, 0, st>>>(pts, n, geom, cell_id, counts);
            SPX_LAUNCH_CHECK();
        };
        auto finish_level = [&](int level, float cell, const int dims[3], size_t ncells, bool order) {
            SPX_CUDA(cudaMallocAsync(&ix->start[level], (ncells + 1) * 4, st));
            SPX_CUDA(cudaMallocAsync(&ix->sorted[level], (size_t)bb.finite * sizeof(float4), st));
            exclusive_scan_u32(st, counts, ix->start[level], ncells + 1, scan_tmp, nullptr);
            SPX_CUDA(cudaMemsetAsync(counts, 0, (ncells + 1) * 4, st));  // reuse as per-cell cursor
            cell_scatter_kernel<<<div_up(n, 256), 256, 0, st>>>(pts, n, cell_id, ix->start[level], counts,
                                                               ix->sorted[level]);
            SPX_LAUNCH_CHECK();
            if (order) {
                cell_order_kernel<<<div_up(ncells, 256), 256, 0, st>>>(ix->start[level], ncells, ix->sorted[level]);
                SPX_LAUNCH_CHECK();
            }
            GridView& v = L.lv[level];
            v.ox = lo[0]; v.oy = lo[1]; v.oz = lo[2];
            v.cell = cell;
            v.inv = 1.0f / cell;
            v.dx = dims[0]; v.dy = dims[1]; v.dz = dims[2];
            v.margin = 1e-3f * cell + 2e-6f * (max_abs + max_ext);
            v.start = ix->start[level];
            v.pts = ix->sorted[level];
            v.n = bb.finite;
            ix->ncells[level] = ncells;
        };

        // finest level: ~2 cells per point over the (thickened) bounding box to start with, then
        // adapted to the measured occupancy (LiDAR clouds are surfaces; volume heuristics are off)
        float cell = cell_size;
        const bool adaptive = !(cell_size > 0.0f);
        if (adaptive) {
            double vol = 1.0;
            for (int a = 0; a < 3; ++a) vol *= std::max(ext[a], 0.02f * max_ext);
            cell = (float)std::cbrt(vol / (2.0 * (double)bb.finite));
            cell = std::max(cell, 1e-6f * std::max(max_abs, 1.0f));
        }
        int dims[3];
        size_t ncells = 0;
        unsigned long long occupied = 0;
        for (int attempt = 0; attempt < 5; ++attempt) {
            for (;;) {  // respect the dense-grid budget
                const double nc = dims_for(cell, dims);
                if (nc <= (double)MAX_CELLS) {
                    ncells = (size_t)dims[0] * dims[1] * dims[2];
                    break;
                }
                cell *= (float)std::cbrt(nc / (double)MAX_CELLS) * 1.02f;
            }
            count_level(cell, dims, ncells);
            if (!adaptive) break;
            SPX_CUDA(cudaMemsetAsync(occ_dev, 0, 8, st));
            occupied_kernel<<<std::min(div_up(ncells, 256), q->sm_count * 8), 256, 0, st>>>(counts, ncells, occ_dev);
            SPX_LAUNCH_CHECK();
            unsigned long long* hocc = reinterpret_cast<unsigned long long*>(q->pinned_get(64));
            SPX_CUDA(cudaMemcpyAsync(hocc, occ_dev, 8, cudaMemcpyDeviceToHost, st));
            q->sync();
            occupied = *hocc;
            if (attempt == 4) break;
            const double avg = (double)bb.finite / (double)std::max<unsigned long long>(occupied, 1);
            // aim for ~2 points per occupied cell: the points a query has to look at grow with the
            // square of the cell edge on a surface, the row look-ups do not
            if (avg < 1.3 && bb.finite > 8) {
                cell *= (float)std::min(3.0, std::max(1.25, std::sqrt(2.0 / avg)));
            } else if (avg > 3.5) {
                cell *= (float)std::max(0.3, std::min(0.8, std::sqrt(2.0 / avg)));
            } else {
                break;
            }
        }
        ix->occupied = (int64_t)occupied;
        finish_level(0, cell, dims, ncells, true);

        // coarser levels until the coarsest grid is only a few cells wide
        int level = 0;
        while (level + 1 < GRID_MAX_LEVELS && std::max(dims[0], std::max(dims[1], dims[2])) > 4) {
            ++level;
            cell *= (float)GRID_LEVEL_FACTOR;
            dims_for(cell, dims);
            ncells = (size_t)dims[0] * dims[1] * dims[2];
            count_level(cell, dims, ncells);
            finish_level(level, cell, dims, ncells, false);
        }
        L.n_levels = level + 1;
        q->sync();
    });
}

int spx_index_destroy(spx_index_t index) {
    return guard([&] {
        if (!index) return;
        DeviceGuard g(index->q->device);
        for (int l = 0; l < GRID_MAX_LEVELS; ++l) {
            if (index->sorted[l]) cudaFreeAsync(index->sorted[l], index->q->stream);
            if (index->start[l]) cudaFreeAsync(index->start[l], index->q->stream);
        }
        delete index;
    });
}

int spx_index_info(spx_index_t index, float* cell_size, int32_t* dims3, int64_t* occupied_cells, int64_t* n_points) {
    return guard([&] {
        SPX_REQUIRE(index, "[spx_index_info] null index");
        const GridView& v = index->levels.lv[0];
        if (cell_size) *cell_size = v.cell;
        if (dims3) {
            dims3[0] = v.dx;
            dims3[1] = v.dy;
            dims3[2] = v.dz;
        }
        if (occupied_cells) *occupied_cells = index->occupied;
        if (n_points) *n_points = index->n;
    });
}

int spx_index_levels(spx_index_t index, int32_t* n_levels) {
    return guard([&] {
        SPX_REQUIRE(index && n_levels, "[spx_index_levels] null argument");
        *n_levels = index->levels.n_levels;
    });
}

int spx_index_nn_stats(spx_index_t index, const float* queries, size_t nq, const float* T_host, float max_radius,
                       uint32_t* stats4) {
    return guard([&] {
        SPX_REQUIRE(index && queries && stats4, "[spx_index_nn_stats] null argument");
        spx_queue_t q = index->q;
        DeviceGuard g(q->device);
        const Xform T = T_host ? xform_from_colmajor(T_host) : xform_identity();
        const float r = max_radius > 0.0f ? max_radius : INFINITY;
        grid_nn_stats_kernel<<<div_up(nq, GRID_THREADS), GRID_THREADS, 0, q->stream>>>(
            index->levels, reinterpret_cast<const float4*>(queries), (uint32_t)nq, T, T_host != nullptr, r,
            reinterpret_cast<uint4*>(stats4));
        SPX_LAUNCH_CHECK();
    });
}

int spx_index_knn(spx_index_t index, const float* queries, size_t nq, int k, const float* T_host, int32_t* idx,
                  float* dist) {
    return guard([&] {
        SPX_REQUIRE(index, "[KDTree::knn_search_async] null index");
        SPX_REQUIRE(k >= 1 && k <= 128, "[KDTree::knn_search_async] `k` is too large. not support.");
        SPX_REQUIRE(nq < (1ull << 31), "[KDTree::knn_search_async] too many queries");
        if (nq == 0) return;  // empty query -> empty result (kdtree.hpp:429-436)
        SPX_REQUIRE(queries && idx && dist, "[KDTree::knn_search_async] null pointer");
        spx_queue_t q = index->q;
        DeviceGuard g(q->device);
        const Xform T = T_host ? xform_from_colmajor(T_host) : xform_identity();
        const int has_T = T_host != nullptr;
        const unsigned blocks = (unsigned)div_up(nq, GRID_THREADS);
        const float4* qs = reinterpret_cast<const float4*>(queries);
        if (k == 1) {
            grid_knn_kernel<true><<<blocks, GRID_THREADS, 0, q->stream>>>(index->levels, qs, (uint32_t)nq, k, T, has_T,
                                                                         idx, dist);
        } else {
            const size_t smem = (size_t)k * GRID_THREADS * 8;
            static bool attr_set = false;
            if (!attr_set) {
                SPX_CUDA(cudaFuncSetAttribute(grid_knn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              128 * GRID_THREADS * 8));
                attr_set = true;
            }
            grid_knn_kernel<false><<<blocks, GRID_THREADS, smem, q->stream>>>(index->levels, qs, (uint32_t)nq, k, T,
                                                                             has_T, idx, dist);
        }
        SPX_LAUNCH_CHECK();
    });
}

}  // extern "C"
