// Device-side view of the exact nearest-neighbour index (uniform cell grid, points counting-
// sorted by cell) and the ring search over it.  Shared by the KNN entry points (spx_knn.cu) and
// the fused nearest-neighbour + linearise kernels (spx_registration.cu).
//
// Exactness argument: cells are visited in Chebyshev shells r = 0,1,2,...
// around the query's (clamped) cell c.  Every point NOT yet visited after shell r lies in a cell
// whose index differs from c by more than r on some axis a, so |p_a - q_a| >= gap_a(r), the
// distance from q_a to the slab [o_a + (c_a-r) h, o_a + (c_a+r+1) h] (infinite on a side where the
// slab already touches the grid boundary: there are no cells, hence no points, beyond it).  The
// fp32 distance fma(dz,dz,fma(dy,dy,dx*dx)) is monotone in each |d_a|, so every unvisited point
// has dist >= (min_a gap_a(r))^2 up to rounding; the search stops once the current k-th best is
// STRICTLY below (bound - margin)^2, margin covering the rounding of cell assignment and of the
// slab coordinates.  Strictness + the (dist, index) insertion order make ties resolve to the
// lowest index no matter in which order cells are visited — the brute-force contract.
#pragma once

#include "spx_common.cuh"

namespace spx {

struct GridView {
    float ox, oy, oz;    // grid origin (bbox min of the indexed points)
    float cell, inv;     // cell edge and its reciprocal
    float margin;        // slab-coordinate / cell-assignment rounding allowance
    int dx, dy, dz;      // grid dimensions
    const uint32_t* __restrict__ start;  // [dx*dy*dz + 1] first sorted position of each cell
    const float4* __restrict__ pts;      // sorted points, w = original index (int bits)
    uint32_t n;                          // indexed (finite) points
};

#ifdef __CUDACC__

__device__ __forceinline__ int grid_coord(float q, float o, float inv, int dim) {
    // identical expression for build and query: floor((q - o) * inv), clamped into the grid
    float g = floorf(__fmul_rn(__fsub_rn(q, o), inv));
    g = fminf(fmaxf(g, 0.0f), (float)(dim - 1));
    return (int)g;
}

// distance from q to the slab of cell index ci on one axis (0 inside), minus nothing; >= 0
__device__ __forceinline__ float axis_gap(float q, float o, float cell, int ci) {
    const float lo = __fmaf_rn((float)ci, cell, o);
    const float hi = __fmaf_rn((float)(ci + 1), cell, o);
    return fmaxf(fmaxf(lo - q, q - hi), 0.0f);
}

// Lower bound on |p_a - q_a| for points in cells farther than r from c on axis a (see header).
__device__ __forceinline__ float shell_bound_axis(float q, float o, float cell, int c, int r, int dim) {
    const float INF = __int_as_float(0x7f800000);
    const float lo = (c - r > 0) ? q - __fmaf_rn((float)(c - r), cell, o) : INF;
    const float hi = (c + r < dim - 1) ? __fmaf_rn((float)(c + r + 1), cell, o) - q : INF;
    return fminf(lo, hi);
}

// One candidate list per thread.  K1: a register pair.  General k: the thread's row of the output
// arrays (global memory, L1/L2-resident; touched only on insertion, which is rare once warm).
struct Best1 {
    float d;
    int i;      // original index
    uint32_t p; // sorted position (for gathers in fused kernels)
    __device__ __forceinline__ void init() { d = FLT_MAX; i = -1; p = 0; }
    __device__ __forceinline__ float worst() const { return d; }
    __device__ __forceinline__ void offer(float ds, int idx, uint32_t pos) {
        if (ds < d || (ds == d && (i < 0 || idx < i))) { d = ds; i = idx; p = pos; }
    }
};

struct BestK {
    float* d;   // [k]
    int* i;     // [k]
    int k;
    float wd;   // cached d[k-1]
    int wi;     // cached i[k-1]
    __device__ __forceinline__ void init() {
        for (int j = 0; j < k; ++j) { d[j] = FLT_MAX; i[j] = -1; }
        wd = FLT_MAX; wi = -1;
    }
    __device__ __forceinline__ float worst() const { return wd; }
    __device__ __forceinline__ void offer(float ds, int idx, uint32_t) {
        if (!lex_less(ds, idx, wd, wi)) return;
        int pos = k - 1;
        while (pos > 0 && lex_less(ds, idx, d[pos - 1], i[pos - 1])) {
            d[pos] = d[pos - 1];
            i[pos] = i[pos - 1];
            --pos;
        }
        d[pos] = ds;
        i[pos] = idx;
        wd = d[k - 1];
        wi = i[k - 1];
    }
};

constexpr int GRID_SEG_CHUNK = 8;

// Ring search.  max_radius: stop expanding once every point within it has been seen (ICP use:
// correspondences farther than max_correspondence_distance are rejected anyway,
// registration.hpp:584); pass +inf for an unbounded exact search.  r_max: shells after which an
// unbounded search gives up (returns false -> caller falls back to a full scan).
template <typename Best>
__device__ inline bool grid_search(const GridView& g, float qx, float qy, float qz, Best& best, float max_radius,
                                   int r_max) {
    const int cx = grid_coord(qx, g.ox, g.inv, g.dx);
    const int cy = grid_coord(qy, g.oy, g.inv, g.dy);
    const int cz = grid_coord(qz, g.oz, g.inv, g.dz);
    const float INF = __int_as_float(0x7f800000);
    // rounding allowance: grid part (build) + the query's own magnitude (q - slab coordinate)
    const float margin = g.margin + 1e-6f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz));

    uint32_t seg_lo[GRID_SEG_CHUNK], seg_hi[GRID_SEG_CHUNK];
    int nseg = 0;

    auto flush = [&]() {
        uint32_t s[GRID_SEG_CHUNK], e[GRID_SEG_CHUNK];
#pragma unroll
        for (int t = 0; t < GRID_SEG_CHUNK; ++t) {
            s[t] = 0; e[t] = 0;
            if (t < nseg) {
                s[t] = __ldg(g.start + seg_lo[t]);
                e[t] = __ldg(g.start + seg_hi[t] + 1);
            }
        }
#pragma unroll
        for (int t = 0; t < GRID_SEG_CHUNK; ++t) {
            if (t < nseg) {
#pragma unroll 4
                for (uint32_t j = s[t]; j < e[t]; ++j) {
                    const float4 p = __ldg(g.pts + j);
                    const float ds = dist_sq(qx, qy, qz, p.x, p.y, p.z);
                    best.offer(ds, __float_as_int(p.w), j);
                }
            }
        }
        nseg = 0;
    };

    for (int r = 0;; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dy - 1);
        for (int zz = z0; zz <= z1; ++zz) {
            const float gz = axis_gap(qz, g.oz, g.cell, zz);
            const bool ez = (zz - cz == r) || (cz - zz == r);
            for (int yy = y0; yy <= y1; ++yy) {
                const float gy = axis_gap(qy, g.oy, g.cell, yy);
                // prune rows that cannot hold anything better than the current k-th best
                const float gyz = fmaxf(sqrtf(__fmaf_rn(gz, gz, __fmul_rn(gy, gy))) - margin, 0.0f);
                if (__fmul_rn(gyz, gyz) > best.worst()) continue;
                const bool edge = ez || (yy - cy == r) || (cy - yy == r);
                const uint32_t row = ((uint32_t)zz * (uint32_t)g.dy + (uint32_t)yy) * (uint32_t)g.dx;
                if (edge) {
                    const int xa = max(cx - r, 0), xb = min(cx + r, g.dx - 1);
                    seg_lo[nseg] = row + xa;
                    seg_hi[nseg] = row + xb;
                    if (++nseg == GRID_SEG_CHUNK) flush();
                } else {
                    if (cx - r >= 0) {
                        seg_lo[nseg] = seg_hi[nseg] = row + (cx - r);
                        if (++nseg == GRID_SEG_CHUNK) flush();
                    }
                    if (cx + r <= g.dx - 1) {
                        seg_lo[nseg] = seg_hi[nseg] = row + (cx + r);
                        if (++nseg == GRID_SEG_CHUNK) flush();
                    }
                }
            }
        }
        if (nseg) flush();

        const float bound = fminf(fminf(shell_bound_axis(qx, g.ox, g.cell, cx, r, g.dx),
                                        shell_bound_axis(qy, g.oy, g.cell, cy, r, g.dy)),
                                  shell_bound_axis(qz, g.oz, g.cell, cz, r, g.dz));
        if (bound == INF) return true;  // whole grid visited
        const float bs = bound - margin;
        if (bs > 0.0f && best.worst() < __fmul_rn(bs, bs)) return true;
        if (bs >= max_radius) return true;  // everything within max_radius has been seen
        if (r >= r_max) return false;
    }
}

#endif  // __CUDACC__

}  // namespace spx

// host-side handle
struct spx_index_s {
    spx_queue_t q = nullptr;
    size_t n_total = 0;   // points given to build
    uint32_t n = 0;       // finite points indexed
    float4* sorted = nullptr;
    uint32_t* start = nullptr;
    size_t ncells = 0;
    int64_t occupied = 0;
    spx::GridView view{};
};
