// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of fateshelled/sycl_points'
// per-iteration registration hot path (voxel grid -> KNN -> covariance -> linearise+reduce
// -> GN/LM/dog-leg update).  The product library never links or calls this file; tests,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it as the
// checker and as the timed CPU baseline ("port": the SYCL reference itself cannot be built
// in this image — no SYCL compiler, no Eigen).
//
// Parity pin: every known answer the reference's own tests hold for this path is checked
// in tests/test_oracle_golden.py (I/ = /root/reference/cpp/include/sycl_points/,
// T/ = /root/reference/cpp/tests/).  Covariance / linearise / H,b / pose values are not
// pinned by any reference test ("parity unpinned" there) — see DESIGN.md §oracle.
//
// Build: make -C oracle   (g++ -O2 -ffp-contract=off -fopenmp; no -ffast-math)
//
// Conventions: points/normals float[n][4]; covariances float[n][16] column-major 4x4
// (Eigen::Matrix4f layout, I/points/types.hpp:11-18); transforms float[16] column-major.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <queue>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_math.hpp"

using namespace orc;

namespace {

constexpr float FMAX = std::numeric_limits<float>::max();

inline M4 load_T(const float* T16) {
    M4 T;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) T(i, j) = T16[j * 4 + i];
    return T;
}
inline void store_T(const M4& T, float* T16) {
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) T16[j * 4 + i] = T(i, j);
}
inline V4 load_p(const float* p) {
    V4 v;
    for (int i = 0; i < 4; ++i) v(i) = p[i];
    return v;
}
inline M4 load_cov(const float* c) { return load_T(c); }

// I/algorithms/common/transform.hpp:32-37 (multiply<4,4>, eigen_utils.hpp:113-127)
inline V4 transform_point(const M4& T, const V4& p) { return mul<4, 4>(T, p); }

// squared distance as the KD-tree leaf scan computes it: dot<4>(q - p, q - p)
// I/algorithms/knn/kdtree.hpp:509-511, eigen_utils.hpp:245-253
inline float dist_sq(const V4& q, const float* p) {
    V4 d;
    for (int i = 0; i < 4; ++i) d(i) = q(i) - p[i];
    return dot<4>(d, d);
}

// ------------------------------------------------------------------ top-k containers
struct BestK {
    float* d;
    int32_t* id;
    int k;
    void init() {
        for (int i = 0; i < k; ++i) {
            d[i] = FMAX;
            id[i] = -1;
        }
    }
    // Oracle semantics: ordered by (dist, index) — what I/algorithms/knn/bruteforce.hpp:71-83
    // produces when targets are visited in index order with a strict '<'.
    inline void insert_lex(float ds, int32_t idx) {
        const int last = k - 1;
        if (!(ds < d[last] || (ds == d[last] && (id[last] < 0 || idx < id[last])))) return;
        int pos = last;
        while (pos > 0 && (ds < d[pos - 1] || (ds == d[pos - 1] && (id[pos - 1] < 0 || idx < id[pos - 1])))) {
            d[pos] = d[pos - 1];
            id[pos] = id[pos - 1];
            --pos;
        }
        d[pos] = ds;
        id[pos] = idx;
    }
    // Reference KD-tree semantics: reject when dist >= worst, strict '<' shifts
    // (I/algorithms/knn/kdtree.hpp:119-137) — ties keep the first visited.
    inline void insert_first_visited(float ds, int32_t idx) {
        if (k == 1) {
            if (ds < d[0]) {
                d[0] = ds;
                id[0] = idx;
            }
            return;
        }
        if (ds >= d[k - 1]) return;
        int pos = k - 1;
        while (pos > 0 && ds < d[pos - 1]) {
            d[pos] = d[pos - 1];
            id[pos] = id[pos - 1];
            --pos;
        }
        d[pos] = ds;
        id[pos] = idx;
    }
};

// ------------------------------------------------------------------ KD-tree
// I/algorithms/knn/kdtree.hpp:34-45
struct KDNode {
    float pt[4];
    int32_t idx;
    int32_t left = -1;
    int32_t right = -1;
    uint8_t axis = 0;
    uint8_t is_leaf = 0;
    uint8_t valid = 1;
    uint8_t pad = 0;
};

struct KDTree {
    std::vector<KDNode> nodes;
};

// I/algorithms/knn/kdtree.hpp:62-91
uint8_t find_axis_range(const float* pts, const std::vector<uint32_t>& ind, uint32_t start, uint32_t end) {
    const int64_t size = (int64_t)end - (int64_t)start + 1;
    if (size <= 1) return 0;
    float mn[3] = {FMAX, FMAX, FMAX};
    float mx[3] = {std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(),
                   std::numeric_limits<float>::lowest()};
    const size_t step = (size_t)std::max<int64_t>(size / 100, 1);
    for (size_t i = start; i <= end; i += step) {
        const float* p = pts + 4 * (size_t)ind[i];
        for (int a = 0; a < 3; ++a) {
            mn[a] = std::min(mn[a], p[a]);
            mx[a] = std::max(mx[a], p[a]);
        }
    }
    const float r0 = mx[0] - mn[0], r1 = mx[1] - mn[1], r2 = mx[2] - mn[2];
    if (r0 >= r1 && r0 >= r2) return 0;
    if (r1 >= r0 && r1 >= r2) return 1;
    return 2;
}

// I/algorithms/knn/kdtree.hpp:292-413  host median-split build with leaf blocks
KDTree* kd_build(const float* pts, size_t n, size_t leaf_threshold) {
    KDTree* t = new KDTree();
    if (n == 0) return t;
    std::vector<KDNode>& tree = t->nodes;
    tree.resize(n * 2 + 2);
    std::vector<uint32_t> ind(n);
    std::iota(ind.begin(), ind.end(), 0u);

    struct Task {
        uint32_t node, start, end;
    };
    std::vector<Task> stack;
    stack.reserve(64);
    stack.push_back({0u, 0u, (uint32_t)(n - 1)});
    uint32_t next = 1;

    while (!stack.empty()) {
        const Task task = stack.back();
        stack.pop_back();
        const uint32_t count = task.end - task.start + 1;
        if (task.start > task.end || count == 0) continue;
        if ((size_t)next + count + 2 > tree.size()) tree.resize(tree.size() * 2);

        if (count <= leaf_threshold) {
            const uint32_t leaf_start = next;
            next += count;
            for (uint32_t i = 0; i < count; ++i) {
                const uint32_t pi = ind[task.start + i];
                KDNode& m = tree[leaf_start + i];
                std::memcpy(m.pt, pts + 4 * (size_t)pi, 16);
                m.idx = (int32_t)pi;
                m.is_leaf = 1;
                m.axis = 0;
                m.left = m.right = -1;
                m.valid = 1;
            }
            KDNode& node = tree[task.node];
            node.is_leaf = 1;
            node.valid = 1;
            node.idx = -1;
            node.axis = 0;
            node.left = (int32_t)leaf_start;
            node.right = (int32_t)count;
            continue;
        }

        const uint8_t axis = find_axis_range(pts, ind, task.start, task.end);
        const uint32_t median = task.start + count / 2;
        std::nth_element(ind.begin() + task.start, ind.begin() + median, ind.begin() + task.end + 1,
                         [&](uint32_t a, uint32_t b) { return pts[4 * (size_t)a + axis] < pts[4 * (size_t)b + axis]; });
        const uint32_t pi = ind[median];
        int32_t left = -1, right = -1;
        if (task.start < median) {
            left = (int32_t)next++;
            stack.push_back({(uint32_t)left, task.start, median - 1});
        }
        if (median < task.end) {
            right = (int32_t)next++;
            stack.push_back({(uint32_t)right, median + 1, task.end});
        }
        KDNode& node = tree[task.node];
        node.is_leaf = 0;
        node.valid = 1;
        std::memcpy(node.pt, pts + 4 * (size_t)pi, 16);
        node.idx = (int32_t)pi;
        node.axis = axis;
        node.left = left;
        node.right = right;
    }
    tree.resize(next);
    return t;
}

struct StackEntry {
    int32_t node;
    float d;
};

// I/algorithms/knn/kdtree.hpp:463-553.  EXACT=false: the reference's search verbatim in
// behaviour (two 16-entry stacks, far pushes dropped when full, first-visited ties).
// EXACT=true: the oracle contract — unlimited stacks, '<=' far test, (dist,index) order.
template <bool EXACT>
void kd_search_one(const KDTree& t, const V4& q, int k, float* out_d, int32_t* out_i) {
    BestK best{out_d, out_i, k};
    best.init();
    const int32_t tree_size = (int32_t)t.nodes.size();
    constexpr int HALF = 16;
    StackEntry near_fixed[HALF], far_fixed[HALF];
    std::vector<StackEntry> near_v, far_v;
    int near_n = 0, far_n = 0;
    auto push_near = [&](StackEntry e) {
        if (EXACT) {
            near_v.push_back(e);
        } else if (near_n < HALF) {
            near_fixed[near_n++] = e;
        }
    };
    auto push_far = [&](StackEntry e) {
        if (EXACT) {
            far_v.push_back(e);
        } else if (far_n < HALF) {
            far_fixed[far_n++] = e;
        }
    };
    auto near_size = [&]() { return EXACT ? (int)near_v.size() : near_n; };
    auto far_size = [&]() { return EXACT ? (int)far_v.size() : far_n; };
    auto pop = [&]() {
        StackEntry e;
        if (near_size() > 0) {
            if (EXACT) {
                e = near_v.back();
                near_v.pop_back();
            } else {
                e = near_fixed[--near_n];
            }
        } else {
            if (EXACT) {
                e = far_v.back();
                far_v.pop_back();
            } else {
                e = far_fixed[--far_n];
            }
        }
        return e;
    };
    if (tree_size == 0) return;
    push_near({0, 0.0f});
    while (near_size() > 0 || far_size() > 0) {
        const StackEntry cur = pop();
        if (cur.d > best.d[k - 1]) continue;
        if (cur.node == -1 || cur.node >= tree_size) continue;
        const KDNode& node = t.nodes[cur.node];
        if (node.is_leaf != 0) {
            for (int32_t li = 0; li < node.right; ++li) {
                const KDNode& m = t.nodes[node.left + li];
                const float ds = m.valid ? dist_sq(q, m.pt) : FMAX;
                if (EXACT) {
                    if (m.valid) best.insert_lex(ds, m.idx);
                } else {
                    best.insert_first_visited(ds, m.idx);
                }
            }
            continue;
        }
        const float ds = node.valid ? dist_sq(q, node.pt) : FMAX;
        if (EXACT) {
            if (node.valid) best.insert_lex(ds, node.idx);
        } else {
            best.insert_first_visited(ds, node.idx);
        }
        const float ax = q(node.axis) - node.pt[node.axis];
        const int32_t nearer = (ax <= 0) ? node.left : node.right;
        const int32_t further = (ax <= 0) ? node.right : node.left;
        const float split = ax * ax;
        const bool go_far = EXACT ? (split <= best.d[k - 1]) : (split < best.d[k - 1]);
        if (go_far && further != -1) push_far({further, split});
        if (nearer != -1) push_near({nearer, 0.0f});
    }
}

// ------------------------------------------------------------------ covariance / normals
// I/algorithms/feature/covariance.hpp:16-47
inline M4 estimate_cov(const float* pts, int k, const int32_t* idx_row) {
    M4 ret = M4::zero();
    V3 sp = V3::zero();
    M3 so = M3::zero();
    size_t cnt = 0;
    for (int j = 0; j < k; ++j) {
        const int32_t id = idx_row[j];
        if (id < 0) continue;
        V3 p;
        p(0) = pts[4 * (size_t)id + 0];
        p(1) = pts[4 * (size_t)id + 1];
        p(2) = pts[4 * (size_t)id + 2];
        for (int a = 0; a < 3; ++a) sp(a) += p(a);
        const M3 o = outer<3>(p, p);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) so(r, c) += o(r, c);
        ++cnt;
    }
    if (cnt < 4) {
        ret(0, 0) = ret(1, 1) = ret(2, 2) = 1.0f;
        return ret;
    }
    const float inv = 1.0f / cnt;
    const V3 mean = scale(sp, inv);
    const M3 c = ensure_symmetric<3>(sub(scale(so, inv), outer<3>(mean, mean)));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ret(i, j) = c(i, j);
    return ret;
}

inline M3 block3(const M4& m) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = m(i, j);
    return r;
}

// I/algorithms/feature/covariance.hpp:49-65
inline V4 extract_normal(const float* pt, const M4& cov) {
    V3 ev;
    M3 evec;
    eigen3(block3(cov), ev, evec);
    V3 n, p;
    for (int i = 0; i < 3; ++i) {
        n(i) = evec(i, 0);
        p(i) = pt[i];
    }
    V4 out;
    if (dot<3>(n, p) <= 1.0) {
        out(0) = n(0); out(1) = n(1); out(2) = n(2);
    } else {
        out(0) = -n(0); out(1) = -n(1); out(2) = -n(2);
    }
    out(3) = 0.0f;
    return out;
}

// I/algorithms/feature/covariance.hpp:67-74
inline void update_covariance_plane(M4& cov) {
    V3 ev;
    M3 evec;
    eigen3(block3(cov), ev, evec);
    M3 D = M3::zero();
    D(0, 0) = 1e-3f; D(1, 1) = 1.0f; D(2, 2) = 1.0f;
    const M3 r = mul<3, 3, 3>(mul<3, 3, 3>(evec, D), transpose(evec));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov(i, j) = r(i, j);
}

// I/utils/eigen_utils.hpp:646-677  log / exp of a symmetric 3x3 through its eigen-decomposition
inline M3 spd_function(const M3& A, bool is_log) {
    V3 ev;
    M3 evec;
    eigen3(A, ev, evec);
    M3 D = M3::zero();
    for (int i = 0; i < 3; ++i) D(i, i) = is_log ? cr_log(std::fmax(ev(i), 1e-6f)) : cr_exp(ev(i));
    return ensure_symmetric<3>(mul<3, 3, 3>(mul<3, 3, 3>(evec, D), transpose(evec)));
}

// ------------------------------------------------------------------ robust kernels
enum Loss { L_NONE = 0, L_HUBER, L_TUKEY, L_CAUCHY, L_GM };
enum Reg { R_P2P = 0, R_P2PLANE = 1, R_P2D = 2, R_GICP = 3, R_GENZ = 4 };

// I/algorithms/robust/robust.hpp:56-90
inline float robust_weight(int loss, float r, float s) {
    if (loss == L_NONE) return 1.0f;
    if (r <= 1e-8f) return 1.0f;
    const float x = r / s;
    switch (loss) {
        case L_HUBER: return std::min(1.0f, 1.0f / x);
        case L_TUKEY: {
            if (x >= 1.0f) return 0.0f;
            const float f = 1.0f - x * x;
            return f * f;
        }
        case L_CAUCHY: return 1.0f / (1.0f + x * x);
        case L_GM: {
            const float d = 1.0f + x * x;
            return 1.0f / (d * d);
        }
    }
    return 1.0f;
}

// I/algorithms/robust/robust.hpp:96-114
inline float robust_error(int loss, float r, float s) {
    switch (loss) {
        case L_NONE: return 0.5f * r * r;
        case L_HUBER: return r <= s ? 0.5f * r * r : s * (r - 0.5f * s);
        case L_TUKEY:
            return r <= s ? (s * s / 6.0f) * (1.0f - cr_cube(1.0f - ((r * r) / (s * s)))) : s * s / 6.0f;
        case L_CAUCHY: return 0.5f * s * s * cr_log(1.0f + ((r * r) / (s * s)));
        case L_GM: return 0.5f * (s * s * r * r) / (s * s + r * r);
    }
    return 0.5f * r * r;
}

// ------------------------------------------------------------------ factors
struct PointTerm {
    M6 H;
    V6 b;
    float sq_err;
    float res_norm;
};

// ---- M-estimated covariance — I/algorithms/feature/covariance.hpp:97-134 (estimate_weighted), :143-173
// (compute_median), :182-222 (estimate_robust)
inline bool estimate_cov_weighted(const float* pts, int k, const int32_t* idx_row, const float* w, M4& ret, V3& mean) {
    ret = M4::zero();
    V3 sp = V3::zero();
    M3 so = M3::zero();
    size_t cnt = 0;
    float tw = 0.0f;
    for (int j = 0; j < k; ++j) {
        const int32_t id = idx_row[j];
        if (id < 0) continue;
        V3 p;
        p(0) = pts[4 * (size_t)id + 0];
        p(1) = pts[4 * (size_t)id + 1];
        p(2) = pts[4 * (size_t)id + 2];
        for (int a = 0; a < 3; ++a) sp(a) += p(a) * w[j];
        const M3 o = outer<3>(p, p);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) so(r, c) += o(r, c) * w[j];
        ++cnt;
        tw += w[j];
    }
    if (cnt < 4 || tw < std::numeric_limits<float>::epsilon()) {
        ret(0, 0) = ret(1, 1) = ret(2, 2) = 1.0f;
        return false;
    }
    mean = scale(sp, 1.0f / tw);
    const M3 c = ensure_symmetric<3>(sub(scale(so, 1.0f / tw), outer<3>(mean, mean)));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ret(i, j) = c(i, j);
    return true;
}

inline M4 estimate_cov_robust(const float* pts, int k, const int32_t* idx_row, int loss, float mad_scale, float min_scale,
                              int max_iter) {
    constexpr int MAX_K = 64;
    float w[MAX_K], d2[MAX_K];
    std::fill(w, w + MAX_K, 1.0f);
    std::fill(d2, d2 + MAX_K, 0.0f);
    M4 cov;
    V3 mean = V3::zero();
    bool ok = estimate_cov_weighted(pts, k, idx_row, w, cov, mean);
    for (int it = 0; ok && it < max_iter; ++it) {
        const M3 ci = inverse(block3(cov));
        for (int j = 0; j < k; ++j) {
            const int32_t id = idx_row[j];
            if (id < 0) continue;
            V3 d;
            for (int a = 0; a < 3; ++a) d(a) = pts[4 * (size_t)id + a] - mean(a);
            d2[j] = dot<3>(d, mul<3, 3>(ci, d));  // dot<4> / multiply<4,4> with zero 4th components: the same chain
        }
        for (int j = 0; j < k; ++j) w[j] = d2[j];
        for (int a = 1; a < k; ++a) {  // insertion sort in the weights buffer
            const float key = w[a];
            int b = a;
            while (b > 0 && w[b - 1] > key) {
                w[b] = w[b - 1];
                --b;
            }
            w[b] = key;
        }
        const int mid = k / 2;
        const float median = (k % 2 == 0) ? (w[mid - 1] + w[mid]) * 0.5f : w[mid];
        float rs = mad_scale * median;
        if (rs < min_scale) rs = min_scale;
        for (int j = 0; j < k; ++j) w[j] = robust_weight(loss, d2[j], rs);
        ok = estimate_cov_weighted(pts, k, idx_row, w, cov, mean);
    }
    return cov;
}

// I/algorithms/registration/factor.hpp:69-84 (+ :100-104 with identity weight: exact no-op)
inline Mat<4, 6> se3_jacobian(const M4& T, const V4& p) {
    Mat<4, 6> J = Mat<4, 6>::zero();
    const M3 S = skew(p(0), p(1), p(2));
    const M3 R = block3(T);
    const M3 RS = mul<3, 3, 3>(R, S);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            J(i, j) = RS(i, j);
            J(i, 3 + j) = -R(i, j);
        }
    return J;
}

inline V4 residual_of(const M4& T, const V4& ps, const V4& pt) {
    const V4 tp = transform_point(T, ps);
    V4 r;
    r(0) = pt(0) - tp(0); r(1) = pt(1) - tp(1); r(2) = pt(2) - tp(2); r(3) = 0.0f;
    return r;
}

// I/algorithms/registration/factor.hpp:130-149
inline PointTerm lin_p2p(const M4& T, const V4& ps, const V4& pt) {
    const V4 r = residual_of(T, ps, pt);
    const Mat<4, 6> J = se3_jacobian(T, ps);
    const Mat<6, 4> JT = transpose(J);
    PointTerm o;
    o.H = ensure_symmetric<6>(mul<6, 4, 6>(JT, J));
    o.b = mul<6, 4>(JT, r);
    o.sq_err = dot<4>(r, r);
    o.res_norm = std::sqrt(o.sq_err);
    return o;
}

// I/algorithms/registration/factor.hpp:172-210
inline PointTerm lin_p2plane(const M4& T, const V4& ps, const V4& pt, const V4& nrm) {
    const V4 r = residual_of(T, ps, pt);
    V3 n, r3;
    for (int i = 0; i < 3; ++i) {
        n(i) = nrm(i);
        r3(i) = r(i);
    }
    const float d = dot<3>(n, r3);
    V4 pe = V4::zero();
    for (int i = 0; i < 3; ++i) pe(i) = n(i) * d;
    const Mat<4, 6> J0 = se3_jacobian(T, ps);
    Mat<1, 3> nT;
    Mat<3, 6> J3;
    Mat<3, 1> nc;
    for (int i = 0; i < 3; ++i) {
        nT(0, i) = n(i);
        nc(i, 0) = n(i);
        for (int j = 0; j < 6; ++j) J3(i, j) = J0(i, j);
    }
    const Mat<1, 6> row = mul<1, 3, 6>(nT, J3);
    const Mat<3, 6> Jp = mul<3, 1, 6>(nc, row);
    Mat<4, 6> J = Mat<4, 6>::zero();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) J(i, j) = Jp(i, j);
    const Mat<6, 4> JT = transpose(J);
    PointTerm o;
    o.H = ensure_symmetric<6>(mul<6, 4, 6>(JT, J));
    o.b = mul<6, 4>(JT, pe);
    o.sq_err = d * d;
    o.res_norm = std::fabs(d);
    return o;
}

// I/algorithms/common/transform.hpp:14-22
inline M4 transform_cov(const M4& C, const M4& T) { return mul<4, 4, 4>(T, mul<4, 4, 4>(C, transpose(T))); }

// ---- GenZ (factor.hpp:378-449): a correspondence whose TARGET neighbourhood is planar (PCA normalised curvature
// l0 / (l0 + l1 + l2) below the threshold) takes the point-to-plane factor weighted by alpha, the others the
// point-to-point factor weighted by 1 - alpha; alpha = planar inliers / inliers of the current correspondences
// (registration.hpp:464-511, recomputed at every linearisation :519).
static float g_genz_planarity_threshold = 0.2f;  // RegistrationParams::genz.planarity_threshold (registration_params.hpp:51-53)

// I/algorithms/registration/factor.hpp:378-385
inline float genz_curvature(const M4& ct) {
    V3 ev;
    M3 evec;
    eigen3(block3(ct), ev, evec);
    const float sum = ev(0) + ev(1) + ev(2);
    return (sum > 1e-12f) ? ev(0) / sum : 1.0f;
}
inline bool genz_planar(const M4& ct) { return genz_curvature(ct) < g_genz_planarity_threshold; }  // :391-393

// I/algorithms/registration/registration.hpp:464-511
inline float genz_alpha_of(const float* tgt_covs, size_t ns, const int32_t* idx, const float* dist, float max_corr_sq) {
    uint32_t inl = 0, plane = 0;
    const M4 ident = M4::identity();
    for (size_t i = 0; i < ns; ++i) {
        if (dist[i] > max_corr_sq) continue;
        if (genz_planar(tgt_covs ? load_cov(tgt_covs + 16 * (size_t)idx[i]) : ident)) ++plane;
        ++inl;
    }
    return inl == 0 ? 1.0f : (float)plane / (float)inl;
}

// I/algorithms/registration/factor.hpp:111-123 + covariance.hpp:136-141
inline M4 gicp_mahalanobis_inv(const M4& cs_in, const M4& ct_in, const M4& T) {
    M4 cs = cs_in, ct = ct_in;
    update_covariance_plane(cs);
    update_covariance_plane(ct);
    const M4 tc = transform_cov(cs, T);
    const M3 rcr = add(block3(tc), block3(ct));
    const M3 inv = inverse(rcr);
    M4 out = M4::zero();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out(i, j) = inv(i, j);
    return out;
}

// I/algorithms/registration/factor.hpp:239-278
inline PointTerm lin_gicp(const M4& T, const V4& ps, const M4& cs, const V4& pt, const M4& ct) {
    const V4 r = residual_of(T, ps, pt);
    const M4 Minv = gicp_mahalanobis_inv(cs, ct, T);
    const Mat<4, 6> J = se3_jacobian(T, ps);
    const Mat<6, 4> JTM = mul<6, 4, 4>(transpose(J), Minv);
    PointTerm o;
    o.H = ensure_symmetric<6>(mul<6, 4, 6>(JTM, J));
    o.b = mul<6, 4>(JTM, r);
    o.sq_err = dot<4>(r, mul<4, 4>(Minv, r));
    o.res_norm = std::sqrt(o.sq_err);
    return o;
}

// I/algorithms/registration/factor.hpp:311-317: Mahalanobis = inverse of the RAW target covariance
inline M4 p2d_mahalanobis(const M4& ct) {
    const M3 inv = inverse(block3(ct));
    M4 out = M4::zero();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out(i, j) = inv(i, j);
    return out;
}

// I/algorithms/registration/factor.hpp:326-354 (identity-weighted Jacobian = the plain SE(3) Jacobian)
inline PointTerm lin_p2d(const M4& T, const V4& ps, const V4& pt, const M4& ct) {
    const V4 r = residual_of(T, ps, pt);
    const M4 Minv = p2d_mahalanobis(ct);
    const Mat<4, 6> J = se3_jacobian(T, ps);
    const Mat<6, 4> JTM = mul<6, 4, 4>(transpose(J), Minv);
    PointTerm o;
    o.H = ensure_symmetric<6>(mul<6, 4, 6>(JTM, J));
    o.b = mul<6, 4>(JTM, r);
    o.sq_err = dot<4>(r, mul<4, 4>(Minv, r));
    o.res_norm = std::sqrt(o.sq_err);
    return o;
}

// I/algorithms/registration/factor.hpp:156-164, 218-230, 287-306, 362-373
inline float err_only(int reg, const M4& T, const V4& ps, const M4& cs, const V4& pt, const M4& ct, const V4& nrm) {
    const V4 r = residual_of(T, ps, pt);
    if (reg == R_P2P || (reg == R_GENZ && !genz_planar(ct))) return dot<4>(r, r);
    if (reg == R_P2PLANE || reg == R_GENZ) {
        V3 n, r3;
        for (int i = 0; i < 3; ++i) {
            n(i) = nrm(i);
            r3(i) = r(i);
        }
        const float d = dot<3>(n, r3);
        return d * d;
    }
    const M4 Minv = reg == R_P2D ? p2d_mahalanobis(ct) : gicp_mahalanobis_inv(cs, ct, T);
    return dot<4>(r, mul<4, 4>(Minv, r));
}

// ---- rotation constraint (rotation_constraint.hpp:15-121): Jensen-Bregman LogDet divergence between the rotated
// source covariance and the target covariance, D = log det(0.5 (R Cs R^T + Ct)) - 0.5 (log det Cs + log det Ct),
// residual r = max(D, 0), Jacobian (rotation block only) g = -R^T vex([Cs', M^-1]); added to every correspondence's
// H / b / error with weight rotation_constraint.weight x robust weight (registration.hpp:629-649,757-764).
static int g_rot_enable = 0;
static float g_rot_weight = 1.0f, g_rot_scale = 10.0f;

inline float rot_logdet(const M3& m) { return cr_log(std::fmax(determinant(m), 1e-10f)); }

struct RotTerm {
    float D;
    V3 grad;
};
inline RotTerm rot_divergence(const M4& cs4, const M4& ct4, const M4& T) {
    const M3 R = block3(T), Cs = block3(cs4), Ct = block3(ct4);
    const M3 Csp = mul<3, 3, 3>(R, mul<3, 3, 3>(Cs, transpose(R)));
    const M3 M = scale(add(Csp, Ct), 0.5f);
    const float log_det_M = rot_logdet(M);
    const float log_det_ref = 0.5f * (rot_logdet(Cs) + rot_logdet(Ct));
    RotTerm o;
    o.D = std::fmax(log_det_M - log_det_ref, 0.0f);
    const M3 Minv = inverse(M);
    const M3 comm = sub(mul<3, 3, 3>(Csp, Minv), mul<3, 3, 3>(Minv, Csp));
    V3 g;
    g(0) = -0.5f * (comm(2, 1) - comm(1, 2));
    g(1) = -0.5f * (comm(0, 2) - comm(2, 0));
    g(2) = -0.5f * (comm(1, 0) - comm(0, 1));
    o.grad = mul<3, 3>(transpose(R), g);
    return o;
}
// linearize_rotation_constraint_logdet — rotation_constraint.hpp:84-105
inline PointTerm lin_rot(const M4& cs, const M4& ct, const M4& T) {
    const RotTerm d = rot_divergence(cs, ct, T);
    PointTerm o;
    o.H = M6::zero();
    o.b = V6::zero();
    for (int i = 0; i < 3; ++i) {
        o.b(i) = d.D * d.grad(i);
        for (int j = 0; j < 3; ++j) o.H(i, j) = d.grad(i) * d.grad(j);
    }
    o.sq_err = 0.5f * d.D * d.D;
    o.res_norm = std::sqrt(o.sq_err);
    return o;
}
// calculate_logdet_divergence_squared — rotation_constraint.hpp:15-44
inline float err_rot(const M4& cs, const M4& ct, const M4& T) {
    const float D = rot_divergence(cs, ct, T).D;
    return 0.5f * D * D;
}

// linearize_geometry<GENZ> — factor.hpp:425-443: the selected factor's H, b scaled by the GenZ weight, the residual
// norm left unweighted; returns the weight
inline float genz_term(const M4& T, const V4& ps, const V4& pt, const M4& ct, const V4& nrm, float alpha, PointTerm& t) {
    const bool planar = genz_planar(ct);
    const float gw = planar ? alpha : (1.0f - alpha);
    t = planar ? lin_p2plane(T, ps, pt, nrm) : lin_p2p(T, ps, pt);
    for (int r = 0; r < 6; ++r) {
        for (int q = 0; q < 6; ++q) t.H(r, q) = t.H(r, q) * gw;
        t.b(r) = t.b(r) * gw;
    }
    t.sq_err = t.sq_err * gw;
    return gw;
}

struct Clouds {
    const float* src_pts;
    const float* src_covs;  // nullable -> identity
    size_t ns;
    const float* tgt_pts;
    const float* tgt_covs;     // nullable -> identity
    const float* tgt_normals;  // nullable -> zero
};

struct Linearized {
    float H[36];
    float b[6];
    float error;
    uint32_t inlier;
};

// I/algorithms/registration/registration.hpp:513-664 (gating :584, weighting :613-620,
// error :626, reduction :653-659).  mode 0: fp32 running sum in index order; mode 1: each
// per-point fp32 term accumulated in fp64 (the "truth" the GPU is compared against).
Linearized linearize(int reg, int loss, const Clouds& c, const int32_t* idx, const float* dist, const M4& T,
                     float max_corr_sq, float scale, int mode) {
    const M4 ident = M4::identity();
    const V4 zero4 = V4::zero();
    const float alpha = reg == R_GENZ ? genz_alpha_of(c.tgt_covs, c.ns, idx, dist, max_corr_sq) : 1.0f;
    Linearized out;
    if (mode == 0) {
        float H[36] = {0}, b[6] = {0}, err = 0.0f;
        uint32_t inl = 0;
        for (size_t i = 0; i < c.ns; ++i) {
            if (dist[i] > max_corr_sq) continue;
            const int32_t ti = idx[i];
            const V4 ps = load_p(c.src_pts + 4 * i);
            const V4 pt = load_p(c.tgt_pts + 4 * (size_t)ti);
            PointTerm t;
            float gw = 1.0f;  // GenZ weight of this correspondence (1 for every other factor)
            if (reg == R_P2P) {
                t = lin_p2p(T, ps, pt);
            } else if (reg == R_P2PLANE) {
                t = lin_p2plane(T, ps, pt, c.tgt_normals ? load_p(c.tgt_normals + 4 * (size_t)ti) : zero4);
            } else if (reg == R_P2D) {
                t = lin_p2d(T, ps, pt, c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident);
            } else if (reg == R_GENZ) {
                gw = genz_term(T, ps, pt, c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident,
                               c.tgt_normals ? load_p(c.tgt_normals + 4 * (size_t)ti) : zero4, alpha, t);
            } else {
                t = lin_gicp(T, ps, c.src_covs ? load_cov(c.src_covs + 16 * i) : ident, pt,
                             c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident);
            }
            const float w = robust_weight(loss, t.res_norm, scale);
            for (int r = 0; r < 6; ++r) {
                for (int q = 0; q < 6; ++q) H[r * 6 + q] += w * t.H(r, q);
                b[r] += w * t.b(r);
            }
            err += gw * robust_error(loss, t.res_norm, scale);
            if (g_rot_enable) {  // registration.hpp:629-649
                const PointTerm rt = lin_rot(c.src_covs ? load_cov(c.src_covs + 16 * i) : ident,
                                             c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident, T);
                const float wr = robust_weight(loss, rt.res_norm, g_rot_scale);
                for (int r = 0; r < 6; ++r) {
                    for (int q = 0; q < 6; ++q) H[r * 6 + q] += g_rot_weight * wr * rt.H(r, q);
                    b[r] += g_rot_weight * wr * rt.b(r);
                }
                err += g_rot_weight * robust_error(loss, rt.res_norm, g_rot_scale);
            }
            ++inl;
        }
        std::memcpy(out.H, H, sizeof(H));
        std::memcpy(out.b, b, sizeof(b));
        out.error = err;
        out.inlier = inl;
        return out;
    }
    double H[36] = {0}, b[6] = {0}, err = 0.0;
    uint64_t inl = 0;
#pragma omp parallel
    {
        double Hl[36] = {0}, bl[6] = {0}, el = 0.0;
        uint64_t il = 0;
#pragma omp for schedule(static)
        for (int64_t ii = 0; ii < (int64_t)c.ns; ++ii) {
            const size_t i = (size_t)ii;
            if (dist[i] > max_corr_sq) continue;
            const int32_t ti = idx[i];
            const V4 ps = load_p(c.src_pts + 4 * i);
            const V4 pt = load_p(c.tgt_pts + 4 * (size_t)ti);
            PointTerm t;
            float gw = 1.0f;  // GenZ weight of this correspondence (1 for every other factor)
            if (reg == R_P2P) {
                t = lin_p2p(T, ps, pt);
            } else if (reg == R_P2PLANE) {
                t = lin_p2plane(T, ps, pt, c.tgt_normals ? load_p(c.tgt_normals + 4 * (size_t)ti) : zero4);
            } else if (reg == R_P2D) {
                t = lin_p2d(T, ps, pt, c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident);
            } else if (reg == R_GENZ) {
                gw = genz_term(T, ps, pt, c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident,
                               c.tgt_normals ? load_p(c.tgt_normals + 4 * (size_t)ti) : zero4, alpha, t);
            } else {
                t = lin_gicp(T, ps, c.src_covs ? load_cov(c.src_covs + 16 * i) : ident, pt,
                             c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident);
            }
            const float w = robust_weight(loss, t.res_norm, scale);
            for (int r = 0; r < 6; ++r) {
                for (int q = 0; q < 6; ++q) Hl[r * 6 + q] += (double)(w * t.H(r, q));
                bl[r] += (double)(w * t.b(r));
            }
            el += (double)(gw * robust_error(loss, t.res_norm, scale));
            if (g_rot_enable) {  // registration.hpp:629-649
                const PointTerm rt = lin_rot(c.src_covs ? load_cov(c.src_covs + 16 * i) : ident,
                                             c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident, T);
                const float wr = robust_weight(loss, rt.res_norm, g_rot_scale);
                for (int r = 0; r < 6; ++r) {
                    for (int q = 0; q < 6; ++q) Hl[r * 6 + q] += (double)(g_rot_weight * wr * rt.H(r, q));
                    bl[r] += (double)(g_rot_weight * wr * rt.b(r));
                }
                el += (double)(g_rot_weight * robust_error(loss, rt.res_norm, g_rot_scale));
            }
            ++il;
        }
#pragma omp critical
        {
            for (int q = 0; q < 36; ++q) H[q] += Hl[q];
            for (int q = 0; q < 6; ++q) b[q] += bl[q];
            err += el;
            inl += il;
        }
    }
    for (int q = 0; q < 36; ++q) out.H[q] = (float)H[q];
    for (int q = 0; q < 6; ++q) out.b[q] = (float)b[q];
    out.error = (float)err;
    out.inlier = (uint32_t)inl;
    return out;
}

// I/algorithms/registration/registration.hpp:678-777
void error_sum(int reg, int loss, const Clouds& c, const int32_t* idx, const float* dist, const M4& T,
               float max_corr_sq, float scale, int mode, float* err_out, uint32_t* inl_out) {
    const M4 ident = M4::identity();
    const V4 zero4 = V4::zero();
    // GenZ: alpha of the (frozen) correspondences — what the last linearisation on them stored (registration.hpp:370,519)
    const float alpha = reg == R_GENZ ? genz_alpha_of(c.tgt_covs, c.ns, idx, dist, max_corr_sq) : 1.0f;
    auto genz_w = [&](int32_t ti) {
        if (reg != R_GENZ) return 1.0f;
        return genz_planar(c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident) ? alpha : (1.0f - alpha);
    };
    if (mode == 0) {
        float err = 0.0f;
        uint32_t inl = 0;
        for (size_t i = 0; i < c.ns; ++i) {
            if (dist[i] > max_corr_sq) continue;
            const int32_t ti = idx[i];
            const float e2 = err_only(reg, T, load_p(c.src_pts + 4 * i),
                                      c.src_covs ? load_cov(c.src_covs + 16 * i) : ident,
                                      load_p(c.tgt_pts + 4 * (size_t)ti),
                                      c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident,
                                      c.tgt_normals ? load_p(c.tgt_normals + 4 * (size_t)ti) : zero4);
            err += genz_w(ti) * robust_error(loss, std::sqrt(e2), scale);
            if (g_rot_enable)
                err += g_rot_weight * robust_error(loss, std::sqrt(err_rot(c.src_covs ? load_cov(c.src_covs + 16 * i) : ident,
                                                                           c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident, T)),
                                                   g_rot_scale);
            ++inl;
        }
        *err_out = err;
        *inl_out = inl;
        return;
    }
    double err = 0.0;
    uint64_t inl = 0;
#pragma omp parallel for schedule(static) reduction(+ : err, inl)
    for (int64_t ii = 0; ii < (int64_t)c.ns; ++ii) {
        const size_t i = (size_t)ii;
        if (dist[i] > max_corr_sq) continue;
        const int32_t ti = idx[i];
        const float e2 =
            err_only(reg, T, load_p(c.src_pts + 4 * i), c.src_covs ? load_cov(c.src_covs + 16 * i) : ident,
                     load_p(c.tgt_pts + 4 * (size_t)ti), c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident,
                     c.tgt_normals ? load_p(c.tgt_normals + 4 * (size_t)ti) : zero4);
        err += (double)(genz_w(ti) * robust_error(loss, std::sqrt(e2), scale));
        if (g_rot_enable)
            err += (double)(g_rot_weight * robust_error(loss, std::sqrt(err_rot(c.src_covs ? load_cov(c.src_covs + 16 * i) : ident,
                                                                               c.tgt_covs ? load_cov(c.tgt_covs + 16 * (size_t)ti) : ident, T)),
                                                       g_rot_scale));
        ++inl;
    }
    *err_out = (float)err;
    *inl_out = (uint32_t)inl;
}

// ------------------------------------------------------------------ 6x6 solves
// Eigen::LDLT<Matrix<float,6,6>> call sites: registration.hpp:791-801, dogleg_step.hpp:43-50.
// Eigen is not in the image (unpinned third-party, needs >= 3.4); restated as the same
// algorithm family — LDL^T with largest-diagonal pivoting — evaluated in fp64 and cast.
struct LDLT6 {
    double L[6][6];
    double D[6];
    int perm[6];
    bool ok;
    void compute(const double A_in[6][6]) {
        double A[6][6];
        std::memcpy(A, A_in, sizeof(A));
        for (int i = 0; i < 6; ++i) perm[i] = i;
        ok = true;
        for (int k = 0; k < 6; ++k) {
            int piv = k;
            double best = std::fabs(A[k][k]);
            for (int i = k + 1; i < 6; ++i)
                if (std::fabs(A[i][i]) > best) {
                    best = std::fabs(A[i][i]);
                    piv = i;
                }
            if (piv != k) {
                for (int j = 0; j < 6; ++j) std::swap(A[k][j], A[piv][j]);
                for (int j = 0; j < 6; ++j) std::swap(A[j][k], A[j][piv]);
                std::swap(perm[k], perm[piv]);
            }
            const double d = A[k][k];
            if (d == 0.0) {
                for (int i = k + 1; i < 6; ++i)
                    if (A[i][k] != 0.0) ok = false;
                for (int i = k + 1; i < 6; ++i) A[i][k] = 0.0;
                continue;
            }
            for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
            for (int i = k + 1; i < 6; ++i)
                for (int j = k + 1; j <= i; ++j) {
                    A[i][j] -= A[i][k] * d * A[j][k];
                    A[j][i] = A[i][j];
                }
        }
        for (int i = 0; i < 6; ++i) {
            D[i] = A[i][i];
            for (int j = 0; j < 6; ++j) L[i][j] = (j < i) ? A[i][j] : (i == j ? 1.0 : 0.0);
        }
    }
    void solve(const double rhs[6], double x[6]) const {
        double y[6];
        for (int i = 0; i < 6; ++i) y[i] = rhs[perm[i]];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < i; ++j) y[i] -= L[i][j] * y[j];
        for (int i = 0; i < 6; ++i) y[i] = (std::fabs(D[i]) > std::numeric_limits<double>::min()) ? y[i] / D[i] : 0.0;
        for (int i = 5; i >= 0; --i)
            for (int j = i + 1; j < 6; ++j) y[i] -= L[j][i] * y[j];
        for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
    }
    double min_d() const { return *std::min_element(D, D + 6); }
};

// registration.hpp:791-801  (H + lambda I) delta = -b
inline bool solve_system(const float* H36, const float* b6, float lambda, float* delta) {
    double A[6][6], rhs[6], x[6];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) A[i][j] = (double)H36[i * 6 + j];
        A[i][i] = (double)(H36[i * 6 + i] + lambda);
        rhs[i] = -(double)b6[i];
    }
    LDLT6 f;
    f.compute(A);
    if (!f.ok) {
        for (int i = 0; i < 6; ++i) delta[i] = 0.0f;
        return false;
    }
    f.solve(rhs, x);
    for (int i = 0; i < 6; ++i) delta[i] = (float)x[i];
    return true;
}

inline float norm3(const float* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

struct Dogleg {
    float p[6];
    float step_norm;
    float predicted_reduction;
};

// I/algorithms/registration/dogleg_step.hpp:34-102 (N = 6)
Dogleg dogleg_step(const float* H, const float* g, float radius) {
    Dogleg r;
    std::memset(&r, 0, sizeof(r));
    float p_gn[6] = {0};
    float n_gn = 0.0f;
    bool valid_gn = false;
    {
        double A[6][6], rhs[6], x[6];
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) A[i][j] = (double)H[i * 6 + j];
            rhs[i] = -(double)g[i];
        }
        LDLT6 f;
        f.compute(A);
        if (f.ok && f.min_d() > 0.0) {
            f.solve(rhs, x);
            float s = 0.0f;
            for (int i = 0; i < 6; ++i) {
                p_gn[i] = (float)x[i];
                s += p_gn[i] * p_gn[i];
            }
            n_gn = std::sqrt(s);
            valid_gn = std::isfinite(n_gn);
        }
    }
    float g2 = 0.0f, Hg[6], gHg = 0.0f;
    for (int i = 0; i < 6; ++i) g2 += g[i] * g[i];
    for (int i = 0; i < 6; ++i) {
        float s = 0.0f;
        for (int j = 0; j < 6; ++j) s += H[i * 6 + j] * g[j];
        Hg[i] = s;
    }
    for (int i = 0; i < 6; ++i) gHg += g[i] * Hg[i];
    float p_sd[6];
    for (int i = 0; i < 6; ++i) p_sd[i] = -g[i];
    if (gHg > std::numeric_limits<float>::epsilon()) {
        const float alpha = g2 / gHg;
        if (std::isfinite(alpha))
            for (int i = 0; i < 6; ++i) p_sd[i] = -alpha * g[i];
    }
    float n_sd = 0.0f;
    for (int i = 0; i < 6; ++i) n_sd += p_sd[i] * p_sd[i];
    const float n_sd2 = n_sd;
    n_sd = std::sqrt(n_sd);

    if (valid_gn && n_gn <= radius) {
        std::memcpy(r.p, p_gn, sizeof(p_gn));
        r.step_norm = n_gn;
    } else if (n_sd >= radius) {
        if (n_sd > std::numeric_limits<float>::epsilon())
            for (int i = 0; i < 6; ++i) r.p[i] = (radius / n_sd) * p_sd[i];
        r.step_norm = radius;
    } else if (valid_gn) {
        float diff[6], a = 0.0f, bq = 0.0f;
        for (int i = 0; i < 6; ++i) {
            diff[i] = p_gn[i] - p_sd[i];
            a += diff[i] * diff[i];
            bq += p_sd[i] * diff[i];
        }
        bq *= 2.0f;
        const float cq = n_sd2 - radius * radius;
        float disc = std::max(bq * bq - 4.0f * a * cq, 0.0f);
        float tau = 0.0f;
        if (a > std::numeric_limits<float>::epsilon()) tau = (-bq + std::sqrt(disc)) / (2.0f * a);
        tau = std::clamp(tau, 0.0f, 1.0f);
        float s = 0.0f;
        for (int i = 0; i < 6; ++i) {
            r.p[i] = p_sd[i] + tau * diff[i];
            s += r.p[i] * r.p[i];
        }
        r.step_norm = std::sqrt(s);
    } else {
        std::memcpy(r.p, p_sd, sizeof(p_sd));
        if (n_sd > radius && n_sd > std::numeric_limits<float>::epsilon()) {
            for (int i = 0; i < 6; ++i) r.p[i] *= radius / n_sd;
            r.step_norm = radius;
        } else {
            r.step_norm = n_sd;
        }
    }
    float gp = 0.0f, pHp = 0.0f;
    for (int i = 0; i < 6; ++i) {
        gp += g[i] * r.p[i];
        float s = 0.0f;
        for (int j = 0; j < 6; ++j) s += H[i * 6 + j] * r.p[j];
        pHp += r.p[i] * s;
    }
    r.predicted_reduction = -(gp + 0.5f * pHp);
    return r;
}

}  // namespace

// =====================================================================================
// C interface (ctypes)
// =====================================================================================
extern "C" {

struct orc_reg_params {
    int32_t reg_type;        // 0 P2P, 1 P2PLANE, 2 P2D, 3 GICP, 4 GENZ   (factor.hpp:18-32)
    int32_t loss;            // 0 NONE 1 HUBER 2 TUKEY 3 CAUCHY 4 GEMAN_MCCLURE (robust.hpp:14-20)
    int32_t opt_method;      // 0 GN, 1 LM, 2 dog-leg      (registration_params.hpp:17-21)
    int32_t max_iterations;  // 20
    float max_corr_dist;     // 2.0
    float robust_default_scale;  // 10
    float crit_translation;  // 1e-3
    float crit_rotation;     // 1e-3
    float gn_lambda;         // 1.0
    int32_t lm_max_inner;    // 10
    float lm_lambda_factor;  // 2
    float lm_init_lambda;    // 1
    float lm_max_lambda;     // 1e3
    float lm_min_lambda;     // 1e-6
    float dl_init_radius;    // 1
    float dl_min_radius;     // 1e-4
    float dl_max_radius;     // 10
    float dl_eta1;           // .25
    float dl_eta2;           // .75
    float dl_gamma_dec;      // .25
    float dl_gamma_inc;      // 2
    int32_t sum_mode;        // oracle only: 0 fp32 sequential, 1 fp64 accumulate
    int32_t knn_mode;        // oracle only: 0 exact KD, 1 reference-faithful KD, 2 brute force
};

struct orc_reg_result {
    float T[16];
    int32_t converged;
    int32_t iterations;
    float H[36];
    float b[6];
    float error;
    float H_raw[36];
    float b_raw[6];
    float error_raw;
    uint32_t inlier;
};

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// T/test_kdtree.cpp:69-75 — the fixture's generator: one std::mt19937 shared across clouds,
// uniform_real_distribution<float>(-range, range) re-created per cloud, w = 1.
void* orc_rng_create(uint32_t seed) { return new std::mt19937(seed); }
void orc_rng_destroy(void* g) { delete static_cast<std::mt19937*>(g); }
void orc_rng_uniform_points(void* g, size_t n, float range, float* out) {
    std::mt19937& gen = *static_cast<std::mt19937*>(g);
    std::uniform_real_distribution<float> dist(-range, range);
    for (size_t i = 0; i < n; ++i) {
        out[4 * i + 0] = dist(gen);
        out[4 * i + 1] = dist(gen);
        out[4 * i + 2] = dist(gen);
        out[4 * i + 3] = 1.0f;
    }
}
// uniform box [lo,hi] per axis (SURVEY §8(d) cfg 3 generator)
void orc_rng_box_points(void* g, size_t n, const float* lo, const float* hi, float* out) {
    std::mt19937& gen = *static_cast<std::mt19937*>(g);
    std::uniform_real_distribution<float> dx(lo[0], hi[0]), dy(lo[1], hi[1]), dz(lo[2], hi[2]);
    for (size_t i = 0; i < n; ++i) {
        out[4 * i + 0] = dx(gen);
        out[4 * i + 1] = dy(gen);
        out[4 * i + 2] = dz(gen);
        out[4 * i + 3] = 1.0f;
    }
}

// I/algorithms/filter/preprocess_operator/random_sampling_operator.hpp:36-50 — partial
// Fisher-Yates with a persistent mt19937, then order-preserving compaction
// (I/algorithms/common/filter_by_flags.hpp:43-49).  flags_out[i] = 1 keeps point i.
void orc_random_sampling_flags(void* g, size_t n, size_t num, uint8_t* flags_out) {
    std::mt19937& gen = *static_cast<std::mt19937*>(g);
    if (n <= num) {
        std::fill(flags_out, flags_out + n, (uint8_t)1);
        return;
    }
    std::fill(flags_out, flags_out + n, (uint8_t)0);
    std::vector<size_t> ind(n);
    std::iota(ind.begin(), ind.end(), (size_t)0);
    for (size_t i = 0; i < num; ++i) {
        std::uniform_int_distribution<size_t> d(i, n - 1);
        const size_t j = d(gen);
        std::swap(ind[i], ind[j]);
    }
    for (size_t i = 0; i < num; ++i) flags_out[ind[i]] = 1;
}

// preprocess_operator/mixed_random_sampling_operator.hpp:29-107 (weights already validated by the caller)
void orc_mixed_random_sampling_flags(void* g, const float* weights, size_t n, size_t sampling_num, float weighted_ratio,
                                     uint8_t* flags_out) {
    std::mt19937& gen = *static_cast<std::mt19937*>(g);
    if (n <= sampling_num) {
        std::fill(flags_out, flags_out + n, (uint8_t)1);
        return;
    }
    const size_t weighted_target = static_cast<size_t>(std::floor(static_cast<double>(sampling_num) * weighted_ratio));
    std::fill(flags_out, flags_out + n, (uint8_t)0);
    using KI = std::pair<float, size_t>;
    std::priority_queue<KI, std::vector<KI>, std::greater<KI>> selected;
    std::uniform_real_distribution<float> wd(std::numeric_limits<float>::min(), 1.0f);
    for (size_t i = 0; i < n; ++i) {
        const float w = weights[i];
        if (w <= 0.0f || weighted_target == 0) continue;
        const float key = std::log(wd(gen)) / w;
        if (selected.size() < weighted_target) {
            selected.emplace(key, i);
            continue;
        }
        if (!selected.empty() && selected.top().first < key) {
            selected.pop();
            selected.emplace(key, i);
        }
    }
    while (!selected.empty()) {
        flags_out[selected.top().second] = 1;
        selected.pop();
    }
    std::vector<size_t> rest;
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i) {
        if (flags_out[i]) ++cnt;
        else rest.push_back(i);
    }
    const size_t ut = std::min(sampling_num - cnt, rest.size());
    for (size_t i = 0; i < ut; ++i) {
        std::uniform_int_distribution<size_t> d(i, rest.size() - 1);
        const size_t j = d(gen);
        std::swap(rest[i], rest[j]);
        flags_out[rest[i]] = 1;
    }
}

// preprocess_operator/weighted_sampling_operator.hpp:29-96 (weights already validated by the caller)
void orc_weighted_random_sampling_flags(void* g, const float* weights, size_t n, size_t sampling_num, uint8_t* flags_out) {
    std::mt19937& gen = *static_cast<std::mt19937*>(g);
    if (n <= sampling_num) {
        std::fill(flags_out, flags_out + n, (uint8_t)1);
        return;
    }
    std::fill(flags_out, flags_out + n, (uint8_t)0);
    using KI = std::pair<float, size_t>;
    std::priority_queue<KI, std::vector<KI>, std::greater<KI>> selected;
    std::uniform_real_distribution<float> wd(std::numeric_limits<float>::min(), 1.0f);
    for (size_t i = 0; i < n; ++i) {
        if (weights[i] <= 0.0f) continue;
        const float key = std::log(wd(gen)) / weights[i];
        if (selected.size() < sampling_num) {
            selected.emplace(key, i);
            continue;
        }
        if (selected.top().first < key) {
            selected.pop();
            selected.emplace(key, i);
        }
    }
    while (!selected.empty()) {
        flags_out[selected.top().second] = 1;
        selected.pop();
    }
}

// preprocess_operator/farthest_point_sampling_operator.hpp:27-94
void orc_farthest_point_sampling_flags(void* g, const float* pts, size_t n, size_t sampling_num, uint8_t* flags_out) {
    std::mt19937& gen = *static_cast<std::mt19937*>(g);
    if (n <= sampling_num) {
        std::fill(flags_out, flags_out + n, (uint8_t)1);
        return;
    }
    std::fill(flags_out, flags_out + n, (uint8_t)0);
    std::vector<float> dist(n, FMAX);
    std::uniform_int_distribution<size_t> d0(0, n - 1);
    size_t sel = d0(gen);
    flags_out[sel] = 1;
    for (size_t it = 1; it < sampling_num; ++it) {
        const float* c = pts + 4 * sel;
        for (size_t i = 0; i < n; ++i) {
            const float* p = pts + 4 * i;
            const float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2], dw = p[3] - c[3];
            const float d = std::fma(dw, dw, std::fma(dz, dz, std::fma(dy, dy, dx * dx)));
            dist[i] = std::fmin(dist[i], d);
        }
        sel = (size_t)(std::max_element(dist.begin(), dist.end()) - dist.begin());
        flags_out[sel] = 1;
    }
}

// preprocess_operator/angle_incidence_filter_operator.hpp:57-103 (normals, or extract_normal of covs when normals == NULL)
void orc_angle_incidence_flags(const float* pts, const float* normals, const float* covs, size_t n, float min_angle,
                               float max_angle, uint8_t* flags_out) {
    const float max_cos = std::cos(min_angle), min_cos = std::cos(max_angle);
    for (size_t i = 0; i < n; ++i) {
        const float* p = pts + 4 * i;
        flags_out[i] = 0;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && std::isfinite(p[3]))) continue;
        V3 nr, pv;
        if (normals) {
            for (int a = 0; a < 3; ++a) nr(a) = normals[4 * i + a];
        } else {
            const V4 e = extract_normal(p, load_cov(covs + 16 * i));
            for (int a = 0; a < 3; ++a) nr(a) = e(a);
        }
        for (int a = 0; a < 3; ++a) pv(a) = p[a];
        const float d = dot<3>(pv, nr);
        const float denom = std::sqrt(dot<3>(pv, pv)) * std::sqrt(dot<3>(nr, nr));
        if (denom <= 1e-6f) continue;
        const float ac = std::fabs(d / denom);
        flags_out[i] = !(ac < min_cos || ac > max_cos);
    }
}

void orc_transform_points(const float* T16, const float* pts, size_t n, float* out) {
    const M4 T = load_T(T16);
    for (size_t i = 0; i < n; ++i) {
        const V4 r = transform_point(T, load_p(pts + 4 * i));
        for (int a = 0; a < 4; ++a) out[4 * i + a] = r(a);
    }
}

// transform::transform_async — I/algorithms/common/transform.hpp:45-94: covariances T C T^T (:14-22),
// normals T n (:24-30; the kernel's normalize<4>() result is DISCARDED there, so the reference's device
// path leaves normals un-normalised — restated as is), points T p (:32-37).  Any of covs / normals
// may be NULL.  Column-major 4x4 covariances.
void orc_transform_cloud(const float* T16, const float* pts, const float* covs, const float* normals, size_t n,
                         float* out_pts, float* out_covs, float* out_normals) {
    const M4 T = load_T(T16);
    for (size_t i = 0; i < n; ++i) {
        if (covs) {
            const M4 r = transform_cov(load_cov(covs + 16 * i), T);
            for (int j = 0; j < 4; ++j)
                for (int a = 0; a < 4; ++a) out_covs[16 * i + j * 4 + a] = r(a, j);
        }
        if (normals) {
            const V4 r = mul<4, 4>(T, load_p(normals + 4 * i));
            for (int a = 0; a < 4; ++a) out_normals[4 * i + a] = r(a);
        }
        const V4 r = transform_point(T, load_p(pts + 4 * i));
        for (int a = 0; a < 4; ++a) out_pts[4 * i + a] = r(a);
    }
}

// deskew::deskew_point_cloud_constant_velocity — I/algorithms/deskew/relative_pose_deskew.hpp:121-174: per point
// tau = clamp(t_ms * 1e-3 / duration, 0, 1), motion = se3_exp(tau * twist) (:140-146), point = motion * p (:149),
// normal = R n (:155-160), covariance = R (C R^T) (:161-168) with R = quat_to_rot(so3_exp(tau * omega)); a
// non-finite timestamp copies the element (:127-137).  normals / covs (column-major 4x4) may be NULL.
void orc_deskew_constant_velocity(const float* pts, const float* normals, const float* covs, const float* ts_ms, size_t n,
                                  const float* twist6, float duration, float* out_pts, float* out_normals,
                                  float* out_covs) {
    for (size_t i = 0; i < n; ++i) {
        const float t_s = ts_ms[i] * 1e-3f;
        if (!std::isfinite(t_s)) {
            for (int a = 0; a < 4; ++a) out_pts[4 * i + a] = pts[4 * i + a];
            if (normals) for (int a = 0; a < 4; ++a) out_normals[4 * i + a] = normals[4 * i + a];
            if (covs) for (int a = 0; a < 16; ++a) out_covs[16 * i + a] = covs[16 * i + a];
            continue;
        }
        const float tau = std::fmin(std::fmax(t_s / duration, 0.0f), 1.0f);
        V6 a;
        for (int k = 0; k < 6; ++k) a(k) = twist6[k] * tau;
        const M4 motion = se3_exp(a);
        const V4 r = mul<4, 4>(motion, load_p(pts + 4 * i));
        for (int k = 0; k < 4; ++k) out_pts[4 * i + k] = r(k);
        V3 om; om(0) = a(0); om(1) = a(1); om(2) = a(2);
        const M3 R = quat_to_rot(so3_exp(om));
        if (normals) {
            V3 nv; nv(0) = normals[4 * i]; nv(1) = normals[4 * i + 1]; nv(2) = normals[4 * i + 2];
            const V3 o = mul<3, 3>(R, nv);
            out_normals[4 * i] = o(0); out_normals[4 * i + 1] = o(1); out_normals[4 * i + 2] = o(2);
            out_normals[4 * i + 3] = 0.0f;
        }
        if (covs) {
            M3 Cm, Rt;
            for (int c = 0; c < 3; ++c)
                for (int rr = 0; rr < 3; ++rr) {
                    Cm(rr, c) = covs[16 * i + c * 4 + rr];
                    Rt(rr, c) = R(c, rr);
                }
            const M3 o = mul<3, 3, 3>(R, mul<3, 3, 3>(Cm, Rt));
            for (int a2 = 0; a2 < 16; ++a2) out_covs[16 * i + a2] = 0.0f;
            for (int c = 0; c < 3; ++c)
                for (int rr = 0; rr < 3; ++rr) out_covs[16 * i + c * 4 + rr] = o(rr, c);
        }
    }
}

// I/algorithms/knn/bruteforce.hpp:24-96 with the oracle's fixed distance formula (the
// reference's sycl::dot leaves contraction to the SYCL implementation):
// dist = fma(dz,dz,fma(dy,dy,dx*dx)) on the transformed query, order (dist, index).
// T16 may be NULL (the reference's brute force takes no transform).
void orc_knn_bruteforce(const float* queries, size_t nq, const float* targets, size_t nt, int k, const float* T16,
                        int32_t* idx, float* dist) {
    const M4 T = T16 ? load_T(T16) : M4::identity();
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t qi = 0; qi < (int64_t)nq; ++qi) {
        const V4 q0 = load_p(queries + 4 * (size_t)qi);
        const V4 q = T16 ? transform_point(T, q0) : q0;
        BestK best{dist + (size_t)qi * k, idx + (size_t)qi * k, k};
        best.init();
        for (size_t j = 0; j < nt; ++j) {
            const float ds = dist_sq(q, targets + 4 * j);
            if (ds < best.d[k - 1]) best.insert_first_visited(ds, (int32_t)j);  // index order => (dist,index)
        }
    }
}

void* orc_kdtree_build(const float* pts, size_t n, size_t leaf_threshold) { return kd_build(pts, n, leaf_threshold); }
void orc_kdtree_destroy(void* t) { delete static_cast<KDTree*>(t); }
size_t orc_kdtree_size(void* t) { return static_cast<KDTree*>(t)->nodes.size(); }

// mode 0: exact (oracle contract), mode 1: reference-faithful traversal (kdtree.hpp:463-553)
void orc_kdtree_knn(void* tree, const float* queries, size_t nq, int k, const float* T16, int mode, int32_t* idx,
                    float* dist) {
    const KDTree& t = *static_cast<KDTree*>(tree);
    const M4 T = T16 ? load_T(T16) : M4::identity();
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t qi = 0; qi < (int64_t)nq; ++qi) {
        const V4 q = transform_point(T, load_p(queries + 4 * (size_t)qi));
        if (mode == 0)
            kd_search_one<true>(t, q, k, dist + (size_t)qi * k, idx + (size_t)qi * k);
        else
            kd_search_one<false>(t, q, k, dist + (size_t)qi * k, idx + (size_t)qi * k);
    }
}

void orc_eigen3(const float* A9_rowmajor, float* evals3, float* evecs9_rowmajor) {
    M3 A, V;
    V3 e;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A(i, j) = A9_rowmajor[i * 3 + j];
    eigen3(A, e, V);
    for (int i = 0; i < 3; ++i) {
        evals3[i] = e(i);
        for (int j = 0; j < 3; ++j) evecs9_rowmajor[i * 3 + j] = V(i, j);
    }
}

void orc_inverse3(const float* A9_rowmajor, float* out9_rowmajor) {
    M3 A;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A(i, j) = A9_rowmajor[i * 3 + j];
    const M3 r = inverse(A);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out9_rowmajor[i * 3 + j] = r(i, j);
}

void orc_covariance(const float* pts, size_t n, const int32_t* idx, int k, float* covs) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const M4 c = estimate_cov(pts, k, idx + (size_t)i * k);
        store_T(c, covs + 16 * (size_t)i);
    }
}

// covariance.hpp:323-373
void orc_covariance_robust(const float* pts, size_t n, const int32_t* idx, int k, int loss, float mad_scale, float min_scale,
                           int max_iter, float* covs) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const M4 c = loss == L_NONE ? estimate_cov(pts, k, idx + (size_t)i * k)
                                    : estimate_cov_robust(pts, k, idx + (size_t)i * k, loss, mad_scale, min_scale, max_iter);
        store_T(c, covs + 16 * (size_t)i);
    }
}

// covariance.hpp:417-442
void orc_normals(const float* pts, size_t n, const int32_t* idx, int k, float* normals) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const M4 c = estimate_cov(pts, k, idx + (size_t)i * k);
        const V4 nr = extract_normal(pts + 4 * (size_t)i, c);
        for (int a = 0; a < 4; ++a) normals[4 * (size_t)i + a] = nr(a);
    }
}

// covariance.hpp:467-495
void orc_normals_from_covs(const float* pts, const float* covs, size_t n, float* normals) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const V4 nr = extract_normal(pts + 4 * (size_t)i, load_cov(covs + 16 * (size_t)i));
        for (int a = 0; a < 4; ++a) normals[4 * (size_t)i + a] = nr(a);
    }
}

void orc_update_covariance_plane(const float* covs_in, size_t n, float* covs_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        M4 c = load_cov(covs_in + 16 * (size_t)i);
        update_covariance_plane(c);
        store_T(c, covs_out + 16 * (size_t)i);
    }
}

void orc_set_genz_planarity_threshold(float t) { g_genz_planarity_threshold = t; }
// RegistrationParams::rotation_constraint {enable, weight, robust.default_scale} (registration_params.hpp:54-62)
void orc_set_rotation_constraint(int enable, float weight, float robust_scale) {
    g_rot_enable = enable;
    g_rot_weight = weight;
    g_rot_scale = robust_scale;
}
float orc_genz_alpha(const float* tgt_covs, size_t ns, const int32_t* idx, const float* dist, float max_corr_sq) {
    return genz_alpha_of(tgt_covs, ns, idx, dist, max_corr_sq);
}
float orc_robust_weight(int loss, float r, float s) { return robust_weight(loss, r, s); }
float orc_robust_error(int loss, float r, float s) { return robust_error(loss, r, s); }

void orc_linearize(int reg, int loss, const float* src_pts, const float* src_covs, size_t ns, const float* tgt_pts,
                   const float* tgt_covs, const float* tgt_normals, const int32_t* idx, const float* dist,
                   const float* T16, float max_corr_sq, float scale, int mode, float* H36, float* b6, float* err,
                   uint32_t* inlier) {
    const Clouds c{src_pts, src_covs, ns, tgt_pts, tgt_covs, tgt_normals};
    const Linearized l = linearize(reg, loss, c, idx, dist, load_T(T16), max_corr_sq, scale, mode);
    std::memcpy(H36, l.H, sizeof(l.H));
    std::memcpy(b6, l.b, sizeof(l.b));
    *err = l.error;
    *inlier = l.inlier;
}

void orc_error(int reg, int loss, const float* src_pts, const float* src_covs, size_t ns, const float* tgt_pts,
               const float* tgt_covs, const float* tgt_normals, const int32_t* idx, const float* dist,
               const float* T16, float max_corr_sq, float scale, int mode, float* err, uint32_t* inlier) {
    const Clouds c{src_pts, src_covs, ns, tgt_pts, tgt_covs, tgt_normals};
    error_sum(reg, loss, c, idx, dist, load_T(T16), max_corr_sq, scale, mode, err, inlier);
}

// registration.hpp:412-462 — per-point robust weights in source order (0 when gated out)
void orc_robust_weights(int reg, int loss, const float* src_pts, const float* src_covs, size_t ns,
                        const float* tgt_pts, const float* tgt_covs, const float* tgt_normals, const int32_t* idx,
                        const float* dist, const float* T16, float max_corr_sq, float scale, float* weights) {
    const M4 T = load_T(T16);
    const M4 ident = M4::identity();
    const V4 zero4 = V4::zero();
    for (size_t i = 0; i < ns; ++i) {
        float w = 0.0f;
        if (dist[i] <= max_corr_sq) {
            const int32_t ti = idx[i];
            const float e2 = err_only(reg, T, load_p(src_pts + 4 * i), src_covs ? load_cov(src_covs + 16 * i) : ident,
                                      load_p(tgt_pts + 4 * (size_t)ti),
                                      tgt_covs ? load_cov(tgt_covs + 16 * (size_t)ti) : ident,
                                      tgt_normals ? load_p(tgt_normals + 4 * (size_t)ti) : zero4);
            w = robust_weight(loss, std::sqrt(e2), scale);
        }
        weights[i] = w;
    }
}

void orc_se3_exp(const float* twist6, float* T16) {
    V6 a;
    for (int i = 0; i < 6; ++i) a(i) = twist6[i];
    store_T(se3_exp(a), T16);
}
void orc_se3_log(const float* T16, float* twist6) {
    const V6 r = se3_log(load_T(T16));
    for (int i = 0; i < 6; ++i) twist6[i] = r(i);
}
void orc_so3_exp(const float* om3, float* quat4) {
    V3 o;
    for (int i = 0; i < 3; ++i) o(i) = om3[i];
    const V4 q = so3_exp(o);
    for (int i = 0; i < 4; ++i) quat4[i] = q(i);
}
void orc_so3_log(const float* quat4, float* om3) {
    const V3 o = so3_log(load_p(quat4));
    for (int i = 0; i < 3; ++i) om3[i] = o(i);
}

int orc_solve6(const float* H36, const float* b6, float lambda, float* delta6) {
    return solve_system(H36, b6, lambda, delta6) ? 1 : 0;
}

void orc_dogleg_step(const float* H36, const float* g6, float radius, float* p6, float* step_norm,
                     float* predicted_reduction) {
    const Dogleg d = dogleg_step(H36, g6, radius);
    std::memcpy(p6, d.p, sizeof(d.p));
    *step_norm = d.step_norm;
    *predicted_reduction = d.predicted_reduction;
}

// I/algorithms/common/voxel_constants.hpp:36-62
uint64_t orc_voxel_key(const float* p, float inv) {
    constexpr uint64_t invalid = std::numeric_limits<uint64_t>::max();
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) return invalid;
    const int64_t off = 1 << 20, mask = (1 << 21) - 1;
    const int64_t c0 = (int64_t)std::floor(p[0] * inv) + off;
    const int64_t c1 = (int64_t)std::floor(p[1] * inv) + off;
    const int64_t c2 = (int64_t)std::floor(p[2] * inv) + off;
    if (c0 < 0 || mask < c0 || c1 < 0 || mask < c1 || c2 < 0 || mask < c2) return invalid;
    return ((uint64_t)(c0 & mask)) | ((uint64_t)(c1 & mask) << 21) | ((uint64_t)(c2 & mask) << 42);
}

// ------------------------------------------------------------------ mapping::VoxelHashMap
// I/algorithms/mapping/voxel_hash_map.hpp:22-1066, restated SEQUENTIALLY: points are inserted in index order (the
// reference's atomics leave the accumulation order and the winner of a slot race unspecified), every slot
// accumulates in fp32 in that order.  Same table: double hashing :607-612, 100 probes :498, capacities :481-482.
struct OrcVoxelMap {
    static constexpr uint64_t INVALID = std::numeric_limits<uint64_t>::max();
    float voxel = 0, inv = 0;
    size_t cap = 30029;
    std::vector<uint64_t> key;
    std::vector<float> core;  // sx sy sz
    std::vector<uint32_t> count;
    std::vector<float> cov;   // xx xy xz yy yz zz
    std::vector<float> color, intensity;
    std::vector<uint32_t> last;
    uint32_t staleness = 0, max_staleness = 100, cycle = 10, min_num_point = 1;
    float rehash_threshold = 0.7f;
    size_t voxel_num = 0;
    bool has_cov = false, has_rgb = false, has_int = false;

    void alloc(size_t c) {
        cap = c;
        key.assign(c, INVALID);
        core.assign(3 * c, 0.f);
        count.assign(c, 0u);
        cov.assign(6 * c, 0.f);
        color.assign(4 * c, 0.f);
        intensity.assign(c, 0.f);
        last.assign(c, 0u);
    }
    size_t slot_of(uint64_t k, size_t probe) const {  // :607-612
        const uint64_t h2 = (cap - 2) - (k % (cap - 2));
        return (size_t)((k + probe * h2) % cap);
    }
    // global_reduction :574-605
    void insert(uint64_t k, const float* c3, uint32_t n, const float* cv6, const float* rgba, float inten, uint32_t stamp,
                bool hc, bool hr, bool hi) {
        if (k == INVALID) return;
        for (size_t j = 0; j < 100; ++j) {
            const size_t s = slot_of(k, j);
            if (key[s] == INVALID) {
                key[s] = k;
                ++voxel_num;
            }
            if (key[s] == k) {
                for (int a = 0; a < 3; ++a) core[3 * s + a] += c3[a];
                count[s] += n;
                if (hc) for (int a = 0; a < 6; ++a) cov[6 * s + a] += cv6[a];
                if (hr) for (int a = 0; a < 4; ++a) color[4 * s + a] += rgba[a];
                if (hi) intensity[s] += inten;
                last[s] = stamp;
                return;
            }
        }
    }
    void remove_old() {  // :794-845
        if (staleness <= max_staleness) return;
        const uint32_t rs = staleness - max_staleness;
        size_t kept = 0;
        for (size_t i = 0; i < cap; ++i) {
            if (key[i] == INVALID) continue;
            if (last[i] >= rs) {
                ++kept;
                continue;
            }
            key[i] = INVALID;
            for (int a = 0; a < 3; ++a) core[3 * i + a] = 0;
            count[i] = 0;
            for (int a = 0; a < 6; ++a) cov[6 * i + a] = 0;
            for (int a = 0; a < 4; ++a) color[4 * i + a] = 0;
            intensity[i] = 0;
            last[i] = 0;
        }
        set_voxel_num(kept);
    }
    void set_voxel_num(size_t v) {  // :510-517
        voxel_num = v;
        if (v == 0) has_cov = has_rgb = has_int = false;
    }
    void rehash(size_t nc) {  // :847-934
        if (cap >= nc) return;
        OrcVoxelMap old = *this;
        alloc(nc);
        voxel_num = 0;
        for (size_t i = 0; i < old.cap; ++i) {
            if (old.key[i] == INVALID) continue;
            insert(old.key[i], &old.core[3 * i], old.count[i], &old.cov[6 * i], &old.color[4 * i], old.intensity[i],
                   old.last[i], has_cov, has_rgb, has_int);
        }
        set_voxel_num(voxel_num);
    }
};

// I/algorithms/filter/voxel_downsampling.hpp:50-62,146-218.  Oracle contract for the order
// the reference leaves to std::sort: stable (key, original index); fp32 running Vector4f
// sum in that order, then sum / sum.w.  out must hold n points; returns the voxel count.
size_t orc_voxel_downsample(const float* pts, size_t n, float voxel_size, size_t min_voxel_count, float* out) {
    const float inv = 1.0f / voxel_size;
    std::vector<uint64_t> keys(n);
    for (size_t i = 0; i < n; ++i) keys[i] = orc_voxel_key(pts + 4 * i, inv);
    std::vector<size_t> order;
    order.reserve(n);
    for (size_t i = 0; i < n; ++i)
        if (keys[i] != std::numeric_limits<uint64_t>::max()) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    const float minc = (float)min_voxel_count;
    size_t m = 0, g = 0;
    while (g < order.size()) {
        const uint64_t key = keys[order[g]];
        float s[4] = {0, 0, 0, 0};
        size_t e = g;
        while (e < order.size() && keys[order[e]] == key) {
            for (int a = 0; a < 4; ++a) s[a] += pts[4 * order[e] + a];
            ++e;
        }
        if (s[3] >= minc) {
            for (int a = 0; a < 4; ++a) out[4 * m + a] = s[a] / s[3];
            ++m;
        }
        g = e;
    }
    return m;
}

// The reference's behaviour as shipped: unstable std::sort over the index vector
// (voxel_downsampling.hpp:169-171).  Timed as part of the CPU baseline; same centroids up to
// fp32 summation order.
size_t orc_voxel_downsample_unstable(const float* pts, size_t n, float voxel_size, size_t min_voxel_count,
                                     float* out) {
    const float inv = 1.0f / voxel_size;
    std::vector<uint64_t> keys(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) keys[i] = orc_voxel_key(pts + 4 * (size_t)i, inv);
    std::vector<size_t> order;
    order.reserve(n);
    for (size_t i = 0; i < n; ++i)
        if (keys[i] != std::numeric_limits<uint64_t>::max()) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    const float minc = (float)min_voxel_count;
    size_t m = 0, g = 0;
    while (g < order.size()) {
        const uint64_t key = keys[order[g]];
        float s[4] = {0, 0, 0, 0};
        size_t e = g;
        while (e < order.size() && keys[order[e]] == key) {
            for (int a = 0; a < 4; ++a) s[a] += pts[4 * order[e] + a];
            ++e;
        }
        if (s[3] >= minc) {
            for (int a = 0; a < 4; ++a) out[4 * m + a] = s[a] / s[3];
            ++m;
        }
        g = e;
    }
    return m;
}

// voxel_downsampling.hpp:64-79,220-288 — attribute handling of the cloud overload on the
// same stable order: mean RGB, median intensity (:82-98), mean timestamp.  Any attribute
// pointer may be NULL.  Returns the voxel count.
static size_t aggregate_by_key(const std::vector<uint64_t>& keys, const float* pts, size_t n, size_t min_voxel_count,
                               const float* rgb, const float* intensity, const float* timestamps, float* out_pts,
                               float* out_rgb, float* out_intensity, float* out_timestamps) {
    std::vector<size_t> order;
    for (size_t i = 0; i < n; ++i)
        if (keys[i] != std::numeric_limits<uint64_t>::max()) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    const float minc = (float)min_voxel_count;
    size_t m = 0, g = 0;
    std::vector<float> vals;
    while (g < order.size()) {
        const uint64_t key = keys[order[g]];
        float s[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0}, ts = 0.0f;
        vals.clear();
        size_t e = g;
        while (e < order.size() && keys[order[e]] == key) {
            const size_t id = order[e];
            for (int a = 0; a < 4; ++a) s[a] += pts[4 * id + a];
            if (rgb)
                for (int a = 0; a < 4; ++a) c[a] += rgb[4 * id + a];
            if (intensity) vals.push_back(intensity[id]);
            if (timestamps) ts += timestamps[id];
            ++e;
        }
        if (s[3] >= minc) {
            for (int a = 0; a < 4; ++a) out_pts[4 * m + a] = s[a] / s[3];
            if (rgb)
                for (int a = 0; a < 4; ++a) out_rgb[4 * m + a] = c[a] / s[3];
            if (intensity) {
                float med = 0.0f;
                if (!vals.empty()) {
                    const size_t mid = vals.size() / 2;
                    std::nth_element(vals.begin(), vals.begin() + mid, vals.end());
                    const float up = vals[mid];
                    if (vals.size() % 2 != 0) {
                        med = up;
                    } else {
                        std::nth_element(vals.begin(), vals.begin() + (mid - 1), vals.begin() + mid);
                        med = 0.5f * (vals[mid - 1] + up);
                    }
                }
                out_intensity[m] = med;
            }
            if (timestamps) out_timestamps[m] = ts / s[3];
            ++m;
        }
        g = e;
    }
    return m;
}

size_t orc_voxel_downsample_attrs(const float* pts, size_t n, float voxel_size, size_t min_voxel_count,
                                  const float* rgb, const float* intensity, const float* timestamps, float* out_pts,
                                  float* out_rgb, float* out_intensity, float* out_timestamps) {
    const float inv = 1.0f / voxel_size;
    std::vector<uint64_t> keys(n);
    for (size_t i = 0; i < n; ++i) keys[i] = orc_voxel_key(pts + 4 * i, inv);
    return aggregate_by_key(keys, pts, n, min_voxel_count, rgb, intensity, timestamps, out_pts, out_rgb, out_intensity,
                            out_timestamps);
}

// I/algorithms/filter/polar_downsampling.hpp:30-108 (coord_system 0 LIDAR, 1 CAMERA).  atan2 correctly rounded
// (fp64 then cast), squared sums as plain fp32 multiplies and adds.
void* orc_voxelmap_create(float voxel_size) {
    auto* m = new OrcVoxelMap;
    m->voxel = voxel_size;
    m->inv = 1.0f / voxel_size;
    m->alloc(30029);
    return m;
}
void orc_voxelmap_destroy(void* h) { delete static_cast<OrcVoxelMap*>(h); }
void orc_voxelmap_set_params(void* h, float voxel_size, uint32_t max_staleness, uint32_t cycle, float rehash_threshold,
                             uint32_t min_num_point) {
    auto* m = static_cast<OrcVoxelMap*>(h);
    m->voxel = voxel_size;
    m->inv = 1.0f / voxel_size;
    m->max_staleness = max_staleness;
    m->cycle = cycle;
    m->rehash_threshold = rehash_threshold;
    m->min_num_point = min_num_point;
}
// add_point_cloud :117-140, add_point_cloud_impl :614-792 (load_entry :661-704)
void orc_voxelmap_add(void* h, const float* pts, const float* covs, const float* rgb, const float* intens, size_t n,
                      const float* T16) {
    static const size_t caps[11] = {30029, 60013, 120011, 240007, 480013, 960017, 1920001, 3840007, 7680017, 15360013, 30720007};
    auto* m = static_cast<OrcVoxelMap*>(h);
    if (m->rehash_threshold < (float)m->voxel_num / (float)m->cap) {
        size_t next = m->cap;
        for (size_t c : caps)
            if (c > m->cap) {
                next = c;
                break;
            }
        if (next > m->cap) m->rehash(next);
    }
    if (n > 0) {
        m->has_cov |= covs != nullptr;
        m->has_rgb |= rgb != nullptr;
        m->has_int |= intens != nullptr;
        const M4 T = load_T(T16);
        for (size_t i = 0; i < n; ++i) {
            const V4 w = transform_point(T, load_p(pts + 4 * i));
            const float wp[4] = {w(0), w(1), w(2), w(3)};
            const uint64_t k = orc_voxel_key(wp, m->inv);
            float cv[6] = {0, 0, 0, 0, 0, 0};
            if (covs) {  // rotate_covariance_upper_triangle :420-456, encode_covariance_for_aggregation :458-476
                const float* c = covs + 16 * i;
                const float cxx = c[0], cxy = c[4], cxz = c[8], cyy = c[5], cyz = c[9], czz = c[10];
                const float r00 = T(0, 0), r01 = T(0, 1), r02 = T(0, 2), r10 = T(1, 0), r11 = T(1, 1), r12 = T(1, 2),
                            r20 = T(2, 0), r21 = T(2, 1), r22 = T(2, 2);
                const float a00 = std::fma(r02, cxz, std::fma(r01, cxy, r00 * cxx));
                const float a01 = std::fma(r02, cyz, std::fma(r01, cyy, r00 * cxy));
                const float a02 = std::fma(r02, czz, std::fma(r01, cyz, r00 * cxz));
                const float a10 = std::fma(r12, cxz, std::fma(r11, cxy, r10 * cxx));
                const float a11 = std::fma(r12, cyz, std::fma(r11, cyy, r10 * cxy));
                const float a12 = std::fma(r12, czz, std::fma(r11, cyz, r10 * cxz));
                const float a20 = std::fma(r22, cxz, std::fma(r21, cxy, r20 * cxx));
                const float a21 = std::fma(r22, cyz, std::fma(r21, cyy, r20 * cxy));
                const float a22 = std::fma(r22, czz, std::fma(r21, cyz, r20 * cxz));
                M3 S;
                S(0, 0) = std::fma(a02, r02, std::fma(a01, r01, a00 * r00));
                S(0, 1) = S(1, 0) = std::fma(a02, r12, std::fma(a01, r11, a00 * r10));
                S(0, 2) = S(2, 0) = std::fma(a02, r22, std::fma(a01, r21, a00 * r20));
                S(1, 1) = std::fma(a12, r12, std::fma(a11, r11, a10 * r10));
                S(1, 2) = S(2, 1) = std::fma(a12, r22, std::fma(a11, r21, a10 * r20));
                S(2, 2) = std::fma(a22, r22, std::fma(a21, r21, a20 * r20));
                const M3 L = spd_function(S, true);
                cv[0] = L(0, 0); cv[1] = L(0, 1); cv[2] = L(0, 2); cv[3] = L(1, 1); cv[4] = L(1, 2); cv[5] = L(2, 2);
            }
            const float zero4[4] = {0, 0, 0, 0};
            m->insert(k, wp, 1u, cv, rgb ? rgb + 4 * i : zero4, intens ? intens[i] : 0.f, m->staleness, covs != nullptr,
                      rgb != nullptr, intens != nullptr);
        }
    }
    if (m->cycle > 0 && (m->staleness % m->cycle) == 0) m->remove_old();
    ++m->staleness;
}
void orc_voxelmap_remove_old(void* h) { static_cast<OrcVoxelMap*>(h)->remove_old(); }
void orc_voxelmap_info(void* h, uint64_t* cap, uint64_t* voxel_num, uint32_t* staleness, int* flags3) {
    auto* m = static_cast<OrcVoxelMap*>(h);
    *cap = m->cap;
    *voxel_num = m->voxel_num;
    *staleness = m->staleness;
    flags3[0] = m->has_cov; flags3[1] = m->has_rgb; flags3[2] = m->has_int;
}
// downsampling :146-188, downsampling_impl :936-1065 (slot order), compute_averaged_attributes :348-393
size_t orc_voxelmap_downsample(void* h, const float* center3, float distance, float* out_pts, float* out_covs,
                               float* out_rgb, float* out_int, uint64_t* out_keys) {
    auto* m = static_cast<OrcVoxelMap*>(h);
    if (m->voxel_num == 0) return 0;
    const float lo[3] = {center3[0] - distance, center3[1] - distance, center3[2] - distance};
    const float hi[3] = {center3[0] + distance, center3[1] + distance, center3[2] + distance};
    size_t o = 0;
    for (size_t i = 0; i < m->cap; ++i) {
        if (m->key[i] == OrcVoxelMap::INVALID || m->count[i] < m->min_num_point || m->count[i] == 0) continue;
        const float ic = 1.0f / (float)m->count[i];
        const float c[3] = {m->core[3 * i] * ic, m->core[3 * i + 1] * ic, m->core[3 * i + 2] * ic};
        if (!(c[0] >= lo[0] && c[0] <= hi[0] && c[1] >= lo[1] && c[1] <= hi[1] && c[2] >= lo[2] && c[2] <= hi[2])) continue;
        out_pts[4 * o] = c[0]; out_pts[4 * o + 1] = c[1]; out_pts[4 * o + 2] = c[2]; out_pts[4 * o + 3] = 1.0f;
        if (m->has_cov && out_covs) {
            const float* v = &m->cov[6 * i];
            M3 S;
            S(0, 0) = v[0] * ic; S(0, 1) = S(1, 0) = v[1] * ic; S(0, 2) = S(2, 0) = v[2] * ic;
            S(1, 1) = v[3] * ic; S(1, 2) = S(2, 1) = v[4] * ic; S(2, 2) = v[5] * ic;
            const M3 E = spd_function(S, false);
            float* oc = out_covs + 16 * o;
            for (int a = 0; a < 16; ++a) oc[a] = 0.f;
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc) oc[cc * 4 + r] = E(r, cc);
        }
        if (m->has_rgb && out_rgb)
            for (int a = 0; a < 4; ++a) out_rgb[4 * o + a] = m->color[4 * i + a] * ic;
        if (m->has_int && out_int) out_int[o] = m->intensity[i] * ic;
        if (out_keys) out_keys[o] = m->key[i];
        ++o;
    }
    return o;
}
// compute_overlap_ratio :194-246
float orc_voxelmap_overlap_ratio(void* h, const float* pts, size_t n, const float* T16) {
    auto* m = static_cast<OrcVoxelMap*>(h);
    if (n == 0 || m->voxel_num == 0) return 0.0f;
    const M4 T = load_T(T16);
    uint32_t hits = 0;
    for (size_t i = 0; i < n; ++i) {
        const V4 w = transform_point(T, load_p(pts + 4 * i));
        const float wp[4] = {w(0), w(1), w(2), w(3)};
        const uint64_t k = orc_voxel_key(wp, m->inv);
        if (k == OrcVoxelMap::INVALID) continue;
        for (size_t j = 0; j < 100; ++j) {
            const size_t s = m->slot_of(k, j);
            if (m->key[s] == k) {
                if (m->count[s] >= m->min_num_point) ++hits;
                break;
            }
            if (m->key[s] == OrcVoxelMap::INVALID) break;
        }
    }
    return (float)hits / (float)n;
}
// eigen_utils.hpp:646-677 on row-major 3x3
void orc_spd_function(const float* A9_rowmajor, int is_log, float* out9_rowmajor) {
    M3 A;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A(i, j) = A9_rowmajor[3 * i + j];
    const M3 R = spd_function(A, is_log != 0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out9_rowmajor[3 * i + j] = R(i, j);
}

uint64_t orc_polar_key(const float* p, float dist_inv, float elev_inv, float azim_inv, int coord_system) {
    constexpr uint64_t invalid = std::numeric_limits<uint64_t>::max();
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) return invalid;
    const float xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
    const float r = std::sqrt(xx + yy + zz);
    if (r == 0.0f) return invalid;
    float azimuth, elevation;
    if (coord_system == 0) {
        const float x2y2 = xx + yy;
        if (x2y2 == 0.0f) return invalid;
        azimuth = cr_atan2(p[1], p[0]);
        elevation = cr_atan2(p[2], std::sqrt(x2y2));
    } else if (coord_system == 1) {
        const float x2z2 = xx + zz;
        if (x2z2 == 0.0f) return invalid;
        azimuth = cr_atan2(p[0], p[2]);
        elevation = cr_atan2(-p[1], std::sqrt(x2z2));
    } else {
        return invalid;
    }
    const int64_t off = 1 << 20, mask = (1 << 21) - 1;
    const float f0 = std::floor(r * dist_inv), f1 = std::floor(elevation * elev_inv), f2 = std::floor(azimuth * azim_inv);
    if (!(std::fabs(f0) < 4e9f && std::fabs(f1) < 4e9f && std::fabs(f2) < 4e9f)) return invalid;
    const int64_t c0 = (int64_t)f0 + off, c1 = (int64_t)f1 + off, c2 = (int64_t)f2 + off;
    if (c0 < 0 || mask < c0 || c1 < 0 || mask < c1 || c2 < 0 || mask < c2) return invalid;
    return ((uint64_t)(c0 & mask)) | ((uint64_t)(c1 & mask) << 21) | ((uint64_t)(c2 & mask) << 42);
}

// filter::PolarGrid::downsampling — polar_downsampling.hpp:186-198 (cloud overload), :317-338 (sort), :375-452
// (aggregation, the same as the voxel grid's).  Stable (key, index) order like the voxel oracle.
size_t orc_polar_downsample_attrs(const float* pts, size_t n, float dist_size, float elev_size, float azim_size,
                                  int coord_system, size_t min_voxel_count, const float* rgb, const float* intensity,
                                  const float* timestamps, float* out_pts, float* out_rgb, float* out_intensity,
                                  float* out_timestamps) {
    const float di = 1.0f / dist_size, ei = 1.0f / elev_size, ai = 1.0f / azim_size;
    std::vector<uint64_t> keys(n);
    for (size_t i = 0; i < n; ++i) keys[i] = orc_polar_key(pts + 4 * i, di, ei, ai, coord_system);
    return aggregate_by_key(keys, pts, n, min_voxel_count, rgb, intensity, timestamps, out_pts, out_rgb, out_intensity,
                            out_timestamps);
}

// I/algorithms/filter/preprocess_operator/box_filter_operator.hpp:36-44, common.hpp:15-25,
// I/algorithms/common/filter_by_flags.hpp:43-49 — order-preserving; returns kept count.
size_t orc_box_filter(const float* pts, size_t n, float min_d, float max_d, float* out) {
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        const float* p = pts + 4 * i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && std::isfinite(p[3]))) continue;
        const float linf = std::max(std::fabs(p[0]), std::max(std::fabs(p[1]), std::fabs(p[2])));
        if (linf < min_d || linf > max_d) continue;
        std::memcpy(out + 4 * m, p, 16);
        ++m;
    }
    return m;
}

void orc_default_params(orc_reg_params* p) {
    p->reg_type = R_GICP;
    p->loss = L_NONE;
    p->opt_method = 0;
    p->max_iterations = 20;
    p->max_corr_dist = 2.0f;
    p->robust_default_scale = 10.0f;
    p->crit_translation = 1e-3f;
    p->crit_rotation = 1e-3f;
    p->gn_lambda = 1.0f;
    p->lm_max_inner = 10;
    p->lm_lambda_factor = 2.0f;
    p->lm_init_lambda = 1.0f;
    p->lm_max_lambda = 1e3f;
    p->lm_min_lambda = 1e-6f;
    p->dl_init_radius = 1.0f;
    p->dl_min_radius = 1e-4f;
    p->dl_max_radius = 10.0f;
    p->dl_eta1 = 0.25f;
    p->dl_eta2 = 0.75f;
    p->dl_gamma_dec = 0.25f;
    p->dl_gamma_inc = 2.0f;
    p->sum_mode = 1;
    p->knn_mode = 0;
}

// I/algorithms/registration/registration.hpp:201-276 (loop), :803-828 (GN), :830-895 (LM),
// :897-964 (dog-leg), :407-410 (convergence).  `tree` is an orc_kdtree built on tgt_pts
// (ignored for knn_mode 2).  robust_scale <= 0 selects params->robust_default_scale (:217-218).
// trace_T (nullable): max_iterations*16 floats receiving the pose after every outer iteration.
// ------------------------------------------------------------------ solver add-ons (default off)
// DegenerateRegularizationParams degenerate_regularization.hpp:35-40, MapPriorParams map_prior.hpp:15-21
struct orc_addons {
    int32_t degenerate_type;  // 0 none, 1 nl_reg
    float rot_thr, trans_thr, base_factor;
    int32_t map_prior_enabled;
    float rot_vel_sigma, trans_vel_sigma, rot_base_sigma, trans_base_sigma;
};
static orc_addons g_addons = {0, 10.0f, 1.0f, 1.0f, 0, 1.0f, 1.0f, 3.16e-2f, 1e-2f};
static bool g_prior_active = false;
static double g_prior_omega[6][6];
static M4 g_prior_T_pred_inv;

// symmetric 3x3 eigen-decomposition in fp64 (stands in for Eigen::SelfAdjointEigenSolver<Matrix3f>: third-party,
// unpinned — SURVEY.md §8(c)); cyclic Jacobi, eigenvectors in the columns of V
static void orc_jacobi3(const double A[3][3], double ev[3], double V[3][3]) {
    double a[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[i][j] = 0.5 * (A[i][j] + A[j][i]);
            V[i][j] = i == j;
        }
    for (int sweep = 0; sweep < 64; ++sweep) {
        if (a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2] < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - sn * y; a[k][q] = sn * x + c * y; }
                for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - sn * y; a[q][k] = sn * x + c * y; }
                for (int k = 0; k < 3; ++k) { const double x = V[k][p], y = V[k][q]; V[k][p] = c * x - sn * y; V[k][q] = sn * x + c * y; }
            }
    }
    for (int i = 0; i < 3; ++i) ev[i] = a[i][i];
}

static M4 isometry_inverse(const M4& T) {
    M4 r = M4::identity();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) r(i, j) = T(j, i);
        r(i, 3) = -(T(0, i) * T(0, 3) + T(1, i) * T(1, 3) + T(2, i) * T(2, 3));
    }
    return r;
}

// degenerate_regularization.hpp:58-112 on row-major H[36], b[6]
static void orc_nl_reg(const orc_addons& A, float* H, float* b, uint32_t inlier, const M4& T_cur, const M4& T_init) {
    if (inlier == 0 || A.degenerate_type != 1) return;
    const float lambda = A.base_factor * (float)inlier;
    float Hp[36] = {};
    for (int blk = 0; blk < 2; ++blk) {
        const float thr = blk == 0 ? A.rot_thr : A.trans_thr;
        if (!(thr > 0.0f)) continue;
        double B[3][3], ev[3], V[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) B[i][j] = H[(3 * blk + i) * 6 + 3 * blk + j];
        orc_jacobi3(B, ev, V);
        for (int k = 0; k < 3; ++k) {
            if (!((float)ev[k] / (float)inlier < thr)) continue;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Hp[(3 * blk + i) * 6 + 3 * blk + j] += lambda * ((float)V[i][k] * (float)V[j][k]);
        }
    }
    const V6 tw = se3_log(isometry_mul(isometry_inverse(T_init), T_cur));
    for (int i = 0; i < 6; ++i) {
        float acc = 0.0f;
        for (int j = 0; j < 6; ++j) acc += Hp[i * 6 + j] * tw(j);
        b[i] += acc;
    }
    for (int i = 0; i < 36; ++i) H[i] += Hp[i];
}

// map_prior.hpp:119-146
static float orc_prior_terms(const M4& T, float* omega_e) {
    const V6 e = se3_log(isometry_mul(g_prior_T_pred_inv, T));
    float dotv = 0.0f;
    for (int i = 0; i < 6; ++i) {
        float acc = 0.0f;
        for (int j = 0; j < 6; ++j) acc += (float)g_prior_omega[i][j] * e(j);
        if (omega_e) omega_e[i] = acc;
        dotv += e(i) * acc;
    }
    return 0.5f * dotv;
}

void orc_align(const orc_reg_params* P, const float* src_pts, const float* src_covs, size_t ns, const float* tgt_pts,
               const float* tgt_covs, const float* tgt_normals, size_t nt, void* tree, const float* T_init16,
               float robust_scale_opt, orc_reg_result* R, float* trace_T) {
    std::memset(R, 0, sizeof(*R));
    M4 T = load_T(T_init16);
    const M4 T_initial = T;
    store_T(T, R->T);
    R->error = FMAX;
    R->error_raw = FMAX;
    if (ns == 0) return;
    int loss = P->loss;
    if (loss != L_NONE && P->robust_default_scale <= 0.0f) loss = L_NONE;  // registration.hpp:186-192
    const Clouds c{src_pts, src_covs, ns, tgt_pts, tgt_covs, tgt_normals};
    const float scale = robust_scale_opt > 0.0f ? robust_scale_opt : P->robust_default_scale;
    const float max2 = P->max_corr_dist * P->max_corr_dist;
    float radius = P->dl_init_radius;
    float lambda = P->lm_init_lambda;
    std::vector<int32_t> idx(ns);
    std::vector<float> dist(ns);
    auto converged = [&](const float* d) {
        return norm3(d) < P->crit_rotation && norm3(d + 3) < P->crit_translation;
    };
    auto apply = [&](const M4& Tc, const float* d) {
        V6 a;
        for (int i = 0; i < 6; ++i) a(i) = d[i];
        return isometry_mul(Tc, se3_exp(a));
    };
    for (int iter = 0; iter < P->max_iterations; ++iter) {
        float T16[16];
        store_T(T, T16);
        if (P->knn_mode == 2)
            orc_knn_bruteforce(src_pts, ns, tgt_pts, nt, 1, T16, idx.data(), dist.data());
        else
            orc_kdtree_knn(tree, src_pts, ns, 1, T16, P->knn_mode, idx.data(), dist.data());
        Linearized lin = linearize(P->reg_type, loss, c, idx.data(), dist.data(), T, max2, scale, P->sum_mode);
        std::memcpy(R->H_raw, lin.H, sizeof(lin.H));
        std::memcpy(R->b_raw, lin.b, sizeof(lin.b));
        R->error_raw = lin.error;
        Linearized& linm = lin;
        orc_nl_reg(g_addons, linm.H, linm.b, lin.inlier, T, T_initial);  // registration.hpp:249-250
        if (g_prior_active) {                                            // :253
            float oe[6];
            linm.error += orc_prior_terms(T, oe);
            for (int i = 0; i < 36; ++i) linm.H[i] += (float)g_prior_omega[i / 6][i % 6];
            for (int i = 0; i < 6; ++i) linm.b[i] += oe[i];
        }

        if (P->opt_method == 0) {  // Gauss-Newton
            float d[6];
            const bool ok = solve_system(lin.H, lin.b, P->gn_lambda, d);
            R->converged = ok ? converged(d) : 0;
            T = apply(T, d);
            R->iterations = iter;
            std::memcpy(R->H, lin.H, sizeof(lin.H));
            std::memcpy(R->b, lin.b, sizeof(lin.b));
            R->error = lin.error;
            R->inlier = lin.inlier;
        } else if (P->opt_method == 1) {  // Levenberg-Marquardt
            const float cur = lin.error;
            float last = FMAX;
            float d[6];
            for (int in = 0; in < P->lm_max_inner; ++in) {
                const bool ok = solve_system(lin.H, lin.b, lambda, d);
                R->converged = ok ? converged(d) : 0;
                const M4 Tn = apply(T, d);
                float ne;
                uint32_t inl;
                error_sum(P->reg_type, loss, c, idx.data(), dist.data(), Tn, max2, scale, P->sum_mode, &ne, &inl);
                if (g_prior_active) ne += orc_prior_terms(Tn, nullptr);  // registration.hpp:854
                if (ne <= cur) {
                    R->converged = converged(d);
                    T = Tn;
                    R->error = ne;
                    R->inlier = inl;
                    lambda = std::clamp(lambda / P->lm_lambda_factor, P->lm_min_lambda, P->lm_max_lambda);
                    break;
                } else if (std::fabs(ne - last) <= 1e-6f) {
                    R->converged = converged(d);
                    T = Tn;
                    R->error = ne;
                    R->inlier = inl;
                    break;
                } else {
                    lambda = std::clamp(lambda * P->lm_lambda_factor, P->lm_min_lambda, P->lm_max_lambda);
                }
                last = ne;
            }
            R->iterations = iter;
            std::memcpy(R->H, lin.H, sizeof(lin.H));
            std::memcpy(R->b, lin.b, sizeof(lin.b));
        } else {  // Powell dog-leg
            std::memcpy(R->H, lin.H, sizeof(lin.H));
            std::memcpy(R->b, lin.b, sizeof(lin.b));
            R->error = lin.error;
            R->inlier = lin.inlier;
            R->iterations = iter;
            auto clampr = [&](float r) { return std::clamp(r, P->dl_min_radius, P->dl_max_radius); };
            radius = clampr(radius);
            const Dogleg dl = dogleg_step(lin.H, lin.b, radius);
            if (dl.predicted_reduction <= 0.0f) {
                radius = clampr(radius * P->dl_gamma_dec);
            } else {
                const M4 Tn = apply(T, dl.p);
                float ne;
                uint32_t inl;
                error_sum(P->reg_type, loss, c, idx.data(), dist.data(), Tn, max2, scale, P->sum_mode, &ne, &inl);
                if (g_prior_active) ne += orc_prior_terms(Tn, nullptr);  // registration.hpp:933
                const float rho = (lin.error - ne) / dl.predicted_reduction;
                if (rho < P->dl_eta1) {
                    radius = clampr(radius * P->dl_gamma_dec);
                } else {
                    R->converged = converged(dl.p);
                    T = Tn;
                    R->error = ne;
                    R->inlier = inl;
                    if (rho > P->dl_eta2 && dl.step_norm >= radius * 0.99f) radius = clampr(radius * P->dl_gamma_inc);
                }
            }
        }
        store_T(T, R->T);
        if (trace_T) store_T(T, trace_T + 16 * (size_t)iter);
        if (R->converged) break;
    }
}

void orc_set_addons(const orc_addons* a) {
    g_addons = *a;
    g_prior_active = false;  // MapPrior::set_params, map_prior.hpp:25-28
}
void orc_degenerate_regularize(const orc_addons* a, float* H36, float* b6, uint32_t inlier, const float* T_cur16,
                               const float* T_init16) {
    orc_nl_reg(*a, H36, b6, inlier, load_T(T_cur16), load_T(T_init16));
}
// MapPrior::update — map_prior.hpp:30-117 (6x6 algebra in fp64; Eigen::LDLT of H + R restated as an exact solve)
int orc_set_map_prior_state(const orc_reg_result* prev, const float* T_pred16, float* omega36_out) {
    g_prior_active = false;
    const orc_addons& A = g_addons;
    if (!A.map_prior_enabled) return 0;
    const float dof = 3.0f * (float)prev->inlier - 6.0f;
    if (dof <= 0.0f) return 0;
    if (!std::isfinite(prev->error_raw) || prev->error_raw < 0.0f) return 0;
    const float s_sq = std::max(1.0f, 2.0f * prev->error_raw / dof);
    const M4 Tp = load_T(T_pred16), To = load_T(prev->T);
    M4 Rrel = M4::identity();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rrel(i, j) = To(0, i) * Tp(0, j) + To(1, i) * Tp(1, j) + To(2, i) * Tp(2, j);
    const V6 tw = se3_log(Rrel);
    float dt[3], dtb[3];
    for (int i = 0; i < 3; ++i) dt[i] = Tp(i, 3) - To(i, 3);
    for (int i = 0; i < 3; ++i) dtb[i] = Tp(0, i) * dt[0] + Tp(1, i) * dt[1] + Tp(2, i) * dt[2];
    double q[6];
    for (int i = 0; i < 3; ++i) {
        q[i] = std::fabs(tw(i)) * A.rot_vel_sigma * A.rot_vel_sigma + A.rot_base_sigma * A.rot_base_sigma;
        q[3 + i] = std::fabs(dtb[i]) * A.trans_vel_sigma * A.trans_vel_sigma + A.trans_base_sigma * A.trans_base_sigma;
    }
    double Ad[6][6] = {}, Hs[6][6], HA[6][6], Hc[6][6];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ad[i][j] = Ad[3 + i][3 + j] = Rrel(i, j);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) Hs[i][j] = (double)prev->H_raw[i * 6 + j] / s_sq;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0;
            for (int k = 0; k < 6; ++k) acc += Hs[i][k] * Ad[k][j];
            HA[i][j] = acc;
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0;
            for (int k = 0; k < 6; ++k) acc += Ad[k][i] * HA[k][j];
            Hc[i][j] = acc;
        }
    // X = (H + R)^-1 R by Gaussian elimination with partial pivoting
    double M[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            M[i][j] = Hc[i][j] + (i == j ? 1.0 / q[i] : 0.0);
            M[i][6 + j] = i == j ? 1.0 / q[i] : 0.0;
        }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r)
            if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
        if (!(std::fabs(M[piv][c]) > 1e-300)) return 0;
        for (int k = 0; k < 12; ++k) std::swap(M[c][k], M[piv][k]);
        const double inv = 1.0 / M[c][c];
        for (int k = 0; k < 12; ++k) M[c][k] *= inv;
        for (int r = 0; r < 6; ++r)
            if (r != c) {
                const double f = M[r][c];
                for (int k = 0; k < 12; ++k) M[r][k] -= f * M[c][k];
            }
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            g_prior_omega[i][j] = (double)(float)((i == j ? 1.0 / q[i] : 0.0) - (1.0 / q[i]) * M[i][6 + j]);
            if (omega36_out) omega36_out[i * 6 + j] = (float)g_prior_omega[i][j];
        }
    g_prior_T_pred_inv = isometry_inverse(Tp);
    g_prior_active = true;
    return 1;
}

// I/algorithms/registration/pipeline/robust.hpp:84-87,106-110 — geometric robust-scale schedule
void orc_robust_schedule(float init_scale, float min_scale, int levels, float* scales_out) {
    const float factor = levels > 1 ? std::pow(min_scale / init_scale, 1.0f / (float)(levels - 1)) : 1.0f;
    float s = init_scale;
    for (int l = 0; l < levels; ++l) {
        scales_out[l] = s;
        s *= factor;
    }
}

// robust.hpp:42-114 with auto_scale enabled and a non-NONE loss: one align per level, each
// seeded with the previous level's pose.
void orc_align_robust(const orc_reg_params* P, const float* src_pts, const float* src_covs, size_t ns,
                      const float* tgt_pts, const float* tgt_covs, const float* tgt_normals, size_t nt, void* tree,
                      const float* T_init16, float init_scale, float min_scale, int levels, orc_reg_result* R) {
    std::vector<float> scales((size_t)std::max(levels, 1));
    orc_robust_schedule(init_scale, min_scale, std::max(levels, 1), scales.data());
    float T16[16];
    std::memcpy(T16, T_init16, sizeof(T16));
    for (int l = 0; l < std::max(levels, 1); ++l) {
        orc_align(P, src_pts, src_covs, ns, tgt_pts, tgt_covs, tgt_normals, nt, tree, T16, scales[l], R, nullptr);
        std::memcpy(T16, R->T, sizeof(T16));
    }
}

}  // extern "C"
