// oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A float64 CPU restatement of the arithmetic the reference's ICP plugin
// delegates to Open3D (legacy CPU pipeline).  The reference itself holds no
// arithmetic for this path: every call below is reached from
//   src/slam/scripts/point_cloud_processors/icp_processor.py:59-74,90-113,130-146
// and is executed by Open3D, an un-pinned, un-vendored conda dependency
// (environment.yml:14) that is neither on disk nor installable here.
//
// PARITY UNPINNED: the reference ships no golden vector / known-answer test
// for this path (src/slam/test/ holds only lint tests) and Open3D could not be
// run here.  The semantics restated below follow the published Open3D 0.17-0.19
// legacy pipeline as summarised in SURVEY.md Appendix A (A.1 .. A.8); each
// function names the Open3D entry point and the reference call site it serves.
// Known-answer tests for this file are authored in tests/ (analytic cases) and
// it is cross-checked against an independent NumPy/SciPy twin (oracle/twin.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product path
// (3d-lidar-slam_b200/) never links, imports or calls it.
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp -ffp-contract=off).  FP
// contraction is disabled so that the fixed operation order written here is
// the operation order executed (x86-64 baseline Open3D builds have no FMA).

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 {
    double x, y, z;
};

inline V3 ld(const double* p, int64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// ----------------------------------------------------------------------------
// A.7  PointCloud::Transform — h = T*(p,1); p <- h.xyz / h.w
// (icp_processor.py:73; also applied inside registration_icp, A.3)
// Eigen evaluates the 4x4 * 4x1 product as a left-to-right sum of scaled
// columns: ((c0*x + c1*y) + c2*z) + c3*1.
// ----------------------------------------------------------------------------
inline V3 xform(const double* T, V3 p) {
    double hx = ((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3];
    double hy = ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7];
    double hz = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11];
    double hw = ((T[12] * p.x + T[13] * p.y) + T[14] * p.z) + T[15];
    return {hx / hw, hy / hw, hz / hw};
}

inline void mat4_mul(const double* A, const double* B, double* C) {
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            r[4 * i + j] = s;
        }
    std::memcpy(C, r, sizeof(r));
}

inline void mat4_identity(double* T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

inline bool mat4_is_identity(const double* T) {
    // Eigen::MatrixBase::isIdentity(prec = 1e-12): off-diagonals must be
    // "much smaller than 1", diagonals approximately 1.
    const double prec = 1e-12;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double v = T[4 * i + j];
            if (i == j) {
                if (std::fabs(v - 1.0) > prec * std::min(std::fabs(v), 1.0)) return false;
            } else {
                if (std::fabs(v) > prec) return false;
            }
        }
    return true;
}

// ----------------------------------------------------------------------------
// KD-tree (stands in for nanoflann inside open3d::geometry::KDTreeFlann, A.4).
// Exact search; results ordered by (d2, index) so exact-distance ties resolve
// to the lowest index (Open3D leaves ties to traversal order = unspecified).
// Squared distance is accumulated dimension by dimension from zero, like
// nanoflann's L2_Simple_Adaptor.
// ----------------------------------------------------------------------------
struct KDTree {
    struct Node {
        int32_t left = -1, right = -1;  // children (internal) or [begin,end) (leaf)
        int32_t dim = -1;               // -1 => leaf
        double split_lo = 0, split_hi = 0;
    };
    const double* pts = nullptr;
    int64_t n = 0;
    std::vector<int32_t> order;
    std::vector<Node> nodes;
    static constexpr int kLeaf = 12;

    void build(const double* p, int64_t count) {
        pts = p;
        n = count;
        order.resize(n);
        std::iota(order.begin(), order.end(), 0);
        nodes.clear();
        nodes.reserve(n / 4 + 8);
        if (n > 0) build_rec(0, (int32_t)n);
    }

    int32_t build_rec(int32_t b, int32_t e) {
        int32_t id = (int32_t)nodes.size();
        nodes.emplace_back();
        if (e - b <= kLeaf) {
            nodes[id].left = b;
            nodes[id].right = e;
            return id;
        }
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int32_t i = b; i < e; ++i)
            for (int d = 0; d < 3; ++d) {
                double v = pts[3 * (int64_t)order[i] + d];
                lo[d] = std::min(lo[d], v);
                hi[d] = std::max(hi[d], v);
            }
        int dim = 0;
        for (int d = 1; d < 3; ++d)
            if (hi[d] - lo[d] > hi[dim] - lo[dim]) dim = d;
        int32_t m = b + (e - b) / 2;
        std::nth_element(order.begin() + b, order.begin() + m, order.begin() + e,
                         [&](int32_t a, int32_t c) {
                             return pts[3 * (int64_t)a + dim] < pts[3 * (int64_t)c + dim];
                         });
        double lmax = -1e300, rmin = 1e300;
        for (int32_t i = b; i < m; ++i) lmax = std::max(lmax, pts[3 * (int64_t)order[i] + dim]);
        for (int32_t i = m; i < e; ++i) rmin = std::min(rmin, pts[3 * (int64_t)order[i] + dim]);
        nodes[id].dim = dim;
        nodes[id].split_lo = lmax;
        nodes[id].split_hi = rmin;
        int32_t l = build_rec(b, m);
        int32_t r = build_rec(m, e);
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
    }

    static inline double dist2(const double* a, V3 q) {
        double r = 0.0;
        double d0 = q.x - a[0];
        r += d0 * d0;
        double d1 = q.y - a[1];
        r += d1 * d1;
        double d2 = q.z - a[2];
        r += d2 * d2;
        return r;
    }

    // Bounded result set: keeps the k best (d2, idx) pairs in ascending order.
    struct KBest {
        int k;
        int cnt = 0;
        double* d2;
        int32_t* idx;
        double worst() const { return cnt < k ? std::numeric_limits<double>::infinity() : d2[cnt - 1]; }
        void push(double d, int32_t i) {
            if (cnt == k) {
                if (d > d2[k - 1] || (d == d2[k - 1] && i > idx[k - 1])) return;
                --cnt;
            }
            int pos = cnt;
            while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
                d2[pos] = d2[pos - 1];
                idx[pos] = idx[pos - 1];
                --pos;
            }
            d2[pos] = d;
            idx[pos] = i;
            ++cnt;
        }
    };

    void knn_rec(int32_t id, V3 q, KBest& best) const {
        const Node& nd = nodes[id];
        if (nd.dim < 0) {
            for (int32_t i = nd.left; i < nd.right; ++i) {
                int32_t j = order[i];
                best.push(dist2(pts + 3 * (int64_t)j, q), j);
            }
            return;
        }
        double v = nd.dim == 0 ? q.x : (nd.dim == 1 ? q.y : q.z);
        double gl = v - nd.split_lo;  // distance to the left half-space (<=0: inside it)
        double gr = nd.split_hi - v;  // distance to the right half-space
        if (gl < 0) gl = 0;
        if (gr < 0) gr = 0;
        bool left_first = gl <= gr;
        int32_t first = left_first ? nd.left : nd.right;
        int32_t second = left_first ? nd.right : nd.left;
        double gap_second = left_first ? gr : gl;
        knn_rec(first, q, best);
        // "<=" keeps equal-distance candidates reachable so the (d2, idx) tie
        // rule is honoured.
        if (gap_second * gap_second <= best.worst()) knn_rec(second, q, best);
    }

    int knn(V3 q, int k, double* d2, int32_t* idx) const {
        if (n == 0 || k <= 0) return 0;
        KBest best{k, 0, d2, idx};
        knn_rec(0, q, best);
        return best.cnt;
    }

    void radius_count_rec(int32_t id, V3 q, double r2, int64_t& cnt) const {
        const Node& nd = nodes[id];
        if (nd.dim < 0) {
            for (int32_t i = nd.left; i < nd.right; ++i)
                if (dist2(pts + 3 * (int64_t)order[i], q) < r2) ++cnt;  // nanoflann RadiusResultSet: dist < radius
            return;
        }
        double v = nd.dim == 0 ? q.x : (nd.dim == 1 ? q.y : q.z);
        double dl = v - nd.split_lo, dr = nd.split_hi - v;
        if (dl <= 0 || dl * dl <= r2) radius_count_rec(nd.left, q, r2, cnt);
        if (dr <= 0 || dr * dr <= r2) radius_count_rec(nd.right, q, r2, cnt);
    }
};

// A.4  KDTreeFlann::SearchHybrid(q, radius, max_nn): kNN(max_nn) ascending,
// then keep the prefix with d2 < radius^2 (strict).
inline int search_hybrid(const KDTree& t, V3 q, double radius, int max_nn, double* d2, int32_t* idx) {
    int k = t.knn(q, max_nn, d2, idx);
    double r2 = radius * radius;
    int keep = 0;
    while (keep < k && d2[keep] < r2) ++keep;
    return keep;
}

// ----------------------------------------------------------------------------
// 3x3 helpers
// ----------------------------------------------------------------------------
inline double det3(const double M[9]) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
           M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// Two-sided Jacobi SVD of a real 3x3 (the algorithm family of
// Eigen::JacobiSVD): M = U * diag(s) * V^T, s >= 0, descending.
void svd3_two_sided(const double Min[9], double U[9], double S[3], double V[9]) {
    double M[9];
    std::memcpy(M, Min, sizeof(M));
    for (int i = 0; i < 9; ++i) U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    auto at = [](double* A, int r, int c) -> double& { return A[3 * r + c]; };
    const double eps = std::numeric_limits<double>::epsilon();
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool done = true;
        double maxdiag = std::max({std::fabs(M[0]), std::fabs(M[4]), std::fabs(M[8])});
        double thresh = std::max(2.0 * eps * maxdiag, std::numeric_limits<double>::min());
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                if (std::fabs(at(M, p, q)) <= thresh && std::fabs(at(M, q, p)) <= thresh) continue;
                done = false;
                // 2x2 block [[a,b],[c,d]] on rows/cols (q,p) ordered as (p,q) sub-problem
                double a = at(M, p, p), b = at(M, p, q), c = at(M, q, p), d = at(M, q, q);
                // Step 1: rotation R1 (left) that symmetrises the block.
                double t = a + d, dd = c - b;  // rot1 = [[cs, sn],[-sn, cs]]
                double cs1, sn1;
                if (std::fabs(dd) < std::numeric_limits<double>::min()) {
                    cs1 = 1.0;
                    sn1 = 0.0;
                } else {
                    double u = t / dd;
                    double tmp = std::sqrt(1.0 + u * u);
                    sn1 = 1.0 / tmp;
                    cs1 = u / tmp;
                }
                // B = R1 * block ; rows: row_p' = cs*row_p + sn*row_q ; row_q' = -sn*row_p + cs*row_q
                double ap = cs1 * a + sn1 * c, bp = cs1 * b + sn1 * d;
                double cp = -sn1 * a + cs1 * c, dp = -sn1 * b + cs1 * d;
                (void)cp;
                // Step 2: symmetric 2x2 Jacobi on [[ap,bp],[bp,dp]] -> rotation J (right), J^T (left)
                double cs2, sn2;
                if (std::fabs(bp) < std::numeric_limits<double>::min()) {
                    cs2 = 1.0;
                    sn2 = 0.0;
                } else {
                    double tau = (dp - ap) / (2.0 * bp);
                    double tt = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                    cs2 = 1.0 / std::sqrt(1.0 + tt * tt);
                    sn2 = tt * cs2;
                }
                // Left rotation L = J^T * R1, right rotation J.
                // J = [[cs2, sn2],[-sn2, cs2]] acting on columns (p,q).
                double l00 = cs2 * cs1 - sn2 * (-sn1), l01 = cs2 * sn1 - sn2 * cs1;
                double l10 = sn2 * cs1 + cs2 * (-sn1), l11 = sn2 * sn1 + cs2 * cs1;
                for (int k = 0; k < 3; ++k) {  // M <- L * M on rows p,q ; U <- U * L^T on cols p,q
                    double mp = at(M, p, k), mq = at(M, q, k);
                    at(M, p, k) = l00 * mp + l01 * mq;
                    at(M, q, k) = l10 * mp + l11 * mq;
                    double up = at(U, k, p), uq = at(U, k, q);
                    at(U, k, p) = l00 * up + l01 * uq;
                    at(U, k, q) = l10 * up + l11 * uq;
                }
                for (int k = 0; k < 3; ++k) {  // M <- M * J on cols p,q ; V <- V * J
                    double mp = at(M, k, p), mq = at(M, k, q);
                    at(M, k, p) = cs2 * mp - sn2 * mq;
                    at(M, k, q) = sn2 * mp + cs2 * mq;
                    double vp = at(V, k, p), vq = at(V, k, q);
                    at(V, k, p) = cs2 * vp - sn2 * vq;
                    at(V, k, q) = sn2 * vp + cs2 * vq;
                }
            }
        if (done) break;
    }
    for (int i = 0; i < 3; ++i) {
        double s = M[4 * i];
        if (s < 0) {
            s = -s;
            for (int k = 0; k < 3; ++k) at(U, k, i) = -at(U, k, i);
        }
        S[i] = s;
    }
    for (int i = 0; i < 2; ++i) {  // sort descending, permuting columns of U and V
        int best = i;
        for (int j = i + 1; j < 3; ++j)
            if (S[j] > S[best]) best = j;
        if (best != i) {
            std::swap(S[i], S[best]);
            for (int k = 0; k < 3; ++k) {
                std::swap(at(U, k, i), at(U, k, best));
                std::swap(at(V, k, i), at(V, k, best));
            }
        }
    }
}

// ----------------------------------------------------------------------------
// A.5  TransformationEstimationPointToPoint (Eigen::umeyama, no scaling)
// (icp_processor.py:107)
// ----------------------------------------------------------------------------
void umeyama_from_corr(const double* src, const double* tgt, const int32_t* corr_src,
                       const int32_t* corr_tgt, int64_t m, double* T) {
    mat4_identity(T);
    if (m == 0) return;
    double one_over_n = 1.0 / (double)m;
    V3 ms{0, 0, 0}, md{0, 0, 0};
    for (int64_t k = 0; k < m; ++k) {
        V3 p = ld(src, corr_src[k]), q = ld(tgt, corr_tgt[k]);
        ms.x += p.x; ms.y += p.y; ms.z += p.z;
        md.x += q.x; md.y += q.y; md.z += q.z;
    }
    ms = {ms.x * one_over_n, ms.y * one_over_n, ms.z * one_over_n};
    md = {md.x * one_over_n, md.y * one_over_n, md.z * one_over_n};
    double Sg[9] = {0};
    for (int64_t k = 0; k < m; ++k) {
        V3 p = ld(src, corr_src[k]), q = ld(tgt, corr_tgt[k]);
        double a[3] = {q.x - md.x, q.y - md.y, q.z - md.z};
        double b[3] = {p.x - ms.x, p.y - ms.y, p.z - ms.z};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Sg[3 * i + j] += a[i] * b[j];
    }
    for (double& v : Sg) v *= one_over_n;
    double U[9], S[3], V[9];
    svd3_two_sided(Sg, U, S, V);
    double s[3] = {1.0, 1.0, 1.0};
    if (det3(U) * det3(V) < 0) s[2] = -1.0;
    double R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += U[3 * i + k] * s[k] * V[3 * j + k];
            R[3 * i + j] = acc;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
    double mdv[3] = {md.x, md.y, md.z}, msv[3] = {ms.x, ms.y, ms.z};
    for (int i = 0; i < 3; ++i) {
        double rs = 0.0;
        for (int j = 0; j < 3; ++j) rs += R[3 * i + j] * msv[j];
        T[4 * i + 3] = mdv[i] - rs;
    }
}

// LDLT-style solve of a symmetric 6x6 system (Eigen: JTJ.ldlt().solve(-JTr)).
// Plain LDL^T without pivoting is adequate for the SPD systems ICP produces;
// a non-finite or vanishing pivot reports failure like Open3D's
// SolveLinearSystemPSD does through its determinant check.
bool ldlt6_solve(const double A_in[36], const double b[6], double x[6]) {
    double L[36] = {0}, D[6];
    for (int j = 0; j < 6; ++j) {
        double d = A_in[6 * j + j];
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k] * D[k];
        if (!(std::fabs(d) > 0.0) || !std::isfinite(d)) return false;
        D[j] = d;
        L[6 * j + j] = 1.0;
        for (int i = j + 1; i < 6; ++i) {
            double v = A_in[6 * i + j];
            for (int k = 0; k < j; ++k) v -= L[6 * i + k] * L[6 * j + k] * D[k];
            L[6 * i + j] = v / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= L[6 * i + k] * y[k];
        y[i] = v;
    }
    for (int i = 0; i < 6; ++i) y[i] /= D[i];
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < 6; ++k) v -= L[6 * k + i] * x[k];
        x[i] = v;
    }
    for (int i = 0; i < 6; ++i)
        if (!std::isfinite(x[i])) return false;
    return true;
}

// A.6  utility::TransformVector6dToMatrix4d: Rz(x2) * Ry(x1) * Rx(x0), t = x3..5
void vec6_to_mat4(const double x[6], double* T) {
    double ca = std::cos(x[0]), sa = std::sin(x[0]);
    double cb = std::cos(x[1]), sb = std::sin(x[1]);
    double cg = std::cos(x[2]), sg = std::sin(x[2]);
    double Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca};
    double Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb};
    double Rz[9] = {cg, -sg, 0, sg, cg, 0, 0, 0, 1};
    double ZY[9], R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Rz[3 * i + k] * Ry[3 * k + j];
            ZY[3 * i + j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += ZY[3 * i + k] * Rx[3 * k + j];
            R[3 * i + j] = s;
        }
    mat4_identity(T);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = x[3 + i];
    }
}

// A.6  TransformationEstimationPointToPlane::ComputeTransformation
void point_to_plane_from_corr(const double* src, const double* tgt, const double* tgt_normals,
                              const int32_t* corr_src, const int32_t* corr_tgt, int64_t m, double* T) {
    mat4_identity(T);
    if (m == 0) return;
    double JTJ[36] = {0}, JTr[6] = {0};
    for (int64_t k = 0; k < m; ++k) {
        V3 p = ld(src, corr_src[k]), q = ld(tgt, corr_tgt[k]), nq = ld(tgt_normals, corr_tgt[k]);
        double r = (p.x - q.x) * nq.x + (p.y - q.y) * nq.y + (p.z - q.z) * nq.z;
        double J[6] = {p.y * nq.z - p.z * nq.y, p.z * nq.x - p.x * nq.z, p.x * nq.y - p.y * nq.x,
                       nq.x, nq.y, nq.z};
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) JTJ[6 * i + j] += J[i] * J[j];
            JTr[i] += J[i] * r;
        }
    }
    double rhs[6], x[6];
    for (int i = 0; i < 6; ++i) rhs[i] = -JTr[i];
    // utility::SolveLinearSystemPSD with its defaults (check_det = false): x = JTJ.ldlt().solve(-JTr);
    // a break-down of the factorisation yields the identity update.
    if (!ldlt6_solve(JTJ, rhs, x)) return;
    vec6_to_mat4(x, T);
}

// ----------------------------------------------------------------------------
// A.2  closed-form smallest-eigenvalue eigenvector of a symmetric 3x3
// (Open3D FastEigen3x3; Eberly, "A Robust Eigensolver for 3x3 Symmetric Matrices")
// ----------------------------------------------------------------------------
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }

V3 eigvec0(const double A[9], double eval0) {
    V3 r0{A[0] - eval0, A[1], A[2]}, r1{A[1], A[4] - eval0, A[5]}, r2{A[2], A[5], A[8] - eval0};
    V3 c01 = cross(r0, r1), c02 = cross(r0, r2), c12 = cross(r1, r2);
    double d0 = dot(c01, c01), d1 = dot(c02, c02), d2 = dot(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    if (imax == 0) return scale(c01, 1.0 / std::sqrt(d0));
    if (imax == 1) return scale(c02, 1.0 / std::sqrt(d1));
    return scale(c12, 1.0 / std::sqrt(d2));
}

V3 eigvec1(const double A[9], V3 e0, double eval1) {
    V3 U;
    if (std::fabs(e0.x) > std::fabs(e0.y)) {
        double inv = 1.0 / std::sqrt(e0.x * e0.x + e0.z * e0.z);
        U = {-e0.z * inv, 0.0, e0.x * inv};
    } else {
        double inv = 1.0 / std::sqrt(e0.y * e0.y + e0.z * e0.z);
        U = {0.0, e0.z * inv, -e0.y * inv};
    }
    V3 Vv = cross(e0, U);
    V3 AU{A[0] * U.x + A[1] * U.y + A[2] * U.z, A[1] * U.x + A[4] * U.y + A[5] * U.z,
          A[2] * U.x + A[5] * U.y + A[8] * U.z};
    V3 AV{A[0] * Vv.x + A[1] * Vv.y + A[2] * Vv.z, A[1] * Vv.x + A[4] * Vv.y + A[5] * Vv.z,
          A[2] * Vv.x + A[5] * Vv.y + A[8] * Vv.z};
    double m00 = dot(U, AU) - eval1, m01 = dot(U, AV), m11 = dot(Vv, AV) - eval1;
    double am00 = std::fabs(m00), am01 = std::fabs(m01), am11 = std::fabs(m11);
    if (am00 >= am11) {
        double mx = std::max(am00, am01);
        if (mx > 0) {
            if (am00 >= am01) {
                m01 /= m00;
                m00 = 1.0 / std::sqrt(1.0 + m01 * m01);
                m01 *= m00;
            } else {
                m00 /= m01;
                m01 = 1.0 / std::sqrt(1.0 + m00 * m00);
                m00 *= m01;
            }
            return {m01 * U.x - m00 * Vv.x, m01 * U.y - m00 * Vv.y, m01 * U.z - m00 * Vv.z};
        }
        return U;
    } else {
        double mx = std::max(am11, am01);
        if (mx > 0) {
            if (am11 >= am01) {
                m01 /= m11;
                m11 = 1.0 / std::sqrt(1.0 + m01 * m01);
                m01 *= m11;
            } else {
                m11 /= m01;
                m01 = 1.0 / std::sqrt(1.0 + m11 * m11);
                m11 *= m01;
            }
            return {m11 * U.x - m01 * Vv.x, m11 * U.y - m01 * Vv.y, m11 * U.z - m01 * Vv.z};
        }
        return U;
    }
}

V3 fast_eigen3x3_smallest(const double Cin[9]) {
    double A[9];
    std::memcpy(A, Cin, sizeof(A));
    double max_coeff = A[0];
    for (int i = 1; i < 9; ++i) max_coeff = std::max(max_coeff, A[i]);
    if (max_coeff == 0) return {0, 0, 0};
    for (double& v : A) v /= max_coeff;
    double norm = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (norm > 0) {
        double q = (A[0] + A[4] + A[8]) / 3.0;
        double b00 = A[0] - q, b11 = A[4] - q, b22 = A[8] - q;
        double p = std::sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0);
        double c00 = b11 * b22 - A[5] * A[5];
        double c01 = A[1] * b22 - A[5] * A[2];
        double c02 = A[1] * A[5] - b11 * A[2];
        double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = std::min(std::max(det * 0.5, -1.0), 1.0);
        double angle = std::acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = std::cos(angle) * 2.0;
        double beta0 = std::cos(angle + two_thirds_pi) * 2.0;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (half_det >= 0) {
            V3 v2 = eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            V3 v1 = eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v1, v2);
        } else {
            V3 v0 = eigvec0(A, e0);
            if (e0 < e1 && e0 < e2) return v0;
            V3 v1 = eigvec1(A, v0, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v0, v1);
        }
    }
    if (A[0] < A[4] && A[0] < A[8]) return {1, 0, 0};
    if (A[4] < A[0] && A[4] < A[8]) return {0, 1, 0};
    return {0, 0, 1};
}

struct KeyHash {
    size_t operator()(const std::array<int32_t, 3>& k) const {
        // open3d::utility::hash_eigen
        size_t seed = 0;
        for (int i = 0; i < 3; ++i) seed ^= std::hash<int32_t>()(k[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};

}  // namespace

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// GetMinBound / GetMaxBound
void orc_min_max_bound(const double* pts, int64_t n, double* mn, double* mx) {
    for (int d = 0; d < 3; ++d) {
        mn[d] = n ? pts[d] : 0.0;
        mx[d] = n ? pts[d] : 0.0;
    }
    for (int64_t i = 1; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            mn[d] = std::min(mn[d], pts[3 * i + d]);
            mx[d] = std::max(mx[d], pts[3 * i + d]);
        }
}

// A.1 key: floor((p - origin) / v) per axis, int32.  origin == NULL selects
// Open3D's own origin, min_bound - v/2.
void orc_voxel_keys(const double* pts, int64_t n, double voxel, const double* origin, int32_t* keys) {
    double org[3];
    if (origin) {
        std::memcpy(org, origin, sizeof(org));
    } else {
        double mn[3], mx[3];
        orc_min_max_bound(pts, n, mn, mx);
        for (int d = 0; d < 3; ++d) org[d] = mn[d] - voxel * 0.5;
    }
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) keys[3 * i + d] = (int32_t)std::floor((pts[3 * i + d] - org[d]) / voxel);
}

// A.1  PointCloud::VoxelDownSample(v)  (icp_processor.py:90-91,132-133)
// Accumulates sum(p) in float64 in INPUT order per voxel (unordered_map of
// AccumulatedPoint), emits sum/n.  Output order is canonicalised to
// lexicographic (ix,iy,iz) — Open3D's is hash-map iteration order.
// Returns the number of voxels; out_pts [n,3], out_keys [n,3], out_counts [n]
// must hold n entries (any of the three may be NULL).
int64_t orc_voxel_downsample(const double* pts, int64_t n, double voxel, const double* origin,
                             double* out_pts, int32_t* out_keys, int32_t* out_counts) {
    if (n == 0) return 0;
    std::vector<int32_t> keys(3 * n);
    orc_voxel_keys(pts, n, voxel, origin, keys.data());
    struct Acc {
        double sx = 0, sy = 0, sz = 0;
        int32_t cnt = 0;
    };
    std::unordered_map<std::array<int32_t, 3>, Acc, KeyHash> vox;
    vox.reserve(n);
    for (int64_t i = 0; i < n; ++i) {
        Acc& a = vox[{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]}];
        a.sx += pts[3 * i];
        a.sy += pts[3 * i + 1];
        a.sz += pts[3 * i + 2];
        a.cnt += 1;
    }
    std::vector<std::pair<std::array<int32_t, 3>, Acc>> items(vox.begin(), vox.end());
    std::sort(items.begin(), items.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    int64_t m = (int64_t)items.size();
    for (int64_t j = 0; j < m; ++j) {
        const Acc& a = items[j].second;
        if (out_pts) {
            out_pts[3 * j] = a.sx / (double)a.cnt;
            out_pts[3 * j + 1] = a.sy / (double)a.cnt;
            out_pts[3 * j + 2] = a.sz / (double)a.cnt;
        }
        if (out_keys)
            for (int d = 0; d < 3; ++d) out_keys[3 * j + d] = items[j].first[d];
        if (out_counts) out_counts[j] = a.cnt;
    }
    return m;
}

// A.7  PointCloud::Transform
void orc_transform(const double* pts, int64_t n, const double* T, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        V3 r = xform(T, ld(pts, i));
        out[3 * i] = r.x;
        out[3 * i + 1] = r.y;
        out[3 * i + 2] = r.z;
    }
}

// A.3 Eval: GetRegistrationResultAndCorrespondences — per source point
// SearchHybrid(p, max_corr, 1).  idx[i] = target index or -1; d2[i] valid when
// idx[i] >= 0.  Returns the number of correspondences.
int64_t orc_nn_search(const double* src, int64_t ns, const double* tgt, int64_t nt, double radius,
                      int32_t* idx, double* d2) {
    KDTree tree;
    tree.build(tgt, nt);
    int64_t cnt = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cnt)
    for (int64_t i = 0; i < ns; ++i) {
        double d;
        int32_t j;
        int k = search_hybrid(tree, ld(src, i), radius, 1, &d, &j);
        idx[i] = k ? j : -1;
        if (d2) d2[i] = k ? d : 0.0;
        cnt += k;
    }
    return cnt;
}

// Brute-force twin of orc_nn_search for small known-answer tests.
int64_t orc_nn_search_brute(const double* src, int64_t ns, const double* tgt, int64_t nt, double radius,
                            int32_t* idx, double* d2) {
    int64_t cnt = 0;
    double r2 = radius * radius;
    for (int64_t i = 0; i < ns; ++i) {
        double best = std::numeric_limits<double>::infinity();
        int32_t bj = -1;
        V3 q = ld(src, i);
        for (int64_t j = 0; j < nt; ++j) {
            double d = KDTree::dist2(tgt + 3 * j, q);
            if (d < best) {
                best = d;
                bj = (int32_t)j;
            }
        }
        if (bj >= 0 && best < r2) {
            idx[i] = bj;
            if (d2) d2[i] = best;
            ++cnt;
        } else {
            idx[i] = -1;
            if (d2) d2[i] = 0.0;
        }
    }
    return cnt;
}

// kNN with optional radius cap (radius <= 0: plain SearchKNN).  out_idx/out_d2
// are [n, k]; out_cnt [n].
void orc_knn(const double* query, int64_t nq, const double* tgt, int64_t nt, int k, double radius,
             int32_t* out_idx, double* out_d2, int32_t* out_cnt) {
    KDTree tree;
    tree.build(tgt, nt);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; ++i) {
        int32_t* ii = out_idx + (int64_t)k * i;
        double* dd = out_d2 + (int64_t)k * i;
        int c = radius > 0 ? search_hybrid(tree, ld(query, i), radius, k, dd, ii)
                           : tree.knn(ld(query, i), k, dd, ii);
        for (int j = c; j < k; ++j) {
            ii[j] = -1;
            dd[j] = 0.0;
        }
        out_cnt[i] = c;
    }
}

// Candidate count of the 27-cell stencil used for the algorithmic-bytes model
// (SURVEY.md §8d): cell = floor((p - origin)/cell_size); returns sum over
// queries of the number of target points in the 27 neighbouring cells.
int64_t orc_stencil_candidates(const double* query, int64_t nq, const double* tgt, int64_t nt,
                               double cell_size, const double* origin) {
    std::unordered_map<std::array<int32_t, 3>, int32_t, KeyHash> cells;
    cells.reserve(nt);
    auto key = [&](const double* p) {
        return std::array<int32_t, 3>{(int32_t)std::floor((p[0] - origin[0]) / cell_size),
                                      (int32_t)std::floor((p[1] - origin[1]) / cell_size),
                                      (int32_t)std::floor((p[2] - origin[2]) / cell_size)};
    };
    for (int64_t j = 0; j < nt; ++j) cells[key(tgt + 3 * j)] += 1;
    int64_t total = 0;
    for (int64_t i = 0; i < nq; ++i) {
        auto c = key(query + 3 * i);
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    auto it = cells.find({c[0] + dx, c[1] + dy, c[2] + dz});
                    if (it != cells.end()) total += it->second;
                }
    }
    return total;
}

// A.5 standalone (for known-answer tests): corr given as two index arrays.
void orc_umeyama(const double* src, const double* tgt, const int32_t* corr_src, const int32_t* corr_tgt,
                 int64_t m, double* T) {
    umeyama_from_corr(src, tgt, corr_src, corr_tgt, m, T);
}

void orc_point_to_plane_step(const double* src, const double* tgt, const double* tgt_normals,
                             const int32_t* corr_src, const int32_t* corr_tgt, int64_t m, double* T) {
    point_to_plane_from_corr(src, tgt, tgt_normals, corr_src, corr_tgt, m, T);
}

void orc_svd3(const double* M, double* U, double* S, double* V) { svd3_two_sided(M, U, S, V); }

void orc_eig3_smallest(const double* C, double* v) {
    V3 r = fast_eigen3x3_smallest(C);
    v[0] = r.x;
    v[1] = r.y;
    v[2] = r.z;
}

// A.2  PointCloud::EstimateNormals(KDTreeSearchParamHybrid(radius, max_nn))
// (icp_processor.py:94-99).  No prior normals => no orientation step.
void orc_estimate_normals(const double* pts, int64_t n, double radius, int max_nn, double* normals) {
    KDTree tree;
    tree.build(pts, n);
#pragma omp parallel
    {
        std::vector<double> d2(max_nn);
        std::vector<int32_t> idx(max_nn);
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            int k = search_hybrid(tree, ld(pts, i), radius, max_nn, d2.data(), idx.data());
            double C[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            if (k >= 3) {
                double c[9] = {0};
                for (int j = 0; j < k; ++j) {
                    V3 p = ld(pts, idx[j]);
                    c[0] += p.x; c[1] += p.y; c[2] += p.z;
                    c[3] += p.x * p.x; c[4] += p.x * p.y; c[5] += p.x * p.z;
                    c[6] += p.y * p.y; c[7] += p.y * p.z; c[8] += p.z * p.z;
                }
                for (double& v : c) v /= (double)k;
                C[0] = c[3] - c[0] * c[0];
                C[4] = c[6] - c[1] * c[1];
                C[8] = c[8] - c[2] * c[2];
                C[1] = C[3] = c[4] - c[0] * c[1];
                C[2] = C[6] = c[5] - c[0] * c[2];
                C[5] = C[7] = c[7] - c[1] * c[2];
            }
            V3 nv = fast_eigen3x3_smallest(C);
            double nn = std::sqrt(dot(nv, nv));
            if (nn == 0.0) nv = {0, 0, 1};
            normals[3 * i] = nv.x;
            normals[3 * i + 1] = nv.y;
            normals[3 * i + 2] = nv.z;
        }
    }
}

// A.3  RegistrationICP  (icp_processor.py:103-113)
// mode 0 = point-to-point (A.5), 1 = point-to-plane (A.6; needs tgt_normals).
// corr_out [ns] (optional) receives the final correspondence map i -> j / -1.
// Returns 0, or -1 when mode 1 is requested without target normals (Open3D
// raises in that case).
int orc_registration_icp(const double* src, int64_t ns, const double* tgt, int64_t nt,
                         const double* tgt_normals, double max_corr, const double* init, int mode,
                         int max_iteration, double rel_fitness, double rel_rmse, double* T_out,
                         double* fitness_out, double* rmse_out, int32_t* iters_out, int32_t* corr_out) {
    if (mode == 1 && !tgt_normals) return -1;
    double T[16];
    std::memcpy(T, init, sizeof(T));
    KDTree tree;
    tree.build(tgt, nt);
    std::vector<double> pcd(src, src + 3 * ns);
    if (!mat4_is_identity(init)) orc_transform(src, ns, init, pcd.data());
    std::vector<int32_t> idx(ns);
    std::vector<double> d2(ns);
    std::vector<int32_t> cs, ct;
    struct Res {
        double fitness = 0, rmse = 0;
        int64_t n = 0;
    };
    auto eval = [&]() {
        Res r;
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < ns; ++i) {
            double d;
            int32_t j;
            int k = search_hybrid(tree, ld(pcd.data(), i), max_corr, 1, &d, &j);
            idx[i] = k ? j : -1;
            d2[i] = k ? d : 0.0;
        }
        cs.clear();
        ct.clear();
        double err2 = 0.0;
        for (int64_t i = 0; i < ns; ++i)
            if (idx[i] >= 0) {
                cs.push_back((int32_t)i);
                ct.push_back(idx[i]);
                err2 += d2[i];
            }
        r.n = (int64_t)cs.size();
        if (r.n > 0) {
            r.fitness = (double)r.n / (double)ns;
            r.rmse = std::sqrt(err2 / (double)r.n);
        }
        return r;
    };
    Res result = eval();
    int it = 0;
    for (; it < max_iteration; ++it) {
        double upd[16];
        if (mode == 0)
            umeyama_from_corr(pcd.data(), tgt, cs.data(), ct.data(), result.n, upd);
        else
            point_to_plane_from_corr(pcd.data(), tgt, tgt_normals, cs.data(), ct.data(), result.n, upd);
        mat4_mul(upd, T, T);
        orc_transform(pcd.data(), ns, upd, pcd.data());
        Res backup = result;
        result = eval();
        if (std::fabs(backup.fitness - result.fitness) < rel_fitness &&
            std::fabs(backup.rmse - result.rmse) < rel_rmse) {
            ++it;
            break;
        }
    }
    std::memcpy(T_out, T, sizeof(T));
    if (fitness_out) *fitness_out = result.fitness;
    if (rmse_out) *rmse_out = result.rmse;
    if (iters_out) *iters_out = it;
    if (corr_out) std::memcpy(corr_out, idx.data(), sizeof(int32_t) * ns);
    return 0;
}

// A.8  RemoveStatisticalOutliers(nb_neighbors, std_ratio)  (icp_processor.py:136,145)
// keep[i] = 1/0; returns the number kept.
int64_t orc_statistical_outlier_mask(const double* pts, int64_t n, int nb_neighbors, double std_ratio,
                                     uint8_t* keep) {
    if (n == 0) return 0;
    KDTree tree;
    tree.build(pts, n);
    std::vector<double> avg(n);
    int64_t valid = 0;
#pragma omp parallel
    {
        std::vector<double> d2(nb_neighbors);
        std::vector<int32_t> idx(nb_neighbors);
#pragma omp for schedule(dynamic, 256) reduction(+ : valid)
        for (int64_t i = 0; i < n; ++i) {
            int k = tree.knn(ld(pts, i), nb_neighbors, d2.data(), idx.data());
            double m = -1.0;
            if (k > 0) {
                valid += 1;
                double s = 0.0;
                for (int j = 0; j < k; ++j) s += std::sqrt(d2[j]);
                m = s / (double)k;
            }
            avg[i] = m;
        }
    }
    if (valid == 0) {
        std::memset(keep, 0, n);
        return 0;
    }
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (avg[i] > 0) sum += avg[i];
    double mean = sum / (double)valid;
    double sq = 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (avg[i] > 0) sq += (avg[i] - mean) * (avg[i] - mean);
    double sd = std::sqrt(sq / (double)(valid - 1));
    double thr = mean + std_ratio * sd;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) {
        keep[i] = (avg[i] > 0 && avg[i] < thr) ? 1 : 0;
        kept += keep[i];
    }
    return kept;
}

// A.8  RemoveRadiusOutliers(nb_points, radius)  (icp_processor.py:139)
int64_t orc_radius_outlier_mask(const double* pts, int64_t n, int nb_points, double radius, uint8_t* keep) {
    KDTree tree;
    tree.build(pts, n);
    int64_t kept = 0;
    double r2 = radius * radius;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : kept)
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = 0;
        tree.radius_count_rec(0, ld(pts, i), r2, c);
        keep[i] = c > nb_points ? 1 : 0;
        kept += keep[i];
    }
    return kept;
}


// ----------------------------------------------------------------------------
// Resident-map CPU leg (the "resident-index CPU variant" of BASELINE.md section 2).
// Same arithmetic as the reference sequence (A.1 down-sample, A.3/A.5 ICP, A.7
// transform, A.1 merge at a FIXED origin) but the map is kept as per-voxel
// (sum, n) accumulators in an open-addressing hash that doubles as the search
// grid, so nothing is re-voxelised or re-indexed per scan: the like-for-like
// CPU counterpart of b2_slam_scan, and the exact C++ twin of
// oracle/pipeline.py:ResidentSequence (tests/test_oracle_kat.py compares them).
// Nearest neighbour = exhaustive scan of the (2r+1)^3 voxel stencil,
// r = ceil(max_corr / voxel); ties resolve to the lexicographically smallest
// voxel key = the lowest index of the canonical (key-ordered) export that
// orc_registration_icp would see.  Centroids are rounded to float32 as the GPU
// stores them.  icp_processor.py:43-119 is the sequence being restated.
// ----------------------------------------------------------------------------
struct ResidentMap {
    double voxel, max_corr, ds_voxel, rel_fitness, rel_rmse;
    double origin[3];
    int max_iteration;
    std::vector<int64_t> tab_key;   // packed key or -1
    std::vector<int32_t> tab_slot;
    uint64_t mask = 0;
    std::vector<std::array<int32_t, 3>> keys;
    std::vector<double> acc;        // [n,4] sum x, y, z, count
    std::vector<double> cent;       // [n,3] float32-rounded centroids, widened
    double pose[16];
    std::vector<double> loc;        // scan-local (sum, n) scratch, zero outside fuse()

    static inline int64_t pack(int32_t x, int32_t y, int32_t z) {
        return ((int64_t)(x & 0x1FFFFF) << 42) | ((int64_t)(y & 0x1FFFFF) << 21) | (int64_t)(z & 0x1FFFFF);
    }
    static inline uint64_t mix(int64_t k) {
        uint64_t h = (uint64_t)k * 0x9E3779B97F4A7C15ull;
        return h ^ (h >> 29);
    }
    void rehash(size_t want) {
        size_t cap = 1024;
        while (cap < want * 2) cap <<= 1;
        tab_key.assign(cap, -1);
        tab_slot.assign(cap, -1);
        mask = cap - 1;
        for (size_t s = 0; s < keys.size(); ++s) {
            uint64_t h = mix(pack(keys[s][0], keys[s][1], keys[s][2])) & mask;
            while (tab_key[h] != -1) h = (h + 1) & mask;
            tab_key[h] = pack(keys[s][0], keys[s][1], keys[s][2]);
            tab_slot[h] = (int32_t)s;
        }
    }
    inline int32_t find(int32_t x, int32_t y, int32_t z) const {
        const int64_t k = pack(x, y, z);
        uint64_t h = mix(k) & mask;
        while (true) {
            const int64_t t = tab_key[h];
            if (t == k) return tab_slot[h];
            if (t == -1) return -1;
            h = (h + 1) & mask;
        }
    }
    int32_t find_or_add(int32_t x, int32_t y, int32_t z) {
        if ((keys.size() + 1) * 2 > tab_key.size()) rehash((keys.size() + 1) * 2);
        const int64_t k = pack(x, y, z);
        uint64_t h = mix(k) & mask;
        while (tab_key[h] != -1) {
            if (tab_key[h] == k) return tab_slot[h];
            h = (h + 1) & mask;
        }
        tab_key[h] = k;
        tab_slot[h] = (int32_t)keys.size();
        keys.push_back({x, y, z});
        acc.insert(acc.end(), {0.0, 0.0, 0.0, 0.0});
        cent.insert(cent.end(), {0.0, 0.0, 0.0});
        return tab_slot[h];
    }
    // A.7 + A.1 at the map's origin: per-scan sums in input order, then one merge per touched voxel
    void fuse(const float* pts, int64_t n, const double* T) {
        std::vector<int32_t> slot(n);
        std::vector<double> moved(3 * n);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            V3 r = xform(T, V3{(double)pts[3 * i], (double)pts[3 * i + 1], (double)pts[3 * i + 2]});
            moved[3 * i] = r.x; moved[3 * i + 1] = r.y; moved[3 * i + 2] = r.z;
        }
        std::vector<int32_t> k3(3 * n);
        orc_voxel_keys(moved.data(), n, voxel, origin, k3.data());
        for (int64_t i = 0; i < n; ++i) slot[i] = find_or_add(k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]);
        // scan-local sums (input order), merged into the accumulators afterwards like the GPU's fuse
        std::vector<int32_t> touched;
        if (loc.size() < 4 * keys.size()) loc.resize(4 * keys.size() + 4 * 65536, 0.0);  // (kept zero between scans)
        for (int64_t i = 0; i < n; ++i) {
            double* a = &loc[4 * (size_t)slot[i]];
            if (a[3] == 0.0) touched.push_back(slot[i]);
            a[0] += moved[3 * i]; a[1] += moved[3 * i + 1]; a[2] += moved[3 * i + 2]; a[3] += 1.0;
        }
        for (int32_t s : touched) {
            double* m = &acc[4 * (size_t)s];
            double* a = &loc[4 * (size_t)s];
            m[0] += a[0]; m[1] += a[1]; m[2] += a[2]; m[3] += a[3];
            a[0] = a[1] = a[2] = a[3] = 0.0;
            for (int d = 0; d < 3; ++d) cent[3 * (size_t)s + d] = (double)(float)(m[d] / m[3]);
        }
    }
};

void* orc_resident_create(double map_voxel, const double* origin, double ds_voxel, double max_corr,
                          int max_iteration, double rel_fitness, double rel_rmse) {
    ResidentMap* m = new ResidentMap();
    m->voxel = map_voxel; m->max_corr = max_corr; m->ds_voxel = ds_voxel;
    m->rel_fitness = rel_fitness; m->rel_rmse = rel_rmse; m->max_iteration = max_iteration;
    std::memcpy(m->origin, origin, sizeof(m->origin));
    mat4_identity(m->pose);
    m->rehash(1024);
    return m;
}
void orc_resident_destroy(void* h) { delete (ResidentMap*)h; }
int64_t orc_resident_size(void* h) { return (int64_t)((ResidentMap*)h)->keys.size(); }
void orc_resident_set_pose(void* h, const double* T) { std::memcpy(((ResidentMap*)h)->pose, T, 16 * sizeof(double)); }
void orc_resident_get_pose(void* h, double* T) { std::memcpy(T, ((ResidentMap*)h)->pose, 16 * sizeof(double)); }
void orc_resident_fuse(void* h, const float* pts, int64_t n, const double* T) { ((ResidentMap*)h)->fuse(pts, n, T); }

// canonical (key-ordered) export: keys [n,3] int32, centroids [n,3] float32, counts [n] int32
void orc_resident_export(void* h, int32_t* out_keys, float* out_cent, int32_t* out_counts) {
    ResidentMap* m = (ResidentMap*)h;
    std::vector<int32_t> order(m->keys.size());
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return m->keys[a] < m->keys[b]; });
    for (size_t j = 0; j < order.size(); ++j) {
        const size_t s = (size_t)order[j];
        for (int d = 0; d < 3; ++d) {
            if (out_keys) out_keys[3 * j + d] = m->keys[s][d];
            if (out_cent) out_cent[3 * j + d] = (float)m->cent[3 * s + d];
        }
        if (out_counts) out_counts[j] = (int32_t)m->acc[4 * s + 3];
    }
}

// One scan (float32 [n,3], sensor frame): first scan -> fuse only (returns 1); otherwise
// down-sample (A.1, Open3D origin) -> ICP against the resident map seeded with the kept
// pose (A.3/A.5) -> fuse the RAW scan with the new pose.  Returns 0 and fills the result.
int orc_resident_process(void* h, const float* pts, int64_t n, double* T_out, double* fitness_out,
                         double* rmse_out, int32_t* iters_out) {
    ResidentMap* m = (ResidentMap*)h;
    if (m->keys.empty()) {
        m->fuse(pts, n, m->pose);
        return 1;
    }
    std::vector<double> src(3 * n);
    for (int64_t i = 0; i < 3 * n; ++i) src[i] = (double)pts[i];
    int64_t ns = n;
    if (m->ds_voxel > 0) {
        std::vector<double> ds(3 * n);
        ns = orc_voxel_downsample(src.data(), n, m->ds_voxel, nullptr, ds.data(), nullptr, nullptr);
        ds.resize(3 * ns);
        for (double& v : ds) v = (double)(float)v;  // float32 storage of the down-sampled cloud
        src.swap(ds);
    }
    double T[16];
    std::memcpy(T, m->pose, sizeof(T));
    std::vector<double> pcd(src);
    if (!mat4_is_identity(T)) orc_transform(src.data(), ns, T, pcd.data());
    const int r = (int)std::ceil(m->max_corr / m->voxel - 1e-12);
    const double r2 = m->max_corr * m->max_corr;
    std::vector<int32_t> idx(ns), cs, ct;
    std::vector<double> d2(ns);
    struct Res {
        double fitness = 0, rmse = 0;
        int64_t n = 0;
    };
    auto eval = [&]() {
        Res res;
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < ns; ++i) {
            const V3 q = ld(pcd.data(), i);
            int32_t c[3];
            orc_voxel_keys(&pcd[3 * i], 1, m->voxel, m->origin, c);
            double best = std::numeric_limits<double>::infinity();
            int32_t bj = -1;
            for (int dx = -r; dx <= r; ++dx)          // lexicographic key order: the first minimum wins ties
                for (int dy = -r; dy <= r; ++dy)
                    for (int dz = -r; dz <= r; ++dz) {
                        const int32_t s = m->find(c[0] + dx, c[1] + dy, c[2] + dz);
                        if (s < 0) continue;
                        const double d = KDTree::dist2(&m->cent[3 * (size_t)s], q);
                        if (d < best) {
                            best = d;
                            bj = s;
                        }
                    }
            const bool ok = bj >= 0 && best < r2;  // strict, A.4
            idx[i] = ok ? bj : -1;
            d2[i] = ok ? best : 0.0;
        }
        cs.clear();
        ct.clear();
        double err2 = 0.0;
        for (int64_t i = 0; i < ns; ++i)
            if (idx[i] >= 0) {
                cs.push_back((int32_t)i);
                ct.push_back(idx[i]);
                err2 += d2[i];
            }
        res.n = (int64_t)cs.size();
        if (res.n > 0) {
            res.fitness = (double)res.n / (double)ns;
            res.rmse = std::sqrt(err2 / (double)res.n);
        }
        return res;
    };
    Res result = eval();
    int it = 0;
    for (; it < m->max_iteration; ++it) {
        double upd[16];
        umeyama_from_corr(pcd.data(), m->cent.data(), cs.data(), ct.data(), result.n, upd);
        mat4_mul(upd, T, T);
        orc_transform(pcd.data(), ns, upd, pcd.data());
        Res backup = result;
        result = eval();
        if (std::fabs(backup.fitness - result.fitness) < m->rel_fitness &&
            std::fabs(backup.rmse - result.rmse) < m->rel_rmse) {
            ++it;
            break;
        }
    }
    std::memcpy(m->pose, T, sizeof(T));
    if (T_out) std::memcpy(T_out, T, sizeof(T));
    if (fitness_out) *fitness_out = result.fitness;
    if (rmse_out) *rmse_out = result.rmse;
    if (iters_out) *iters_out = it;
    m->fuse(pts, n, T);
    return 0;
}

}  // extern "C"
