// b2_icp.cu — the ICP iteration: correspondence search fused with the
// normal-equation reduction and the on-device SE(3) solve.
//
// Replaces open3d::pipelines::registration::RegistrationICP (SURVEY.md A.3-A.6;
// reference call site icp_processor.py:103-113).  One launch of
// icp_iter_kernel == one "Eval + ComputeTransformation" step of that loop:
//
//   per query (one warp each, 27 lanes probe the 27 stencil cells in parallel):
//     P  = T * p                       (fp64, un-fused, fixed order — A.7)
//     j* = argmin_j |P - q_j|^2 over the stencil, strict d^2 < r^2, ties -> lowest
//          index                       (KDTreeFlann::SearchHybrid(P, r, 1), A.4)
//     lane k accumulates term k of the point-to-point (17 sums: Sp, Sq, Sqp^T, Sd2,
//     n) or point-to-plane (29 sums: upper JtJ, Jtr, Sd2, n) system in fp64
//   per CTA: fixed-order reduction over warps -> one 32-double partial
//   last CTA of the pair (atomic ticket): fixed-order reduction over CTAs, then
//     thread 0 evaluates fitness/rmse, the convergence test of A.3, the Umeyama
//     rotation (one-sided Jacobi SVD, A.5) or the 6x6 solve + Rz*Ry*Rx (A.6), and
//     composes T <- dT * T in place.  Nothing returns to the host inside the loop;
//     pairs that have converged turn later launches into no-ops.
//
// Algorithmic bytes of one launch: 16*Ns (queries) +
// 16*sum_q cand(q) (candidate points) + 16*27*Ns (cell records) + 256*CTAs.
#include <climits>
#include <cstdlib>

#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 640;   // 20 warps, one CTA per SM (<= 96 registers per thread)
constexpr int kCtasPerSm = 1;
constexpr int kWarps = kThreads / 32;
constexpr int kTerms = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------
// small dense solvers (single thread, fp64)
// ---------------------------------------------------------------------------

// Rotation maximising tr(R * C^T) for C = sum (q - mq)(p - mp)^T / n: R = U diag(1,1,det(U)det(V)) V^T.
// One-sided (Hestenes) Jacobi: columns of A = C*V are driven orthogonal; U's third column is never
// taken from data (it is u1 x u2), which keeps R a rotation when the correspondences are coplanar.
__device__ __noinline__ void rotation_from_covariance(const double C[9], double R[9]) {
    double A[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        A[i] = C[i];
        V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0;
            const int q = pq == 0 ? 1 : 2;
            double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                alpha += A[3 * k + p] * A[3 * k + p];
                beta += A[3 * k + q] * A[3 * k + q];
                gamma += A[3 * k + p] * A[3 * k + q];
            }
            // 8 ulp: below that the columns are orthogonal to working precision (a tighter test
            // never settles — rounding noise keeps it rotating for all 60 sweeps)
            if (gamma != 0.0 && gamma * gamma > 3e-30 * (alpha * beta)) {
                rotated = true;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double ap = A[3 * k + p], aq = A[3 * k + q];
                    A[3 * k + p] = cs * ap - sn * aq;
                    A[3 * k + q] = sn * ap + cs * aq;
                    double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = cs * vp - sn * vq;
                    V[3 * k + q] = sn * vp + cs * vq;
                }
            }
        }
        if (!rotated) break;
    }
    double s[3];
    int ord[3] = {0, 1, 2};
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = sqrt(A[i] * A[i] + A[3 + i] * A[3 + i] + A[6 + i] * A[6 + i]);
    if (s[ord[0]] < s[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    if (s[ord[1]] < s[ord[2]]) { int t = ord[1]; ord[1] = ord[2]; ord[2] = t; }
    if (s[ord[0]] < s[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    const double s0 = s[ord[0]], s1 = s[ord[1]];
    if (!(s0 > 0.0)) return;
    double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u1[k] = A[3 * k + ord[0]] / s0;
        v1[k] = V[3 * k + ord[0]];
        v2[k] = V[3 * k + ord[1]];
        v3[k] = V[3 * k + ord[2]];
    }
    if (s1 > 1e-14 * s0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) u2[k] = A[3 * k + ord[1]] / s1;
    } else {  // rank one: any unit vector orthogonal to u1
        int m = fabs(u1[0]) <= fabs(u1[1]) ? (fabs(u1[0]) <= fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
        u2[0] = u2[1] = u2[2] = 0.0;
        u2[m] = 1.0;
    }
    double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    double nn = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u2[k] -= d * u1[k];
        nn += u2[k] * u2[k];
    }
    nn = 1.0 / sqrt(nn);
#pragma unroll
    for (int k = 0; k < 3; ++k) u2[k] *= nn;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    const double detV = v1[0] * (v2[1] * v3[2] - v2[2] * v3[1]) - v2[0] * (v1[1] * v3[2] - v1[2] * v3[1]) +
                        v3[0] * (v1[1] * v2[2] - v1[2] * v2[1]);
    const double sg = detV < 0 ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cidx = 0; cidx < 3; ++cidx)
            R[3 * r + cidx] = u1[r] * v1[cidx] + u2[r] * v2[cidx] + sg * u3[r] * v3[cidx];
}

// Gaussian elimination with partial pivoting on the symmetric 6x6 normal equations.
__device__ __noinline__ bool solve6(const double* JtJ_upper /*21*/, const double* rhs /*6*/, double x[6]) {
    double M[6][7];
    int t = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) {
            M[i][j] = JtJ_upper[t];
            M[j][i] = JtJ_upper[t];
            ++t;
        }
    for (int i = 0; i < 6; ++i) M[i][6] = rhs[i];
    for (int col = 0; col < 6; ++col) {
        int piv = col;
        for (int r = col + 1; r < 6; ++r)
            if (fabs(M[r][col]) > fabs(M[piv][col])) piv = r;
        if (!(fabs(M[piv][col]) > 0.0) || !isfinite(M[piv][col])) return false;
        if (piv != col)
            for (int k = 0; k < 7; ++k) {
                double tmp = M[col][k];
                M[col][k] = M[piv][k];
                M[piv][k] = tmp;
            }
        for (int r = col + 1; r < 6; ++r) {
            double f = M[r][col] / M[col][col];
            for (int k = col; k < 7; ++k) M[r][k] -= f * M[col][k];
        }
    }
    for (int i = 5; i >= 0; --i) {
        double v = M[i][6];
        for (int k = i + 1; k < 6; ++k) v -= M[i][k] * x[k];
        x[i] = v / M[i][i];
        if (!isfinite(x[i])) return false;
    }
    return true;
}

// Polar factor of a 3x3 with positive determinant by Frobenius-scaled Newton iteration,
// X <- (g X + (g X)^-T) / 2: when det(C) > 0 the orthogonal polar factor IS U V^T, the Umeyama
// rotation.  Everything stays in registers; the scale factor only steers convergence, so it is
// evaluated in float32.  Returns false (caller falls back to the Jacobi SVD) for det <= 0 — the
// reflection / rank-deficient cases — or if the iteration does not settle.
__device__ __forceinline__ bool polar_rotation(double a, double b, double c, double d, double e, double f, double g,
                                               double h, double i, double* R) {
    {
        const double nrm = sqrt(a * a + b * b + c * c + d * d + e * e + f * f + g * g + h * h + i * i);
        if (!(nrm > 0.0) || !isfinite(nrm)) return false;
        const double s = 1.7320508075688772 / nrm;
        a *= s; b *= s; c *= s; d *= s; e *= s; f *= s; g *= s; h *= s; i *= s;
    }
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < 40; ++it) {
        const double c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
        const double c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
        const double c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
        const double det = a * c00 + b * c01 + c * c02;
        if (!(det > 1e-12)) return false;  // (scaled so that |X|_F = sqrt(3): an orthogonal X has det 1)
        const double nx2 = a * a + b * b + c * c + d * d + e * e + f * f + g * g + h * h + i * i;
        const double nc2 = c00 * c00 + c01 * c01 + c02 * c02 + c10 * c10 + c11 * c11 + c12 * c12 + c20 * c20 +
                           c21 * c21 + c22 * c22;
        // gamma = (|X^-T|_F / |X|_F)^(1/2), |X^-T|_F = sqrt(nc2)/det
        const float gam = sqrtf(sqrtf((float)(nc2 / (det * det * nx2))));
        const double ga = 0.5 * (double)gam, gb = 0.5 / ((double)gam * det);
        const double na = ga * a + gb * c00, nb = ga * b + gb * c01, nc = ga * c + gb * c02;
        const double nd = ga * d + gb * c10, ne = ga * e + gb * c11, nf = ga * f + gb * c12;
        const double ng = ga * g + gb * c20, nh = ga * h + gb * c21, ni = ga * i + gb * c22;
        const double diff = (na - a) * (na - a) + (nb - b) * (nb - b) + (nc - c) * (nc - c) + (nd - d) * (nd - d) +
                            (ne - e) * (ne - e) + (nf - f) * (nf - f) + (ng - g) * (ng - g) + (nh - h) * (nh - h) +
                            (ni - i) * (ni - i);
        a = na; b = nb; c = nc; d = nd; e = ne; f = nf; g = ng; h = nh; i = ni;
        if (diff < 1e-22) {  // quadratic convergence: the next update is below 1e-20
            ok = true;
            break;
        }
    }
    if (!ok) return false;
    {  // one unscaled polishing step
        const double c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
        const double c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
        const double c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
        const double hd = 0.5 / (a * c00 + b * c01 + c * c02);
        R[0] = 0.5 * a + hd * c00; R[1] = 0.5 * b + hd * c01; R[2] = 0.5 * c + hd * c02;
        R[3] = 0.5 * d + hd * c10; R[4] = 0.5 * e + hd * c11; R[5] = 0.5 * f + hd * c12;
        R[6] = 0.5 * g + hd * c20; R[7] = 0.5 * h + hd * c21; R[8] = 0.5 * i + hd * c22;
    }
    return true;
}

// Evaluate one A.3 loop step from the reduced sums.  Runs on one thread; written so that every
// small matrix stays in registers (a single GPU thread retires roughly one dependent instruction
// per 6-10 cycles, so the instruction count of this tail is what the iteration latency is made of).
template <int MODE>
__device__ __noinline__ void icp_step(const double* S, int ns, const IcpParams prm, const IcpState& cur,
                                      IcpState* st, IcpResultDev* res, int* ndone) {
    const double n = MODE == 0 ? S[16] : S[28];
    const double sd2 = MODE == 0 ? S[15] : S[27];
    const double fitness = ns > 0 ? n / (double)ns : 0.0;
    const double rmse = n > 0 ? sqrt(sd2 / n) : 0.0;
    bool done = false;
    if (cur.has_prev && fabs(cur.prev_fitness - fitness) < prm.rel_fitness &&
        fabs(cur.prev_rmse - rmse) < prm.rel_rmse)
        done = true;
    if (cur.iter >= prm.max_iter) done = true;
    st->fitness = fitness;
    st->rmse = rmse;
    st->n_corr = (int)n;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = cur.T[k];
    if (done) {
        st->done = 1;
#pragma unroll
        for (int k = 0; k < 12; ++k) res->T[k] = T[k];
        res->T[12] = 0.0; res->T[13] = 0.0; res->T[14] = 0.0; res->T[15] = 1.0;
        res->fitness = fitness;
        res->rmse = rmse;
        res->iterations = cur.iter;
        res->n_corr = (int)n;
        atomicAdd(ndone, 1);
        return;
    }
    double U[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // update, 3x4
    if (n > 0) {
        if (MODE == 0) {
            const double inv = 1.0 / n;
            double mp0 = S[0] * inv, mp1 = S[1] * inv, mp2 = S[2] * inv;
            double mq0 = S[3] * inv, mq1 = S[4] * inv, mq2 = S[5] * inv;
            double C[9], R[9];
            C[0] = S[6] * inv - mq0 * mp0;  C[1] = S[7] * inv - mq0 * mp1;  C[2] = S[8] * inv - mq0 * mp2;
            C[3] = S[9] * inv - mq1 * mp0;  C[4] = S[10] * inv - mq1 * mp1; C[5] = S[11] * inv - mq1 * mp2;
            C[6] = S[12] * inv - mq2 * mp0; C[7] = S[13] * inv - mq2 * mp1; C[8] = S[14] * inv - mq2 * mp2;
            if (!polar_rotation(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8], R))
                rotation_from_covariance(C, R);
            const double p0 = cur.pivot[0], p1 = cur.pivot[1], p2 = cur.pivot[2];
            mp0 += p0; mp1 += p1; mp2 += p2;
            mq0 += p0; mq1 += p1; mq2 += p2;
            U[0] = R[0]; U[1] = R[1]; U[2] = R[2];  U[3] = mq0 - (R[0] * mp0 + R[1] * mp1 + R[2] * mp2);
            U[4] = R[3]; U[5] = R[4]; U[6] = R[5];  U[7] = mq1 - (R[3] * mp0 + R[4] * mp1 + R[5] * mp2);
            U[8] = R[6]; U[9] = R[7]; U[10] = R[8]; U[11] = mq2 - (R[6] * mp0 + R[7] * mp1 + R[8] * mp2);
            st->pivot[0] = mq0; st->pivot[1] = mq1; st->pivot[2] = mq2;
        } else {
            double rhs[6], x[6];
            for (int k = 0; k < 6; ++k) rhs[k] = -S[21 + k];
            if (solve6(S, rhs, x)) {
                const double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]);
                const double cg = cos(x[2]), sg = sin(x[2]);
                U[0] = cg * cb; U[1] = cg * sb * sa - sg * ca; U[2] = cg * sb * ca + sg * sa; U[3] = x[3];
                U[4] = sg * cb; U[5] = sg * sb * sa + cg * ca; U[6] = sg * sb * ca - cg * sa; U[7] = x[4];
                U[8] = -sb;     U[9] = cb * sa;                U[10] = cb * ca;               U[11] = x[5];
            }
        }
    }
    // T <- U * T for affine 3x4 blocks (the last row stays 0,0,0,1)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double u0 = U[4 * r], u1 = U[4 * r + 1], u2 = U[4 * r + 2], u3 = U[4 * r + 3];
        st->T[4 * r + 0] = u0 * T[0] + u1 * T[4] + u2 * T[8];
        st->T[4 * r + 1] = u0 * T[1] + u1 * T[5] + u2 * T[9];
        st->T[4 * r + 2] = u0 * T[2] + u1 * T[6] + u2 * T[10];
        st->T[4 * r + 3] = u0 * T[3] + u1 * T[7] + u2 * T[11] + u3;
    }
    st->iter = cur.iter + 1;
    st->prev_fitness = fitness;
    st->prev_rmse = rmse;
    st->has_prev = 1;
}

// ---------------------------------------------------------------------------
// the hot kernels
// ---------------------------------------------------------------------------
// Work decomposition: a warp serves EIGHT queries at a time, four lanes each, so that a 21.7k-point
// source is one round over the 148 x 160 groups of the grid (one 640-thread CTA per SM).  Lane g of a
// group probes stencil cells {g, g+4, ..., g+24 (<27)} in batches of four hash probes that are issued
// back to back (memory-level parallelism 4 per lane, 128 per warp) before any of them is consumed,
// and the group's arg-min is a 2-step shuffle butterfly.  The group leader (lane g == 0) owns the query:
// it loads and transforms the point, and adds the query's contribution to its private column of
// the shared-memory accumulators, so no per-query warp-wide reduction is needed.
//
// Two drivers share that body:
//   icp_iter_kernel     one launch per A.3 step, gridDim.y = pairs (batches; last CTA of a pair solves)
//   icp_persist_kernel  one cooperative launch per registration of a single pair: the CTAs stay
//                       resident, keep their query points in registers, and meet once per step at a
//                       ticket barrier whose last arriver reduces, solves and publishes T — no
//                       kernel boundary, no relaunch, immediate exit on convergence.
#ifdef B2_TAIL_TIMING
__device__ unsigned long long g_tail_ts[8 + 2 * 1024];
__device__ unsigned long long g_pts[300 * 8];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define B2_TS(slot) do { if (threadIdx.x == 0) g_tail_ts[slot] = gtime(); } while (0)
#define B2_TS_BLOCK(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_tail_ts[8 + 2 * blockIdx.x + (k)] = gtime(); } while (0)
#else
#define B2_TS(slot) do {} while (0)
#define B2_TS_BLOCK(k) do {} while (0)
#endif
constexpr int kGroup = 4;
constexpr int kCellsPerLane = (27 + kGroup - 1) / kGroup;  // 7
constexpr int kProbeBatch = 4;  // probes of a lane in flight together
constexpr int kLeaders = kThreads / kGroup;  // accumulator columns per CTA

struct SearchArgs {
    const CellEntry* entries;
    const VoxEntry* vox;
    const float4* gpts;
    const float4* tgt_normals;
    unsigned mask;
    int tgt_base;
    unsigned pair;
    double r2;
    int* corr_out;
    double* d2_out;
    int search_only;
    const unsigned char* walk;  // sweep kernels: walk order table in shared memory
};

// One round: the warp's four groups each register one query (p = the leader's source point, valid
// per group) against the grid and add its normal-equation terms to column `col` of sAccL.
template <int MODE, int KIND>
__device__ __forceinline__ void search_round(const SearchArgs& a, const double* __restrict__ sT,
                                             const double* __restrict__ sPivot, const GridMeta& meta, float4 p,
                                             bool valid, int q, double (*sAccL)[kLeaders], int col) {
    const int lane = threadIdx.x & 31;
    const int g = lane & (kGroup - 1);
    const int lead = lane & ~(kGroup - 1);
    // ---- leader: transform the query, find its cell ---------------------------------------------
    double Px = 0.0, Py = 0.0, Pz = 0.0;
    int cx = 0, cy = 0, cz = 0;
    if (g == 0 && valid) {
        const D3 P = xform_strict(sT, (double)p.x, (double)p.y, (double)p.z);
        Px = P.x; Py = P.y; Pz = P.z;
        cx = voxel_coord(Px, meta.origin[0], meta.cell);
        cy = voxel_coord(Py, meta.origin[1], meta.cell);
        cz = voxel_coord(Pz, meta.origin[2], meta.cell);
        if (KIND == 1 && meta.brick > 1) {
            cx = floor_div(cx, meta.brick);
            cy = floor_div(cy, meta.brick);
            cz = floor_div(cz, meta.brick);
        }
        // map grids: one range test per query instead of one per probe (a query outside the 2^21-cell
        // key space cannot have a neighbour); poisoning cx makes every lane of the group skip the probes
        if (KIND) {
            const int lim = (1 << 20) - 2;
            if (cx < -lim || cx > lim || cy < -lim || cy > lim || cz < -lim || cz > lim) cx = INT_MIN;
        }
    }
    Px = __shfl_sync(kFull, Px, lead);
    Py = __shfl_sync(kFull, Py, lead);
    Pz = __shfl_sync(kFull, Pz, lead);
    cx = __shfl_sync(kFull, cx, lead);
    cy = __shfl_sync(kFull, cy, lead);
    cz = __shfl_sync(kFull, cz, lead);

    // ---- the group's 27 stencil cells: 4 (3 for g >= 3) per lane, probes issued together ----------
    double best = INFINITY;
    int best_idx = 0x7fffffff;
    float bx = 0.f, by = 0.f, bz = 0.f;
    // pair / slab grids: float32 screening state of this lane and the cells it resolved
    const float Pxf = (float)Px, Pyf = (float)Py, Pzf = (float)Pz;
    float bf = INFINITY, sf = INFINITY;
    unsigned bpos = 0;
    unsigned coff[kCellsPerLane + kProbeBatch];
    int ccnt[kCellsPerLane + kProbeBatch];
#pragma unroll
    for (int k = 0; k < kCellsPerLane + kProbeBatch; ++k) {
        coff[k] = 0;
        ccnt[k] = 0;
    }
    if (valid && cx != INT_MIN) {
#pragma unroll
        for (int k0 = 0; k0 < kCellsPerLane; k0 += kProbeBatch) {
            unsigned long long key[kProbeBatch];
            unsigned hpos[kProbeBatch];
            bool live[kProbeBatch];
#pragma unroll
            for (int k = 0; k < kProbeBatch; ++k) {
                const int c = g + kGroup * (k0 + k);
                const int c3 = (c * 171) >> 9;  // c / 3 for c < 128
                const int c9 = (c * 57) >> 9;   // c / 9
                const int nx = cx + c9 - 1, ny = cy + (c3 - 3 * c9) - 1, nz = cz + (c - 3 * c3) - 1;
                const bool in_stencil = (k0 + k) < kCellsPerLane && c < 27;
                if (KIND) {
                    live[k] = in_stencil;
                    key[k] = map_cell_key(nx, ny, nz);
                } else {
                    live[k] = in_stencil && nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536;
                    key[k] = pair_cell_key(a.pair, nx, ny, nz);
                }
                hpos[k] = cell_hash(key[k]) & a.mask;
            }
            if (KIND == 2) {
                unsigned long long ek[kProbeBatch], exy[kProbeBatch], ezn[kProbeBatch], epad[kProbeBatch];
#pragma unroll
                for (int k = 0; k < kProbeBatch; ++k)
                    if (live[k]) ldg256(a.vox + hpos[k], ek[k], exy[k], ezn[k], epad[k]);
#pragma unroll
                for (int k = 0; k < kProbeBatch; ++k) {
                    if (!live[k]) continue;
                    unsigned h = hpos[k];
                    while (ek[k] != key[k] && ek[k] != kEmptyKey) {  // collision chain (rare at load <= 0.5)
                        h = (h + 1) & a.mask;
                        ldg256(a.vox + h, ek[k], exy[k], ezn[k], epad[k]);
                    }
                    if (ek[k] == key[k] && (unsigned)(ezn[k] >> 32) != 0u) {
                        const float x = __uint_as_float((unsigned)exy[k]);
                        const float y = __uint_as_float((unsigned)(exy[k] >> 32));
                        const float z = __uint_as_float((unsigned)ezn[k]);
                        const double d2 = dist2_strict(Px, Py, Pz, (double)x, (double)y, (double)z);
                        if (d2 < best || (d2 == best && (int)h < best_idx)) {
                            best = d2;
                            best_idx = (int)h;
                            bx = x; by = y; bz = z;
                        }
                    }
                }
            } else {
                uint4 raw[kProbeBatch];
#pragma unroll
                for (int k = 0; k < kProbeBatch; ++k)
                    if (live[k]) raw[k] = __ldg(reinterpret_cast<const uint4*>(a.entries + hpos[k]));
#pragma unroll
                for (int k = 0; k < kProbeBatch; ++k) {
                    if (!live[k]) continue;
                    unsigned h = hpos[k];
                    unsigned long long kk = ((unsigned long long)raw[k].y << 32) | raw[k].x;
                    while (kk != key[k] && kk != kEmptyKey) {
                        h = (h + 1) & a.mask;
                        raw[k] = __ldg(reinterpret_cast<const uint4*>(a.entries + h));
                        kk = ((unsigned long long)raw[k].y << 32) | raw[k].x;
                    }
                    if (kk != key[k]) continue;
                    const unsigned off = raw[k].z;
                    const int cnt = (int)raw[k].w;
                    coff[k0 + k] = off;
                    ccnt[k0 + k] = cnt;
                    // float32 screening pass over the cell's points (uniform, no fp64): keep this lane's
                    // best and second-best float32 squared distance
#pragma unroll 4
                    for (int j = 0; j < cnt; ++j) {
                        const float4 t = __ldg(a.gpts + off + j);
                        const float fx = Pxf - t.x, fy = Pyf - t.y, fz = Pzf - t.z;
                        const float d = fx * fx + fy * fy + fz * fz;
                        if (d < bf) {
                            sf = bf;
                            bf = d;
                            bpos = off + j;
                        } else if (d < sf) {
                            sf = d;
                        }
                    }
                }
            }
        }
    }
    if (KIND != 2) {
        // Exact phase.  Let M bound |d2_f32 - d2| for every candidate (rounding of P to float32 plus the
        // float32 arithmetic): every candidate's d2_f32 >= d2* - M, where d2* is the true minimum, so a
        // candidate with d2_f32 > gmin + 2M has d2 > d2* and cannot be the nearest neighbour.  Only the
        // contenders (normally one per group) are evaluated in fp64, in the strict order of the oracle.
        float gmin = bf;
#pragma unroll
        for (int o = 1; o < kGroup; o <<= 1) gmin = fminf(gmin, __shfl_xor_sync(kFull, gmin, o));
        if (valid && gmin < INFINITY) {
            const float e_abs = (fmaxf(fmaxf(fabsf(Pxf), fabsf(Pyf)), fabsf(Pzf)) + 4.0f * (float)meta.cell *
                                                                                         (float)meta.brick + 1e-3f) * 1.2e-7f;
            const float M = 3.5f * e_abs * sqrtf(gmin) + 3.0f * e_abs * e_abs + 1e-6f * gmin;
            const float thr = (gmin + 2.0f * M) * 1.000002f + 1e-30f;
            if (bf <= thr) {
                if (sf > thr) {  // one contender in this lane
                    const float4 t = __ldg(a.gpts + bpos);
                    best = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                    best_idx = __float_as_int(t.w);
                    bx = t.x; by = t.y; bz = t.z;
                } else {  // several within the margin (near-ties): exact pass over this lane's cells
#pragma unroll
                    for (int k = 0; k < kCellsPerLane; ++k) {
                        for (int j = 0; j < ccnt[k]; ++j) {
                            const float4 t = __ldg(a.gpts + coff[k] + j);
                            const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                            const int idx = __float_as_int(t.w);
                            if (d2 < best || (d2 == best && idx < best_idx)) {
                                best = d2;
                                best_idx = idx;
                                bx = t.x; by = t.y; bz = t.z;
                            }
                        }
                    }
                }
            }
        }
    }
    // ---- arg-min over the lanes of the group: (d2, idx) lexicographic -----------------------------
    double gbest = best;
    int gidx = best_idx;
#pragma unroll
    for (int o = 1; o < kGroup; o <<= 1) {
        const double od = __shfl_xor_sync(kFull, gbest, o);
        const int oi = __shfl_xor_sync(kFull, gidx, o);
        if (od < gbest || (od == gbest && oi < gidx)) {
            gbest = od;
            gidx = oi;
        }
    }
    const bool found = valid && gbest < a.r2;
    const unsigned winners = __ballot_sync(kFull, found && best == gbest && best_idx == gidx);
    const int wl = lead + __ffs((winners >> lead) & ((1u << kGroup) - 1u)) - 1;  // meaningful only when found
    const float Qxf = __shfl_sync(kFull, bx, wl & 31);
    const float Qyf = __shfl_sync(kFull, by, wl & 31);
    const float Qzf = __shfl_sync(kFull, bz, wl & 31);
    if (g != 0 || !valid) return;
    // ---- leader: outputs + contribution of this correspondence --------------------------------------
    if (a.corr_out) a.corr_out[q] = found ? gidx : -1;
    if (a.d2_out) a.d2_out[q] = found ? gbest : 0.0;
    if (a.search_only || !found) return;
    const double Qx = Qxf, Qy = Qyf, Qz = Qzf;
    if (MODE == 0) {
        const double px = Px - sPivot[0], py = Py - sPivot[1], pz = Pz - sPivot[2];
        const double qx = Qx - sPivot[0], qy = Qy - sPivot[1], qz = Qz - sPivot[2];
        sAccL[0][col] += px;       sAccL[1][col] += py;       sAccL[2][col] += pz;
        sAccL[3][col] += qx;       sAccL[4][col] += qy;       sAccL[5][col] += qz;
        sAccL[6][col] += qx * px;  sAccL[7][col] += qx * py;  sAccL[8][col] += qx * pz;
        sAccL[9][col] += qy * px;  sAccL[10][col] += qy * py; sAccL[11][col] += qy * pz;
        sAccL[12][col] += qz * px; sAccL[13][col] += qz * py; sAccL[14][col] += qz * pz;
        sAccL[15][col] += gbest;
        sAccL[16][col] += 1.0;
    } else {
        const float4 nq = __ldg(a.tgt_normals + a.tgt_base + gidx);
        const double nx = nq.x, ny = nq.y, nz = nq.z;
        double J[6];
        J[0] = Py * nz - Pz * ny;
        J[1] = Pz * nx - Px * nz;
        J[2] = Px * ny - Py * nx;
        J[3] = nx; J[4] = ny; J[5] = nz;
        const double r = (Px - Qx) * nx + (Py - Qy) * ny + (Pz - Qz) * nz;
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) sAccL[t++][col] += J[i] * J[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) sAccL[21 + i][col] += J[i] * r;
        sAccL[27][col] += gbest;
        sAccL[28][col] += 1.0;
    }
}

// CTA-wide fixed-order fold of the leader columns into one partial per term (warp w: terms w, w+16).
template <int NT, int COLS>
__device__ __forceinline__ void cta_fold(const double* __restrict__ acc, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int t = w; t < NT; t += (int)(blockDim.x >> 5)) {
        double s = 0.0;
#pragma unroll
        for (int c0 = 0; c0 < COLS; c0 += 32) s += acc[t * COLS + c0 + lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        if (lane == 0) out[t] = s;
    }
}

// Fixed-order sum of the per-CTA partials (all loads of a thread in flight before the first add).
template <int NT>
__device__ __forceinline__ void grid_fold(const double* __restrict__ partial, int n_cta, double (*sRed)[kTerms],
                                          double* __restrict__ sSum) {
    constexpr int kParts = kThreads / 32;
    constexpr int kMaxPer = (kCtasPerSm * kSMs + kParts - 1) / kParts;
    const int j = threadIdx.x & 31, part = threadIdx.x >> 5;
    double s = 0.0;
    if (j < NT && part < kParts) {  // (a sweep CTA may have more warps than kParts: the extra ones idle)
        if (n_cta <= kCtasPerSm * kSMs) {
            double v[kMaxPer];
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) {
                const int b = part + k * kParts;
                v[k] = b < n_cta ? __ldcg(partial + (size_t)b * kTerms + j) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) s += v[k];
        } else {
            for (int b = part; b < n_cta; b += kParts) s += __ldcg(partial + (size_t)b * kTerms + j);
        }
    }
    if (part < kParts) sRed[part][j] = s;
    __syncthreads();
    if (threadIdx.x < kTerms) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kParts; ++k) t += sRed[k][threadIdx.x];
        sSum[threadIdx.x] = t;
    }
    __syncthreads();
}

template <int MODE, int KIND>  // KIND 0: pair grid, 1: resident map with slabs, 2: dense resident map
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
icp_iter_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs, const int* __restrict__ ns_dev,
                int ns_max, const CellEntry* __restrict__ entries, unsigned mask,
                const float4* __restrict__ gpts, const VoxEntry* __restrict__ vox,
                const GridMeta* __restrict__ metas, const float4* __restrict__ tgt_normals,
                const int* __restrict__ tgt_offs, IcpParams prm, IcpState* __restrict__ states,
                double* __restrict__ partial, int* __restrict__ tickets, IcpResultDev* __restrict__ results,
                int* __restrict__ corr_out, int* __restrict__ ndone, double* __restrict__ d2_out,
                int search_only, unsigned long long loop_handle) {
    constexpr int NT = MODE == 0 ? 17 : 29;  // live accumulator terms
    const int pair = blockIdx.y;
    IcpState* st = states + pair;

    __shared__ IcpState sState;  // snapshot of the pair's state: every later read is on-chip
    __shared__ GridMeta sMeta;
    __shared__ double sAccL[NT][kLeaders];
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ int sLast;
    static_assert(sizeof(IcpState) % 8 == 0 && sizeof(GridMeta) % 8 == 0, "copied as 8-byte words");

    B2_TS_BLOCK(0);
    // Programmatic dependent launch: let the next step's grid be scheduled as soon as our CTAs retire
    // (all but the solving CTA leave right after the search), and do everything that does not depend
    // on the previous step — grid metadata, accumulator reset, this group's query point — before
    // waiting for the previous grid's writes (the pair state) to land.
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] =
            reinterpret_cast<const double*>(metas + (KIND ? 0 : pair))[threadIdx.x - 32];
    for (int i = threadIdx.x; i < NT * kLeaders; i += kThreads) (&sAccL[0][0])[i] = 0.0;
    const int lane = threadIdx.x & 31;
    const int col = threadIdx.x / kGroup;
    const int groups_total = gridDim.x * kLeaders;
    const int warp_first = blockIdx.x * kLeaders + col - lane / kGroup;
    float4 p_first = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!src_offs) {  // single pair: prefetch the first query (the buffer holds ns_max rows; validity is decided later)
        const int q = warp_first + lane / kGroup;
        if ((lane & (kGroup - 1)) == 0 && q < ns_max) p_first = __ldg(src + q);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // all remaining prologue loads are independent of each other: one memory round trip, not a chain
    if (threadIdx.x < sizeof(IcpState) / 8)
        reinterpret_cast<double*>(&sState)[threadIdx.x] = reinterpret_cast<const double*>(st)[threadIdx.x];
    int sb, se;
    if (src_offs) {
        sb = src_offs[pair];
        se = src_offs[pair + 1];
    } else {
        sb = 0;
        se = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    }
    __syncthreads();
    if (sState.done) {  // (as the body of a graph WHILE node: make sure the loop ends)
        if (loop_handle && threadIdx.x == 0 && blockIdx.x == 0)
            cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, 0u);
        return;
    }
    const double* sT = sState.T;
    const double* sPivot = sState.pivot;

    SearchArgs a;
    a.entries = entries; a.vox = vox; a.gpts = gpts; a.tgt_normals = tgt_normals; a.mask = mask;
    a.tgt_base = tgt_offs ? tgt_offs[pair] : 0;
    a.pair = (unsigned)pair;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = d2_out; a.search_only = search_only;

    // queries are dealt to groups round-robin over (round, CTA, group)
    bool first = !src_offs;
    for (int q0 = sb + warp_first; q0 < se; q0 += groups_total) {  // warp-uniform trip count
        const int q = q0 + lane / kGroup;
        const bool valid = q < se;
        float4 p = p_first;
        if (!first && (lane & (kGroup - 1)) == 0 && valid) p = __ldg(src + q);
        first = false;
        search_round<MODE, KIND>(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
    }
    if (search_only) return;

    __syncthreads();
    B2_TS_BLOCK(1);
    cta_fold<NT, kLeaders>(&sAccL[0][0], partial + ((size_t)pair * gridDim.x + blockIdx.x) * kTerms);
    __syncthreads();
    if (threadIdx.x == 0) {
        // one gpu-scope fence by the signalling thread: the CTA barrier above made every warp's
        // partial visible to it, and the fence is cumulative
        __threadfence();
        int t = atomicAdd(&tickets[pair], 1);
        sLast = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!sLast) return;
    B2_TS(0);
    grid_fold<NT>(partial + (size_t)pair * gridDim.x * kTerms, (int)gridDim.x, sRed, sSum);
    B2_TS(1);
    if (threadIdx.x == 0) {
        tickets[pair] = 0;
        icp_step<MODE>(sSum, se - sb, prm, sState, st, results + pair, ndone);
        // body of a CUDA-graph WHILE node (single pair): run another step unless the registration is done
        if (loop_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, st->done ? 0u : 1u);
    }
    B2_TS(2);
}

// ---- pair / slab grids: one thread per query ------------------------------------------------------
// A pair or slab grid holds several points per cell (~6 at a 2 cm target under a 5 cm cell), so a query
// meets ~57 candidates under the full stencil and the walk — not memory latency — is the cost.  Here a
// thread owns a query and walks the stencil nearest-first with a bound: its own cell, then the 6 face,
// 12 edge and 8 corner neighbours, skipping every cell whose box (squared distance from the query to
// the box, shrunk by a safety margin) is farther than the best candidate so far or than the radius.
// After convergence ~3 of 27 cells and ~15 of 57 candidates survive.  Candidates are screened in
// float32; only the contenders — those within the float32 error bound of the best — are evaluated in
// fp64 in the oracle's strict order, so the result is the exact nearest neighbour of the full stencil,
// ties included.  Queries of a warp are consecutive in voxel-key order, i.e. spatial neighbours: their
// probes and candidate reads hit the same L1 lines.
// Walk order: 0 the query's cell, 1-6 faces, 7-18 edges, 19-26 corners, as a 6-bit offset code
// (dx+1) | (dy+1) << 2 | (dz+1) << 4.
__host__ __device__ constexpr int offset_code(int dx, int dy, int dz) {
    return (dx + 1) | ((dy + 1) << 2) | ((dz + 1) << 4);
}
__host__ __device__ constexpr int walk_code(int b) {
    if (b == 0) return offset_code(0, 0, 0);
    if (b < 7) {
        const int f = b - 1, s = (f & 1) ? 1 : -1;
        return (f >> 1) == 0 ? offset_code(s, 0, 0) : ((f >> 1) == 1 ? offset_code(0, s, 0) : offset_code(0, 0, s));
    }
    if (b < 19) {
        const int e = b - 7, s1 = (e & 2) ? 1 : -1, s2 = (e & 1) ? 1 : -1;
        return (e >> 2) == 0 ? offset_code(s1, s2, 0) : ((e >> 2) == 1 ? offset_code(s1, 0, s2) : offset_code(0, s1, s2));
    }
    const int c = b - 19;
    return offset_code((c & 4) ? 1 : -1, (c & 2) ? 1 : -1, (c & 1) ? 1 : -1);
}

#ifndef B2_SWEEP_T0
#define B2_SWEEP_T0 1024  // p2p: 64 registers per thread, 32 warps per SM (19 % faster than 640 threads at 96)
#define B2_SWEEP_T1 768   // p2l: 29 accumulator terms per thread limit the CTA through shared memory
#endif
template <int MODE>
struct SweepCfg {  // threads per sweep CTA (one CTA per SM): the register budget per thread is 64 K / threads
    static constexpr int kT = MODE == 0 ? B2_SWEEP_T0 : B2_SWEEP_T1;
    static constexpr int kNT = MODE == 0 ? 17 : 29;
    static constexpr size_t kSmem = sizeof(double) * kNT * kT;  // one accumulator column per thread
};

template <int MODE, int KIND>
__device__ __forceinline__ void search_one(const SearchArgs& a, const double* __restrict__ sT,
                                           const double* __restrict__ sPivot, const GridMeta& meta, float4 p, bool valid,
                                           int q, double* __restrict__ acc, unsigned long long* __restrict__ wbest,
                                           unsigned* __restrict__ wsecond) {
    // (warp-collective: all 32 lanes enter; `valid` marks the lanes that hold a query)
    const int lane = threadIdx.x & 31;
    const D3 P = xform_strict(sT, (double)p.x, (double)p.y, (double)p.z);
    const double Px = P.x, Py = P.y, Pz = P.z;
    int cx = voxel_coord(Px, meta.origin[0], meta.cell);
    int cy = voxel_coord(Py, meta.origin[1], meta.cell);
    int cz = voxel_coord(Pz, meta.origin[2], meta.cell);
    if (KIND == 1 && meta.brick > 1) {
        cx = floor_div(cx, meta.brick);
        cy = floor_div(cy, meta.brick);
        cz = floor_div(cz, meta.brick);
    }
    bool searching = valid;
    if (KIND) {  // a query outside the 2^21-cell key space of a map cannot have a neighbour
        const int lim = (1 << 20) - 2;
        searching = searching && !(cx < -lim || cx > lim || cy < -lim || cy > lim || cz < -lim || cz > lim);
    }
    const double Ed = KIND == 1 ? meta.cell * (double)meta.brick : meta.cell;
    const float E = (float)Ed;
    const float ax = (float)(Px - (meta.origin[0] + (double)cx * Ed));  // offset inside the cell, in [0, E]
    const float ay = (float)(Py - (meta.origin[1] + (double)cy * Ed));
    const float az = (float)(Pz - (meta.origin[2] + (double)cz * Ed));
    const float Pxf = (float)Px, Pyf = (float)Py, Pzf = (float)Pz;
    const float slack = 1e-5f * E;  // >> the float32 rounding of ax and E - ax
    const float e_abs = (fmaxf(fmaxf(fabsf(Pxf), fabsf(Pyf)), fabsf(Pzf)) + 4.0f * E + 1e-3f) * 1.2e-7f;
    auto margin = [&](float d2f) -> float {  // bound of |d2_f32 - d2| for a candidate at float32 distance^2 d2f
        return 3.5f * e_abs * sqrtf(d2f) + 3.0f * e_abs * e_abs + 1e-6f * d2f;
    };
    // keys are additive in the cell offsets while no field overflows: always for map keys (range-checked
    // above), and for pair keys when the query's cell is not on the border of the 16-bit coordinate range
    const bool interior = KIND || (cx >= 1 && cx <= 65534 && cy >= 1 && cy <= 65534 && cz >= 1 && cz <= 65534);
    const unsigned long long kbase = KIND ? map_cell_key(cx - 1, cy - 1, cz - 1) : pair_cell_key(a.pair, cx - 1, cy - 1, cz - 1);
    auto key_delta = [](int code) -> unsigned long long {
        constexpr int sx = KIND ? 42 : 32, sy = KIND ? 21 : 16;
        return ((unsigned long long)(code & 3) << sx) + ((unsigned long long)((code >> 2) & 3) << sy) +
               (unsigned long long)(code >> 4);
    };
    auto key_of = [&](int code, bool& ok) -> unsigned long long {
        if (interior) return kbase + key_delta(code);
        const int nx = cx + (code & 3) - 1, ny = cy + ((code >> 2) & 3) - 1, nz = cz + (code >> 4) - 1;
        ok = nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536;
        return pair_cell_key(a.pair, nx, ny, nz);
    };
    auto probe = [&](unsigned long long key, unsigned& off, int& cnt) -> bool {  // is the cell occupied?
        unsigned h = cell_hash(key) & a.mask;
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(a.entries + h));
        unsigned long long kk = ((unsigned long long)raw.y << 32) | raw.x;
        while (kk != key && kk != kEmptyKey) {  // collision chain
            h = (h + 1) & a.mask;
            raw = __ldg(reinterpret_cast<const uint4*>(a.entries + h));
            kk = ((unsigned long long)raw.y << 32) | raw.x;
        }
        off = raw.z;
        cnt = (int)raw.w;
        return kk == key;
    };

    float bf = INFINITY, sf = INFINITY;  // best / second-best float32 squared distance
    unsigned bpos = 0;
    float ub = (float)a.r2 * 1.00001f;  // nothing at or beyond the radius can be a correspondence
    // float32 screening of the points [off, off + cnt) for a query at (x, y, z): running best two
    auto screen_cell = [&](unsigned off, int cnt, float x, float y, float z, float& b1, float& b2, unsigned& bp) {
        const float4* cp = a.gpts + off;
        for (int j = 0; j < cnt; ++j) {
            const float4 t = __ldg(cp + j);
            const float fx = x - t.x, fy = y - t.y, fz = z - t.z;
            const float d = fx * fx + fy * fy + fz * fz;
            bp = d < b1 ? off + j : bp;
            b2 = fminf(b2, fmaxf(b1, d));  // second best of what has been seen
            b1 = fminf(b1, d);
        }
    };
    unsigned m = 0;  // this query's list: bit b = walk position b can still hold a closer point
    float xm = 0.f, xp = 0.f, ym = 0.f, yp = 0.f, zm = 0.f, zp = 0.f;
    if (searching) {
        {  // the query's own cell first: it usually holds the neighbour and gives the bound for the rest
            bool ok = true;
            const unsigned long long key = key_of(offset_code(0, 0, 0), ok);
            unsigned off;
            int cnt;
            if (ok && probe(key, off, cnt)) {
                screen_cell(off, cnt, Pxf, Pyf, Pzf, bf, sf, bpos);
                if (bf < INFINITY) ub = fminf(ub, (bf + 2.0f * margin(bf)) * 1.00001f);  // true d2 <= bf + M
            }
        }
        // conservative (too small) squared distance from the query to the slab of cells at -1 / +1 per axis
        auto sq = [](float v) { return v * v * 0.9999f; };
        xm = sq(fmaxf(ax - slack, 0.f)); xp = sq(fmaxf(E - ax - slack, 0.f));
        ym = sq(fmaxf(ay - slack, 0.f)); yp = sq(fmaxf(E - ay - slack, 0.f));
        zm = sq(fmaxf(az - slack, 0.f)); zp = sq(fmaxf(E - az - slack, 0.f));
        // (a face neighbour qualifies when its slab does; an edge or corner neighbour only if all its faces do)
        const bool fxm = xm <= ub, fxp = xp <= ub, fym = ym <= ub, fyp = yp <= ub, fzm = zm <= ub, fzp = zp <= ub;
        m = (fxm ? 1u << 1 : 0u) | (fxp ? 1u << 2 : 0u) | (fym ? 1u << 3 : 0u) | (fyp ? 1u << 4 : 0u) |
            (fzm ? 1u << 5 : 0u) | (fzp ? 1u << 6 : 0u);
        if (((fxm || fxp) ? 1 : 0) + ((fym || fyp) ? 1 : 0) + ((fzm || fzp) ? 1 : 0) >= 2) {
#pragma unroll
            for (int b = 7; b < 27; ++b) {
                const int code = walk_code(b), dx = (code & 3) - 1, dy = ((code >> 2) & 3) - 1, dz = (code >> 4) - 1;
                const float lb = (dx < 0 ? xm : (dx > 0 ? xp : 0.f)) + (dy < 0 ? ym : (dy > 0 ? yp : 0.f)) +
                                 (dz < 0 ? zm : (dz > 0 ? zp : 0.f));
                if (lb <= ub) m |= 1u << b;
            }
        }
    }
    // ---- the neighbour cells, pooled over the warp ---------------------------------------------------
    // List lengths are very uneven (most queries need 1-3 neighbours, a query without a close point needs
    // all 26), and a warp that walks per lane waits for its longest list.  So the (query, cell) pairs of the
    // whole warp form one pool of tasks, 32 at a time: a lane takes a task, fetches the owner's key base and
    // coordinates by shuffle, probes the cell, screens its points and merges its best two into the owner's
    // slot in shared memory (an atomic min on (d2, position) keeps the best; whatever loses goes to the
    // second best).  The float32 arithmetic is the owner's, so the result does not depend on who computed it.
    const unsigned mp = interior ? m : 0u;  // (border cells of the 16-bit pair key space: walked by the owner below)
    const int c_own = __popc(mp);
    int pre = c_own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, pre, o);
        if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(kFull, pre, 31);
    pre -= c_own;  // exclusive prefix: first task of this lane
    wbest[lane] = ((unsigned long long)__float_as_uint(bf) << 32) | bpos;
    wsecond[lane] = __float_as_uint(sf);
    __syncwarp();
    for (int g0 = 0; g0 < total; g0 += 32) {  // warp-uniform trip count
        const int g = g0 + lane;
        int o = 0;  // owner: the last lane whose first task is <= g
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const int cand = o + step;
            const int pv = __shfl_sync(kFull, pre, cand & 31);
            if (cand < 32 && pv <= g) o = cand;
        }
        const int opre = __shfl_sync(kFull, pre, o);
        const unsigned om = __shfl_sync(kFull, mp, o);
        const unsigned long long okb = __shfl_sync(kFull, kbase, o);
        const float ox = __shfl_sync(kFull, Pxf, o), oy = __shfl_sync(kFull, Pyf, o), oz = __shfl_sync(kFull, Pzf, o);
        if (g < total) {
            const int b = __fns(om, 0, g - opre + 1);  // the (g - opre)-th cell of the owner's list
            unsigned off;
            int cnt;
            if (probe(okb + key_delta(a.walk[b]), off, cnt)) {
                float l1 = INFINITY, l2 = INFINITY;
                unsigned lp = 0;
                screen_cell(off, cnt, ox, oy, oz, l1, l2, lp);
                if (l1 < INFINITY) {
                    const unsigned long long mine = ((unsigned long long)__float_as_uint(l1) << 32) | lp;
                    const unsigned long long old = atomicMin(wbest + o, mine);
                    // non-negative floats order like their bit patterns
                    const unsigned loser = mine < old ? (unsigned)(old >> 32) : __float_as_uint(l1);
                    atomicMin(wsecond + o, min(loser, __float_as_uint(l2)));
                }
            }
        }
    }
    __syncwarp();
    if (mp) {
        const unsigned long long w = wbest[lane];
        bf = __uint_as_float((unsigned)(w >> 32));
        bpos = (unsigned)w;
        sf = __uint_as_float(wsecond[lane]);
    }
    __syncwarp();  // the slots are reused by the warp's next round
    if (!interior && m) {  // rare: per-lane walk with range-checked keys
        unsigned mm = m;
        while (mm) {
            const int b = __ffs(mm) - 1;
            mm &= mm - 1;
            bool ok = true;
            const unsigned long long key = key_of(a.walk[b], ok);
            unsigned off;
            int cnt;
            if (ok && probe(key, off, cnt)) screen_cell(off, cnt, Pxf, Pyf, Pzf, bf, sf, bpos);
        }
    }
    if (!valid) return;
    // Exact phase.  Let M bound |d2_f32 - d2| for every candidate (rounding of P to float32 plus the float32
    // arithmetic): every candidate's d2_f32 >= d2* - M, where d2* is the true minimum, so a candidate with
    // d2_f32 > bf + 2M has d2 > d2* and cannot be the nearest neighbour.
    double best = INFINITY;
    int best_idx = 0x7fffffff;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (bf < INFINITY) {
        const float thr = (bf + 2.0f * margin(bf)) * 1.000002f + 1e-30f;
        if (sf > thr) {  // a single contender
            const float4 t = __ldg(a.gpts + bpos);
            best = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
            best_idx = __float_as_int(t.w);
            qx = t.x; qy = t.y; qz = t.z;
        } else {  // near-ties: exact (d2, idx) arg-min over the whole stencil
#pragma unroll 1
            for (int b = 0; b < 27; ++b) {
                bool ok = true;
                const unsigned long long key = key_of(a.walk[b], ok);
                if (!ok) continue;
                const CellEntry e = grid_lookup(a.entries, a.mask, key);
                if (e.key != key) continue;
                for (unsigned j = 0; j < e.cnt; ++j) {
                    const float4 t = __ldg(a.gpts + e.off + j);
                    const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                    const int idx = __float_as_int(t.w);
                    if (d2 < best || (d2 == best && idx < best_idx)) {
                        best = d2;
                        best_idx = idx;
                        qx = t.x; qy = t.y; qz = t.z;
                    }
                }
            }
        }
    }
    const bool found = best < a.r2;
    if (a.corr_out) a.corr_out[q] = found ? best_idx : -1;
    if (a.d2_out) a.d2_out[q] = found ? best : 0.0;
    if (a.search_only || !found) return;
    const double Qx = qx, Qy = qy, Qz = qz;
    constexpr int S = SweepCfg<MODE>::kT;  // acc[t * S]: term t of this thread's column
    if (MODE == 0) {
        const double px = Px - sPivot[0], py = Py - sPivot[1], pz = Pz - sPivot[2];
        const double ux = Qx - sPivot[0], uy = Qy - sPivot[1], uz = Qz - sPivot[2];
        acc[0 * S] += px;       acc[1 * S] += py;       acc[2 * S] += pz;
        acc[3 * S] += ux;       acc[4 * S] += uy;       acc[5 * S] += uz;
        acc[6 * S] += ux * px;  acc[7 * S] += ux * py;  acc[8 * S] += ux * pz;
        acc[9 * S] += uy * px;  acc[10 * S] += uy * py; acc[11 * S] += uy * pz;
        acc[12 * S] += uz * px; acc[13 * S] += uz * py; acc[14 * S] += uz * pz;
        acc[15 * S] += best;
        acc[16 * S] += 1.0;
    } else {
        const float4 nq = __ldg(a.tgt_normals + a.tgt_base + best_idx);
        const double nx = nq.x, ny = nq.y, nz = nq.z;
        double J[6];
        J[0] = Py * nz - Pz * ny;
        J[1] = Pz * nx - Px * nz;
        J[2] = Px * ny - Py * nx;
        J[3] = nx; J[4] = ny; J[5] = nz;
        const double r = (Px - Qx) * nx + (Py - Qy) * ny + (Pz - Qz) * nz;
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) acc[(t++) * S] += J[i] * J[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[(21 + i) * S] += J[i] * r;
        acc[27 * S] += best;
        acc[28 * S] += 1.0;
    }
}

template <int MODE, int KIND>  // KIND 0: pair grid, 1: resident map with slabs
__global__ void __launch_bounds__(SweepCfg<MODE>::kT, 1)
icp_sweep_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs, const int* __restrict__ ns_dev,
                 int ns_max, const CellEntry* __restrict__ entries, unsigned mask, const float4* __restrict__ gpts,
                 const GridMeta* __restrict__ metas, const float4* __restrict__ tgt_normals,
                 const int* __restrict__ tgt_offs, IcpParams prm, IcpState* __restrict__ states,
                 double* __restrict__ partial, int* __restrict__ tickets, IcpResultDev* __restrict__ results,
                 int* __restrict__ corr_out, int* __restrict__ ndone, double* __restrict__ d2_out, int search_only) {
    constexpr int NT = SweepCfg<MODE>::kNT, kT = SweepCfg<MODE>::kT;
    const int pair = blockIdx.y;
    IcpState* st = states + pair;
    extern __shared__ double sAcc[];  // [NT][kT]: one accumulator column per thread
    __shared__ unsigned char sWalk[32];
    __shared__ unsigned long long sBest[SweepCfg<MODE>::kT];  // per warp: 32 slots of (best d2 bits, position)
    __shared__ unsigned sSecond[SweepCfg<MODE>::kT];          //           and of the second-best d2 bits
    __shared__ IcpState sState;
    __shared__ GridMeta sMeta;
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ int sLast;

    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] =
            reinterpret_cast<const double*>(metas + (KIND ? 0 : pair))[threadIdx.x - 32];
#pragma unroll
    for (int t = 0; t < NT; ++t) sAcc[t * kT + threadIdx.x] = 0.0;
    if (threadIdx.x < 27) sWalk[threadIdx.x] = (unsigned char)walk_code(threadIdx.x);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < sizeof(IcpState) / 8)
        reinterpret_cast<double*>(&sState)[threadIdx.x] = reinterpret_cast<const double*>(st)[threadIdx.x];
    int sb, se;
    if (src_offs) {
        sb = src_offs[pair];
        se = src_offs[pair + 1];
    } else {
        sb = 0;
        se = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    }
    __syncthreads();
    if (sState.done) return;

    SearchArgs a;
    a.walk = sWalk;
    a.entries = entries; a.vox = nullptr; a.gpts = gpts; a.tgt_normals = tgt_normals; a.mask = mask;
    a.tgt_base = tgt_offs ? tgt_offs[pair] : 0;
    a.pair = (unsigned)pair;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = d2_out; a.search_only = search_only;

    // 32-query blocks are dealt to the warps round-robin over (round, warp, CTA), so that a source smaller
    // than the grid still spreads over every SM
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (kT / 32);
    for (int blk = (threadIdx.x >> 5) * gridDim.x + blockIdx.x; sb + blk * 32 < se; blk += warps_total) {
        const int q = sb + blk * 32 + lane;
        const bool valid = q < se;
        const float4 p = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        search_one<MODE, KIND>(a, sState.T, sState.pivot, sMeta, p, valid, q, sAcc + threadIdx.x,
                               sBest + (threadIdx.x & ~31), sSecond + (threadIdx.x & ~31));
    }
    if (search_only) return;

    __syncthreads();
    cta_fold<NT, kT>(sAcc, partial + ((size_t)pair * gridDim.x + blockIdx.x) * kTerms);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int t = atomicAdd(&tickets[pair], 1);
        sLast = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!sLast) return;
    grid_fold<NT>(partial + (size_t)pair * gridDim.x * kTerms, (int)gridDim.x, sRed, sSum);
    if (threadIdx.x == 0) {
        tickets[pair] = 0;
        icp_step<MODE>(sSum, se - sb, prm, sState, st, results + pair, ndone);
    }
}

// ---- persistent single-pair driver ------------------------------------------------------------------
struct PersistSync {
    unsigned int ticket;  // cumulative CTA arrivals
    unsigned int epoch;   // steps published so far
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kPersistRounds = 2;  // query points a group keeps in registers across steps

template <int MODE, int KIND>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
icp_persist_kernel(const float4* __restrict__ src, const int* __restrict__ ns_dev, int ns_max,
                   const CellEntry* __restrict__ entries, unsigned mask, const float4* __restrict__ gpts,
                   const VoxEntry* __restrict__ vox, const GridMeta* __restrict__ metas,
                   const float4* __restrict__ tgt_normals, IcpParams prm, IcpState* __restrict__ st,
                   double* __restrict__ partial, PersistSync* __restrict__ sync, IcpResultDev* __restrict__ results,
                   int* __restrict__ corr_out, int* __restrict__ ndone) {
    constexpr int NT = MODE == 0 ? 17 : 29;
    __shared__ IcpState sState;  // every CTA keeps a snapshot of the pair's state: the solver never re-reads it
    __shared__ GridMeta sMeta;
    __shared__ double sAccL[NT][kLeaders];
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ int sFlag;  // 1: this CTA was the last arriver of the step
    const double* sT = sState.T;
    const double* sPivot = sState.pivot;

    const int se = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    if (threadIdx.x < sizeof(IcpState) / 8)
        reinterpret_cast<double*>(&sState)[threadIdx.x] = reinterpret_cast<const double*>(st)[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] = reinterpret_cast<const double*>(metas)[threadIdx.x - 32];

    SearchArgs a;
    a.entries = entries; a.vox = vox; a.gpts = gpts; a.tgt_normals = tgt_normals; a.mask = mask;
    a.tgt_base = 0;
    a.pair = 0u;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = nullptr; a.search_only = 0;

    const int lane = threadIdx.x & 31;
    const int col = threadIdx.x / kGroup;
    const bool leader = (lane & (kGroup - 1)) == 0;
    const int groups_total = gridDim.x * kLeaders;
    const int warp_first = blockIdx.x * kLeaders + col - lane / kGroup;
    // the queries of a group do not change between steps: keep the first rounds in registers
    float4 pq[kPersistRounds];
#pragma unroll
    for (int r = 0; r < kPersistRounds; ++r) {
        const int q = warp_first + r * groups_total + lane / kGroup;
        pq[r] = (leader && q < se) ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const unsigned n_cta = gridDim.x;

    for (unsigned step = 0;; ++step) {
#ifdef B2_TAIL_TIMING
#define B2_PTS(k) do { if (step == 3 && threadIdx.x == 0 && blockIdx.x < 300) g_pts[blockIdx.x * 8 + (k)] = gtime(); } while (0)
#else
#define B2_PTS(k) do {} while (0)
#endif
        B2_PTS(0);
        for (int i = threadIdx.x; i < NT * kLeaders; i += kThreads) (&sAccL[0][0])[i] = 0.0;
        __syncthreads();  // sT / sPivot / sMeta of this step and the zeroed accumulators are visible
#pragma unroll
        for (int r = 0; r < kPersistRounds; ++r) {
            const int q0 = warp_first + r * groups_total;
            if (q0 < se) {
                const int q = q0 + lane / kGroup;
                search_round<MODE, KIND>(a, sT, sPivot, sMeta, pq[r], q < se, q, sAccL, col);
            }
        }
        for (int q0 = warp_first + kPersistRounds * groups_total; q0 < se; q0 += groups_total) {
            const int q = q0 + lane / kGroup;
            const bool valid = q < se;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (leader && valid) p = __ldg(src + q);
            search_round<MODE, KIND>(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
        }
        __syncthreads();
        B2_PTS(1);
        cta_fold<NT, kLeaders>(&sAccL[0][0], partial + (size_t)blockIdx.x * kTerms);
        __syncthreads();
        B2_PTS(2);
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned t = atomicAdd(&sync->ticket, 1u);
            sFlag = (t == (step + 1) * n_cta - 1);
        }
        __syncthreads();
        B2_PTS(3);
        if (sFlag) {  // last arriver of this step: the ticket ordered every CTA's partial before this point
            grid_fold<NT>(partial, (int)n_cta, sRed, sSum);
            B2_PTS(4);
            if (threadIdx.x == 0) {
                icp_step<MODE>(sSum, se, prm, sState, st, results, ndone);
                B2_PTS(5);
                __threadfence();
                st_release_u32(&sync->epoch, step + 1);
            }
        } else if (threadIdx.x == 0) {
            while (ld_acquire_u32(&sync->epoch) < step + 1) __nanosleep(20);
        }
        __syncthreads();
        B2_PTS(6);
        // pick up the published state (written by another CTA: read it from L2)
        if (threadIdx.x < sizeof(IcpState) / 8)
            reinterpret_cast<double*>(&sState)[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(st) + threadIdx.x);
        __syncthreads();
        if (sState.done) break;
    }
}

__global__ void icp_init_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs,
                                const int* __restrict__ ns_dev, int ns_max, const double* __restrict__ init,
                                int init_stride, IcpState* __restrict__ states, int* __restrict__ tickets,
                                int* __restrict__ ndone, int n_pairs) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair == 0) *ndone = 0;
    if (pair >= n_pairs) return;
    IcpState s;
    for (int i = 0; i < 16; ++i) s.T[i] = init ? init[(size_t)pair * init_stride + i] : ((i % 5 == 0) ? 1.0 : 0.0);
    s.fitness = s.rmse = s.prev_fitness = s.prev_rmse = 0.0;
    int sb = src_offs ? src_offs[pair] : 0;
    int se = src_offs ? src_offs[pair + 1] : (ns_dev ? min(*ns_dev, ns_max) : ns_max);
    if (se > sb) {
        float4 p = src[sb + (se - sb) / 2];
        D3 P = xform_strict(s.T, (double)p.x, (double)p.y, (double)p.z);
        s.pivot[0] = P.x; s.pivot[1] = P.y; s.pivot[2] = P.z;
    } else {
        s.pivot[0] = s.pivot[1] = s.pivot[2] = 0.0;
    }
    s.iter = 0;
    s.done = 0;
    s.n_corr = 0;
    s.has_prev = 0;
    states[pair] = s;
    tickets[pair] = 0;
}

template <int MODE, int IS_MAP>
void launch_iter(Context* c, dim3 grid, const float4* src, const int* src_offs, const int* ns_dev, int ns_max,
                 const GridView& g, const float4* nrm, const int* tgt_offs, const IcpParams& prm, IcpState* st,
                 double* partial, int* tickets, IcpResultDev* res, int* corr, int* ndone, double* d2_out,
                 int search_only, cudaStream_t stream, unsigned long long loop_handle) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("B2_NO_PDL") != nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    cudaLaunchKernelEx(&cfg, icp_iter_kernel<MODE, IS_MAP>, src, src_offs, ns_dev, ns_max,
                       (const CellEntry*)g.entries, g.mask, (const float4*)g.pts, (const VoxEntry*)g.vox,
                       (const GridMeta*)g.meta, nrm, tgt_offs, prm, st, partial, tickets, res, corr, ndone, d2_out,
                       search_only, loop_handle);
}

template <int MODE, int KIND>
void launch_sweep(Context* c, dim3 grid, const float4* src, const int* src_offs, const int* ns_dev, int ns_max,
                  const GridView& g, const float4* nrm, const int* tgt_offs, const IcpParams& prm, IcpState* st,
                  double* partial, int* tickets, IcpResultDev* res, int* corr, int* ndone, double* d2_out,
                  int search_only) {
    constexpr size_t smem = SweepCfg<MODE>::kSmem;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(SweepCfg<MODE>::kT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("B2_NO_PDL") != nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    cudaLaunchKernelEx(&cfg, icp_sweep_kernel<MODE, KIND>, src, src_offs, ns_dev, ns_max, (const CellEntry*)g.entries,
                       g.mask, (const float4*)g.pts, (const GridMeta*)g.meta, nrm, tgt_offs, prm, st, partial, tickets,
                       res, corr, ndone, d2_out, search_only);
}

// Candidate census of the 27-cell stencil (the data-dependent term of the byte model); kind as in GridView.
__global__ void __launch_bounds__(256)
stencil_count_kernel(const float4* __restrict__ src, int ns, const double* __restrict__ T,
                     const CellEntry* __restrict__ entries, const VoxEntry* __restrict__ vox, unsigned mask,
                     const GridMeta* __restrict__ meta, int kind, unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (q >= ns) return;
    const float4 p = __ldg(src + q);
    D3 P{(double)p.x, (double)p.y, (double)p.z};
    if (T) P = xform_strict(T, P.x, P.y, P.z);
    const int brick = meta->brick;
    int cx = voxel_coord(P.x, meta->origin[0], meta->cell);
    int cy = voxel_coord(P.y, meta->origin[1], meta->cell);
    int cz = voxel_coord(P.z, meta->origin[2], meta->cell);
    if (kind == 1 && brick > 1) {
        cx = floor_div(cx, brick);
        cy = floor_div(cy, brick);
        cz = floor_div(cz, brick);
    }
    unsigned cnt = 0;
    if (lane < 27) {
        const int nx = cx + lane / 9 - 1, ny = cy + (lane / 3) % 3 - 1, nz = cz + lane % 3 - 1;
        if (kind == 0) {
            if (nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536) {
                const unsigned long long key = pair_cell_key(0u, nx, ny, nz);
                const CellEntry e = grid_lookup(entries, mask, key);
                if (e.key == key) cnt = e.cnt;
            }
        } else {
            const unsigned long long key = map_cell_key(nx, ny, nz);
            if (kind == 2) {
                float x, y, z;
                unsigned slot;
                cnt = vox_lookup(vox, mask, key, x, y, z, slot) ? 1u : 0u;
            } else {
                const CellEntry e = grid_lookup(entries, mask, key);
                if (e.key == key) cnt = e.cnt;
            }
        }
    }
    cnt = __reduce_add_sync(kFull, cnt);
    if (lane == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

}  // namespace

int map_stencil_count(Context* c, const float4* src, int ns, const double* T_dev, const GridView& g,
                      unsigned long long* total_dev) {
    stencil_count_kernel<<<div_up((long long)ns * 32, 256), 256, 0, c->stream>>>(src, ns, T_dev, g.entries, g.vox,
                                                                                  g.mask, g.meta, g.is_map, total_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

namespace {

template <int MODE, int KIND>
int launch_persist(Context* c, const float4* src, const int* ns_dev, int ns_max, const GridView& g,
                   const float4* nrm, const IcpParams& prm, IcpState* st, double* partial, PersistSync* psync,
                   IcpResultDev* res, int* corr, int* ndone) {
    static int max_ctas = 0;  // co-resident CTAs of this instantiation (cooperative launch limit)
    if (max_ctas == 0) {
        int per_sm = 0, sms = 0;
        B2_CUDA_OK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_persist_kernel<MODE, KIND>, kThreads, 0));
        B2_CUDA_OK(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        max_ctas = per_sm * sms;
        if (max_ctas < 1) return set_error(c, B2_ERR_CUDA, "persistent ICP kernel cannot be made resident");
    }
    int grid = div_up(ns_max, kLeaders);  // one query per 8-lane group when it fits
    if (grid > max_ctas) grid = max_ctas;
    if (grid > kCtasPerSm * kSMs) grid = kCtasPerSm * kSMs;
    if (grid < 1) grid = 1;
    const CellEntry* entries = g.entries;
    unsigned mask = g.mask;
    const float4* gpts = g.pts;
    const VoxEntry* vox = g.vox;
    const GridMeta* meta = g.meta;
    IcpParams prm_v = prm;
    void* args[] = {(void*)&src, (void*)&ns_dev, (void*)&ns_max, (void*)&entries, (void*)&mask, (void*)&gpts,
                    (void*)&vox, (void*)&meta, (void*)&nrm, (void*)&prm_v, (void*)&st, (void*)&partial,
                    (void*)&psync, (void*)&res, (void*)&corr, (void*)&ndone};
    B2_CUDA_OK(c, cudaLaunchCooperativeKernel((const void*)icp_persist_kernel<MODE, KIND>, dim3(grid), dim3(kThreads),
                                              args, 0, c->stream));
    c->launches++;
    return B2_OK;
}

}  // namespace

// The sweep kernels keep one accumulator column per thread in dynamic shared memory (> 48 KB): opt in once
// per context (device attribute, not a stream operation — done at creation so it never lands in a capture).
int icp_configure(Context* c) {
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<0>::kSmem));
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<0>::kSmem));
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<1>::kSmem));
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<1>::kSmem));
    return B2_OK;
}

constexpr int kChainSteps = 40;  // captured scan step: steps launched back to back before the WHILE node takes over
constexpr long long kSweepMinQueries = 2LL * kCtasPerSm * kSMs * kThreads;  // ~190 k queries in flight

int icp_grid_x(int ns_max, int n_pairs, int kind) {
    const int pairs = n_pairs > 0 ? n_pairs : 1;
    int cap = (kCtasPerSm * kSMs) / pairs;
    if (cap < 2) cap = 2;
    if (pairs == 1) cap = kCtasPerSm * kSMs;
    int gx;
    if (kind == 2) {
        // dense map, single pair: one query per 4-lane group when it fits (160 groups per CTA, one CTA per SM)
        gx = div_up(ns_max / pairs + 1, kLeaders);
    } else if (pairs == 1) {
        gx = div_up(ns_max, 32);  // one thread per query: spread the 32-query blocks over every SM
    } else {
        gx = div_up(ns_max / pairs + 1, 640);  // batches fill the machine through gridDim.y
    }
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    return gx;
}

int icp_run(Context* c, const float4* src, int ns_max, const int* src_offsets, const int* ns_dev, int n_pairs,
            const GridView& grid, const float4* tgt_normals, const int* tgt_offsets, const double* init_dev,
            const IcpParams& prm, IcpResultDev* results_dev, int* corr_out, double* d2_out, int search_only,
            int check_every) {
    if (n_pairs <= 0 || ns_max <= 0) return set_error(c, B2_ERR_INVALID, "icp: empty source");
    if (prm.mode == 1 && !tgt_normals && !search_only)
        return set_error(c, B2_ERR_INVALID, "point-to-plane ICP requires target normals");
    if (!(prm.max_corr > 0.0)) return set_error(c, B2_ERR_INVALID, "max_correspondence_distance must be > 0");
    // Pair / slab grids have two drivers: the 4-lane-group kernel (parallel probes: best when the queries do
    // not fill the machine and latency is the cost) and the thread-per-query sweep (pruned walk: best when
    // they do).  b2_set_icp_driver forces one (parity tests run both; results are identical).
    bool sweep = grid.is_map != 2 && (long long)ns_max >= kSweepMinQueries;
    if (c->icp_driver != 0 && grid.is_map != 2) sweep = c->icp_driver == 2;
    const int gx = icp_grid_x(ns_max, n_pairs, sweep ? grid.is_map : 2);
    B2_WS(c, states, IcpState, WS_ICP_STATE, sizeof(IcpState) * (size_t)n_pairs);
    B2_WS(c, partial, double, WS_ICP_PARTIAL, sizeof(double) * kTerms * (size_t)(gx > kCtasPerSm * kSMs ? gx : kCtasPerSm * kSMs) * n_pairs);
    B2_WS(c, tickets, int, WS_ICP_TICKET, sizeof(int) * ((size_t)n_pairs + 1));
    int* ndone = tickets + n_pairs;
    icp_init_kernel<<<div_up(n_pairs, 128), 128, 0, c->stream>>>(src, src_offsets, ns_dev, ns_max, init_dev, 16,
                                                                 states, tickets, ndone, n_pairs);
    B2_LAUNCH_CHECK(c);
    dim3 g(gx, n_pairs);
    const int launches = search_only ? 1 : prm.max_iter + 1;
    // (opt-in: on B200 the software grid barrier + solve of a step costs more than a kernel boundary)
    static const bool persist = getenv("B2_PERSIST") != nullptr;
    if (persist && n_pairs == 1 && !src_offsets && !search_only && !d2_out && !c->capturing) {
        // single registration: one cooperative launch runs the whole A.3 loop
        B2_WS(c, psync, PersistSync, WS_ICP_SYNC, sizeof(PersistSync));
        B2_CUDA_OK(c, cudaMemsetAsync(psync, 0, sizeof(PersistSync), c->stream));
        int rc = B2_OK;
        if (prm.mode == 0) {
            if (grid.is_map == 2) rc = launch_persist<0, 2>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
            else if (grid.is_map == 1) rc = launch_persist<0, 1>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
            else rc = launch_persist<0, 0>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
        } else {
            if (grid.is_map == 2) rc = launch_persist<1, 2>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
            else if (grid.is_map == 1) rc = launch_persist<1, 1>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
            else rc = launch_persist<1, 0>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
        }
#ifdef B2_TAIL_TIMING
        if (getenv("B2_TIMING")) {
            cudaStreamSynchronize(c->stream);
            static unsigned long long h[300 * 8];
            cudaMemcpyFromSymbol(h, g_pts, sizeof(h));
            int nb = div_up(ns_max, kLeaders); if (nb > 296) nb = 296;
            unsigned long long t0 = ~0ull; for (int b = 0; b < nb; ++b) if (h[b * 8] < t0) t0 = h[b * 8];
            unsigned long long mx[7] = {0}, mn[7]; for (int k = 0; k < 7; ++k) mn[k] = ~0ull;
            for (int b = 0; b < nb; ++b) for (int k = 0; k < 7; ++k) { unsigned long long v = h[b * 8 + k]; if (!v) continue; v -= t0; if (v > mx[k]) mx[k] = v; if (v < mn[k]) mn[k] = v; }
            fprintf(stderr, "[b2 persist step3 ns] start %llu..%llu | main done %llu..%llu | folded %llu..%llu | ticketed %llu..%llu | grid-fold %llu | step %llu | released %llu..%llu\n",
                    mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], mn[3], mx[3], mx[4], mx[5], mn[6], mx[6]);
        }
#endif
        return rc;
    }
    int* host_done = (int*)c->host_mailbox + 512;  // scratch word in the pinned mailbox
    // B2_TIMING=1: bracket every launch with CUDA events and print the per-launch device times
    static const bool timing_env = getenv("B2_TIMING") != nullptr;
    const bool timing = timing_env && !c->capturing;
    std::vector<cudaEvent_t> evs;
    if (timing) {
        evs.resize(launches + 1);
        for (auto& e : evs) cudaEventCreate(&e);
        cudaEventRecord(evs[0], c->stream);
    }
    // Inside a captured scan step with a long iteration budget (the reference default is 500) only the first
    // kChainSteps steps are chained launches (programmatic dependent launch: the cheapest per step); the rest
    // of the A.3 loop is a CUDA-graph WHILE node whose body is ONE step kernel.  The solving CTA clears the
    // loop condition when the registration is done, so a replay runs the steps it needs instead of
    // max_iteration + 1 launches of which most return at once (500: 985 -> ~495 us per scan).
    // B2_ICP_WHILE=0 disables the loop node, =1 uses it for every step.
    static const char* while_env = getenv("B2_ICP_WHILE");
    const int while_mode = while_env ? atoi(while_env) : -1;
    const bool can_while = c->capturing && n_pairs == 1 && !sweep && !search_only && !corr_out && !d2_out && while_mode != 0;
    const int chained = !can_while ? launches : (while_mode == 1 ? 0 : (launches > kChainSteps ? kChainSteps : launches));
    auto launch_one = [&](cudaStream_t launch_stream, unsigned long long loop_handle) {
#define B2_ICP_LAUNCH(M, K)                                                                                   \
    launch_iter<M, K>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states, partial, \
                      tickets, results_dev, corr_out, ndone, d2_out, search_only, launch_stream, loop_handle)
#define B2_ICP_SWEEP(M, K)                                                                                     \
    launch_sweep<M, K>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states, partial, \
                       tickets, results_dev, corr_out, ndone, d2_out, search_only)
        if (prm.mode == 0) {
            if (grid.is_map == 2) B2_ICP_LAUNCH(0, 2);
            else if (!sweep && grid.is_map == 1) B2_ICP_LAUNCH(0, 1);
            else if (!sweep) B2_ICP_LAUNCH(0, 0);
            else if (grid.is_map == 1) B2_ICP_SWEEP(0, 1);
            else B2_ICP_SWEEP(0, 0);
        } else {
            if (grid.is_map == 2) B2_ICP_LAUNCH(1, 2);
            else if (!sweep && grid.is_map == 1) B2_ICP_LAUNCH(1, 1);
            else if (!sweep) B2_ICP_LAUNCH(1, 0);
            else if (grid.is_map == 1) B2_ICP_SWEEP(1, 1);
            else B2_ICP_SWEEP(1, 0);
        }
#undef B2_ICP_LAUNCH
#undef B2_ICP_SWEEP
    };
    for (int it = 0; it < chained; ++it) {
        launch_one(c->stream, 0ull);
        B2_LAUNCH_CHECK(c);
        if (timing) cudaEventRecord(evs[it + 1], c->stream);
        if (check_every > 0 && !search_only && !timing && (it + 1) % check_every == 0 && it + 1 < launches) {
            B2_CUDA_OK(c, cudaMemcpyAsync(host_done, ndone, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
            if (*host_done >= n_pairs) break;
        }
    }
    if (chained < launches) {  // the remaining steps: a WHILE node after the chained launches
        cudaStreamCaptureStatus cs;
        cudaGraph_t parent = nullptr;
        const cudaGraphNode_t* deps = nullptr;
        size_t n_deps = 0;
        B2_CUDA_OK(c, cudaStreamGetCaptureInfo(c->stream, &cs, nullptr, &parent, &deps, &n_deps));
        cudaGraphConditionalHandle handle;
        B2_CUDA_OK(c, cudaGraphConditionalHandleCreate(&handle, parent, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t loop_node;
        B2_CUDA_OK(c, cudaGraphAddNode(&loop_node, parent, deps, n_deps, &np));
        // the body is captured through the copy stream (capture only: nothing runs on that stream)
        B2_CUDA_OK(c, cudaStreamBeginCaptureToGraph(c->copy_stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                                    cudaStreamCaptureModeThreadLocal));
        launch_one(c->copy_stream, (unsigned long long)handle);
        cudaGraph_t body = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->copy_stream, &body);
        B2_LAUNCH_CHECK(c);
        B2_CUDA_OK(c, e);
        B2_CUDA_OK(c, cudaStreamUpdateCaptureDependencies(c->stream, &loop_node, 1, cudaStreamSetCaptureDependencies));
    }
    if (timing) {
        cudaStreamSynchronize(c->stream);
        fprintf(stderr, "[b2 timing] icp launches (us):");
        for (int it = 0; it < launches; ++it) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, evs[it], evs[it + 1]);
            fprintf(stderr, " %.1f", ms * 1e3f);
        }
        fprintf(stderr, "\n");
#ifdef B2_TAIL_TIMING
        {
            static unsigned long long h[8 + 2 * 1024];
            cudaMemcpyFromSymbol(h, g_tail_ts, sizeof(h));
            unsigned long long t0 = ~0ull, tmain_min = ~0ull, tmain_max = 0, tstart_max = 0;
            for (int b = 0; b < gx && b < 1024; ++b) {
                if (h[8 + 2 * b] < t0) t0 = h[8 + 2 * b];
                if (h[8 + 2 * b] > tstart_max) tstart_max = h[8 + 2 * b];
                if (h[8 + 2 * b + 1] < tmain_min) tmain_min = h[8 + 2 * b + 1];
                if (h[8 + 2 * b + 1] > tmain_max) tmain_max = h[8 + 2 * b + 1];
            }
            fprintf(stderr, "[b2 blocks main-done-minus-start ns]");
            for (int b = 0; b < gx && b < 1024; ++b) fprintf(stderr, " %llu", h[8 + 2 * b + 1] - h[8 + 2 * b]);
            fprintf(stderr, "\n");
            fprintf(stderr, "[b2 tail] last launch (ns from first CTA start): last CTA start %llu | main done min %llu max %llu | "
                            "last-CTA: ticket won %llu, reduced %llu, step done %llu\n",
                    tstart_max - t0, tmain_min - t0, tmain_max - t0, h[0] - t0, h[1] - t0, h[2] - t0);
        }
#endif
        for (auto& e : evs) cudaEventDestroy(e);
    }
    return B2_OK;
}

}  // namespace b2
