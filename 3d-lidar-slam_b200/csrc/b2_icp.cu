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
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 512;
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
            if (gamma != 0.0 && fabs(gamma) > 1e-16 * sqrt(alpha * beta)) {
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

__device__ void mat4_mul_inplace_left(const double U[16], double T[16]) {
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += U[4 * i + k] * T[4 * k + j];
            r[4 * i + j] = s;
        }
    for (int i = 0; i < 16; ++i) T[i] = r[i];
}

// Evaluate one A.3 loop step from the reduced sums.  Runs on one thread.
template <int MODE>
__device__ __noinline__ void icp_step(const double* S, int ns, const IcpParams prm, IcpState* st, IcpResultDev* res,
                         int* ndone) {
    const double n = MODE == 0 ? S[16] : S[28];
    const double sd2 = MODE == 0 ? S[15] : S[27];
    const double fitness = ns > 0 ? n / (double)ns : 0.0;
    const double rmse = n > 0 ? sqrt(sd2 / n) : 0.0;
    bool done = false;
    if (st->has_prev && fabs(st->prev_fitness - fitness) < prm.rel_fitness &&
        fabs(st->prev_rmse - rmse) < prm.rel_rmse)
        done = true;
    if (st->iter >= prm.max_iter) done = true;
    st->fitness = fitness;
    st->rmse = rmse;
    st->n_corr = (int)n;
    if (done) {
        st->done = 1;
        for (int i = 0; i < 16; ++i) res->T[i] = st->T[i];
        res->fitness = fitness;
        res->rmse = rmse;
        res->iterations = st->iter;
        res->n_corr = (int)n;
        atomicAdd(ndone, 1);
        return;
    }
    double U[16];
    for (int i = 0; i < 16; ++i) U[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (n > 0) {
        if (MODE == 0) {
            double mp[3], mq[3], C[9], R[9];
            const double inv = 1.0 / n;
            for (int k = 0; k < 3; ++k) {
                mp[k] = S[k] * inv;
                mq[k] = S[3 + k] * inv;
            }
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) C[3 * i + j] = S[6 + 3 * i + j] * inv - mq[i] * mp[j];
            rotation_from_covariance(C, R);
            for (int k = 0; k < 3; ++k) {
                mp[k] += st->pivot[k];
                mq[k] += st->pivot[k];
            }
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) U[4 * i + j] = R[3 * i + j];
                U[4 * i + 3] = mq[i] - (R[3 * i] * mp[0] + R[3 * i + 1] * mp[1] + R[3 * i + 2] * mp[2]);
            }
            for (int k = 0; k < 3; ++k) st->pivot[k] = mq[k];
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
    mat4_mul_inplace_left(U, st->T);
    st->iter += 1;
    st->prev_fitness = fitness;
    st->prev_rmse = rmse;
    st->has_prev = 1;
}

// ---------------------------------------------------------------------------
// the hot kernel
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_lower_bound_gt(int incl, int c) {
    // smallest lane l with incl[l] > c  (incl is non-decreasing over lanes)
    int l = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        int probe = l + step - 1;
        int v = __shfl_sync(kFull, incl, probe & 31);
        if (v <= c) l += step;
    }
    return l;
}

template <int MODE, int IS_MAP>
__global__ void __launch_bounds__(kThreads, 2)
icp_iter_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs, const int* __restrict__ ns_dev,
                int ns_max, const CellEntry* __restrict__ entries, unsigned mask,
                const float4* __restrict__ gpts, const GridMeta* __restrict__ metas,
                const float4* __restrict__ tgt_normals, const int* __restrict__ tgt_offs, IcpParams prm,
                IcpState* __restrict__ states, double* __restrict__ partial, int* __restrict__ tickets,
                IcpResultDev* __restrict__ results, int* __restrict__ corr_out, int* __restrict__ ndone,
                double* __restrict__ d2_out, int search_only) {
    const int pair = blockIdx.y;
    IcpState* st = states + pair;
    if (st->done) return;

    __shared__ double sT[12];
    __shared__ double sPivot[3];
    __shared__ GridMeta sMeta;
    __shared__ double sVals[kWarps][16];
    __shared__ double sAcc[kWarps][kTerms];
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ int sLast;

    if (threadIdx.x < 12) sT[threadIdx.x] = st->T[threadIdx.x];
    if (threadIdx.x < 3) sPivot[threadIdx.x] = st->pivot[threadIdx.x];
    if (threadIdx.x == 0) sMeta = metas[IS_MAP ? 0 : pair];
    __syncthreads();

    int sb, se;
    if (src_offs) {
        sb = src_offs[pair];
        se = src_offs[pair + 1];
    } else {
        sb = 0;
        se = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    }
    const int tgt_base = tgt_offs ? tgt_offs[pair] : 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double r2 = dmul(prm.max_corr, prm.max_corr);
    const double cell = sMeta.cell;
    const int brick = sMeta.brick;
    const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;

    // Which product this lane accumulates: term = vals[ia] * vals[ib].
    // vals layout p2p: 0:1 1..3:P' 4..6:Q' 7:d2   p2l: 0:1 1..6:J 7:r 8:d2
    int ia = 0, ib = 0;
    if (MODE == 0) {
        if (lane < 3) ia = 1 + lane;
        else if (lane < 6) ia = 4 + (lane - 3);
        else if (lane < 15) { ia = 4 + (lane - 6) / 3; ib = 1 + (lane - 6) % 3; }
        else if (lane == 15) ia = 7;
        else if (lane == 16) ia = 0;
        else ia = -1;
    } else {
        if (lane < 21) {
            int t = lane, i = 0;
            while (t >= 6 - i) { t -= 6 - i; ++i; }
            ia = 1 + i;
            ib = 1 + i + t;
        } else if (lane < 27) { ia = 1 + (lane - 21); ib = 7; }
        else if (lane == 27) ia = 8;
        else if (lane == 28) ia = 0;
        else ia = -1;
    }
    double acc = 0.0;

    const int warp_global = blockIdx.x * kWarps + w;
    const int warp_stride = gridDim.x * kWarps;
    for (int q = sb + warp_global; q < se; q += warp_stride) {
        const float4 p = __ldg(src + q);
        const D3 P = xform_strict(sT, (double)p.x, (double)p.y, (double)p.z);
        int cx = voxel_coord(P.x, sMeta.origin[0], cell);
        int cy = voxel_coord(P.y, sMeta.origin[1], cell);
        int cz = voxel_coord(P.z, sMeta.origin[2], cell);
        if (IS_MAP && brick > 1) {
            cx = floor_div(cx, brick);
            cy = floor_div(cy, brick);
            cz = floor_div(cz, brick);
        }
        unsigned off = 0;
        int cnt = 0;
        if (lane < 27) {
            const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
            bool ok;
            unsigned long long key;
            if (IS_MAP) {
                ok = nx > -(1 << 20) && nx < (1 << 20) && ny > -(1 << 20) && ny < (1 << 20) && nz > -(1 << 20) &&
                     nz < (1 << 20);
                key = map_cell_key(nx, ny, nz);
            } else {
                ok = nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536;
                key = pair_cell_key((unsigned)pair, nx, ny, nz);
            }
            if (ok) {
                const CellEntry e = grid_lookup(entries, mask, key);
                if (e.key == key) {
                    off = e.off;
                    cnt = (int)e.cnt;
                }
            }
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        const int excl = incl - cnt;

        double best = INFINITY;
        int best_idx = 0x7fffffff;
        float bx = 0.f, by = 0.f, bz = 0.f;
        for (int c0 = 0; c0 < total; c0 += 32) {
            const int cnd = c0 + lane;
            const int owner = warp_lower_bound_gt(incl, cnd);
            const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
            const int o_excl = __shfl_sync(kFull, excl, owner & 31);
            if (cnd < total) {
                const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                const double d2 = dist2_strict(P.x, P.y, P.z, (double)t.x, (double)t.y, (double)t.z);
                const int idx = __float_as_int(t.w);
                if (d2 < best || (d2 == best && idx < best_idx)) {
                    best = d2;
                    best_idx = idx;
                    bx = t.x; by = t.y; bz = t.z;
                }
            }
        }
        // warp argmin over (d2, idx): d2 >= 0 so its bit pattern orders like the value
        const unsigned long long bits = (unsigned long long)__double_as_longlong(best);
        const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
        const unsigned mhi = __reduce_min_sync(kFull, hi);
        const unsigned mlo = __reduce_min_sync(kFull, hi == mhi ? lo : 0xffffffffu);
        const bool tie = hi == mhi && lo == mlo;
        const int midx = (int)__reduce_min_sync(kFull, tie ? (unsigned)best_idx : 0x7fffffffu);
        const unsigned win = __ballot_sync(kFull, tie && best_idx == midx);
        const int wl = __ffs(win) - 1;
        const double wbest = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
        const bool found = total > 0 && wbest < r2;
        if (corr_out && lane == 0) corr_out[q] = found ? midx : -1;
        if (d2_out && lane == 0) d2_out[q] = found ? wbest : 0.0;
        if (search_only) continue;
        if (found) {  // warp-uniform
            const double Qx = (double)__shfl_sync(kFull, bx, wl);
            const double Qy = (double)__shfl_sync(kFull, by, wl);
            const double Qz = (double)__shfl_sync(kFull, bz, wl);
            if (MODE == 0) {
                if (lane == 0) {
                    double* v = sVals[w];
                    v[0] = 1.0;
                    v[1] = P.x - sPivot[0]; v[2] = P.y - sPivot[1]; v[3] = P.z - sPivot[2];
                    v[4] = Qx - sPivot[0];  v[5] = Qy - sPivot[1];  v[6] = Qz - sPivot[2];
                    v[7] = wbest;
                }
            } else {
                if (lane == 0) {
                    const float4 nq = __ldg(tgt_normals + tgt_base + midx);
                    const double nx = nq.x, ny = nq.y, nz = nq.z;
                    double* v = sVals[w];
                    v[0] = 1.0;
                    v[1] = P.y * nz - P.z * ny;
                    v[2] = P.z * nx - P.x * nz;
                    v[3] = P.x * ny - P.y * nx;
                    v[4] = nx; v[5] = ny; v[6] = nz;
                    v[7] = (P.x - Qx) * nx + (P.y - Qy) * ny + (P.z - Qz) * nz;
                    v[8] = wbest;
                }
            }
            __syncwarp();
            if (ia >= 0) acc += sVals[w][ia] * sVals[w][ib];
            __syncwarp();
        }
    }
    if (search_only) return;

    // CTA reduction in a fixed order
    sAcc[w][lane] = acc;
    __syncthreads();
    if (threadIdx.x < kTerms) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kWarps; ++k) s += sAcc[k][threadIdx.x];
        partial[((size_t)pair * gridDim.x + blockIdx.x) * kTerms + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = atomicAdd(&tickets[pair], 1);
        sLast = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    {
        const int j = threadIdx.x & 31, part = threadIdx.x >> 5;
        double s = 0.0;
        for (int b = part; b < (int)gridDim.x; b += kThreads / 32)
            s += __ldcg(partial + ((size_t)pair * gridDim.x + b) * kTerms + j);
        sRed[part][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < kTerms) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) s += sRed[k][threadIdx.x];
        sAcc[0][threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tickets[pair] = 0;
        icp_step<MODE>(sAcc[0], se - sb, prm, st, results + pair, ndone);
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
                 int search_only) {
    icp_iter_kernel<MODE, IS_MAP><<<grid, kThreads, 0, c->stream>>>(src, src_offs, ns_dev, ns_max, g.entries, g.mask,
                                                                   g.pts, g.meta, nrm, tgt_offs, prm, st, partial,
                                                                   tickets, res, corr, ndone, d2_out, search_only);
}

// Candidate census of the resident-map stencil (the data-dependent term of the byte model).
__global__ void __launch_bounds__(256)
stencil_count_kernel(const float4* __restrict__ src, int ns, const double* __restrict__ T,
                     const CellEntry* __restrict__ entries, unsigned mask, const GridMeta* __restrict__ meta,
                     unsigned long long* __restrict__ total) {
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
    if (brick > 1) {
        cx = floor_div(cx, brick);
        cy = floor_div(cy, brick);
        cz = floor_div(cz, brick);
    }
    unsigned cnt = 0;
    if (lane < 27) {
        const int nx = cx + lane / 9 - 1, ny = cy + (lane / 3) % 3 - 1, nz = cz + lane % 3 - 1;
        const unsigned long long key = map_cell_key(nx, ny, nz);
        const CellEntry e = grid_lookup(entries, mask, key);
        if (e.key == key) cnt = e.cnt;
    }
    cnt = __reduce_add_sync(kFull, cnt);
    if (lane == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

}  // namespace

int map_stencil_count(Context* c, const float4* src, int ns, const double* T_dev, const GridView& g,
                      unsigned long long* total_dev) {
    stencil_count_kernel<<<div_up((long long)ns * 32, 256), 256, 0, c->stream>>>(src, ns, T_dev, g.entries, g.mask,
                                                                                  g.meta, total_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

int icp_grid_x(int ns_max, int n_pairs) {
    // one CTA of 16 warps per ~64 queries, capped at two resident CTAs per SM for a
    // single pair; batches already fill the machine through gridDim.y
    int gx = div_up(ns_max / (n_pairs > 0 ? n_pairs : 1) + 1, kWarps * 4);
    int cap = n_pairs >= 2 * kSMs ? 2 : (2 * kSMs) / (n_pairs > 0 ? n_pairs : 1);
    if (cap < 1) cap = 1;
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
    const int gx = icp_grid_x(ns_max, n_pairs);
    B2_WS(c, states, IcpState, WS_ICP_STATE, sizeof(IcpState) * (size_t)n_pairs);
    B2_WS(c, partial, double, WS_ICP_PARTIAL, sizeof(double) * kTerms * (size_t)gx * n_pairs);
    B2_WS(c, tickets, int, WS_ICP_TICKET, sizeof(int) * ((size_t)n_pairs + 1));
    int* ndone = tickets + n_pairs;
    icp_init_kernel<<<div_up(n_pairs, 128), 128, 0, c->stream>>>(src, src_offsets, ns_dev, ns_max, init_dev, 16,
                                                                 states, tickets, ndone, n_pairs);
    B2_LAUNCH_CHECK(c);
    dim3 g(gx, n_pairs);
    const int launches = search_only ? 1 : prm.max_iter + 1;
    int* host_done = (int*)c->host_mailbox + 512;  // scratch word in the pinned mailbox
    for (int it = 0; it < launches; ++it) {
        if (prm.mode == 0) {
            if (grid.is_map)
                launch_iter<0, 1>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states,
                                  partial, tickets, results_dev, corr_out, ndone, d2_out, search_only);
            else
                launch_iter<0, 0>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states,
                                  partial, tickets, results_dev, corr_out, ndone, d2_out, search_only);
        } else {
            if (grid.is_map)
                launch_iter<1, 1>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states,
                                  partial, tickets, results_dev, corr_out, ndone, d2_out, search_only);
            else
                launch_iter<1, 0>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states,
                                  partial, tickets, results_dev, corr_out, ndone, d2_out, search_only);
        }
        B2_LAUNCH_CHECK(c);
        if (check_every > 0 && !search_only && (it + 1) % check_every == 0 && it + 1 < launches) {
            B2_CUDA_OK(c, cudaMemcpyAsync(host_done, ndone, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
            if (*host_done >= n_pairs) break;
        }
    }
    return B2_OK;
}

}  // namespace b2
