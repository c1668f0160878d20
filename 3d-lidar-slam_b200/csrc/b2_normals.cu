// b2_normals.cu — PCA normal estimation.
//
// Replaces PointCloud::EstimateNormals(KDTreeSearchParamHybrid(radius, max_nn))
// (reference call sites icp_processor.py:94-99): for every point
// the (at most) max_nn nearest neighbours with d^2 < radius^2 — the point itself
// included — feed nine cumulants; the covariance's smallest-eigenvalue
// eigenvector is the normal ((0,0,1) when fewer than three neighbours).
//
// One warp serves 32 consecutive points: for each, 27 lanes probe the stencil
// cells of a uniform grid (cell = radius) in parallel, the lanes then stride the
// candidate points (coalesced float4 runs), and a fixed-order shuffle tree folds
// the cumulants.  When more than max_nn neighbours fall inside the radius (rare:
// max_nn = 20 at the reference's settings) a bit-wise bisection on the fp64
// distance pattern finds the max_nn-th nearest exactly (ties -> lowest index).
// Finally every lane solves its own point's 3x3 symmetric eigenproblem (cyclic
// Jacobi, fp64), so the dense algebra runs 32-wide and the float4 normals are
// written coalesced.
// Algorithmic bytes: 16 N (points) + 16 sum_q cand(q) + 16*27 N (cell records) +
// 16 N (normals out).
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ void smallest_eigenvector_sym3(double a00, double a01, double a02, double a11, double a12, double a22,
                                          double n[3]) {
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double scale = fabs(a00) + fabs(a11) + fabs(a22) + fabs(a01) + fabs(a02) + fabs(a12);
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (!(off > 1e-18 * scale)) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0;
            const int q = pq == 0 ? 1 : 2;
            const double apq = A[p][q];
            if (apq == 0.0) continue;
            const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- A * J
                const double akp = A[k][p], akq = A[k][q];
                A[k][p] = cs * akp - sn * akq;
                A[k][q] = sn * akp + cs * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- J^T * A
                const double apk = A[p][k], aqk = A[q][k];
                A[p][k] = cs * apk - sn * aqk;
                A[q][k] = sn * apk + cs * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = cs * vkp - sn * vkq;
                V[k][q] = sn * vkp + cs * vkq;
            }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    double x = V[0][m], y = V[1][m], z = V[2][m];
    const double nn = sqrt(x * x + y * y + z * z);
    if (!(nn > 0.0)) {
        n[0] = 0; n[1] = 0; n[2] = 1;
        return;
    }
    n[0] = x / nn; n[1] = y / nn; n[2] = z / nn;
}

__device__ __forceinline__ int warp_owner(int incl, int c) {
    int l = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        int v = __shfl_sync(kFull, incl, (l + step - 1) & 31);
        if (v <= c) l += step;
    }
    return l;
}

__global__ void __launch_bounds__(kThreads)
normals_kernel(const float4* __restrict__ pts, const int* __restrict__ n_dev, int n_max,
               const CellEntry* __restrict__ entries, unsigned mask, const float4* __restrict__ gpts,
               const GridMeta* __restrict__ meta, double radius, int max_nn, float4* __restrict__ normals) {
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * kWarps;
    const double r2 = dmul(radius, radius);
    const double ox = meta->origin[0], oy = meta->origin[1], oz = meta->origin[2], cell = meta->cell;
    const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
    const unsigned long long r2bits = (unsigned long long)__double_as_longlong(r2);

    for (int base = warp_global * 32; base < n; base += warp_stride * 32) {
        double mine[9];
        int mine_k = 0;
#pragma unroll
        for (int t = 0; t < 9; ++t) mine[t] = 0.0;
        const int qn = min(32, n - base);
        for (int qi = 0; qi < qn; ++qi) {
            const float4 p = __ldg(pts + base + qi);
            const double Px = p.x, Py = p.y, Pz = p.z;
            const int cx = voxel_coord(Px, ox, cell), cy = voxel_coord(Py, oy, cell), cz = voxel_coord(Pz, oz, cell);
            unsigned off = 0;
            int cnt = 0;
            if (lane < 27) {
                const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                if (nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536) {
                    const unsigned long long key = pair_cell_key(0u, nx, ny, nz);
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

            // pass 1: count neighbours inside the radius
            int inside = 0;
            for (int c0 = 0; c0 < total; c0 += 32) {
                const int cnd = c0 + lane;
                const int owner = warp_owner(incl, cnd);
                const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
                const int o_excl = __shfl_sync(kFull, excl, owner & 31);
                bool in = false;
                if (cnd < total) {
                    const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                    in = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z) < r2;
                }
                inside += __popc(__ballot_sync(kFull, in));
            }
            // selection threshold: everything inside the radius, or the max_nn nearest
            unsigned long long thr_bits = r2bits;  // admit d2 with bits < thr_bits ...
            int tie_idx_max = -1;                  // ... plus d2 bits == thr_bits with idx <= tie_idx_max
            int k_used = inside;
            if (inside > max_nn) {
                k_used = max_nn;
                unsigned long long lo = 0, hi = r2bits;
                while (lo < hi) {  // smallest v with count(d2 <= v) >= max_nn
                    const unsigned long long mid = lo + ((hi - lo) >> 1);
                    int cle = 0;
                    for (int c0 = 0; c0 < total; c0 += 32) {
                        const int cnd = c0 + lane;
                        const int owner = warp_owner(incl, cnd);
                        const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
                        const int o_excl = __shfl_sync(kFull, excl, owner & 31);
                        bool le = false;
                        if (cnd < total) {
                            const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                            const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                            le = d2 < r2 && (unsigned long long)__double_as_longlong(d2) <= mid;
                        }
                        cle += __popc(__ballot_sync(kFull, le));
                    }
                    if (cle >= max_nn) hi = mid; else lo = mid + 1;
                }
                thr_bits = lo;
                // ties at the threshold distance: admit the lowest indices until max_nn is reached
                int clt = 0;
                for (int c0 = 0; c0 < total; c0 += 32) {
                    const int cnd = c0 + lane;
                    const int owner = warp_owner(incl, cnd);
                    const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
                    const int o_excl = __shfl_sync(kFull, excl, owner & 31);
                    bool lt = false;
                    if (cnd < total) {
                        const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                        const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                        lt = (unsigned long long)__double_as_longlong(d2) < thr_bits;
                    }
                    clt += __popc(__ballot_sync(kFull, lt));
                }
                const int need = max_nn - clt;
                int ilo = 0, ihi = 0x7ffffffe;
                while (ilo < ihi) {  // smallest index bound admitting `need` ties
                    const int mid = ilo + ((ihi - ilo) >> 1);
                    int ct = 0;
                    for (int c0 = 0; c0 < total; c0 += 32) {
                        const int cnd = c0 + lane;
                        const int owner = warp_owner(incl, cnd);
                        const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
                        const int o_excl = __shfl_sync(kFull, excl, owner & 31);
                        bool hit = false;
                        if (cnd < total) {
                            const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                            const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                            hit = (unsigned long long)__double_as_longlong(d2) == thr_bits && __float_as_int(t.w) <= mid;
                        }
                        ct += __popc(__ballot_sync(kFull, hit));
                    }
                    if (ct >= need) ihi = mid; else ilo = mid + 1;
                }
                tie_idx_max = ilo;
            }
            // pass 2: cumulants over the admitted neighbours
            double cum[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) cum[t] = 0.0;
            if (k_used >= 3) {
                for (int c0 = 0; c0 < total; c0 += 32) {
                    const int cnd = c0 + lane;
                    const int owner = warp_owner(incl, cnd);
                    const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
                    const int o_excl = __shfl_sync(kFull, excl, owner & 31);
                    if (cnd < total) {
                        const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                        const double x = t.x, y = t.y, z = t.z;
                        const double d2 = dist2_strict(Px, Py, Pz, x, y, z);
                        const unsigned long long b = (unsigned long long)__double_as_longlong(d2);
                        const bool take = inside > max_nn
                                              ? (b < thr_bits || (b == thr_bits && __float_as_int(t.w) <= tie_idx_max))
                                              : d2 < r2;
                        if (take) {
                            cum[0] += x; cum[1] += y; cum[2] += z;
                            cum[3] += x * x; cum[4] += x * y; cum[5] += x * z;
                            cum[6] += y * y; cum[7] += y * z; cum[8] += z * z;
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) cum[t] += __shfl_xor_sync(kFull, cum[t], o);
                }
            }
            if (lane == qi) {
#pragma unroll
                for (int t = 0; t < 9; ++t) mine[t] = cum[t];
                mine_k = k_used;
            }
        }
        if (lane < qn) {
            double nrm[3] = {0.0, 0.0, 1.0};
            if (mine_k >= 3) {
                const double k = (double)mine_k;
                double c[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) c[t] = mine[t] / k;
                smallest_eigenvector_sym3(c[3] - c[0] * c[0], c[4] - c[0] * c[1], c[5] - c[0] * c[2],
                                          c[6] - c[1] * c[1], c[7] - c[1] * c[2], c[8] - c[2] * c[2], nrm);
            }
            normals[base + lane] = make_float4((float)nrm[0], (float)nrm[1], (float)nrm[2], 0.f);
        }
    }
}

}  // namespace

int estimate_normals(Context* c, const float4* pts, int n_max, const int* n_dev, double radius, int max_nn,
                     float4* normals_out) {
    if (n_max <= 0) return B2_OK;
    if (!(radius > 0.0) || max_nn < 1) return set_error(c, B2_ERR_INVALID, "estimate_normals: radius > 0 and max_nn >= 1 required");
    GridView g;
    B2_TRY(grid_build(c, pts, n_max, n_dev, nullptr, 1, radius, WS_GRID2_ENTRIES, &g));
    int blocks = div_up(div_up(n_max, 32), kWarps);
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    normals_kernel<<<blocks, kThreads, 0, c->stream>>>(pts, n_dev, n_max, g.entries, g.mask, g.pts, g.meta, radius,
                                                       max_nn, normals_out);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

}  // namespace b2
