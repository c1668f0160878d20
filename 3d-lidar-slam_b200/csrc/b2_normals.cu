// b2_normals.cu — PCA normal estimation.
//
// Replaces PointCloud::EstimateNormals(KDTreeSearchParamHybrid(radius, max_nn))
// (reference call sites icp_processor.py:94-99): for every point
// the (at most) max_nn nearest neighbours with d^2 < radius^2 — the point itself
// included — feed nine cumulants; the covariance's smallest-eigenvalue
// eigenvector is the normal ((0,0,1) when fewer than three neighbours).
//
// One warp serves 32 consecutive points, four at a time: an 8-lane group per point,
// each lane probing 3-4 of the 27 stencil cells of a uniform grid (cell = radius)
// with its hash probes in flight together, folding the cumulants of its own cells'
// points and a 3-step shuffle butterfly folding the group.  When more than max_nn
// neighbours fall inside the radius (rare: max_nn = 20 at the reference's settings)
// a bit-wise bisection on the fp64 distance pattern finds the max_nn-th nearest
// exactly (ties -> lowest index) inside the group.
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
        // a few ulp: a tighter bound is below the rounding noise of the rotations and never settles
        if (!(off > 4e-16 * scale)) break;
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

constexpr int kGroup = 8;     // lanes per query
constexpr int kCellsPerLane = 4;  // ceil(27 / 8)
constexpr int kStage = 64;        // in-radius candidates a group can rank in shared memory

// Walk the candidates of this lane's stencil cells and fold those admitted by `take(d2 bits, idx)`.
template <typename Take>
__device__ __forceinline__ void fold_candidates(const uint4 (&raw)[kCellsPerLane], const bool (&hit)[kCellsPerLane],
                                                const float4* __restrict__ gpts, double Px, double Py, double Pz,
                                                Take take, int& cnt, double (&cum)[9]) {
#pragma unroll
    for (int k = 0; k < kCellsPerLane; ++k) {
        if (!hit[k]) continue;
        const unsigned off = raw[k].z;
        const int n = (int)raw[k].w;
        for (int j = 0; j < n; ++j) {
            const float4 t = __ldg(gpts + off + j);
            const double x = t.x, y = t.y, z = t.z;
            const double d2 = dist2_strict(Px, Py, Pz, x, y, z);
            if (take((unsigned long long)__double_as_longlong(d2), __float_as_int(t.w))) {
                ++cnt;
                cum[0] += x; cum[1] += y; cum[2] += z;
                cum[3] += x * x; cum[4] += x * y; cum[5] += x * z;
                cum[6] += y * y; cum[7] += y * z; cum[8] += z * z;
            }
        }
    }
}

template <typename Pred>
__device__ __forceinline__ int count_candidates(const uint4 (&raw)[kCellsPerLane], const bool (&hit)[kCellsPerLane],
                                                const float4* __restrict__ gpts, double Px, double Py, double Pz,
                                                Pred pred) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < kCellsPerLane; ++k) {
        if (!hit[k]) continue;
        const unsigned off = raw[k].z;
        const int n = (int)raw[k].w;
        for (int j = 0; j < n; ++j) {
            const float4 t = __ldg(gpts + off + j);
            const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
            c += pred((unsigned long long)__double_as_longlong(d2), __float_as_int(t.w)) ? 1 : 0;
        }
    }
    return c;
}

__device__ __forceinline__ int group_sum(int v, unsigned gmask) {
#pragma unroll
    for (int o = 1; o < kGroup; o <<= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreads)
normals_kernel(const float4* __restrict__ pts, const int* __restrict__ n_dev, int n_max,
               const int* __restrict__ offs, int n_clouds, const CellEntry* __restrict__ entries, unsigned mask,
               const float4* __restrict__ gpts, const GridMeta* __restrict__ metas, double radius, int max_nn,
               float4* __restrict__ normals) {
    // staging area of a group for the capped case: (d2 bit pattern, index, position) of its in-radius
    // candidates, ranked against each other on chip
    __shared__ unsigned long long s_d2[kThreads / kGroup][kStage];
    __shared__ int2 s_ip[kThreads / kGroup][kStage];
    // (a batch of clouds: offs[0..n_clouds] delimit them, every cloud has its own grid origin and key prefix)
    const int n = offs ? offs[n_clouds] : (n_dev ? min(*n_dev, n_max) : n_max);
    const int lane = threadIdx.x & 31;
    const int g = lane & (kGroup - 1), gi = lane / kGroup;  // lane in group, group in warp
    const int gslot = threadIdx.x / kGroup;
    const unsigned gmask = ((1u << kGroup) - 1u) << (gi * kGroup);
    const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * kWarps;
    const double r2 = dmul(radius, radius);
    const unsigned long long r2bits = (unsigned long long)__double_as_longlong(r2);

    // a warp owns 32 consecutive points: 8 rounds of 4 concurrent group searches, then every lane
    // solves the eigenproblem of "its" point so the dense algebra runs 32-wide
    for (int base = warp_global * 32; base < n; base += warp_stride * 32) {
        double mine[9];
        int mine_k = 0;
#pragma unroll
        for (int t = 0; t < 9; ++t) mine[t] = 0.0;
        for (int r = 0; r < 32 / (32 / kGroup); ++r) {
            const int q = base + r * (32 / kGroup) + gi;
            const bool valid = q < n;
            double Px = 0.0, Py = 0.0, Pz = 0.0;
            int cx = 0, cy = 0, cz = 0, cloud = 0;
            if (g == 0 && valid) {
                const float4 p = __ldg(pts + q);
                Px = p.x; Py = p.y; Pz = p.z;
                if (offs) {  // offs[cloud] <= q < offs[cloud + 1]
                    int lo = 0, hi = n_clouds;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(offs + mid) <= q) lo = mid; else hi = mid;
                    }
                    cloud = lo;
                }
                const GridMeta* meta = metas + cloud;
                cx = voxel_coord(Px, meta->origin[0], meta->cell);
                cy = voxel_coord(Py, meta->origin[1], meta->cell);
                cz = voxel_coord(Pz, meta->origin[2], meta->cell);
            }
            const int lead = gi * kGroup;
            cloud = __shfl_sync(kFull, cloud, lead);
            Px = __shfl_sync(kFull, Px, lead);
            Py = __shfl_sync(kFull, Py, lead);
            Pz = __shfl_sync(kFull, Pz, lead);
            cx = __shfl_sync(kFull, cx, lead);
            cy = __shfl_sync(kFull, cy, lead);
            cz = __shfl_sync(kFull, cz, lead);

            uint4 raw[kCellsPerLane];
            bool hit[kCellsPerLane];
            unsigned long long key[kCellsPerLane];
            unsigned hpos[kCellsPerLane];
#pragma unroll
            for (int k = 0; k < kCellsPerLane; ++k) {
                const int c = g + kGroup * k;
                const int c3 = (c * 171) >> 9, c9 = (c * 57) >> 9;
                const int nx = cx + c9 - 1, ny = cy + (c3 - 3 * c9) - 1, nz = cz + (c - 3 * c3) - 1;
                hit[k] = valid && c < 27 && nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536;
                key[k] = pair_cell_key((unsigned)cloud, nx, ny, nz);
                hpos[k] = cell_hash(key[k]) & mask;
            }
#pragma unroll
            for (int k = 0; k < kCellsPerLane; ++k)
                if (hit[k]) raw[k] = __ldg(reinterpret_cast<const uint4*>(entries + hpos[k]));
#pragma unroll
            for (int k = 0; k < kCellsPerLane; ++k) {
                if (!hit[k]) continue;
                unsigned h = hpos[k];
                unsigned long long kk = ((unsigned long long)raw[k].y << 32) | raw[k].x;
                while (kk != key[k] && kk != kEmptyKey) {
                    h = (h + 1) & mask;
                    raw[k] = __ldg(reinterpret_cast<const uint4*>(entries + h));
                    kk = ((unsigned long long)raw[k].y << 32) | raw[k].x;
                }
                hit[k] = kk == key[k];
            }
            // optimistic single pass: everything inside the radius
            int cnt = 0;
            double cum[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) cum[t] = 0.0;
            fold_candidates(raw, hit, gpts, Px, Py, Pz,
                            [&](unsigned long long b, int) { return b < r2bits; }, cnt, cum);
            int inside = cnt;
#pragma unroll
            for (int o = 1; o < kGroup; o <<= 1) inside += __shfl_xor_sync(kFull, inside, o);
            int k_used = inside;
            if (inside > max_nn && inside <= kStage) {
                // The cap binds (common on dense, close-range surfaces): keep exactly the max_nn nearest,
                // ties by ascending index.  The group stages its in-radius candidates in shared memory and
                // every lane ranks its share against all of them — K^2/8 compares on chip instead of a
                // bisection that re-reads the candidates ~100 times.
                k_used = max_nn;
                int before = cnt;  // exclusive prefix of the per-lane in-radius counts inside the group
#pragma unroll
                for (int o = 1; o < kGroup; o <<= 1) {
                    const int v = __shfl_up_sync(gmask, before, o, kGroup);
                    if (g >= o) before += v;
                }
                before -= cnt;
                {
                    int w = before;
#pragma unroll
                    for (int k = 0; k < kCellsPerLane; ++k) {
                        if (!hit[k]) continue;
                        const unsigned off = raw[k].z;
                        const int m = (int)raw[k].w;
                        for (int j = 0; j < m; ++j) {
                            const float4 t = __ldg(gpts + off + j);
                            const unsigned long long b = (unsigned long long)__double_as_longlong(
                                dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z));
                            if (b < r2bits) {
                                s_d2[gslot][w] = b;
                                s_ip[gslot][w] = make_int2(__float_as_int(t.w), (int)(off + j));
                                ++w;
                            }
                        }
                    }
                }
                __syncwarp(gmask);
                cnt = 0;
#pragma unroll
                for (int t = 0; t < 9; ++t) cum[t] = 0.0;
                for (int e = g; e < inside; e += kGroup) {
                    const unsigned long long be = s_d2[gslot][e];
                    const int2 ie = s_ip[gslot][e];
                    int rank = 0;
                    for (int f = 0; f < inside; ++f) {
                        const unsigned long long bf = s_d2[gslot][f];
                        rank += (bf < be || (bf == be && s_ip[gslot][f].x < ie.x)) ? 1 : 0;
                    }
                    if (rank < max_nn) {
                        const float4 t = __ldg(gpts + ie.y);
                        const double x = t.x, y = t.y, z = t.z;
                        ++cnt;
                        cum[0] += x; cum[1] += y; cum[2] += z;
                        cum[3] += x * x; cum[4] += x * y; cum[5] += x * z;
                        cum[6] += y * y; cum[7] += y * z; cum[8] += z * z;
                    }
                }
                __syncwarp(gmask);
            } else if (inside > max_nn) {
                // more in-radius candidates than the staging area holds: exact max_nn-th nearest by
                // bisection on the fp64 bit pattern.  Group-uniform, so the shuffles use the group mask.
                k_used = max_nn;
                unsigned long long lo = 0, hi = r2bits;
                while (lo < hi) {  // smallest v with count(d2 <= v, d2 < r2) >= max_nn
                    const unsigned long long mid = lo + ((hi - lo) >> 1);
                    const int cle = group_sum(count_candidates(raw, hit, gpts, Px, Py, Pz,
                                                               [&](unsigned long long b, int) { return b < r2bits && b <= mid; }),
                                              gmask);
                    if (cle >= max_nn) hi = mid; else lo = mid + 1;
                }
                const unsigned long long thr = lo;
                const int clt = group_sum(count_candidates(raw, hit, gpts, Px, Py, Pz,
                                                           [&](unsigned long long b, int) { return b < thr; }), gmask);
                const int need = max_nn - clt;
                int ilo = 0, ihi = 0x7ffffffe;
                while (ilo < ihi) {  // smallest index bound admitting `need` ties
                    const int mid = ilo + ((ihi - ilo) >> 1);
                    const int ct = group_sum(count_candidates(raw, hit, gpts, Px, Py, Pz,
                                                              [&](unsigned long long b, int idx) { return b == thr && idx <= mid; }),
                                             gmask);
                    if (ct >= need) ihi = mid; else ilo = mid + 1;
                }
                const int tie_max = ilo;
                cnt = 0;
#pragma unroll
                for (int t = 0; t < 9; ++t) cum[t] = 0.0;
                fold_candidates(raw, hit, gpts, Px, Py, Pz,
                                [&](unsigned long long b, int idx) { return b < thr || (b == thr && idx <= tie_max); },
                                cnt, cum);
            }
            // fold the group's cumulants (fixed butterfly order) and hand them to the lane that will
            // solve this point: lane r*4 + gi
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
                for (int o = 1; o < kGroup; o <<= 1) cum[t] += __shfl_xor_sync(kFull, cum[t], o);
            }
            const int src = (lane & 3) * kGroup;
            const int kq = __shfl_sync(kFull, k_used, src);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const double v = __shfl_sync(kFull, cum[t], src);
                if ((lane >> 2) == r) mine[t] = v;
            }
            if ((lane >> 2) == r) mine_k = kq;
        }
        if (base + lane < n) {
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

int estimate_normals(Context* c, const float4* pts, int n_max, const int* n_dev, const int* offs, int n_clouds,
                     double radius, int max_nn, float4* normals_out) {
    if (n_max <= 0) return B2_OK;
    if (!(radius > 0.0) || max_nn < 1) return set_error(c, B2_ERR_INVALID, "estimate_normals: radius > 0 and max_nn >= 1 required");
    GridView g;
    B2_TRY(grid_build(c, pts, n_max, n_dev, offs, offs ? n_clouds : 1, radius, WS_GRID2_ENTRIES, &g));
    int blocks = div_up(div_up(n_max, 32), kWarps);
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    normals_kernel<<<blocks, kThreads, 0, c->stream>>>(pts, n_dev, n_max, offs, offs ? n_clouds : 1, g.entries, g.mask,
                                                       g.pts, g.meta, radius, max_nn, normals_out);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

}  // namespace b2
