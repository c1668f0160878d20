// b2_filters.cu — map-maintenance filters of ICPProcessor.do_stuff (icp_processor.py:121-148):
//   PointCloud::RemoveRadiusOutliers(nb_points, radius)         (icp_processor.py:139)
//   PointCloud::RemoveStatisticalOutliers(nb_neighbors, ratio)  (icp_processor.py:136,145)
// plus the order-preserving selection both are followed by.
//
// Both run on the uniform grid of b2_grid.cu, one warp per point:
//   radius filter : cell = radius, 27-cell stencil, count of neighbours with d^2 < r^2 (nanoflann's
//                   RadiusResultSet is strict), keep when count > nb_points (self included).
//   statistical   : exact k-nearest mean distance without a tree.  The warp gathers the squared
//                   distances of every point in the (2m+1)^3 stencil into shared memory, picks the
//                   k-th smallest with an 8-bit MSB radix select on the fp64 bit patterns, and accepts
//                   it once it cannot be beaten from outside the stencil (d_k <= (m*cell)^2);
//                   otherwise the stencil grows by one ring.  A stencil with more candidates than the
//                   buffer holds is swept twice more: histogram of d^2 -> bin of the k-th -> buffer only
//                   the candidates up to that bin.  mean_i = sum of the k smallest
//                   sqrt(d^2) / k; then mu, sigma (Bessel) over points with mean_i > 0 and
//                   keep_i = 0 < mean_i < mu + ratio * sigma — reduced in a fixed order.
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kBufCap = 1024;  // candidate distances a warp can hold in shared memory
constexpr unsigned kFull = 0xffffffffu;
constexpr double kUnresolved = -2.0;  // mean_out marker: k-NN not settled yet

__device__ __forceinline__ int warp_owner(int incl, int c) {
    int l = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        int v = __shfl_sync(kFull, incl, (l + step - 1) & 31);
        if (v <= c) l += step;
    }
    return l;
}

__global__ void __launch_bounds__(256)
radius_count_kernel(const float4* __restrict__ pts, int n, const CellEntry* __restrict__ entries, unsigned mask,
                    const float4* __restrict__ gpts, const GridMeta* __restrict__ meta, double radius, int nb_points,
                    unsigned char* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (i >= n) return;
    const float4 p = __ldg(pts + i);
    const double Px = p.x, Py = p.y, Pz = p.z;
    const double r2 = dmul(radius, radius);
    const int cx = voxel_coord(Px, meta->origin[0], meta->cell);
    const int cy = voxel_coord(Py, meta->origin[1], meta->cell);
    const int cz = voxel_coord(Pz, meta->origin[2], meta->cell);
    unsigned off = 0;
    int cnt = 0;
    if (lane < 27) {
        const int nx = cx + lane / 9 - 1, ny = cy + (lane / 3) % 3 - 1, nz = cz + lane % 3 - 1;
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
    if (lane == 0) keep[i] = inside > nb_points ? 1 : 0;
}

// Walk the (2m+1)^3 stencil of a point with the whole warp: cells are looked up 32 at a time, their points
// are dealt to the lanes in stencil order, and fn(running_index, d2_bits, lane_is_live) is called for every
// candidate — by all 32 lanes together (live == false for the padding lanes of the last trip), so fn may use
// warp collectives.  Returns the number of candidates.
template <typename Fn>
__device__ __forceinline__ int sweep_stencil(int m, int cx, int cy, int cz, double Px, double Py, double Pz,
                                             const CellEntry* __restrict__ entries, unsigned mask,
                                             const float4* __restrict__ gpts, Fn&& fn) {
    const int lane = threadIdx.x & 31;
    const int side = 2 * m + 1;
    const int ncell = side * side * side;
    int total = 0;
    for (int c0 = 0; c0 < ncell; c0 += 32) {
        const int c = c0 + lane;
        unsigned off = 0;
        int cnt = 0;
        if (c < ncell) {
            const int nx = cx + c / (side * side) - m, ny = cy + (c / side) % side - m, nz = cz + c % side - m;
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
        const int sum = __shfl_sync(kFull, incl, 31);
        const int excl = incl - cnt;
        for (int t0 = 0; t0 < sum; t0 += 32) {
            const int cnd = t0 + lane;
            const int owner = warp_owner(incl, cnd);
            const unsigned o_off = __shfl_sync(kFull, off, owner & 31);
            const int o_excl = __shfl_sync(kFull, excl, owner & 31);
            const bool live = cnd < sum;
            unsigned long long bits = 0;
            if (live) {
                const float4 t = __ldg(gpts + (o_off + (unsigned)(cnd - o_excl)));
                bits = (unsigned long long)__double_as_longlong(
                    dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z));
            }
            fn(total + cnd, bits, live);
        }
        total += sum;
    }
    return total;
}

// k-nearest mean distance, one warp per point.
__global__ void __launch_bounds__(kThreads)
knn_mean_kernel(const float4* __restrict__ pts, int n, const CellEntry* __restrict__ entries, unsigned mask,
                const float4* __restrict__ gpts, const GridMeta* __restrict__ meta, const SortCtl* __restrict__ ctl,
                int k_req, int max_ring, double* __restrict__ mean_out) {
    __shared__ unsigned long long sbuf[kWarps][kBufCap];
    __shared__ unsigned shist[kWarps][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * kWarps + w;
    if (i >= n) return;
    if (mean_out[i] != kUnresolved) return;  // settled on a finer grid level
    const int k = min(k_req, n);
    const float4 p = __ldg(pts + i);
    const double Px = p.x, Py = p.y, Pz = p.z;
    const double cell = meta->cell;
    const int cx = voxel_coord(Px, meta->origin[0], cell);
    const int cy = voxel_coord(Py, meta->origin[1], cell);
    const int cz = voxel_coord(Pz, meta->origin[2], cell);
    // grid extent in cells (upper bound): beyond it the stencil holds the whole cloud
    const int ext = 1 << max(ctl->bits[0], max(ctl->bits[1], ctl->bits[2]));
    unsigned long long* buf = sbuf[w];
    unsigned* hist = shist[w];

    for (int m = 1;; ++m) {
        // a sparse neighbourhood is handed to the next, 4x coarser grid level instead of sweeping
        // ever larger shells of empty cells here (the coarsest level has max_ring = INT_MAX)
        if (m > max_ring) return;
        // ---- gather the squared distances of every point in the (2m+1)^3 stencil -------------------
        int total = sweep_stencil(m, cx, cy, cz, Px, Py, Pz, entries, mask, gpts,
                                  [&](int idx, unsigned long long bits, bool live) {
                                      if (live && idx < kBufCap) buf[idx] = bits;
                                  });
        const bool covers_all = m >= ext;
        if (total < k && !covers_all) continue;  // not enough candidates yet: next ring
        __syncwarp();
        const int kk0 = min(k, total);
        bool buffered = total <= kBufCap;
        if (!buffered) {
            // ---- more candidates than the buffer holds (cell coarse for the local density): two more sweeps.
            //      Sweep A histograms the distances on 1024 bins that are linear in d^2 (a surface-like
            //      neighbourhood fills them evenly), which locates the bin of the k-th smallest; sweep B
            //      buffers, in stencil order, only the candidates up to that bin — the k nearest plus a few.
            unsigned* hist1k = reinterpret_cast<unsigned*>(buf);  // the overflowed buffer holds nothing of use
            for (int b = lane; b < 1024; b += 32) hist1k[b] = 0;
            __syncwarp();
            const double reach1 = dmul((double)(m + 1), cell);
            const double scale = 1024.0 / dmul(3.0, dmul(reach1, reach1));  // every candidate has d2 <= 3 reach1^2
            auto bin_of = [&](unsigned long long bits) -> int {
                const double v = __longlong_as_double((long long)bits) * scale;  // monotone in d2
                return v >= 1023.0 ? 1023 : (int)v;
            };
            sweep_stencil(m, cx, cy, cz, Px, Py, Pz, entries, mask, gpts,
                          [&](int, unsigned long long bits, bool live) {
                              if (live) atomicAdd(&hist1k[bin_of(bits)], 1u);
                          });
            __syncwarp();
            // lane l owns bins [32 l, 32 l + 32): find the bin where the running count reaches kk0
            unsigned lsum = 0;
            for (int b = 0; b < 32; ++b) lsum += hist1k[lane * 32 + b];
            unsigned incl = lsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            const bool mine = incl - lsum < (unsigned)kk0 && (unsigned)kk0 <= incl;
            int cut = 0, upto = 0;
            if (mine) {
                unsigned run = incl - lsum;
                for (int b = 0; b < 32; ++b) {
                    run += hist1k[lane * 32 + b];
                    if (run >= (unsigned)kk0) {
                        cut = lane * 32 + b;
                        upto = (int)run;
                        break;
                    }
                }
            }
            const int src = __ffs(__ballot_sync(kFull, mine)) - 1;
            cut = __shfl_sync(kFull, cut, src);
            upto = __shfl_sync(kFull, upto, src);  // candidates in bins <= cut
            __syncwarp();
            if (upto <= kBufCap) {
                int filled = 0;
                sweep_stencil(m, cx, cy, cz, Px, Py, Pz, entries, mask, gpts,
                              [&](int, unsigned long long bits, bool live) {
                                  const bool take = live && bin_of(bits) <= cut;
                                  const unsigned votes = __ballot_sync(kFull, take);
                                  if (take) buf[filled + __popc(votes & ((1u << lane) - 1u))] = bits;
                                  filled += __popc(votes);
                              });
                total = filled;  // == upto: the k nearest are among them, in stencil order
                buffered = true;
                __syncwarp();
            }
        }
        unsigned long long kth = 0;  // bit pattern of the kk0-th smallest d^2
        if (buffered) {
            // ---- MSB radix select over the buffered bit patterns -------------------------------------
            unsigned long long prefix = 0;
            int want = kk0;  // 1-based rank among the elements matching the prefix
            for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
                for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
                __syncwarp();
                for (int e = lane; e < total; e += 32) {
                    const unsigned long long v = buf[e];
                    if (shift == 56 || (v >> (shift + 8)) == (prefix >> (shift + 8)))
                        atomicAdd(&hist[(unsigned)(v >> shift) & 255u], 1u);
                }
                __syncwarp();
                unsigned c8[8];
                unsigned lsum = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    c8[b] = hist[lane * 8 + b];
                    lsum += c8[b];
                }
                unsigned incl = lsum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned v = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += v;
                }
                const unsigned base = incl - lsum;
                const bool mine = base < (unsigned)want && (unsigned)want <= incl;
                int digit = 0, below = 0;
                if (mine) {
                    unsigned run = base;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        if (run < (unsigned)want && (unsigned)want <= run + c8[b]) {
                            digit = lane * 8 + b;
                            below = (int)run;
                        }
                        run += c8[b];
                    }
                }
                const int src = __ffs(__ballot_sync(kFull, mine)) - 1;
                digit = __shfl_sync(kFull, digit, src);
                below = __shfl_sync(kFull, below, src);
                want -= below;
                prefix |= (unsigned long long)digit << shift;
                __syncwarp();
            }
            kth = prefix;
        } else {
            // ---- degenerate neighbourhood (more than the buffer holds inside ONE histogram bin, e.g. thousands
            //      of coincident points): bisection on the bit pattern, one sweep per step.  Slow, exact. -----
            unsigned long long lo = 0, hi = 0x7ff0000000000000ull;
            while (lo < hi) {
                const unsigned long long mid = lo + ((hi - lo) >> 1);
                int cle = 0;
                sweep_stencil(m, cx, cy, cz, Px, Py, Pz, entries, mask, gpts,
                              [&](int, unsigned long long bits, bool live) { cle += (live && bits <= mid) ? 1 : 0; });
                cle = __reduce_add_sync(kFull, cle);
                if (cle >= kk0) hi = mid; else lo = mid + 1;
            }
            kth = lo;
        }
        const double dk = __longlong_as_double((long long)kth);
        const double reach = dmul((double)m, cell);
        if (!covers_all && !(dk <= dmul(reach, reach))) continue;  // a closer point may sit outside: next ring
        // ---- mean of the k smallest distances (ties at the k-th all have the same length) ------------
        double s = 0.0;
        int clt = 0;
        if (buffered) {
            for (int e = lane; e < total; e += 32) {
                const unsigned long long v = buf[e];
                if (v < kth) {
                    s += sqrt(__longlong_as_double((long long)v));
                    ++clt;
                }
            }
        } else {
            sweep_stencil(m, cx, cy, cz, Px, Py, Pz, entries, mask, gpts,
                          [&](int, unsigned long long bits, bool live) {
                              if (live && bits < kth) {
                                  s += sqrt(__longlong_as_double((long long)bits));
                                  ++clt;
                              }
                          });
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            s += __shfl_xor_sync(kFull, s, o);
            clt += __shfl_xor_sync(kFull, clt, o);
        }
        if (lane == 0) {
            const double tot = s + (double)(kk0 - clt) * sqrt(dk);
            mean_out[i] = kk0 > 0 ? tot / (double)kk0 : -1.0;
        }
        return;
    }
}

__global__ void fill_kernel(double* __restrict__ a, int n, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// Fixed-order reduction helpers for mu / sigma: block partials, then one block.
__global__ void __launch_bounds__(256)
stat_partial_kernel(const double* __restrict__ m, int n, const double* __restrict__ mu_dev, double* __restrict__ part) {
    // pass 0 (mu_dev == null): (sum of m_i > 0, count);  pass 1: sum of (m_i - mu)^2 over m_i > 0
    __shared__ double sa[256], sb[256];
    const double mu = mu_dev ? mu_dev[0] : 0.0;
    double a = 0.0, b = 0.0;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int beg = blockIdx.x * per, end = min(n, beg + per);
    for (int i = beg + threadIdx.x; i < end; i += 256) {
        const double v = m[i];
        if (v > 0.0) {
            if (mu_dev) a += (v - mu) * (v - mu);
            else { a += v; b += 1.0; }
        }
    }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            sa[threadIdx.x] += sa[threadIdx.x + o];
            sb[threadIdx.x] += sb[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = sa[0];
        part[2 * blockIdx.x + 1] = sb[0];
    }
}

__global__ void stat_final_kernel(const double* __restrict__ part, int nblocks, int pass, double std_ratio,
                                  double* __restrict__ stats /* [0] mu [1] valid [2] sigma [3] threshold */) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblocks; ++i) {
        a += part[2 * i];
        b += part[2 * i + 1];
    }
    if (pass == 0) {
        stats[1] = b;
        stats[0] = b > 0.0 ? a / b : 0.0;
    } else {
        const double valid = stats[1];
        const double sd = valid > 1.0 ? sqrt(a / (valid - 1.0)) : 0.0;
        stats[2] = sd;
        stats[3] = stats[0] + std_ratio * sd;
    }
}

__global__ void stat_mask_kernel(const double* __restrict__ m, int n, const double* __restrict__ stats,
                                 unsigned char* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = m[i];
    keep[i] = (stats[1] > 0.0 && v > 0.0 && v < stats[3]) ? 1 : 0;
}

// order-preserving compaction
__global__ void __launch_bounds__(256)
select_count_kernel(const unsigned char* __restrict__ keep, int n, int* __restrict__ blockcnt) {
    const int per = ((n + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
    const int beg = min(n, (int)blockIdx.x * per), end = min(n, beg + per);
    int c = 0;
    for (int i = beg + threadIdx.x; i < end; i += 256) c += keep[i] ? 1 : 0;
    __shared__ int s[8];
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int q = 0; q < 8; ++q) t += s[q];
        blockcnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
select_write_kernel(const float4* __restrict__ pts, const unsigned char* __restrict__ keep, int n,
                    const int* __restrict__ blockcnt, float4* __restrict__ out, int* __restrict__ n_out) {
    __shared__ int sbase, swarp[8];
    if (threadIdx.x < 32) {
        int acc = 0, tot = 0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += 32) {
            const int v = blockcnt[q];
            if (q < (int)blockIdx.x) acc += v;
            tot += v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc += __shfl_xor_sync(kFull, acc, o);
            tot += __shfl_xor_sync(kFull, tot, o);
        }
        if (threadIdx.x == 0) {
            sbase = acc;
            if (blockIdx.x == 0) *n_out = tot;
        }
    }
    __syncthreads();
    const int per = ((n + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
    const int beg = min(n, (int)blockIdx.x * per), end = min(n, beg + per);
    int running = sbase;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = beg; base < end; base += 256) {
        const int i = base + threadIdx.x;
        const bool k = i < end && keep[i];
        const unsigned bal = __ballot_sync(kFull, k);
        if (lane == 0) swarp[w] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q < w) woff += swarp[q];
            tot += swarp[q];
        }
        if (k) out[running + woff + __popc(bal & ((1u << lane) - 1u))] = pts[i];
        running += tot;
        __syncthreads();
    }
}

}  // namespace

int radius_outlier_mask(Context* c, const float4* pts, int n, int nb_points, double radius, unsigned char* keep) {
    if (n <= 0) return B2_OK;
    if (!(radius > 0.0) || nb_points < 0) return set_error(c, B2_ERR_INVALID, "radius filter: radius > 0 and nb_points >= 0 required");
    GridView g;
    B2_TRY(grid_build(c, pts, n, nullptr, nullptr, 1, radius, WS_GRID2_ENTRIES, &g));
    radius_count_kernel<<<div_up((long long)n * 32, 256), 256, 0, c->stream>>>(pts, n, g.entries, g.mask, g.pts, g.meta,
                                                                              radius, nb_points, keep);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

int statistical_outlier_mask(Context* c, const float4* pts, int n, int nb_neighbors, double std_ratio, double cell,
                             double extent, unsigned char* keep) {
    if (n <= 0) return B2_OK;
    if (nb_neighbors < 1 || !(cell > 0.0)) return set_error(c, B2_ERR_INVALID, "statistical filter: nb_neighbors >= 1 and cell > 0 required");
    B2_WS(c, mean, double, WS_FILTER_MEAN, sizeof(double) * (size_t)n);
    const int nb = 256;
    B2_WS(c, part, double, WS_FILTER_PART, sizeof(double) * (2 * nb + 8));
    double* stats = part + 2 * nb;
    fill_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(mean, n, kUnresolved);
    B2_LAUNCH_CHECK(c);
    // Multi-resolution sweep: level l uses cells of edge cell * 4^l and looks at most two rings out;
    // points whose k-th neighbour is not settled there are picked up by the next level.  The last
    // level's cells span the whole cloud, so everything is settled at the latest there.
    double level_cell = cell;
    for (int level = 0; level < 16; ++level) {
        const bool last = level_cell >= extent || level == 15;
        GridView g;
        B2_TRY(grid_build(c, pts, n, nullptr, nullptr, 1, level_cell, WS_GRID2_ENTRIES, &g));
        B2_WS(c, ctl, SortCtl, WS_CTL, sizeof(SortCtl));  // still holds this grid's key bit widths
        knn_mean_kernel<<<div_up(n, kWarps), kThreads, 0, c->stream>>>(pts, n, g.entries, g.mask, g.pts, g.meta, ctl,
                                                                      nb_neighbors, last ? 0x7fffffff : 2, mean);
        B2_LAUNCH_CHECK(c);
        if (last) break;
        level_cell *= 4.0;
    }
    stat_partial_kernel<<<nb, 256, 0, c->stream>>>(mean, n, nullptr, part);
    B2_LAUNCH_CHECK(c);
    stat_final_kernel<<<1, 1, 0, c->stream>>>(part, nb, 0, std_ratio, stats);
    B2_LAUNCH_CHECK(c);
    stat_partial_kernel<<<nb, 256, 0, c->stream>>>(mean, n, stats, part);
    B2_LAUNCH_CHECK(c);
    stat_final_kernel<<<1, 1, 0, c->stream>>>(part, nb, 1, std_ratio, stats);
    B2_LAUNCH_CHECK(c);
    stat_mask_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(mean, n, stats, keep);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

int select_points(Context* c, const float4* pts, int n, const unsigned char* keep, float4* out, int* n_out_dev) {
    if (n <= 0) return B2_OK;
    int nb = div_up(n, 256 * 8);
    if (nb > kSMs * 4) nb = kSMs * 4;
    B2_WS(c, blockcnt, int, WS_BLOCKCNT, sizeof(int) * kSMs * 4);
    select_count_kernel<<<nb, 256, 0, c->stream>>>(keep, n, blockcnt);
    B2_LAUNCH_CHECK(c);
    select_write_kernel<<<nb, 256, 0, c->stream>>>(pts, keep, n, blockcnt, out, n_out_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

}  // namespace b2
