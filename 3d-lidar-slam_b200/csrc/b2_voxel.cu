// b2_voxel.cu — voxel keys, sort-by-key segmentation and centroid reduction.
//
// Replaces PointCloud::VoxelDownSample (reference call sites
// icp_processor.py:90-91 and :132-133) and supplies the "sorted by voxel/cell"
// view that the uniform grid (b2_grid.cu) and the resident-map fusion
// (b2_map.cu) are built from.
//
// Pipeline (all sizes device-resident, no host round trip):
//   bbox_kernel   : float64 min/max bound of every cloud (after the optional T)
//   prep_kernel   : origin, key range, packed-key bit widths, radix pass count
//   keys_kernel   : key = floor((p - origin)/v) per axis (IEEE fp64 division,
//                   bit-identical to Open3D's arithmetic on the widened inputs),
//                   packed as [cloud | x | y | z] so an unsigned sort yields the
//                   canonical lexicographic voxel order
//   radix sort    : b2_sort.cu (stable => points of a voxel stay in input order)
//   seg_*_kernel  : unique keys -> segment starts (+ per-cloud output offsets)
//   centroid_kernel: one thread per voxel sums its points sequentially in input
//                   order in fp64 (the order Open3D's AccumulatedPoint sees), so
//                   centroids are reproducible bit for bit.
//
// Algorithmic bytes per launch set: 16 N (read) + 16 N' (centroids)
// [+ 12 N' keys]; the sort moves 12 B/element/pass on top of that.
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kBlock = 256;
constexpr int kSegBlocks = kSMs * 4;

__device__ __forceinline__ void cloud_range(const int* __restrict__ offs, const int* __restrict__ n_dev, int n_max,
                                            int cloud, int& b, int& e) {
    if (offs) {
        b = offs[cloud];
        e = offs[cloud + 1];
    } else {
        b = 0;
        e = n_dev ? min(*n_dev, n_max) : n_max;
    }
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// bbox[cloud*6 + 0..2] = min over order_encode(p), [3..5] = min over ~order_encode(p) (i.e. max)
__global__ void __launch_bounds__(kBlock)
bbox_kernel(const float4* __restrict__ pts, const int* __restrict__ offs, const int* __restrict__ n_dev, int n_max,
            const double* __restrict__ T, unsigned long long* __restrict__ bbox) {
    pdl_enter();
    const int cloud = blockIdx.y;
    int b, e;
    cloud_range(offs, n_dev, n_max, cloud, b, e);
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    double Tl[12];
    if (T)
        for (int k = 0; k < 12; ++k) Tl[k] = T[k];
    for (int i = b + blockIdx.x * kBlock + threadIdx.x; i < e; i += gridDim.x * kBlock) {
        float4 p = __ldg(pts + i);
        D3 q{(double)p.x, (double)p.y, (double)p.z};
        if (T) q = xform_strict(Tl, q.x, q.y, q.z);
        mn[0] = fmin(mn[0], q.x); mn[1] = fmin(mn[1], q.y); mn[2] = fmin(mn[2], q.z);
        mx[0] = fmax(mx[0], q.x); mx[1] = fmax(mx[1], q.y); mx[2] = fmax(mx[2], q.z);
        // fmin / fmax skip NaN: a point without a voxel must not slip through — open the box to infinity, which
        // prep_kernel reports as B2_ERR_KEY_RANGE (Open3D's int32 voxel index is undefined for it as well)
        if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) mn[0] = -INFINITY;
    }
    __shared__ double smn[kBlock / 32][3], smx[kBlock / 32][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double a = warp_min(mn[d]), c = warp_max(mx[d]);
        if (lane == 0) {
            smn[w][d] = a;
            smx[w][d] = c;
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = INFINITY, c = -INFINITY;
        for (int k = 0; k < kBlock / 32; ++k) {
            a = fmin(a, smn[k][threadIdx.x]);
            c = fmax(c, smx[k][threadIdx.x]);
        }
        if (a <= c) {  // block saw at least one point
            atomicMin(&bbox[cloud * 6 + threadIdx.x], order_encode(a));
            atomicMin(&bbox[cloud * 6 + 3 + threadIdx.x], ~order_encode(c));
        }
    }
}

__device__ __forceinline__ int ceil_log2(unsigned v) {  // bits needed to hold values 0..v
    return v == 0 ? 0 : 32 - __clz(v);
}

// One CTA.  origin_mode 0: min_bound - v/2 (Open3D), 1: explicit, 2: min_bound.
__global__ void __launch_bounds__(kBlock)
prep_kernel(const unsigned long long* __restrict__ bbox, const int* __restrict__ offs, const int* __restrict__ n_dev,
            int n_max, int n_clouds, double voxel, int origin_mode, double ox, double oy, double oz,
            int max_passes, VoxelPrep* __restrict__ prep, SortCtl* __restrict__ ctl, int* __restrict__ dev_status) {
    pdl_enter();
    __shared__ int sbits[3];
    __shared__ int sstatus;
    if (threadIdx.x < 3) sbits[threadIdx.x] = 0;
    if (threadIdx.x == 0) sstatus = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < n_clouds; c += kBlock) {
        int b, e;
        cloud_range(offs, n_dev, n_max, c, b, e);
        VoxelPrep vp;
        const double oh[3] = {ox, oy, oz};
        for (int d = 0; d < 3; ++d) {
            double mn = 0.0, mx = 0.0;
            if (e > b) {
                mn = order_decode(bbox[c * 6 + d]);
                mx = order_decode(~bbox[c * 6 + 3 + d]);
            }
            vp.minb[d] = mn;
            vp.maxb[d] = mx;
            double o = origin_mode == 1 ? oh[d] : (origin_mode == 0 ? dsub(mn, dmul(voxel, 0.5)) : mn);
            vp.origin[d] = o;
            double fmn = floor(ddiv(dsub(mn, o), voxel)), fmx = floor(ddiv(dsub(mx, o), voxel));
            if (!(fabs(fmn) < 2147483000.0) || !(fabs(fmx) < 2147483000.0)) {  // Open3D: int32 voxel index
                atomicExch(&sstatus, B2_ERR_KEY_RANGE);
                fmn = fmx = 0.0;
            }
            vp.kmin[d] = (int)fmn;
            unsigned ext = (unsigned)((long long)fmx - (long long)fmn);
            atomicMax(&sbits[d], ceil_log2(ext));
        }
        vp.pad = 0;
        prep[c] = vp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = offs ? offs[n_clouds] : (n_dev ? min(*n_dev, n_max) : n_max);
        int cb = ceil_log2((unsigned)(n_clouds - 1));
        int total = sbits[0] + sbits[1] + sbits[2] + cb;
        int np = (total + 7) / 8;
        if (np < 1) np = 1;
        int st = sstatus;
        if (total > 64 || np > max_passes) st = B2_ERR_KEY_RANGE;
        if (st != 0) {
            n = 0;
            np = 0;
            atomicCAS(dev_status, 0, st);
        }
        ctl->n = n;
        ctl->npasses = np;
        ctl->bits[0] = sbits[0];
        ctl->bits[1] = sbits[1];
        ctl->bits[2] = sbits[2];
        ctl->cloud_bits = cb;
        ctl->nseg = 0;
        ctl->status = st;
    }
}

__device__ __forceinline__ int find_cloud(const int* __restrict__ offs, int n_clouds, int i) {
    int lo = 0, hi = n_clouds;  // offs[lo] <= i < offs[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kBlock)
keys_kernel(const float4* __restrict__ pts, const int* __restrict__ offs, int n_clouds, double voxel,
            const double* __restrict__ T, const VoxelPrep* __restrict__ prep, const SortCtl* __restrict__ ctl,
            unsigned long long* __restrict__ keys) {
    pdl_enter();
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= ctl->n) return;
    const int cloud = (n_clouds > 1) ? find_cloud(offs, n_clouds, i) : 0;
    float4 p = __ldg(pts + i);
    D3 q{(double)p.x, (double)p.y, (double)p.z};
    if (T) q = xform_strict(T, q.x, q.y, q.z);
    const VoxelPrep& vp = prep[cloud];
    unsigned kx = (unsigned)(voxel_coord(q.x, vp.origin[0], voxel) - vp.kmin[0]);
    unsigned ky = (unsigned)(voxel_coord(q.y, vp.origin[1], voxel) - vp.kmin[1]);
    unsigned kz = (unsigned)(voxel_coord(q.z, vp.origin[2], voxel) - vp.kmin[2]);
    const int by = ctl->bits[1], bz = ctl->bits[2], bx = ctl->bits[0];
    unsigned long long k = (unsigned long long)cloud;
    k = (k << bx) | kx;
    k = (k << by) | ky;
    k = (k << bz) | kz;
    keys[i] = k;
}

__device__ __forceinline__ void seg_block_range(int n, int& begin, int& end) {
    int per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + kBlock - 1) / kBlock * kBlock;
    long long b = (long long)blockIdx.x * per;
    begin = (int)(b < n ? b : n);
    end = (int)(b + per < n ? b + per : n);
}

__global__ void __launch_bounds__(kBlock)
seg_count_kernel(const unsigned long long* __restrict__ keysA, const unsigned long long* __restrict__ keysB,
                 const SortCtl* __restrict__ ctl, int* __restrict__ blockcnt) {
    pdl_enter();
    const unsigned long long* keys = (ctl->npasses & 1) ? keysB : keysA;
    int begin, end;
    seg_block_range(ctl->n, begin, end);
    int cnt = 0;
    for (int i = begin + threadIdx.x; i < end; i += kBlock) cnt += (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
    __shared__ int s[kBlock / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kBlock / 32; ++k) t += s[k];
        blockcnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kBlock)
seg_write_kernel(const unsigned long long* __restrict__ keysA, const unsigned long long* __restrict__ keysB,
                 SortCtl* __restrict__ ctl, const int* __restrict__ blockcnt, int* __restrict__ seg_start,
                 int* __restrict__ cloud_out_offs, int n_clouds) {
    pdl_enter();
    const unsigned long long* keys = (ctl->npasses & 1) ? keysB : keysA;
    const int n = ctl->n;
    const int shift = ctl->bits[0] + ctl->bits[1] + ctl->bits[2];
    __shared__ int sbase;
    __shared__ int swarp[kBlock / 32];
    // exclusive prefix of the per-block head counts (<= 592 entries: one warp)
    if (threadIdx.x < 32) {
        int acc = 0, tot = 0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) {
            int v = blockcnt[k];
            if (k < (int)blockIdx.x) acc += v;
            tot += v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        if (threadIdx.x == 0) {
            sbase = acc;
            if (blockIdx.x == 0) {
                ctl->nseg = tot;
                if (cloud_out_offs) {
                    if (n == 0) {
                        for (int c = 0; c <= n_clouds; ++c) cloud_out_offs[c] = 0;
                    } else {
                        int last_cloud = n_clouds > 1 ? (int)(keys[n - 1] >> shift) : 0;
                        for (int c = last_cloud + 1; c <= n_clouds; ++c) cloud_out_offs[c] = tot;
                    }
                }
            }
        }
    }
    __syncthreads();
    int begin, end;
    seg_block_range(n, begin, end);
    int running = sbase;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = begin; base < end; base += kBlock) {
        const int i = base + threadIdx.x;
        bool head = false;
        unsigned long long k = 0, kp = 0;
        if (i < end) {
            k = keys[i];
            if (i == 0) head = true;
            else {
                kp = keys[i - 1];
                head = k != kp;
            }
        }
        unsigned bal = __ballot_sync(0xffffffffu, head);
        int wrank = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) swarp[w] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int q = 0; q < kBlock / 32; ++q) {
            int v = swarp[q];
            if (q < w) woff += v;
            tot += v;
        }
        if (head) {
            int rank = running + woff + wrank;
            seg_start[rank] = i;
            if (cloud_out_offs) {
                int c = n_clouds > 1 ? (int)(k >> shift) : 0;
                int cprev = (i == 0) ? -1 : (n_clouds > 1 ? (int)(kp >> shift) : 0);
                for (int cc = cprev + 1; cc <= c; ++cc) cloud_out_offs[cc] = rank;
            }
        }
        running += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock)
centroid_kernel(const float4* __restrict__ pts, const int* __restrict__ offs, int n_clouds,
                const unsigned long long* __restrict__ keysA, const unsigned long long* __restrict__ keysB,
                const unsigned* __restrict__ valsA, const unsigned* __restrict__ valsB,
                const SortCtl* __restrict__ ctl, const VoxelPrep* __restrict__ prep,
                const int* __restrict__ seg_start, const double* __restrict__ T, float4* __restrict__ out_pts,
                int* __restrict__ out_keys, double* __restrict__ out_sums) {
    pdl_enter();
    const int j = blockIdx.x * kBlock + threadIdx.x;
    const int nseg = ctl->nseg;
    if (j >= nseg) return;
    const bool odd = ctl->npasses & 1;
    const unsigned long long* keys = odd ? keysB : keysA;
    const unsigned* vals = odd ? valsB : valsA;
    const int b = seg_start[j];
    const int e = (j + 1 < nseg) ? seg_start[j + 1] : ctl->n;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int k = b; k < e; ++k) {
        float4 p = __ldg(pts + vals[k]);
        D3 q{(double)p.x, (double)p.y, (double)p.z};
        if (T) q = xform_strict(T, q.x, q.y, q.z);
        sx = dadd(sx, q.x);
        sy = dadd(sy, q.y);
        sz = dadd(sz, q.z);
    }
    const int cnt = e - b;
    const double dn = (double)cnt;
    float4 o;
    o.x = (float)ddiv(sx, dn);
    o.y = (float)ddiv(sy, dn);
    o.z = (float)ddiv(sz, dn);
    o.w = __int_as_float(cnt);
    out_pts[j] = o;
    if (out_keys || out_sums) {
        unsigned long long k = keys[b];
        const int bx = ctl->bits[0], by = ctl->bits[1], bz = ctl->bits[2];
        unsigned kz = (unsigned)(k & ((1ull << bz) - 1ull));
        k >>= bz;
        unsigned ky = (unsigned)(k & ((1ull << by) - 1ull));
        k >>= by;
        unsigned kx = (unsigned)(k & ((1ull << bx) - 1ull));
        k >>= bx;
        const VoxelPrep& vp = prep[n_clouds > 1 ? (int)k : 0];
        if (out_keys) {
            out_keys[3 * j + 0] = (int)kx + vp.kmin[0];
            out_keys[3 * j + 1] = (int)ky + vp.kmin[1];
            out_keys[3 * j + 2] = (int)kz + vp.kmin[2];
        }
        if (out_sums) {
            out_sums[4 * j + 0] = sx;
            out_sums[4 * j + 1] = sy;
            out_sums[4 * j + 2] = sz;
            out_sums[4 * j + 3] = dn;
        }
    }
}

}  // namespace

int voxel_sort(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets, int n_clouds,
               double voxel, int origin_mode, const double* origin_host, const double* T_dev, int* cloud_out_offs,
               VoxelSortOut* out) {
    if (n_max <= 0 || n_clouds <= 0) return set_error(c, B2_ERR_INVALID, "voxel_sort: empty input");
#ifdef B2_EXPERIMENTS
    if (!cloud_offsets && !cloud_out_offs && voxel_small_applicable(n_max, n_clouds))
        return voxel_small(c, pts, n_max, n_dev, voxel, origin_mode, origin_host, T_dev, false, DownsampleOut{}, out);
#endif
    if (!(voxel > 0.0)) return set_error(c, B2_ERR_INVALID, "voxel size must be > 0 (got %g)", voxel);
    B2_WS(c, keysA, unsigned long long, WS_KEYS_A, sizeof(unsigned long long) * (size_t)n_max);
    B2_WS(c, keysB, unsigned long long, WS_KEYS_B, sizeof(unsigned long long) * (size_t)n_max);
    B2_WS(c, valsA, unsigned, WS_VALS_A, sizeof(unsigned) * (size_t)n_max);
    B2_WS(c, valsB, unsigned, WS_VALS_B, sizeof(unsigned) * (size_t)n_max);
    B2_WS(c, ctl, SortCtl, WS_CTL, sizeof(SortCtl));
    B2_WS(c, seg_start, int, WS_SEGSTART, sizeof(int) * (size_t)(n_max + 1));
    B2_WS(c, blockcnt, int, WS_BLOCKCNT, sizeof(int) * kSegBlocks);
    B2_WS(c, bbox, unsigned long long, WS_BBOX, sizeof(unsigned long long) * 6 * (size_t)n_clouds);
    B2_WS(c, prep, VoxelPrep, WS_PREP, sizeof(VoxelPrep) * (size_t)n_clouds);

    B2_CUDA_OK(c, cudaMemsetAsync(bbox, 0xFF, sizeof(unsigned long long) * 6 * (size_t)n_clouds, c->stream));
    int per_cloud = n_max / n_clouds + 1;
    dim3 bgrid(max(1, min(div_up(per_cloud, kBlock * 4), n_clouds > 1 ? 8 : kSMs * 2)), n_clouds);
    pdl_launch(bbox_kernel, bgrid, dim3(kBlock), c->stream, pts, cloud_offsets, n_dev, n_max, T_dev, bbox);
    B2_LAUNCH_CHECK(c);
    double o[3] = {0, 0, 0};
    if (origin_mode == 1) {
        if (!origin_host) return set_error(c, B2_ERR_INVALID, "explicit origin requested but origin is NULL");
        o[0] = origin_host[0]; o[1] = origin_host[1]; o[2] = origin_host[2];
    }
    pdl_launch(prep_kernel, dim3(1), dim3(kBlock), c->stream, bbox, cloud_offsets, n_dev, n_max, n_clouds, voxel,
               origin_mode, o[0], o[1], o[2], c->sort_passes_hint, prep, ctl, c->dev_status);
    B2_LAUNCH_CHECK(c);
    pdl_launch(keys_kernel, dim3(div_up(n_max, kBlock)), dim3(kBlock), c->stream, pts, cloud_offsets, n_clouds, voxel,
               T_dev, (const VoxelPrep*)prep, (const SortCtl*)ctl, keysA);
    B2_LAUNCH_CHECK(c);
    B2_TRY(radix_sort_pairs(c, keysA, keysB, valsA, valsB, ctl, n_max, c->sort_passes_hint));
    int sblocks = min(kSegBlocks, div_up(n_max, kBlock));
    pdl_launch(seg_count_kernel, dim3(sblocks), dim3(kBlock), c->stream, (const unsigned long long*)keysA,
               (const unsigned long long*)keysB, (const SortCtl*)ctl, blockcnt);
    B2_LAUNCH_CHECK(c);
    pdl_launch(seg_write_kernel, dim3(sblocks), dim3(kBlock), c->stream, (const unsigned long long*)keysA,
               (const unsigned long long*)keysB, ctl, (const int*)blockcnt, seg_start, cloud_out_offs, n_clouds);
    B2_LAUNCH_CHECK(c);
    out->keysA = keysA; out->keysB = keysB; out->valsA = valsA; out->valsB = valsB;
    out->ctl = ctl; out->prep = prep; out->seg_start = seg_start;
    return B2_OK;
}

int voxel_downsample(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets_in,
                     int n_clouds, double voxel, int origin_mode, const double* origin_host, const double* T_dev,
                     DownsampleOut out, VoxelSortOut* vs_out) {
#ifdef B2_EXPERIMENTS
    if (!cloud_offsets_in && !out.cloud_offsets && voxel_small_applicable(n_max, n_clouds))
        return voxel_small(c, pts, n_max, n_dev, voxel, origin_mode, origin_host, T_dev, true, out, vs_out);
#endif
    VoxelSortOut vs;
    B2_TRY(voxel_sort(c, pts, n_max, n_dev, cloud_offsets_in, n_clouds, voxel, origin_mode, origin_host, T_dev,
                      out.cloud_offsets, &vs));
    pdl_launch(centroid_kernel, dim3(div_up(n_max, kBlock)), dim3(kBlock), c->stream, pts, cloud_offsets_in, n_clouds,
               (const unsigned long long*)vs.keysA, (const unsigned long long*)vs.keysB, (const unsigned*)vs.valsA,
               (const unsigned*)vs.valsB, (const SortCtl*)vs.ctl, (const VoxelPrep*)vs.prep, (const int*)vs.seg_start,
               T_dev, out.pts, out.keys, out.sums);
    B2_LAUNCH_CHECK(c);
    if (vs_out) *vs_out = vs;
    return B2_OK;
}

}  // namespace b2
