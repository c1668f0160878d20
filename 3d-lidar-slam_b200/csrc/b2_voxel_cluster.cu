// b2_voxel_cluster.cu — the whole voxel sort of ONE scan-sized cloud in a single launch.
//
// A 49k-point scan is far too small to amortise the ~15 kernel launches of the general
// pipeline in b2_voxel.cu (bound box, prep, keys, 4 x (histogram, scatter), segmentation,
// centroids): each of those is a few microseconds of launch latency around a microsecond
// of work.  Here one thread-block CLUSTER of 8 CTAs x 1024 threads owns the cloud
// (<= 65,536 points, 8,192 per CTA) and walks through the same phases separated by
// hardware cluster barriers (barrier.cluster, ~0.2 us) instead of kernel boundaries:
//
//   bound box      per-CTA fp64 min/max  -> exchanged through distributed shared memory
//   prep           every CTA derives origin / key range / bit widths / pass count redundantly
//   keys           packed [x|y|z] key per point (fp64 IEEE division, identical to b2_voxel.cu)
//   radix passes   per-CTA digit histogram in shared memory; after the barrier every CTA reads
//                  the eight histograms over DSMEM and derives its own scatter bases — no
//                  histogram table in global memory, no scan kernel; stable warp match-any
//                  ranking; ping-pong key/value buffers stay in L2
//   segmentation   head flags of the sorted keys, per-CTA counts exchanged over DSMEM
//   centroids      one thread per voxel, fp64 sums in input order (bit-identical results)
//
// Outputs are exactly those of voxel_sort()/voxel_downsample(): VoxelSortOut, SortCtl,
// VoxelPrep, seg_start, centroids / keys / sums, so callers cannot tell the paths apart.
#include <cooperative_groups.h>

#include <cstdlib>

#include "b2_common.cuh"

namespace cg = cooperative_groups;

namespace b2 {

namespace {

constexpr int kCtas = 8;
constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 8192;              // elements per CTA
constexpr int kItems = kTile / kThreads;  // 8
constexpr unsigned kFull = 0xffffffffu;

struct SmallArgs {
    const float4* pts;
    const int* n_dev;
    int n_max;
    double voxel;
    int origin_mode;  // 0: min - v/2, 1: explicit, 2: min
    double ox, oy, oz;
    const double* T;
    unsigned long long *keysA, *keysB;
    unsigned *valsA, *valsB;
    SortCtl* ctl;
    VoxelPrep* prep;
    int* seg_start;
    int* dev_status;
    int do_centroid;
    float4* out_pts;
    int* out_keys;
    double* out_sums;
};

__device__ __forceinline__ int ceil_log2u(unsigned v) { return v == 0 ? 0 : 32 - __clz(v); }

__global__ void __cluster_dims__(kCtas, 1, 1) __launch_bounds__(kThreads, 1)
voxel_small_kernel(SmallArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    __shared__ double s_bb[6];                 // this CTA's min xyz, max xyz
    __shared__ double s_wmn[kWarps][3], s_wmx[kWarps][3];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_running[256];
    __shared__ unsigned s_scan[kWarps];
    __shared__ unsigned short s_wcount[kWarps][256];
    __shared__ int s_cnt;                      // heads in this CTA's tile
    __shared__ int s_bits[3], s_kmin[3], s_np, s_status;
    __shared__ double s_origin[3];
    __shared__ double s_T[12];

    const int n = a.n_dev ? min(*a.n_dev, a.n_max) : a.n_max;
    const int tile_b = min(n, (int)rank * kTile), tile_e = min(n, tile_b + kTile);
    if (a.T && tid < 12) s_T[tid] = a.T[tid];
    if (tid == 0) s_status = 0;
    __syncthreads();

    // ---- phase 0: bound box -------------------------------------------------------------------
    float4 p[kItems];
    {
        double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const int i = tile_b + k * kThreads + tid;
            if (i < tile_e) {
                p[k] = __ldg(a.pts + i);
                D3 q{(double)p[k].x, (double)p[k].y, (double)p[k].z};
                if (a.T) q = xform_strict(s_T, q.x, q.y, q.z);
                mn[0] = fmin(mn[0], q.x); mn[1] = fmin(mn[1], q.y); mn[2] = fmin(mn[2], q.z);
                mx[0] = fmax(mx[0], q.x); mx[1] = fmax(mx[1], q.y); mx[2] = fmax(mx[2], q.z);
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                mn[d] = fmin(mn[d], __shfl_xor_sync(kFull, mn[d], o));
                mx[d] = fmax(mx[d], __shfl_xor_sync(kFull, mx[d], o));
            }
            if (lane == 0) {
                s_wmn[w][d] = mn[d];
                s_wmx[w][d] = mx[d];
            }
        }
        __syncthreads();
        if (tid < 3) {
            double m0 = INFINITY, m1 = -INFINITY;
            for (int k = 0; k < kWarps; ++k) {
                m0 = fmin(m0, s_wmn[k][tid]);
                m1 = fmax(m1, s_wmx[k][tid]);
            }
            s_bb[tid] = m0;
            s_bb[3 + tid] = m1;
        }
    }
    cluster.sync();
    // ---- prep (redundant in every CTA; rank 0 publishes) ------------------------------------------
    if (tid < 3) {
        double mn = INFINITY, mx = -INFINITY;
        for (unsigned r = 0; r < kCtas; ++r) {
            const double* rb = cluster.map_shared_rank(s_bb, r);
            mn = fmin(mn, rb[tid]);
            mx = fmax(mx, rb[3 + tid]);
        }
        if (n == 0) { mn = 0.0; mx = 0.0; }
        const double oh = tid == 0 ? a.ox : (tid == 1 ? a.oy : a.oz);
        const double o = a.origin_mode == 1 ? oh : (a.origin_mode == 0 ? dsub(mn, dmul(a.voxel, 0.5)) : mn);
        double fmn = floor(ddiv(dsub(mn, o), a.voxel)), fmx = floor(ddiv(dsub(mx, o), a.voxel));
        int bad = 0;
        if (!(fabs(fmn) < 2147483000.0) || !(fabs(fmx) < 2147483000.0)) {
            bad = 1;
            fmn = fmx = 0.0;
        }
        s_origin[tid] = o;
        s_kmin[tid] = (int)fmn;
        s_bits[tid] = ceil_log2u((unsigned)((long long)fmx - (long long)fmn));
        if (rank == 0) {
            a.prep->origin[tid] = o;
            a.prep->minb[tid] = mn;
            a.prep->maxb[tid] = mx;
            a.prep->kmin[tid] = (int)fmn;
        }
        if (bad) atomicExch(&s_status, B2_ERR_KEY_RANGE);
    }
    __syncthreads();
    if (tid == 0) {
        const int total = s_bits[0] + s_bits[1] + s_bits[2];
        int np = (total + 7) / 8;
        if (np < 1) np = 1;
        int st = s_status;
        if (total > 64) st = B2_ERR_KEY_RANGE;
        if (st != 0) np = 0;
        s_np = np;
        s_status = st;
        if (rank == 0) {
            if (st != 0) atomicCAS(a.dev_status, 0, st);
            a.ctl->n = st != 0 ? 0 : n;
            a.ctl->npasses = np;
            a.ctl->bits[0] = s_bits[0];
            a.ctl->bits[1] = s_bits[1];
            a.ctl->bits[2] = s_bits[2];
            a.ctl->cloud_bits = 0;
            a.ctl->nseg = 0;
            a.ctl->status = st;
            a.prep->pad = 0;
        }
    }
    __syncthreads();
    const int np = s_np;
    const int live_b = s_status != 0 ? 0 : tile_b, live_e = s_status != 0 ? 0 : tile_e;
    const int n_live = s_status != 0 ? 0 : n;
    const int bx = s_bits[0], by = s_bits[1], bz = s_bits[2];

    // ---- keys (own tile, blocked: warp w owns elements [w*256, w*256+256) of the tile) ----------------
    // re-read in the blocked order the scatter uses (the strided registers above only fed the bound box)
    unsigned long long key[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int i = live_b + w * (32 * kItems) + r * 32 + lane;
        key[r] = 0ull;
        if (i < live_e) {
            const float4 q4 = __ldg(a.pts + i);
            D3 q{(double)q4.x, (double)q4.y, (double)q4.z};
            if (a.T) q = xform_strict(s_T, q.x, q.y, q.z);
            const unsigned kx = (unsigned)(voxel_coord(q.x, s_origin[0], a.voxel) - s_kmin[0]);
            const unsigned ky = (unsigned)(voxel_coord(q.y, s_origin[1], a.voxel) - s_kmin[1]);
            const unsigned kz = (unsigned)(voxel_coord(q.z, s_origin[2], a.voxel) - s_kmin[2]);
            unsigned long long k = kx;
            k = (k << by) | ky;
            k = (k << bz) | kz;
            key[r] = k;
        }
    }
    (void)bx;

    // ---- radix passes ------------------------------------------------------------------------------
    for (int pass = 0; pass < np; ++pass) {
        const int shift = 8 * pass;
        const bool odd = pass & 1;
        const unsigned long long* kin = odd ? a.keysB : a.keysA;
        const unsigned* vin = odd ? a.valsB : a.valsA;
        unsigned long long* kout = odd ? a.keysA : a.keysB;
        unsigned* vout = odd ? a.valsA : a.valsB;
        unsigned val[kItems];
        if (tid < 256) s_hist[tid] = 0;
        for (int k = 0; k < kWarps; k += 1) s_wcount[k][tid & 255] = 0;  // (each of the 4 thread quarters clears all rows)
        __syncthreads();
        unsigned lrank[kItems];
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = live_b + w * (32 * kItems) + r * 32 + lane;
            const bool valid = i < live_e;
            if (pass > 0) {
                key[r] = valid ? kin[i] : 0ull;
                val[r] = valid ? vin[i] : 0u;
            } else {
                val[r] = (unsigned)i;
            }
            const unsigned d = (unsigned)(key[r] >> shift) & 255u;
            const unsigned active = __ballot_sync(kFull, valid);
            lrank[r] = 0;
            if (valid) {
                const unsigned peers = __match_any_sync(active, d);
                const int leader = __ffs(peers) - 1;
                unsigned old = 0;
                if (lane == leader) {
                    old = s_wcount[w][d];
                    s_wcount[w][d] = (unsigned short)(old + __popc(peers));
                }
                old = __shfl_sync(peers, old, leader);
                lrank[r] = old + __popc(peers & ((1u << lane) - 1u));
            }
            __syncwarp();
        }
        __syncthreads();
        if (tid < 256) {  // exclusive scan over the warps of this CTA; CTA histogram
            unsigned acc = 0;
#pragma unroll 8
            for (int k = 0; k < kWarps; ++k) {
                const unsigned t = s_wcount[k][tid];
                s_wcount[k][tid] = (unsigned short)acc;
                acc += t;
            }
            s_hist[tid] = acc;
        }
        cluster.sync();  // every CTA's histogram is complete (and the previous pass's scatter has landed)
        if (tid < 256) {
            unsigned below = 0, total = 0;
            for (unsigned r = 0; r < kCtas; ++r) {
                const unsigned v = cluster.map_shared_rank(s_hist, r)[tid];
                total += v;
                if (r < rank) below += v;
            }
            unsigned incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_scan[w] = incl;
            s_running[tid] = incl - total + below;  // completed below once the warp totals are known
        }
        __syncthreads();
        if (tid < 256) {
            unsigned woff = 0;
            for (int k = 0; k < 8; ++k)
                if (k < w) woff += s_scan[k];
            s_running[tid] += woff;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = live_b + w * (32 * kItems) + r * 32 + lane;
            if (i < live_e) {
                const unsigned d = (unsigned)(key[r] >> shift) & 255u;
                const unsigned dst = s_running[d] + s_wcount[w][d] + lrank[r];
                kout[dst] = key[r];
                vout[dst] = val[r];
            }
        }
        cluster.sync();  // scatter visible cluster-wide; histograms free to be reused
    }
    const bool res_odd = np & 1;
    const unsigned long long* skeys = res_odd ? a.keysB : a.keysA;
    const unsigned* svals = res_odd ? a.valsB : a.valsA;

    // ---- segmentation ------------------------------------------------------------------------------
    // thread t owns elements [tile_b + 8t, +8) of the sorted order
    int flags = 0, cnt = 0;
    {
        const int i0 = live_b + tid * kItems;
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const int i = i0 + k;
            if (i < live_e) {
                const bool head = i == 0 || skeys[i] != skeys[i - 1];
                flags |= head ? (1 << k) : 0;
                cnt += head ? 1 : 0;
            }
        }
    }
    unsigned incl = (unsigned)cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[w] = incl;
    __syncthreads();
    unsigned woff = 0, ctot = 0;
    for (int k = 0; k < kWarps; ++k) {
        const unsigned v = s_scan[k];
        if (k < w) woff += v;
        ctot += v;
    }
    if (tid == 0) s_cnt = (int)ctot;
    cluster.sync();
    int base = 0, nseg = 0;
    for (unsigned r = 0; r < kCtas; ++r) {
        const int v = *cluster.map_shared_rank(&s_cnt, r);
        if (r < rank) base += v;
        nseg += v;
    }
    {
        int pos = base + (int)(woff + incl) - cnt;
        const int i0 = live_b + tid * kItems;
#pragma unroll
        for (int k = 0; k < kItems; ++k)
            if (flags & (1 << k)) a.seg_start[pos++] = i0 + k;
    }
    if (rank == 0 && tid == 0) a.ctl->nseg = nseg;
    cluster.sync();  // seg_start complete (and nobody reads a remote s_cnt after this point)
    if (!a.do_centroid) return;

    // ---- centroids: one thread per voxel, sequential fp64 sums in input order ------------------------
    for (int j = (int)rank * kThreads + tid; j < nseg; j += kCtas * kThreads) {
        const int b = a.seg_start[j];
        const int e = (j + 1 < nseg) ? a.seg_start[j + 1] : n_live;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int k = b; k < e; ++k) {
            const float4 q4 = __ldg(a.pts + svals[k]);
            D3 q{(double)q4.x, (double)q4.y, (double)q4.z};
            if (a.T) q = xform_strict(s_T, q.x, q.y, q.z);
            sx = dadd(sx, q.x);
            sy = dadd(sy, q.y);
            sz = dadd(sz, q.z);
        }
        const int c = e - b;
        const double dn = (double)c;
        a.out_pts[j] = make_float4((float)ddiv(sx, dn), (float)ddiv(sy, dn), (float)ddiv(sz, dn), __int_as_float(c));
        if (a.out_keys || a.out_sums) {
            unsigned long long k = skeys[b];
            const unsigned kz = (unsigned)(k & ((1ull << bz) - 1ull));
            k >>= bz;
            const unsigned ky = (unsigned)(k & ((1ull << by) - 1ull));
            k >>= by;
            const unsigned kx = (unsigned)k;
            if (a.out_keys) {
                a.out_keys[3 * j + 0] = (int)kx + s_kmin[0];
                a.out_keys[3 * j + 1] = (int)ky + s_kmin[1];
                a.out_keys[3 * j + 2] = (int)kz + s_kmin[2];
            }
            if (a.out_sums) {
                a.out_sums[4 * j + 0] = sx;
                a.out_sums[4 * j + 1] = sy;
                a.out_sums[4 * j + 2] = sz;
                a.out_sums[4 * j + 3] = dn;
            }
        }
    }
}

}  // namespace

// Opt-in (B2_CLUSTER_VOXEL=1).  Measured on B200 for a 49k-point scan: 97 us in this single launch
// versus ~50 us for the ~15 launches of b2_voxel.cu inside the step's CUDA graph — the phase chain is
// 1.2 M warp instructions, which eight SMs retire more slowly than 148 SMs retire them plus the launch
// gaps.  Kept because it is parity-tested and is the starting point for a 148-SM cooperative variant.
bool voxel_small_applicable(int n_max, int n_clouds) {
    static const bool on = getenv("B2_CLUSTER_VOXEL") != nullptr;
    return on && n_clouds == 1 && n_max > 0 && n_max <= kCtas * kTile;
}

int voxel_small(Context* c, const float4* pts, int n_max, const int* n_dev, double voxel, int origin_mode,
                const double* origin_host, const double* T_dev, bool do_centroid, DownsampleOut out,
                VoxelSortOut* vs_out) {
    if (!(voxel > 0.0)) return set_error(c, B2_ERR_INVALID, "voxel size must be > 0 (got %g)", voxel);
    B2_WS(c, keysA, unsigned long long, WS_KEYS_A, sizeof(unsigned long long) * (size_t)n_max);
    B2_WS(c, keysB, unsigned long long, WS_KEYS_B, sizeof(unsigned long long) * (size_t)n_max);
    B2_WS(c, valsA, unsigned, WS_VALS_A, sizeof(unsigned) * (size_t)n_max);
    B2_WS(c, valsB, unsigned, WS_VALS_B, sizeof(unsigned) * (size_t)n_max);
    B2_WS(c, ctl, SortCtl, WS_CTL, sizeof(SortCtl));
    B2_WS(c, seg_start, int, WS_SEGSTART, sizeof(int) * (size_t)(n_max + 1));
    B2_WS(c, prep, VoxelPrep, WS_PREP, sizeof(VoxelPrep));
    SmallArgs a;
    a.pts = pts; a.n_dev = n_dev; a.n_max = n_max; a.voxel = voxel; a.origin_mode = origin_mode;
    a.ox = a.oy = a.oz = 0.0;
    if (origin_mode == 1) {
        if (!origin_host) return set_error(c, B2_ERR_INVALID, "explicit origin requested but origin is NULL");
        a.ox = origin_host[0]; a.oy = origin_host[1]; a.oz = origin_host[2];
    }
    a.T = T_dev;
    a.keysA = keysA; a.keysB = keysB; a.valsA = valsA; a.valsB = valsB;
    a.ctl = ctl; a.prep = prep; a.seg_start = seg_start; a.dev_status = c->dev_status;
    a.do_centroid = do_centroid ? 1 : 0;
    a.out_pts = out.pts; a.out_keys = out.keys; a.out_sums = out.sums;
    voxel_small_kernel<<<kCtas, kThreads, 0, c->stream>>>(a);
    B2_LAUNCH_CHECK(c);
    if (vs_out) {
        vs_out->keysA = keysA; vs_out->keysB = keysB; vs_out->valsA = valsA; vs_out->valsB = valsB;
        vs_out->ctl = ctl; vs_out->prep = prep; vs_out->seg_start = seg_start;
    }
    return B2_OK;
}

}  // namespace b2
