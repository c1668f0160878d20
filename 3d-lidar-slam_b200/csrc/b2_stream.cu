// b2_stream.cu — the sort-free voxel passes of the streaming scan step (b2_slam_scan).
//
// The reference runs, per scan, voxel_down_sample(scan, 0.02) in front of registration_icp and folds the
// registered scan into the map (icp_processor.py:73-74, 90).  b2_voxel.cu does both with a
// radix sort (bbox, prep, keys, 3-4 x (hist, scatter), seg_count, seg_write, centroid: 13-15 launches each),
// which is what a scan step spent a third of its time in once the ICP loop had been shortened: 43 launches of
// ~2.5 us for 49k points.  A single scan needs no sort: its few thousand voxels fit a scratch hash
// ("scan table", 2^17 slots for 49k points) filled with atomics —
//
//   down-sample  : ds_bbox -> ds_scatter (key = floor((p - (min - v/2)) / v), insert, fp64 atomicAdd of the raw
//                  float32 coordinates, atomicMin of the point index) -> ds_emit (the voxel's FIRST point writes
//                  the centroid at its rank among the first points: one decoupled look-back scan).
//                  The sum of n float32 values of one voxel is exact in fp64 (n x 24-bit mantissas of nearly equal
//                  exponent), so the order of the atomics does not matter and the centroids are bit-identical to
//                  the oracle's in-order sums.  Output order = order of first appearance in the scan (pixel
//                  order), which keeps the queries of an ICP CTA spatially together like the key order did.
//   fuse         : fuse_scatter (P = T p, q = (P - origin) / v, key = floor(q), insert, 64-bit FIXED-POINT
//                  atomicAdd of frac(q) at 2^-40 voxel: order-independent, hence deterministic) -> fuse_collect
//                  (occupied slots -> (key, sum, count) records; sum = n * corner + v * sum frac; creates the
//                  voxel's brick if the map has none yet) -> the map's own merge kernel (b2_map.cu: one thread per
//                  unique voxel, fp64 accumulators; its last thread block also commits the registered pose).
//                  The per-scan sums differ from in-order fp64 sums by ~1e-15 relative (far inside the 1e-5
//                  centroid tolerance); keys and counts are exact.
//
// 3 + 3 launches instead of 15 + 13.  The down-sample half runs on the library's FRONT stream: it depends neither on
// the map nor on the previous pose, so the down-sample of scan k+1 overlaps the registration of scan k.
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kBlock = 256;
// The down-sample kernels run on the front stream NEXT TO the ICP chain of the previous scan, whose one CTA per SM
// holds 640 x 96 of the SM's 65,536 registers: a CTA of 128 threads x 32 registers is exactly what still fits.  Two of
// them would keep the next ICP CTA out until one retires, so each asks for 100 KB of (unused) dynamic shared memory:
// with the ICP CTA's 32 KB at most one fits, two when the SM is otherwise empty.
constexpr int kFront = 128;
constexpr int kEmitBlock = kFront;
constexpr size_t kFrontSmem = 0;
constexpr double kFixScale = 1099511627776.0;        // 2^40
constexpr double kFixInv = 1.0 / 1099511627776.0;

// Scratch buffers: one per scan parity for the down-sample, [PreHead][agg: kMaxEmitBlocks u32][keys | sums (3
// doubles) | cnt | first] — the down-sample of scan k+1 runs on the front stream while scan k is still being
// registered and its status word must survive until the fusion of ITS scan has read it — and one for the fusion,
// [MainHead][keys | sums | cnt]; each is cleared by one zero-fill on the stream that uses it.
struct PreHead {
    unsigned long long bbox[6];  // max over ~order_encode(min) [0..2], max over order_encode(max) [3..5]; 0 = no point
    int status;                  // B2_ERR_* of the down-sample (handed to the sticky status word by the fusion)
    int pad[3];
};
struct MainHead {
    SortCtl ctl;                 // fuse: nseg = unique voxels of the scan (read by the merge kernels)
    int nlist;                   // occupied table slots so far (the thread that creates a slot appends it to the list)
    int pad[7];
};
constexpr int kMaxEmitBlocks = 2048;

struct ScanTable {
    unsigned long long* keys;  // packed key + 1; 0 = vacant
    unsigned long long* sums;  // [3 * cap]: doubles (down-sample) or int64 fixed point (fuse)
    unsigned* cnt;
    unsigned* first;           // ~(smallest point index)  (zero-fill friendly: atomicMax)
    unsigned mask;
};

__device__ __forceinline__ unsigned table_insert(const ScanTable& t, unsigned long long key, bool* created = nullptr) {
    unsigned h = (unsigned)hash64(key) & t.mask;
#pragma unroll 1
    for (unsigned probe = 0; probe <= t.mask; ++probe) {
        unsigned long long cur = t.keys[h];
        if (cur == 0ull) cur = atomicCAS(&t.keys[h], 0ull, key);
        if (cur == 0ull || cur == key) {
            if (created) *created = cur == 0ull;
            return h;
        }
        h = (h + 1) & t.mask;
    }
    return 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(kFront, 16)
ds_bbox_kernel(const float4* __restrict__ pts, const int* __restrict__ n_dev, int n_max, PreHead* __restrict__ head) {
    pdl_enter();
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
    for (int i = blockIdx.x * kFront + threadIdx.x; i < n; i += gridDim.x * kFront) {
        const float4 p = __ldg(pts + i);
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
        bad |= !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
    }
    if (bad) mn[0] = -INFINITY;  // (fmin skips NaN: a point without a voxel opens the box, reported by ds_scatter)
    __shared__ float smn[kFront / 32][3], smx[kFront / 32][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float a = mn[d], c = mx[d];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
            c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));
        }
        if (lane == 0) { smn[w][d] = a; smx[w][d] = c; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = INFINITY, c = -INFINITY;
        for (int k = 0; k < kFront / 32; ++k) { a = fminf(a, smn[k][threadIdx.x]); c = fmaxf(c, smx[k][threadIdx.x]); }
        if (a <= c) {
            atomicMax(&head->bbox[threadIdx.x], ~order_encode((double)a));
            atomicMax(&head->bbox[3 + threadIdx.x], order_encode((double)c));
        }
    }
}

// A.1 with Open3D's own origin (min_bound - v/2): keys are >= 0.
__global__ void __launch_bounds__(kFront, 16)
ds_scatter_kernel(const float4* __restrict__ pts, const int* __restrict__ n_dev, int n_max, double voxel,
                  PreHead* __restrict__ head, ScanTable t, unsigned* __restrict__ slot_of) {
    pdl_enter();
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int i = blockIdx.x * kFront + threadIdx.x;
    if (i >= n) return;
    double origin[3];
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double mn = order_decode(~head->bbox[d]), mx = order_decode(head->bbox[3 + d]);
        origin[d] = dsub(mn, dmul(voxel, 0.5));
        // Open3D's voxel index is an int32; the packed key holds 21 bits per axis
        ok = ok && isfinite(mn) && isfinite(mx) && floor(ddiv(dsub(mx, origin[d]), voxel)) < 2097151.0;
    }
    int* status = &head->status;
    if (!ok) {
        if (i == 0) atomicCAS(status, 0, B2_ERR_KEY_RANGE);
        return;
    }
    const float4 p = __ldg(pts + i);
    const unsigned long long kx = (unsigned long long)voxel_coord((double)p.x, origin[0], voxel);
    const unsigned long long ky = (unsigned long long)voxel_coord((double)p.y, origin[1], voxel);
    const unsigned long long kz = (unsigned long long)voxel_coord((double)p.z, origin[2], voxel);
    const unsigned h = table_insert(t, ((kx << 42) | (ky << 21) | kz) + 1ull);
    if (h == 0xFFFFFFFFu) {
        atomicCAS(status, 0, B2_ERR_CAPACITY);
        return;
    }
    double* s = reinterpret_cast<double*>(t.sums) + 3 * (size_t)h;
    atomicAdd(s, (double)p.x);
    atomicAdd(s + 1, (double)p.y);
    atomicAdd(s + 2, (double)p.z);
    atomicAdd(&t.cnt[h], 1u);
    atomicMax(&t.first[h], ~(unsigned)i);
    slot_of[i] = h;
}

// The first point of every voxel writes the voxel's centroid at its rank among the first points.
__global__ void __launch_bounds__(kEmitBlock, 16)
ds_emit_kernel(const int* __restrict__ n_dev, int n_max, ScanTable t, const unsigned* __restrict__ slot_of,
               unsigned* __restrict__ agg, float4* __restrict__ out_pts, int* __restrict__ out_count,
               const PreHead* __restrict__ head) {
    pdl_enter();
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int i = blockIdx.x * kEmitBlock + threadIdx.x;
    const bool failed = *reinterpret_cast<const volatile int*>(&head->status) != 0;
    unsigned h = 0;
    bool owner = false;
    if (i < n && !failed) {
        h = slot_of[i];
        owner = t.first[h] == ~(unsigned)i;
    }
    __shared__ unsigned swarp[kEmitBlock / 32];
    __shared__ unsigned sprefix;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned ball = __ballot_sync(0xffffffffu, owner);
    if (lane == 0) swarp[w] = __popc(ball);
    __syncthreads();
    if (w == 0) {
        unsigned v = lane < kEmitBlock / 32 ? swarp[lane] : 0u, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane < kEmitBlock / 32) swarp[lane] = incl - v;  // exclusive warp offsets
        if (lane == 31) {
            // publish this block's aggregate; the predecessors' aggregates give the block's base (look-back)
            __threadfence();
            atomicExch(&agg[blockIdx.x], (incl << 1) | 1u);
        }
    }
    // the threads share the wait for the aggregates of the blocks before this one (blocks are dispatched in index
    // order, so whatever a resident block waits for is resident or has finished)
    unsigned mine = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kEmitBlock) {
        unsigned v;
        do { v = *reinterpret_cast<volatile unsigned*>(agg + b); } while (!(v & 1u));
        mine += v >> 1;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    __shared__ unsigned sbase[kEmitBlock / 32];
    if (lane == 0) sbase[w] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned b = 0;
        for (int k = 0; k < kEmitBlock / 32; ++k) b += sbase[k];
        sprefix = b;
    }
    __syncthreads();
    const unsigned base = sprefix + swarp[w] + __popc(ball & ((1u << lane) - 1u));
    if (owner) {
        const double* s = reinterpret_cast<const double*>(t.sums) + 3 * (size_t)h;
        const unsigned c = t.cnt[h];
        const double dn = (double)c;
        out_pts[base] = make_float4((float)ddiv(s[0], dn), (float)ddiv(s[1], dn), (float)ddiv(s[2], dn),
                                    __int_as_float((int)c));
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kEmitBlock - 1)
        *out_count = failed ? 0 : (int)(base + (owner ? 1u : 0u));
}

__global__ void __launch_bounds__(kBlock)
fuse_scatter_kernel(const float4* __restrict__ pts, const int* __restrict__ n_dev, int n_max,
                    const double* __restrict__ T, double ox, double oy, double oz, double voxel, ScanTable t,
                    MainHead* __restrict__ head, unsigned* __restrict__ list, const PreHead* __restrict__ pre,
                    int* __restrict__ dev_status) {
    pdl_enter();
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (pre) {  // a failed down-sample of this scan fails the step: nothing of it may reach the map
        const int ps = pre->status;
        if (ps != 0) {
            if (i == 0) atomicCAS(dev_status, 0, ps);
            return;
        }
    }
    if (i >= n || *dev_status != 0) return;
    const float4 p = __ldg(pts + i);
    D3 P{(double)p.x, (double)p.y, (double)p.z};
    if (T) P = xform_strict(T, P.x, P.y, P.z);
    const double q[3] = {ddiv(dsub(P.x, ox), voxel), ddiv(dsub(P.y, oy), voxel), ddiv(dsub(P.z, oz), voxel)};
    int k[3];
    long long f[3];
    const double lim = (double)((1 << 20) - 2);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double fl = floor(q[d]);
        if (!(fabs(fl) < lim)) {  // (also catches NaN / Inf)
            atomicCAS(dev_status, 0, B2_ERR_KEY_RANGE);
            return;
        }
        k[d] = (int)fl;
        double fr = q[d] - fl;  // exact (Sterbenz), in [0, 1)
        f[d] = (long long)(fr * kFixScale);
    }
    bool created = false;
    const unsigned h = table_insert(t, map_cell_key(k[0], k[1], k[2]) + 1ull, &created);
    if (h == 0xFFFFFFFFu) {
        atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
        return;
    }
    if (created) list[atomicAdd(&head->nlist, 1)] = h;  // (a few thousand per scan: the scan's unique voxels)
    unsigned long long* s = t.sums + 3 * (size_t)h;
    atomicAdd(s, (unsigned long long)f[0]);
    atomicAdd(s + 1, (unsigned long long)f[1]);
    atomicAdd(s + 2, (unsigned long long)f[2]);
    atomicAdd(&t.cnt[h], 1u);
}

// Occupied slots (listed by the scatter pass) -> (voxel key, (sum x, sum y, sum z, count)) records for the map's
// merge kernel.  When the map is brick-major (btab != null) the thread also makes sure the voxel's brick exists (the
// work of brick_ensure_kernel: the winner of the key CAS takes the next pool brick and publishes its index; nobody
// waits for it here, the merge is a separate launch).
__global__ void __launch_bounds__(kBlock)
fuse_collect_kernel(ScanTable t, double ox, double oy, double oz, double voxel, int* __restrict__ keys3,
                    double* __restrict__ sums, MainHead* __restrict__ head, const unsigned* __restrict__ list,
                    BrickEntry* __restrict__ btab, unsigned bmask, unsigned long long* __restrict__ bkeys,
                    unsigned pool_bricks, long long* __restrict__ counters, int* __restrict__ dev_status) {
    pdl_enter();
    const int j = blockIdx.x * kBlock + threadIdx.x;
    const int n = *dev_status == 0 ? head->nlist : 0;
    if (j == 0) head->ctl.nseg = n;
    if (j >= n) return;
    const unsigned h = list[j];
    const unsigned long long key = t.keys[h];
    const unsigned long long k = key - 1ull;
    const int kx = (int)((k >> 42) & 0x1FFFFF) - (1 << 20);
    const int ky = (int)((k >> 21) & 0x1FFFFF) - (1 << 20);
    const int kz = (int)(k & 0x1FFFFF) - (1 << 20);
    keys3[3 * j] = kx; keys3[3 * j + 1] = ky; keys3[3 * j + 2] = kz;
    const double dn = (double)t.cnt[h];
    const unsigned long long* s = t.sums + 3 * (size_t)h;
    const double o[3] = {ox, oy, oz};
    const int kk[3] = {kx, ky, kz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double corner = dadd(o[d], dmul((double)kk[d], voxel));
        sums[4 * j + d] = dadd(dmul(dn, corner), dmul(dmul((double)(long long)s[d], kFixInv), voxel));
    }
    sums[4 * j + 3] = dn;
    if (btab) {
        const unsigned long long bkey = map_cell_key(kx >> 3, ky >> 3, kz >> 3);  // (arithmetic shift = floor / 8)
        unsigned bh = cell_hash(bkey) & bmask;
        for (unsigned probe = 0; probe < 4096u; ++probe) {
            unsigned long long cur = btab[bh].key;
            if (cur == kEmptyKey) cur = atomicCAS(&btab[bh].key, kEmptyKey, bkey);
            if (cur == kEmptyKey) {  // created here
                const unsigned long long idx = (unsigned long long)atomicAdd((unsigned long long*)&counters[0], 1ull);
                if (idx >= pool_bricks) {
                    atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
                    return;
                }
                bkeys[idx] = bkey;
                btab[bh].index = (unsigned)idx;
                return;
            }
            if (cur == bkey) return;
            bh = (bh + 1) & bmask;
        }
        atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
    }
}

unsigned table_slots(int n_max) {
    unsigned cap = 1024;
    while ((double)cap < 2.5 * (double)n_max) cap <<= 1;
    return cap;
}

size_t table_bytes(unsigned cap) { return (size_t)cap * (8 + 24 + 4 + 4); }

ScanTable table_at(char* base, unsigned cap) {
    ScanTable t;
    t.keys = reinterpret_cast<unsigned long long*>(base);
    t.sums = reinterpret_cast<unsigned long long*>(base + (size_t)cap * 8);
    t.cnt = reinterpret_cast<unsigned*>(base + (size_t)cap * 32);
    t.first = reinterpret_cast<unsigned*>(base + (size_t)cap * 36);
    t.mask = cap - 1;
    return t;
}

struct PreBuffers {   // one per scan parity: [PreHead][agg][table], cleared by one zero-fill
    PreHead* head;
    unsigned* agg;
    ScanTable ds;
    size_t bytes;
};
struct MainBuffers {  // one per scan parity: [MainHead][table] cleared by one zero-fill, then the slot list
    MainHead* head;
    ScanTable fuse;
    unsigned* list;   // [n_max] occupied slots in creation order
    size_t bytes;     // of the zero-filled part
};
constexpr size_t kHeadBytes = 256;
static_assert(sizeof(PreHead) <= kHeadBytes && sizeof(MainHead) <= kHeadBytes, "heads fit their slots");

int pre_buffers(Context* c, int n_max, int par, PreBuffers* pb) {
    const unsigned cap = table_slots(n_max);
    const size_t agg_bytes = sizeof(unsigned) * kMaxEmitBlocks;
    const size_t total = kHeadBytes + agg_bytes + table_bytes(cap);
    B2_WS(c, raw, char, (par & 1) ? WS_STREAM_PRE1 : WS_STREAM_PRE0, total);
    pb->head = reinterpret_cast<PreHead*>(raw);
    pb->agg = reinterpret_cast<unsigned*>(raw + kHeadBytes);
    pb->ds = table_at(raw + kHeadBytes + agg_bytes, cap);
    pb->bytes = total;
    return B2_OK;
}

int main_buffers(Context* c, int n_max, int par, MainBuffers* mb) {
    const unsigned cap = table_slots(n_max);
    const size_t fill = kHeadBytes + table_bytes(cap);
    B2_WS(c, raw, char, (par & 1) ? WS_STREAM_MAIN1 : WS_STREAM_MAIN0, fill + sizeof(unsigned) * (size_t)n_max);
    mb->head = reinterpret_cast<MainHead*>(raw);
    mb->fuse = table_at(raw + kHeadBytes, cap);
    mb->list = reinterpret_cast<unsigned*>(raw + fill);
    mb->bytes = fill;
    return B2_OK;
}

__global__ void propagate_status_kernel(const PreHead* __restrict__ pre, int* __restrict__ dev_status) {
    if (pre->status != 0) atomicCAS(dev_status, 0, pre->status);
}

}  // namespace

bool stream_applicable(int n_max) { return n_max > 0 && div_up(n_max, kEmitBlock) <= kMaxEmitBlocks; }

namespace {
// launch of a front-stream kernel: programmatic stream serialisation + the dynamic shared memory that caps the CTAs
// per SM (see kFrontSmem)
template <typename... KArgs, typename... Args>
cudaError_t front_launch(void (*kernel)(KArgs...), dim3 grid, cudaStream_t stream, Args... args) {
    static const bool off = getenv("B2_NO_PDL") != nullptr;
    // (once per kernel: this function is instantiated per kernel signature and every signature here is one kernel;
    // the first launch of a kernel is always an eager one, outside stream capture)
    static const cudaError_t attr_rc =
        kFrontSmem > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrontSmem)
                               : cudaSuccess;
    if (attr_rc != cudaSuccess) return attr_rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kFront);
    cfg.dynamicSmemBytes = kFrontSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace

// Clear the down-sample's scratch state (table, bounding box, look-back flags, status) of scan parity `par`.
int stream_begin_pre(Context* c, int n_max, int par) {
    PreBuffers pb;
    B2_TRY(pre_buffers(c, n_max, par, &pb));
    B2_CUDA_OK(c, cudaMemsetAsync(pb.head, 0, pb.bytes, c->stream));
    return B2_OK;
}

// Clear the fusion's scratch state (table, voxel count) of scan parity `par` — on c->stream: the front half does it
// for the scans whose down-sample it runs (off the critical path), the main half otherwise.
int stream_begin_main(Context* c, int n_max, int par) {
    MainBuffers mb;
    B2_TRY(main_buffers(c, n_max, par, &mb));
    B2_CUDA_OK(c, cudaMemsetAsync(mb.head, 0, mb.bytes, c->stream));
    return B2_OK;
}

// voxel_down_sample(pts, voxel) with Open3D's origin, sort-free (see the header).  out_pts [n_max] float4
// (w = voxel point count), *out_count_dev = number of voxels (0 when the pass failed: the reason stays in the
// parity's status word until the fusion of the same scan — or stream_propagate_status — hands it on).
// stream_begin_pre(par) must have run.  Launches on c->stream.
int stream_downsample(Context* c, const float4* pts, int n_max, const int* n_dev, double voxel, float4* out_pts,
                      int* out_count_dev, int par) {
    if (!(voxel > 0.0)) return set_error(c, B2_ERR_INVALID, "voxel size must be > 0 (got %g)", voxel);
    PreBuffers sb;
    B2_TRY(pre_buffers(c, n_max, par, &sb));
    PreHead* head = sb.head;
    B2_WS(c, slot_of, unsigned, WS_STREAM_SLOTS, sizeof(unsigned) * (size_t)n_max);
    B2_CUDA_OK(c, front_launch(ds_bbox_kernel, dim3(min(div_up(n_max, kFront * 2), 2 * kSMs)), c->stream, pts, n_dev, n_max,
                               head));
    B2_LAUNCH_CHECK(c);
    B2_CUDA_OK(c, front_launch(ds_scatter_kernel, dim3(div_up(n_max, kFront)), c->stream, pts, n_dev, n_max, voxel, head,
                               sb.ds, slot_of));
    B2_LAUNCH_CHECK(c);
    B2_CUDA_OK(c, front_launch(ds_emit_kernel, dim3(div_up(n_max, kEmitBlock)), c->stream, n_dev, n_max, sb.ds,
                               (const unsigned*)slot_of, sb.agg, out_pts, out_count_dev, (const PreHead*)head));
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

// Hand a failed down-sample's reason to the sticky status word (stand-alone use of stream_downsample).
int stream_propagate_status(Context* c, int n_max, int par) {
    PreBuffers sb;
    B2_TRY(pre_buffers(c, n_max, par, &sb));
    propagate_status_kernel<<<1, 1, 0, c->stream>>>((const PreHead*)sb.head, c->dev_status);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

// The (key, sum, count) records of voxel_down_sample(T * pts, voxel) at an explicit origin, sort-free:
// keys3 [n_max,3], sums [n_max,4], count in *ctl_out->nseg.  stream_begin_main() must have run in this step.
// pre_par >= 0: the scan's down-sample ran under that parity; its failure fails this pass as well.
int stream_voxel_records(Context* c, const float4* pts, int n_max, const int* n_dev, const double* T_dev,
                         const double* origin, double voxel, int* keys3, double* sums, SortCtl** ctl_out,
                         BrickEntry* btab, unsigned bmask, unsigned long long* bkeys, unsigned pool_bricks,
                         long long* counters, int par, int pre_par) {
    MainBuffers sb;
    B2_TRY(main_buffers(c, n_max, par, &sb));
    const PreHead* pre = nullptr;
    if (pre_par >= 0) {
        PreBuffers pb;
        B2_TRY(pre_buffers(c, n_max, pre_par, &pb));
        pre = pb.head;
    }
    pdl_launch(fuse_scatter_kernel, dim3(div_up(n_max, kBlock)), dim3(kBlock), c->stream, pts, n_dev, n_max, T_dev,
               origin[0], origin[1], origin[2], voxel, sb.fuse, sb.head, sb.list, pre, c->dev_status);
    B2_LAUNCH_CHECK(c);
    pdl_launch(fuse_collect_kernel, dim3(div_up(n_max, kBlock)), dim3(kBlock), c->stream, sb.fuse,
               origin[0], origin[1], origin[2], voxel, keys3, sums, sb.head, (const unsigned*)sb.list, btab, bmask, bkeys,
               pool_bricks, counters, c->dev_status);
    B2_LAUNCH_CHECK(c);
    *ctl_out = &sb.head->ctl;
    return B2_OK;
}

}  // namespace b2
