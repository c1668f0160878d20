// b2_common.cuh — shared declarations of the sm_100a hot-path library (libb2slam.so).
//
// Layout rules used by every kernel in this directory:
//   * points live in HBM as float4 (x, y, z, w) — 16 B, one coalesced LDG.128 per
//     point; w carries an int32 tag (original index / slab position) where stated.
//   * every distance, transform, key and reduction is evaluated in float64 with
//     un-fused __dmul_rn/__dadd_rn in a fixed order, which is the arithmetic
//     Open3D (the reference's CPU back-end) performs on the widened inputs —
//     see SURVEY.md Appendix A and oracle/oracle.cpp.
//   * sizes that depend on data (voxel counts, pass counts) stay in device
//     memory; kernels are launched for the host-known upper bound and read the
//     actual count, so pipelines never synchronise with the host mid-flight.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/b2slam.h"

namespace b2 {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// Workspace slots (grown on demand, never shrunk).
enum WsSlot : int {
    WS_KEYS_A = 0, WS_KEYS_B, WS_VALS_A, WS_VALS_B, WS_HIST, WS_CTL, WS_SEGSTART, WS_BLOCKCNT,
    WS_BBOX, WS_PREP, WS_TMP_PTS, WS_TMP_PTS2, WS_TMP_KEYS, WS_SUMS, WS_COUNTS, WS_OFFS_A, WS_OFFS_B,
    WS_GRID_ENTRIES, WS_GRID_PTS, WS_GRID_META, WS_ICP_STATE, WS_ICP_PARTIAL, WS_ICP_TICKET, WS_ICP_SYNC,
    WS_SRC_DOWN, WS_SRC_KEYS, WS_STAGE_XYZ, WS_SCAN_PTS, WS_NORMALS_TMP, WS_GRID2_ENTRIES, WS_GRID2_PTS,
    WS_GRID2_META, WS_MISC, WS_EXPORT_KEYS, WS_EXPORT_PTS, WS_TGT_DOWN, WS_TGT_KEYS, WS_TGT_NORMALS,
    WS_CORR, WS_FILTER_MEAN, WS_FILTER_PART, WS_FILTER_KEEP, WS_SCAN_COUNT, WS_BATCH_A, WS_BATCH_B, WS_BATCH_C, WS_BATCH_D,
    WS_INGEST_RAW, WS_INGEST_TMP, WS_INGEST_KEEP, WS_STAGE_XYZ2, WS_ICP_DENSE, WS_STREAM_PRE0, WS_STREAM_PRE1, WS_STREAM_MAIN0, WS_STREAM_MAIN1, WS_STREAM_SLOTS, WS_SCAN_PTS2, WS_SRC_DOWN2, WS_NUM_SLOTS
};

struct MapState;  // b2_map.cu

// Handles alive on a device in this process.  The persistent dense-map registration wants every CTA of its grid
// resident (one per SM); two such launches from two handles of one process could each hold part of the SMs and
// wait for the rest, so the persistent driver is only taken while the handle is the only one on its device
// (b2_api.cu keeps the count; a change invalidates the captured scan steps through `handles_epoch`).
int live_handles(int device);
unsigned long long handles_epoch();

// Cached CUDA graph of one scan step (b2_api.cu) and the key it was captured for.
struct ScanGraph {
    cudaGraphExec_t exec = nullptr;
    int n = -1;
    double ds_voxel = 0.0;
    const void* scan = nullptr;
    b2_icp_params params{};
    unsigned long long generation = 0;
    unsigned long long handles_epoch = 0;
    int kernel_nodes = 0;
    int passes_hint = 0;  // radix pass budget baked into the captured sorts
};

struct Context {
    int icp_driver = 0;  // B2_ICP_DRIVER_*
    int scan_path = 0;   // B2_SCAN_PATH_*
    int dense_driver = 0;  // B2_DENSE_DRIVER_*
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string last_error;
    DevBuf ws[WS_NUM_SLOTS];
    int* dev_status = nullptr;     // sticky device-side error word (0 = ok)
    void* host_mailbox = nullptr;  // pinned, 4 KiB: results copied back here
    int sort_passes_hint = 4;      // upper bound of 8-bit radix passes launched
    MapState* map = nullptr;
    unsigned long long launches = 0;  // kernels launched through this context
    unsigned long long ws_generation = 0;  // bumped whenever a workspace / map buffer is (re)allocated
    bool capturing = false;                // inside stream capture: no allocation, no synchronisation
    unsigned scan_seq = 0;                 // scans submitted (indexes the pinned count ring)
    ScanGraph front_graph[2], main_graph[2];  // the two halves of the scan step, by scan parity
    // b2_slam_scan with host scans: uploads run on a second stream into two alternating staging buffers, so
    // the copy of scan k+1 overlaps the registration of scan k
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    bool consumed_valid[2] = {false, false};
    cudaEvent_t ev_dev_ready = nullptr;            // main -> front: a device scan produced on the main stream is ready
    unsigned long long launches_after_scan = ~0ull;  // c->launches at the end of the last scan step
    // pageable host scans are staged through a small ring of pinned buffers owned by the library
    static constexpr int kHostStages = 4;
    void* host_stage[kHostStages] = {nullptr, nullptr, nullptr, nullptr};
    size_t host_stage_bytes[kHostStages] = {0, 0, 0, 0};
    cudaEvent_t ev_host_stage[kHostStages] = {nullptr, nullptr, nullptr, nullptr};
    unsigned host_stage_seq = 0;
    // deferred read-back (b2_slam_scan with out == NULL, then b2_slam_results): pinned ring of per-scan results
    void* result_ring = nullptr;
    unsigned char* ring_first = nullptr;   // slot holds a first-scan record (pose only)
    unsigned long long scans_total = 0;
    // live point count of every submitted scan: pinned ring as deep as the result ring (the H2D copy of a
    // count is asynchronous, so a slot must stay untouched until the scan that uses it has run).  The ring is
    // split in four blocks with one event each; the host waits for a block's previous use before reusing it.
    int* count_ring = nullptr;
    cudaEvent_t ev_block[4] = {nullptr, nullptr, nullptr, nullptr};
    bool block_valid[4] = {false, false, false, false};
};
constexpr int kResultRing = 1024;

int set_error(Context* c, int code, const char* fmt, ...);
void* ws_get(Context* c, int slot, size_t bytes);  // nullptr on failure (error set)

#define B2_CUDA_OK(c, expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return b2::set_error((c), B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

#define B2_LAUNCH_CHECK(c)                                                                        \
    do {                                                                                          \
        (c)->launches++;                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess)                                                                    \
            return b2::set_error((c), B2_ERR_CUDA, "kernel launch failed: %s (%s:%d)",           \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

#define B2_TRY(expr)                \
    do {                            \
        int _rc = (expr);           \
        if (_rc != B2_OK) return _rc; \
    } while (0)

#define B2_WS(c, var, type, slot, bytes)                    \
    type* var = (type*)b2::ws_get((c), (slot), (bytes));    \
    if (!var) return B2_ERR_CUDA

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// Launch with programmatic stream serialisation: the grid may be scheduled while its predecessor in
// the stream is still draining (the launch latency of the ~60 small kernels of a scan step overlaps
// with the previous kernel's tail).  Every kernel launched this way begins with pdl_enter().
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    static const bool off = getenv("B2_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// First statement of a pdl_launch'ed kernel: let the successor be scheduled, then wait until the
// predecessor grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// ---------------------------------------------------------------------------
// device-side structs
// ---------------------------------------------------------------------------

// Per-cloud voxelisation parameters produced by prep_kernel (b2_voxel.cu).
struct VoxelPrep {
    double origin[3];  // key = floor((p - origin) / voxel)
    double minb[3], maxb[3];
    int kmin[3];       // smallest key per axis over the cloud
    int pad;
};

// Sort / segmentation control block (device resident).
struct SortCtl {
    int n;          // number of live elements
    int npasses;    // 8-bit passes needed for the live key range
    int bits[3];    // packed bits per axis (shared by all clouds of a launch)
    int cloud_bits; // bits of the cloud id field
    int nseg;       // number of unique keys (written by segmentation)
    int status;     // 0 ok, else B2_ERR_*
};

// Uniform-grid / voxel-hash cell record: one 16 B load returns key and range.
struct __align__(16) CellEntry {
    unsigned long long key;  // kEmptyKey = vacant
    unsigned int off;        // first point of the cell in the grid point array
    unsigned int cnt;        // number of points
};
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// Resident map with one voxel per cell (voxel edge == search radius, e.g. 5 cm / 5 cm, BASELINE config 2):
// brick-major and directly indexed.  The hash holds BRICKS of 8 x 8 x 8 voxels (one 16-byte record per brick);
// a brick is a dense array of 512 float4 cells (centroid x, y, z; w = bit pattern of the int32 point count,
// 0 = empty), z fastest, so the three z-neighbours of a stencil row are 48 contiguous bytes, a 128-byte line
// holds eight consecutive cells, and a whole 27-cell stencil touches 1-8 bricks instead of 27 random hash
// slots.  Bricks are allocated from a pool in creation order (scans arrive along a trajectory, so the
// bricks one scan reads are neighbours in memory as well).
constexpr int kBrickEdge = 8;
constexpr int kBrickCells = kBrickEdge * kBrickEdge * kBrickEdge;  // 512
constexpr unsigned kNoBrick = 0xFFFFFFFFu;
struct __align__(16) BrickEntry {
    unsigned long long key;  // map_cell_key(brick x, y, z); kEmptyKey = vacant
    unsigned index;          // brick number in the pool (kNoBrick while being created / pool exhausted)
    unsigned pad;
};

struct GridMeta {
    double origin[3];
    double cell;     // pair grids: cell edge.  map grids: voxel edge
    int brick;       // map grids: voxels per cell edge (cell edge = brick * voxel); pair grids: 1
    int pad;
};

// Per-pair ICP state, device resident across iterations.
struct IcpState {
    double T[16];        // current source->target transform (row major)
    double fitness, rmse;
    double prev_fitness, prev_rmse;
    double pivot[3];     // accumulation pivot (keeps sums small)
    int iter;            // estimation steps applied so far
    int done;            // 1 once converged / exhausted
    int n_corr;
    int has_prev;
};

// Result record copied back to the host (matches b2_icp_result in b2slam.h).
struct IcpResultDev {
    double T[16];
    double fitness, rmse;
    int iterations, n_corr;
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

struct D3 {
    double x, y, z;
};

// A.7: ((T0*x + T1*y) + T2*z) + T3, un-fused.  The homogeneous divide is the
// identity for a rigid T (last row 0,0,0,1 — validated at the C-ABI).
__device__ __forceinline__ D3 xform_strict(const double* __restrict__ T, double x, double y, double z) {
    D3 r;
    r.x = dadd(dadd(dadd(dmul(T[0], x), dmul(T[1], y)), dmul(T[2], z)), T[3]);
    r.y = dadd(dadd(dadd(dmul(T[4], x), dmul(T[5], y)), dmul(T[6], z)), T[7]);
    r.z = dadd(dadd(dadd(dmul(T[8], x), dmul(T[9], y)), dmul(T[10], z)), T[11]);
    return r;
}

// A.1 key: floor((p - origin) / v) — IEEE division, like Open3D.
__device__ __forceinline__ int voxel_coord(double p, double origin, double v) {
    return (int)floor(ddiv(dsub(p, origin), v));
}

// nanoflann L2_Simple accumulation order: ((dx*dx) + dy*dy) + dz*dz
__device__ __forceinline__ double dist2_strict(double qx, double qy, double qz, double tx, double ty, double tz) {
    double dx = dsub(qx, tx), dy = dsub(qy, ty), dz = dsub(qz, tz);
    return dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
}

__device__ __forceinline__ unsigned long long order_encode(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_decode(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ unsigned long long hash64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// Table index hash of a packed cell key: two 32-bit multiplies + xorshift-multiply finaliser
// (the full 64-bit murmur finaliser costs ~3x the instructions and the probe is issue-bound).
__device__ __forceinline__ unsigned cell_hash(unsigned long long key) {
    unsigned h = (unsigned)key * 0x9E3779B1u ^ (unsigned)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    return h;
}

// pair-grid cell key: [pair:16][x:16][y:16][z:16], coordinates relative to the cloud origin
__device__ __forceinline__ unsigned long long pair_cell_key(unsigned pair, int x, int y, int z) {
    return ((unsigned long long)pair << 48) | ((unsigned long long)(unsigned)x << 32) |
           ((unsigned long long)(unsigned)y << 16) | (unsigned long long)(unsigned)z;
}
// resident-map cell key: [x:21][y:21][z:21] with a 2^20 bias per axis
__device__ __forceinline__ unsigned long long map_cell_key(int x, int y, int z) {
    return ((unsigned long long)(unsigned)(x + (1 << 20)) << 42) | ((unsigned long long)(unsigned)(y + (1 << 20)) << 21) |
           (unsigned long long)(unsigned)(z + (1 << 20));
}

__device__ __forceinline__ int floor_div(int a, int b) {
    int q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

// Probe an open-addressing table (linear probing); returns the entry or a vacant one.
__device__ __forceinline__ CellEntry grid_lookup(const CellEntry* __restrict__ tab, unsigned mask,
                                                 unsigned long long key) {
    unsigned h = cell_hash(key) & mask;
    CellEntry e;
    e.key = kEmptyKey; e.off = 0; e.cnt = 0;
#pragma unroll 1
    for (unsigned probe = 0; probe <= mask; ++probe) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(tab + h));
        unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
        if (k == key) {
            e.key = k; e.off = raw.z; e.cnt = raw.w;
            return e;
        }
        if (k == kEmptyKey) return e;
        h = (h + 1) & mask;
    }
    return e;
}

// One 256-bit load (LDG.E.256 on sm_100a) of a 32-byte-aligned record.
__device__ __forceinline__ void ldg256(const void* p, unsigned long long& a, unsigned long long& b,
                                       unsigned long long& c, unsigned long long& d) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

// Probe the brick table: the brick's pool index, or kNoBrick.
__device__ __forceinline__ unsigned brick_lookup(const BrickEntry* __restrict__ tab, unsigned mask,
                                                 unsigned long long key) {
    unsigned h = cell_hash(key) & mask;
#pragma unroll 1
    for (unsigned probe = 0; probe <= mask; ++probe) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(tab + h));
        const unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
        if (k == key) return raw.z;
        if (k == kEmptyKey) return kNoBrick;
        h = (h + 1) & mask;
    }
    return kNoBrick;
}
// cell number inside a brick: z fastest
__device__ __forceinline__ unsigned brick_local(int x, int y, int z) {
    return ((unsigned)(x & 7) << 6) | ((unsigned)(y & 7) << 3) | (unsigned)(z & 7);
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------
// internal host entry points (one per .cu)
// ---------------------------------------------------------------------------

// b2_sort.cu — stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits.
// Element and pass counts are read from ctl (device).  Result lands in A when
// ctl->npasses is even, else in B; consumers use sort_result_*().
int radix_sort_pairs(Context* c, unsigned long long* keysA, unsigned long long* keysB, unsigned* valsA,
                     unsigned* valsB, const SortCtl* ctl, int n_max, int max_passes);

// b2_voxel.cu
struct VoxelSortOut {
    unsigned long long *keysA, *keysB;  // sorted packed keys live in (ctl->npasses & 1 ? B : A)
    unsigned *valsA, *valsB;            // original point index of every sorted element
    SortCtl* ctl;
    VoxelPrep* prep;                    // [n_clouds]
    int* seg_start;                     // [ctl->nseg] first sorted position of every unique key
};
struct DownsampleOut {
    float4* pts;          // [n_max] centroids, w = bit pattern of the voxel point count
    int* keys;            // [n_max,3] voxel keys (may be null)
    double* sums;         // [n_max,4] (sum x, sum y, sum z, count) (may be null)
    int* cloud_offsets;   // [n_clouds+1] output offsets (may be null)
};
// Sort n_clouds concatenated clouds by voxel key.  cloud_offsets device [n_clouds+1] (null => one cloud of
// *n_dev (or n_max) points).  T_dev (16 doubles, device, or null) is applied to every point first.
// origin_mode 0: per-cloud min_bound - v/2 (Open3D); 1: explicit origin_host[3]; 2: per-cloud min_bound.
int voxel_sort(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets, int n_clouds,
               double voxel, int origin_mode, const double* origin_host, const double* T_dev, int* cloud_out_offs,
               VoxelSortOut* out);
// voxel_sort + centroid reduction.  The voxel count ends up in vs_out->ctl->nseg (device).
int voxel_downsample(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets_in,
                     int n_clouds, double voxel, int origin_mode, const double* origin_host, const double* T_dev,
                     DownsampleOut out, VoxelSortOut* vs_out);

// b2_voxel_cluster.cu — single-launch (8-CTA cluster) variant for one cloud of <= 65,536 points
bool voxel_small_applicable(int n_max, int n_clouds);
int voxel_small(Context* c, const float4* pts, int n_max, const int* n_dev, double voxel, int origin_mode,
                const double* origin_host, const double* T_dev, bool do_centroid, DownsampleOut out,
                VoxelSortOut* vs_out);

// b2_grid.cu — uniform grid (cell hash) over n_clouds concatenated target clouds.
struct GridView {
    CellEntry* entries;
    unsigned mask;        // capacity - 1
    float4* pts;          // cell-sorted copy, w = original (per-cloud) index
    GridMeta* meta;       // [n_clouds]
    int is_map;           // 0: pair grid, 1: resident map with slabs (brick > 1), 2: dense resident map (bricks)
    BrickEntry* btab;     // is_map == 2: brick hash (mask = its capacity - 1), cells = pts
    unsigned pool_bricks; // is_map == 2: bricks in the pool
};
int grid_build(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets, int n_clouds,
               double cell, int slot_base, GridView* out);

// b2_icp.cu
struct IcpParams {
    double max_corr;
    int max_iter;
    double rel_fitness, rel_rmse;
    int mode;  // 0 p2p, 1 p2l
};
// Runs the A.3 loop for n_pairs independent (source, target-grid) pairs.  src_offsets device [n_pairs+1]
// (null => one pair of *ns_dev (or ns_max) points).  init_dev: [n_pairs,16] doubles or null (identity).
// results_dev[pair] is written when the pair finishes.  corr_out/d2_out (optional) receive the per-source
// correspondence of the last evaluated transform.  search_only: one correspondence pass, no estimation.
// check_every > 0: poll the finished-pair counter every that many launches (host sync) and stop early.
int icp_configure(Context* c);
int icp_run(Context* c, const float4* src, int ns_max, const int* src_offsets, const int* ns_dev, int n_pairs,
            const GridView& grid, const float4* tgt_normals, const int* tgt_offsets, const double* init_dev,
            const IcpParams& prm, IcpResultDev* results_dev, int* corr_out, double* d2_out, int search_only,
            int check_every);

int map_stencil_count(Context* c, const float4* src, int ns, const double* T_dev, const GridView& g,
                      unsigned long long* total_dev);

// b2_normals.cu
int estimate_normals(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets, int n_clouds,
                     double radius, int max_nn, float4* normals_out);

// b2_filters.cu
int radius_outlier_mask(Context* c, const float4* pts, int n, int nb_points, double radius, unsigned char* keep);
// extent: upper bound of the cloud's largest bounding-box edge (decides the coarsest search level)
int statistical_outlier_mask(Context* c, const float4* pts, int n, int nb_neighbors, double std_ratio, double cell,
                             double extent, unsigned char* keep);
int select_points(Context* c, const float4* pts, int n, const unsigned char* keep, float4* out, int* n_out_dev);

// b2_stream.cu — sort-free voxel passes of the streaming scan step (scratch hash per scan); everything launches on
// c->stream.  `par` = scan parity: the down-sample of scan k+1 may run (front stream) while scan k is registered.
bool stream_applicable(int n_max);
int stream_begin_pre(Context* c, int n_max, int par);
int stream_begin_main(Context* c, int n_max, int par);
int stream_downsample(Context* c, const float4* pts, int n_max, const int* n_dev, double voxel, float4* out_pts,
                      int* out_count_dev, int par);
int stream_propagate_status(Context* c, int n_max, int par);
// btab != null (brick-major map): the collecting pass also creates the bricks of the scan's voxels
int stream_voxel_records(Context* c, const float4* pts, int n_max, const int* n_dev, const double* T_dev,
                         const double* origin, double voxel, int* keys3, double* sums, SortCtl** ctl_out,
                         BrickEntry* btab, unsigned bmask, unsigned long long* bkeys, unsigned pool_bricks,
                         long long* counters, int par, int pre_par);

// b2_map.cu
void map_free(Context* c);
int map_reset(Context* c);
int map_init(Context* c, double voxel, const double* origin, double max_corr, long long capacity_cells);
int map_grid_view(Context* c, GridView* g);
int map_fuse(Context* c, const float4* pts, int n_max, const int* n_dev, const double* T_dev);
// the same through the sort-free per-scan records of b2_stream.cu (stream_begin() must have run in this step)
// commit_res != null: the merge kernel of a brick-major map also copies commit_res->T to the pose prior (*committed)
int map_fuse_stream(Context* c, const float4* pts, int n_max, const int* n_dev, const double* T_dev, int par,
                    int pre_par, const IcpResultDev* commit_res, bool* committed);
int map_counts(Context* c, long long* cells, long long* voxels);
int map_export(Context* c, float4* out_pts, int* out_keys, double* out_sums, long long max_out, long long* n_out);
int map_merge_records(Context* c, const int* keys3, const double* sums, int n);
double* map_pose_dev(Context* c);
IcpResultDev* map_result_dev(Context* c);
double map_voxel(Context* c);
double map_max_corr(Context* c);
bool map_is_empty_host(Context* c);
void map_set_voxels_bound(Context* c, long long voxels);

}  // namespace b2
