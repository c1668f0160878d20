// b2_grid.cu — uniform grid (cell hash) over one or many target clouds.
//
// Replaces the KD-tree Open3D builds inside registration_icp / estimate_normals
// (geometry::KDTreeFlann; reference call sites icp_processor.py:94-99,103-104).
// The cell edge equals the search radius, so every neighbour with d^2 < r^2 of a
// query lies in the 27 cells around the query's cell.
//
// Build = voxel_sort with voxel := cell, origin := the cloud's min bound, then
//   * grid_fill_kernel: cell-sorted float4 copy of the points (w = index of the
//     point in its own cloud), so a cell is one contiguous, coalesced run;
//   * one open-addressing hash entry {key, first, count} per occupied cell —
//     16 B, fetched with a single LDG.128 by the search.
// Algorithmic bytes: 16 N read + 16 N written + 16 C entries.
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock)
grid_fill_kernel(const float4* __restrict__ pts, const int* __restrict__ offs, int n_clouds, double cell,
                 const unsigned long long* __restrict__ keysA, const unsigned long long* __restrict__ keysB,
                 const unsigned* __restrict__ valsA, const unsigned* __restrict__ valsB,
                 const SortCtl* __restrict__ ctl, const VoxelPrep* __restrict__ prep,
                 const int* __restrict__ seg_start, CellEntry* __restrict__ entries, unsigned mask,
                 float4* __restrict__ grid_pts, GridMeta* __restrict__ meta, int* __restrict__ dev_status) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    const bool odd = ctl->npasses & 1;
    const unsigned long long* keys = odd ? keysB : keysA;
    const unsigned* vals = odd ? valsB : valsA;
    const int n = ctl->n, nseg = ctl->nseg;
    const int bx = ctl->bits[0], by = ctl->bits[1], bz = ctl->bits[2];
    const int shift = bx + by + bz;
    if (i < n_clouds) {
        GridMeta m;
        m.origin[0] = prep[i].origin[0];
        m.origin[1] = prep[i].origin[1];
        m.origin[2] = prep[i].origin[2];
        m.cell = cell;
        m.brick = 1;
        m.pad = 0;
        meta[i] = m;
    }
    if (i == 0 && (bx > 16 || by > 16 || bz > 16 || n_clouds > 65536)) atomicCAS(dev_status, 0, B2_ERR_KEY_RANGE);
    if (i < n) {
        const unsigned v = vals[i];
        const int cloud = n_clouds > 1 ? (int)(keys[i] >> shift) : 0;
        float4 p = __ldg(pts + v);
        p.w = __int_as_float((int)v - (offs ? offs[cloud] : 0));
        grid_pts[i] = p;
    }
    if (i < nseg) {
        const int b = seg_start[i];
        const int e = (i + 1 < nseg) ? seg_start[i + 1] : n;
        unsigned long long k = keys[b];
        unsigned kz = (unsigned)(k & ((1ull << bz) - 1ull));
        k >>= bz;
        unsigned ky = (unsigned)(k & ((1ull << by) - 1ull));
        k >>= by;
        unsigned kx = (unsigned)(k & ((1ull << bx) - 1ull));
        k >>= bx;
        const unsigned long long ckey = pair_cell_key((unsigned)k, (int)kx, (int)ky, (int)kz);
        unsigned h = cell_hash(ckey) & mask;
        for (unsigned probe = 0; probe <= mask; ++probe) {
            unsigned long long prev = atomicCAS(&entries[h].key, kEmptyKey, ckey);
            if (prev == kEmptyKey) {
                entries[h].off = (unsigned)b;
                entries[h].cnt = (unsigned)(e - b);
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

}  // namespace

int grid_build(Context* c, const float4* pts, int n_max, const int* n_dev, const int* cloud_offsets, int n_clouds,
               double cell, int slot_base, GridView* out) {
    if (n_max <= 0) return set_error(c, B2_ERR_INVALID, "grid_build: empty target cloud");
    VoxelSortOut vs;
    B2_TRY(voxel_sort(c, pts, n_max, n_dev, cloud_offsets, n_clouds, cell, /*origin_mode=*/2, nullptr, nullptr,
                      nullptr, &vs));
    unsigned cap = 1024;
    while (cap < 2u * (unsigned)n_max) cap <<= 1;
    B2_WS(c, entries, CellEntry, slot_base + 0, sizeof(CellEntry) * (size_t)cap);
    B2_WS(c, gpts, float4, slot_base + 1, sizeof(float4) * (size_t)n_max);
    B2_WS(c, meta, GridMeta, slot_base + 2, sizeof(GridMeta) * (size_t)n_clouds);
    B2_CUDA_OK(c, cudaMemsetAsync(entries, 0xFF, sizeof(CellEntry) * (size_t)cap, c->stream));
    int work = n_max > n_clouds ? n_max : n_clouds;
    grid_fill_kernel<<<div_up(work, kBlock), kBlock, 0, c->stream>>>(pts, cloud_offsets, n_clouds, cell, vs.keysA,
                                                                     vs.keysB, vs.valsA, vs.valsB, vs.ctl, vs.prep,
                                                                     vs.seg_start, entries, cap - 1, gpts, meta,
                                                                     c->dev_status);
    B2_LAUNCH_CHECK(c);
    out->entries = entries;
    out->mask = cap - 1;
    out->pts = gpts;
    out->meta = meta;
    out->is_map = 0;
    out->btab = nullptr;
    out->pool_bricks = 0;
    return B2_OK;
}

}  // namespace b2
