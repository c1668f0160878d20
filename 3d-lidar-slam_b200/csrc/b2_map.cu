// b2_map.cu — GPU-resident global map: a voxel hash that is also the uniform
// grid of the correspondence search.
//
// The reference keeps the map as a host ndarray of raw points, and for EVERY
// scan re-voxelises it, re-builds a KD-tree over it and copies it host<->Open3D
// (icp_processor.py:64-65,74,91,103-104 — O(map) per scan, README.md:194-195).
// With a fixed voxel origin, voxel_down_sample(concat(scans)) equals the merge
// of the per-scan (sum, count) accumulators, so the map can stay resident:
//
//   cells   : open-addressing hash, one 16 B CellEntry per CELL of `brick`^3
//             voxels (cell edge = brick*voxel >= max_corr, so the 27-cell stencil
//             of b2_icp.cu covers the correspondence radius).  The slab of a cell
//             is addressed by its hash slot — no allocator, no pointers.
//   cent    : [capacity * brick^3] float4 voxel centroids, compact per cell
//             (first `cnt` slots used), w = slab position  -> what ICP reads.
//   acc     : [capacity * brick^3] (sum x, sum y, sum z, n) in fp64 -> exact merge.
//   lut/lid : per-cell voxel-local-id <-> slab-slot maps (bytes).
//   brick == 1 (voxel edge == search radius, BASELINE config 2): brick-major, directly indexed.  The hash
//             holds bricks of 8^3 voxels; a brick is a dense array of 512 float4 cells (z fastest) from a pool
//             filled in creation order, so a 27-cell stencil is 9 runs of 48 contiguous bytes in 1-8 bricks and
//             the ICP kernel stages the cells its queries need with coalesced loads (b2_icp_dense.cuh).  Round 1
//             hashed single voxels into a 1 GiB table of 32-byte records: 27 random sectors per query, and
//             every probe of a warp in a different 128-byte line — the step was bound by L1 wavefronts.
//
// Fusion of a scan = voxel_downsample(T * scan, voxel, origin) [b2_voxel.cu]
// followed by map_merge_kernel: one thread per UNIQUE scan voxel, so a voxel is
// touched by exactly one thread per launch and needs no floating-point atomics
// (results are deterministic); only cell creation (CAS) and slot allocation
// (atomicAdd) are atomic.  Algorithmic bytes per fused voxel: 32 (sums in) +
// 16 (entry) + 2*32 (acc RMW) + 16 (centroid out).
#include "b2_common.cuh"

namespace b2 {

struct MapState {
    double voxel = 0, max_corr = 0;
    double origin[3] = {0, 0, 0};
    int brick = 1, cap_slots = 1, lut_stride = 16;
    unsigned capacity = 0;
    CellEntry* entries = nullptr;   // brick > 1
    BrickEntry* btab = nullptr;     // brick == 1: brick hash [capacity]
    unsigned long long* bkeys = nullptr;  // brick == 1: key of every pool brick [pool_bricks]
    unsigned pool_bricks = 0;       // brick == 1: bricks in the pool; cent = [pool_bricks * 512] cells
    unsigned long long max_voxels = 0;
    unsigned char* lut = nullptr;   // [capacity * lut_stride]  voxel-local id -> slot + 1
    unsigned char* lid = nullptr;   // [capacity * cap_slots]   slot -> voxel-local id
    float4* cent = nullptr;
    double* acc = nullptr;          // [capacity * cap_slots * 4]
    long long* counters = nullptr;  // device: [0] cells, [1] voxels
    GridMeta* meta = nullptr;       // device
    double* pose = nullptr;         // device, 16 doubles: current sensor->map pose
    IcpResultDev* result = nullptr; // device, last registration result
    long long host_voxels_bound = 0;
};

namespace {

constexpr int kBlock = 256;

__global__ void map_clear_kernel(CellEntry* __restrict__ entries, unsigned capacity, int cap_slots) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= capacity) return;
    CellEntry e;
    e.key = kEmptyKey;
    e.off = i * (unsigned)cap_slots;
    e.cnt = 0;
    entries[i] = e;
}

__global__ void map_set_meta_kernel(GridMeta* meta, double ox, double oy, double oz, double voxel, int brick,
                                    double* pose, long long* counters) {
    meta->origin[0] = ox; meta->origin[1] = oy; meta->origin[2] = oz;
    meta->cell = voxel;
    meta->brick = brick;
    meta->pad = 0;
    for (int i = 0; i < 16; ++i) pose[i] = (i % 5 == 0) ? 1.0 : 0.0;
    counters[0] = counters[1] = 0;
}

__global__ void __launch_bounds__(kBlock)
map_merge_kernel(const int* __restrict__ keys3, const double* __restrict__ sums, const SortCtl* __restrict__ ctl,
                 CellEntry* __restrict__ entries, unsigned mask, unsigned char* __restrict__ lut,
                 unsigned char* __restrict__ lid, float4* __restrict__ cent, double* __restrict__ acc,
                 long long* __restrict__ counters, int brick, int cap_slots, int lut_stride,
                 unsigned long long max_cells, int* __restrict__ dev_status) {
    pdl_enter();
    const int j = blockIdx.x * kBlock + threadIdx.x;
    if (j >= ctl->nseg || *dev_status != 0) return;  // a step that already failed must not touch the map
    const int kx = keys3[3 * j], ky = keys3[3 * j + 1], kz = keys3[3 * j + 2];
    const int cx = floor_div(kx, brick), cy = floor_div(ky, brick), cz = floor_div(kz, brick);
    const int lim = (1 << 20) - 2;
    if (cx < -lim || cx > lim || cy < -lim || cy > lim || cz < -lim || cz > lim) {
        atomicCAS(dev_status, 0, B2_ERR_KEY_RANGE);
        return;
    }
    const int local = ((kx - cx * brick) * brick + (ky - cy * brick)) * brick + (kz - cz * brick);
    const unsigned long long ckey = map_cell_key(cx, cy, cz);
    unsigned h = cell_hash(ckey) & mask;
    bool placed = false;
    for (unsigned probe = 0; probe < 4096u; ++probe) {
        unsigned long long prev = atomicCAS(&entries[h].key, kEmptyKey, ckey);
        if (prev == kEmptyKey) {
            unsigned long long cells = (unsigned long long)atomicAdd((unsigned long long*)&counters[0], 1ull) + 1ull;
            if (cells > max_cells) atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
            placed = true;
            break;
        }
        if (prev == ckey) {
            placed = true;
            break;
        }
        h = (h + 1) & mask;
    }
    if (!placed) {
        atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
        return;
    }
    int slot;
    bool fresh = false;
    if (brick == 1) {
        slot = 0;
        if (entries[h].cnt == 0) {
            entries[h].cnt = 1;
            fresh = true;
        }
    } else {
        const unsigned char s = lut[(size_t)h * lut_stride + local];
        if (s == 0) {
            slot = (int)atomicAdd(&entries[h].cnt, 1u);
            lut[(size_t)h * lut_stride + local] = (unsigned char)(slot + 1);
            fresh = true;
        } else {
            slot = (int)s - 1;
        }
    }
    const size_t pos = (size_t)h * cap_slots + slot;
    double sx = sums[4 * j], sy = sums[4 * j + 1], sz = sums[4 * j + 2], n = sums[4 * j + 3];
    if (fresh) {
        lid[pos] = (unsigned char)local;
        atomicAdd((unsigned long long*)&counters[1], 1ull);
    } else {
        sx = dadd(acc[4 * pos], sx);
        sy = dadd(acc[4 * pos + 1], sy);
        sz = dadd(acc[4 * pos + 2], sz);
        n = acc[4 * pos + 3] + n;
    }
    acc[4 * pos] = sx; acc[4 * pos + 1] = sy; acc[4 * pos + 2] = sz; acc[4 * pos + 3] = n;
    float4 cpt;
    cpt.x = (float)ddiv(sx, n);
    cpt.y = (float)ddiv(sy, n);
    cpt.z = (float)ddiv(sz, n);
    cpt.w = __int_as_float((int)pos);
    cent[pos] = cpt;
}

__global__ void brick_clear_kernel(BrickEntry* __restrict__ tab, unsigned capacity) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= capacity) return;
    BrickEntry e;
    e.key = kEmptyKey;
    e.index = kNoBrick;
    e.pad = 0;
    tab[i] = e;
}

__device__ __forceinline__ bool brick_of_voxel(int kx, int ky, int kz, unsigned long long& bkey, unsigned& local) {
    const int lim = (1 << 20) - 2;
    if (kx < -lim || kx > lim || ky < -lim || ky > lim || kz < -lim || kz > lim) return false;
    bkey = map_cell_key(kx >> 3, ky >> 3, kz >> 3);  // arithmetic shift = floor division by 8
    local = brick_local(kx, ky, kz);
    return true;
}

// brick == 1, pass 1: make sure the brick of every unique scan voxel exists.  The winner of the key CAS takes
// the next pool brick and publishes its index; nobody waits for it inside this kernel (pass 2 is a separate
// launch), so there is no spinning on another thread's progress.
__global__ void __launch_bounds__(kBlock)
brick_ensure_kernel(const int* __restrict__ keys3, const SortCtl* __restrict__ ctl, BrickEntry* __restrict__ tab,
                    unsigned mask, unsigned long long* __restrict__ bkeys, unsigned pool_bricks,
                    long long* __restrict__ counters, int* __restrict__ dev_status) {
    pdl_enter();
    const int j = blockIdx.x * kBlock + threadIdx.x;
    if (j >= ctl->nseg || *dev_status != 0) return;  // a step that already failed must not touch the map
    unsigned long long bkey;
    unsigned local;
    if (!brick_of_voxel(keys3[3 * j], keys3[3 * j + 1], keys3[3 * j + 2], bkey, local)) {
        atomicCAS(dev_status, 0, B2_ERR_KEY_RANGE);
        return;
    }
    // sorted scan voxels: the previous one is usually in the same brick and does the work
    if (j > 0) {
        unsigned long long pk;
        unsigned pl;
        if (brick_of_voxel(keys3[3 * j - 3], keys3[3 * j - 2], keys3[3 * j - 1], pk, pl) && pk == bkey) return;
    }
    unsigned h = cell_hash(bkey) & mask;
    for (unsigned probe = 0; probe < 4096u; ++probe) {
        unsigned long long cur = tab[h].key;
        if (cur == kEmptyKey) cur = atomicCAS(&tab[h].key, kEmptyKey, bkey);
        if (cur == kEmptyKey) {  // created here
            const unsigned long long idx = (unsigned long long)atomicAdd((unsigned long long*)&counters[0], 1ull);
            if (idx >= pool_bricks) {
                atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
                return;
            }
            bkeys[idx] = bkey;
            tab[h].index = (unsigned)idx;
            return;
        }
        if (cur == bkey) return;
        h = (h + 1) & mask;
    }
    atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
}

// brick == 1, pass 2: one thread per unique scan voxel adds its (sum, count) to the voxel's accumulator and
// refreshes the centroid cell; a voxel is touched by exactly one thread per launch.
__global__ void __launch_bounds__(kBlock)
brick_merge_kernel(const int* __restrict__ keys3, const double* __restrict__ sums, const SortCtl* __restrict__ ctl,
                   const BrickEntry* __restrict__ tab, unsigned mask, float4* __restrict__ cells,
                   double* __restrict__ acc, long long* __restrict__ counters, unsigned long long max_voxels,
                   int* __restrict__ dev_status, const IcpResultDev* __restrict__ commit_res,
                   double* __restrict__ commit_pose) {
    pdl_enter();
    const int j = blockIdx.x * kBlock + threadIdx.x;
    // scan step: the registered pose becomes the prior of the next scan together with the merge (same gate)
    if (commit_res && j < 16 && *dev_status == 0) commit_pose[j] = commit_res->T[j];
    if (j >= ctl->nseg || *dev_status != 0) return;
    unsigned long long bkey;
    unsigned local;
    if (!brick_of_voxel(keys3[3 * j], keys3[3 * j + 1], keys3[3 * j + 2], bkey, local)) return;
    const unsigned idx = brick_lookup(tab, mask, bkey);
    if (idx == kNoBrick) {
        atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
        return;
    }
    const size_t pos = (size_t)idx * kBrickCells + local;
    double sx = sums[4 * j], sy = sums[4 * j + 1], sz = sums[4 * j + 2], n = sums[4 * j + 3];
    const bool fresh = __float_as_int(cells[pos].w) == 0;
    if (fresh) {
        const unsigned long long v = (unsigned long long)atomicAdd((unsigned long long*)&counters[1], 1ull) + 1ull;
        if (v > max_voxels) atomicCAS(dev_status, 0, B2_ERR_CAPACITY);
    } else {
        sx = dadd(acc[4 * pos], sx);
        sy = dadd(acc[4 * pos + 1], sy);
        sz = dadd(acc[4 * pos + 2], sz);
        n = acc[4 * pos + 3] + n;
    }
    acc[4 * pos] = sx; acc[4 * pos + 1] = sy; acc[4 * pos + 2] = sz; acc[4 * pos + 3] = n;
    cells[pos] = make_float4((float)ddiv(sx, n), (float)ddiv(sy, n), (float)ddiv(sz, n), __int_as_float((int)n));
}

// Occupied voxels of the pool bricks: packed voxel key (for the canonical sort) + cell position.
__global__ void __launch_bounds__(kBlock)
brick_export_collect_kernel(const float4* __restrict__ cells, const unsigned long long* __restrict__ bkeys,
                            const long long* __restrict__ counters, unsigned pool_bricks,
                            unsigned long long* __restrict__ out_keys, unsigned* __restrict__ out_pos,
                            SortCtl* __restrict__ ctl, long long max_out) {
    const size_t pos = (size_t)blockIdx.x * kBlock + threadIdx.x;
    const unsigned long long used = min((unsigned long long)counters[0], (unsigned long long)pool_bricks);
    if (pos >= used * kBrickCells) return;
    if (__float_as_int(cells[pos].w) == 0) return;
    const unsigned long long bk = bkeys[pos / kBrickCells];
    const unsigned local = (unsigned)(pos % kBrickCells);
    const unsigned long long bx = (bk >> 42) & 0x1FFFFF, by = (bk >> 21) & 0x1FFFFF, bz = bk & 0x1FFFFF;
    // voxel coordinate + 2^20 bias = (brick coordinate + 2^20 bias) * 8 + local - 7 * 2^20
    const long long off = 7LL << 20;
    const unsigned long long vx = (unsigned long long)((long long)(bx * 8 + (local >> 6)) - off);
    const unsigned long long vy = (unsigned long long)((long long)(by * 8 + ((local >> 3) & 7)) - off);
    const unsigned long long vz = (unsigned long long)((long long)(bz * 8 + (local & 7)) - off);
    const int base = atomicAdd(&ctl->n, 1);
    if (base >= max_out) return;
    out_keys[base] = (vx << 42) | (vy << 21) | vz;
    out_pos[base] = (unsigned)pos;
}

// Compaction of the occupied voxels: packed 63-bit key (for the canonical sort) + slab position.
__global__ void __launch_bounds__(kBlock)
map_export_collect_kernel(const CellEntry* __restrict__ entries, unsigned capacity,
                          const unsigned char* __restrict__ lid, int brick, int cap_slots,
                          unsigned long long* __restrict__ out_keys, unsigned* __restrict__ out_pos,
                          SortCtl* __restrict__ ctl, long long max_out) {
    const unsigned h = blockIdx.x * kBlock + threadIdx.x;
    if (h >= capacity) return;
    const CellEntry e = entries[h];
    if (e.key == kEmptyKey || e.cnt == 0) return;
    const int cx = (int)((e.key >> 42) & 0x1FFFFF) - (1 << 20);
    const int cy = (int)((e.key >> 21) & 0x1FFFFF) - (1 << 20);
    const int cz = (int)(e.key & 0x1FFFFF) - (1 << 20);
    const int base = atomicAdd(&ctl->n, (int)e.cnt);
    for (unsigned s = 0; s < e.cnt; ++s) {
        if (base + (long long)s >= max_out) break;
        const size_t pos = (size_t)h * cap_slots + s;
        const int l = lid[pos];
        const int vx = cx * brick + l / (brick * brick), vy = cy * brick + (l / brick) % brick, vz = cz * brick + l % brick;
        const unsigned long long k = ((unsigned long long)(unsigned)(vx + (1 << 20)) << 42) |
                                     ((unsigned long long)(unsigned)(vy + (1 << 20)) << 21) |
                                     (unsigned long long)(unsigned)(vz + (1 << 20));
        out_keys[base + s] = k;
        out_pos[base + s] = (unsigned)pos;
    }
}

__global__ void map_export_ctl_kernel(SortCtl* ctl, long long max_out) {
    if (ctl->n > max_out) ctl->n = (int)max_out;
    ctl->npasses = 8;
    ctl->nseg = ctl->n;
    ctl->status = 0;
}

__global__ void __launch_bounds__(kBlock)
map_export_write_kernel(const unsigned long long* __restrict__ keysA, const unsigned long long* __restrict__ keysB,
                        const unsigned* __restrict__ valsA, const unsigned* __restrict__ valsB,
                        const SortCtl* __restrict__ ctl, const float4* __restrict__ cent,
                        const double* __restrict__ acc, float4* __restrict__ out_pts, int* __restrict__ out_keys,
                        double* __restrict__ out_sums) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= ctl->n) return;
    const bool odd = ctl->npasses & 1;
    const unsigned long long k = (odd ? keysB : keysA)[i];
    const unsigned pos = (odd ? valsB : valsA)[i];
    if (out_pts) {
        float4 p = cent[pos];
        p.w = __int_as_float((int)acc[4 * (size_t)pos + 3]);
        out_pts[i] = p;
    }
    if (out_sums)
        for (int d = 0; d < 4; ++d) out_sums[4 * (size_t)i + d] = acc[4 * (size_t)pos + d];
    if (out_keys) {
        out_keys[3 * i] = (int)((k >> 42) & 0x1FFFFF) - (1 << 20);
        out_keys[3 * i + 1] = (int)((k >> 21) & 0x1FFFFF) - (1 << 20);
        out_keys[3 * i + 2] = (int)(k & 0x1FFFFF) - (1 << 20);
    }
}

__global__ void gather_pos_kernel(const unsigned long long* keysA, const unsigned long long* keysB,
                                  unsigned* valsA, unsigned* valsB, const unsigned* pos, const SortCtl* ctl) {
    // sort ran with implicit iota values: translate sorted element index -> slab position
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ctl->n) return;
    unsigned* v = (ctl->npasses & 1) ? valsB : valsA;
    v[i] = pos[v[i]];
}

}  // namespace

void map_free(Context* c) {
    MapState* m = c->map;
    if (!m) return;
    cudaFree(m->entries); cudaFree(m->btab); cudaFree(m->bkeys); cudaFree(m->lut); cudaFree(m->lid); cudaFree(m->cent); cudaFree(m->acc);
    cudaFree(m->counters); cudaFree(m->meta); cudaFree(m->pose); cudaFree(m->result);
    delete m;
    c->map = nullptr;
}

int map_reset(Context* c) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    if (m->brick == 1) {
        brick_clear_kernel<<<div_up(m->capacity, 256), 256, 0, c->stream>>>(m->btab, m->capacity);
        B2_LAUNCH_CHECK(c);
        // an empty cell is a zero point count; the accumulators need no clearing (a fresh cell overwrites them)
        B2_CUDA_OK(c, cudaMemsetAsync(m->cent, 0, sizeof(float4) * (size_t)m->pool_bricks * kBrickCells, c->stream));
    } else {
        map_clear_kernel<<<div_up(m->capacity, 256), 256, 0, c->stream>>>(m->entries, m->capacity, m->cap_slots);
        B2_LAUNCH_CHECK(c);
        B2_CUDA_OK(c, cudaMemsetAsync(m->lut, 0, (size_t)m->capacity * m->lut_stride, c->stream));
    }
    map_set_meta_kernel<<<1, 1, 0, c->stream>>>(m->meta, m->origin[0], m->origin[1], m->origin[2], m->voxel,
                                                m->brick, m->pose, m->counters);
    B2_LAUNCH_CHECK(c);
    m->host_voxels_bound = 0;
    return B2_OK;
}

int map_init(Context* c, double voxel, const double* origin, double max_corr, long long capacity_cells) {
    if (!(voxel > 0.0) || !(max_corr > 0.0) || !origin || capacity_cells < 1)
        return set_error(c, B2_ERR_INVALID, "map_init: voxel, max_corr must be > 0, origin non-NULL, capacity >= 1");
    int brick = (int)ceil(max_corr / voxel - 1e-9);
    if (brick < 1) brick = 1;
    if (brick > 6) return set_error(c, B2_ERR_INVALID, "map_init: max_corr/voxel = %g > 6 is not supported", max_corr / voxel);
    unsigned long long cap = 1024;
    while (cap < (unsigned long long)capacity_cells) cap <<= 1;
    const int cap_slots = brick * brick * brick;
    if (cap * (unsigned long long)cap_slots >= 0xFFFFFFFFull)
        return set_error(c, B2_ERR_INVALID, "map_init: capacity %llu cells x %d slots exceeds 2^32", cap, cap_slots);
    // brick == 1: `capacity_cells` keeps its meaning as the voxel budget (B2_ERR_CAPACITY beyond 70 %); the pool
    // gets one 8^3 brick per 64 budgeted voxels (a surface fills 60-100 of a brick's 512 cells; 24 KB per brick)
    // and the brick hash four slots per pool brick.
    unsigned long long pool = cap / 64;
    if (pool < 2048) pool = 2048;
    if (pool > (1ull << 22)) pool = 1ull << 22;
    unsigned long long bcap = 1024;
    while (bcap < 4 * pool) bcap <<= 1;
    // the old map goes first (its tables may be most of the memory the new one needs); the new one is only
    // published once every table exists and has been cleared, so a failed allocation leaves NO map (every
    // map call then reports B2_ERR_STATE) instead of a half-initialised one
    map_free(c);
    c->ws_generation++;  // invalidates any captured graph that baked the old table addresses
    MapState* m = new MapState();
    m->voxel = voxel; m->max_corr = max_corr; m->brick = brick; m->cap_slots = cap_slots;
    m->lut_stride = (cap_slots + 15) / 16 * 16;
    m->capacity = (unsigned)(brick == 1 ? bcap : cap);
    m->pool_bricks = brick == 1 ? (unsigned)pool : 0u;
    m->max_voxels = (unsigned long long)(0.7 * (double)cap);
    for (int d = 0; d < 3; ++d) m->origin[d] = origin[d];
    auto alloc_all = [&]() -> cudaError_t {
        cudaError_t e;
        if (brick == 1) {
            if ((e = cudaMalloc(&m->btab, sizeof(BrickEntry) * bcap)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->bkeys, sizeof(unsigned long long) * pool)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->cent, sizeof(float4) * pool * kBrickCells)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->acc, sizeof(double) * 4 * pool * kBrickCells)) != cudaSuccess) return e;
        } else {
            if ((e = cudaMalloc(&m->entries, sizeof(CellEntry) * cap)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->lut, (size_t)cap * m->lut_stride)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->lid, (size_t)cap * cap_slots)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->cent, sizeof(float4) * cap * cap_slots)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&m->acc, sizeof(double) * 4 * cap * cap_slots)) != cudaSuccess) return e;
        }
        if ((e = cudaMalloc(&m->counters, sizeof(long long) * 2)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&m->meta, sizeof(GridMeta))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&m->pose, sizeof(double) * 16)) != cudaSuccess) return e;
        return cudaMalloc(&m->result, sizeof(IcpResultDev));
    };
    const cudaError_t e = alloc_all();
    c->map = m;  // (map_free / map_reset work on c->map)
    if (e != cudaSuccess) {
        map_free(c);
        cudaGetLastError();
        return set_error(c, B2_ERR_CUDA, "map_init: allocating the tables of %llu cells failed: %s", cap,
                         cudaGetErrorString(e));
    }
    const int rc = map_reset(c);
    if (rc != B2_OK) map_free(c);
    return rc;
}

int map_grid_view(Context* c, GridView* g) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    g->entries = m->entries;
    g->mask = m->capacity - 1;
    g->pts = m->cent;
    g->meta = m->meta;
    g->is_map = m->brick == 1 ? 2 : 1;
    g->btab = m->btab;
    g->pool_bricks = m->pool_bricks;
    return B2_OK;
}

double* map_pose_dev(Context* c) { return c->map ? c->map->pose : nullptr; }
IcpResultDev* map_result_dev(Context* c) { return c->map ? c->map->result : nullptr; }
double map_voxel(Context* c) { return c->map ? c->map->voxel : 0.0; }
bool map_is_empty_host(Context* c) { return !c->map || c->map->host_voxels_bound == 0; }
void map_set_voxels_bound(Context* c, long long voxels) { if (c->map) c->map->host_voxels_bound = voxels; }
double map_max_corr(Context* c) { return c->map ? c->map->max_corr : 0.0; }

// The merge of n (upper bound; ctl->nseg live) unique voxels (keys3, sums) into the tables.
static int merge_launch(Context* c, MapState* m, const int* keys3, const double* sums, const SortCtl* ctl, int n) {
    if (m->brick == 1) {
        pdl_launch(brick_ensure_kernel, dim3(div_up(n, kBlock)), dim3(kBlock), c->stream, keys3, ctl, m->btab,
                   m->capacity - 1, m->bkeys, m->pool_bricks, m->counters, c->dev_status);
        B2_LAUNCH_CHECK(c);
        pdl_launch(brick_merge_kernel, dim3(div_up(n, kBlock)), dim3(kBlock), c->stream, keys3, sums, ctl,
                   (const BrickEntry*)m->btab, m->capacity - 1, m->cent, m->acc, m->counters, m->max_voxels,
                   c->dev_status, (const IcpResultDev*)nullptr, (double*)nullptr);
    } else {
        const unsigned long long max_cells = (unsigned long long)(0.7 * (double)m->capacity);
        pdl_launch(map_merge_kernel, dim3(div_up(n, kBlock)), dim3(kBlock), c->stream, keys3, sums, ctl, m->entries,
                   m->capacity - 1, m->lut, m->lid, m->cent, m->acc, m->counters, m->brick, m->cap_slots,
                   m->lut_stride, max_cells, c->dev_status);
    }
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

// Fuse n_max (or *n_dev) points, transformed by T_dev (device, may be null), into the map.
int map_fuse(Context* c, const float4* pts, int n_max, const int* n_dev, const double* T_dev) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    if (n_max <= 0) return B2_OK;
    B2_WS(c, dpts, float4, WS_TMP_PTS, sizeof(float4) * (size_t)n_max);
    B2_WS(c, dkeys, int, WS_TMP_KEYS, sizeof(int) * 3 * (size_t)n_max);
    B2_WS(c, dsums, double, WS_SUMS, sizeof(double) * 4 * (size_t)n_max);
    DownsampleOut out{dpts, dkeys, dsums, nullptr};
    VoxelSortOut vs;
    B2_TRY(voxel_downsample(c, pts, n_max, n_dev, nullptr, 1, m->voxel, /*origin_mode=*/1, m->origin, T_dev, out, &vs));
    B2_TRY(merge_launch(c, m, dkeys, dsums, vs.ctl, n_max));
    m->host_voxels_bound += n_max;
    return B2_OK;
}

int map_fuse_stream(Context* c, const float4* pts, int n_max, const int* n_dev, const double* T_dev, int par,
                    int pre_par, const IcpResultDev* commit_res, bool* committed) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    if (n_max <= 0) return B2_OK;
    B2_WS(c, dkeys, int, WS_TMP_KEYS, sizeof(int) * 3 * (size_t)n_max);
    B2_WS(c, dsums, double, WS_SUMS, sizeof(double) * 4 * (size_t)n_max);
    SortCtl* ctl = nullptr;
    if (m->brick == 1) {  // the collecting pass creates the bricks: only the merge is left
        B2_TRY(stream_voxel_records(c, pts, n_max, n_dev, T_dev, m->origin, m->voxel, dkeys, dsums, &ctl, m->btab,
                                    m->capacity - 1, m->bkeys, m->pool_bricks, m->counters, par, pre_par));
        pdl_launch(brick_merge_kernel, dim3(div_up(n_max, kBlock)), dim3(kBlock), c->stream, (const int*)dkeys,
                   (const double*)dsums, (const SortCtl*)ctl, (const BrickEntry*)m->btab, m->capacity - 1, m->cent, m->acc,
                   m->counters, m->max_voxels, c->dev_status, commit_res, m->pose);
        B2_LAUNCH_CHECK(c);
        if (committed) *committed = commit_res != nullptr;
    } else {
        B2_TRY(stream_voxel_records(c, pts, n_max, n_dev, T_dev, m->origin, m->voxel, dkeys, dsums, &ctl, nullptr, 0,
                                    nullptr, 0, nullptr, par, pre_par));
        B2_TRY(merge_launch(c, m, dkeys, dsums, ctl, n_max));
    }
    m->host_voxels_bound += n_max;
    return B2_OK;
}

namespace {
__global__ void set_nseg_kernel(SortCtl* ctl, int n) {
    ctl->n = n;
    ctl->nseg = n;
    ctl->status = 0;
}
}  // namespace

// Merge n exported voxel records (unique keys; (sum x, sum y, sum z, count) per voxel) into the map:
// acc += record, centroid refreshed — the K10 centroid merge weighted by the record's count.
int map_merge_records(Context* c, const int* keys3, const double* sums, int n) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    if (n <= 0) return B2_OK;
    B2_WS(c, ctl, SortCtl, WS_CTL, sizeof(SortCtl));
    set_nseg_kernel<<<1, 1, 0, c->stream>>>(ctl, n);
    B2_LAUNCH_CHECK(c);
    B2_TRY(merge_launch(c, m, keys3, sums, ctl, n));
    m->host_voxels_bound += n;
    return B2_OK;
}

int map_counts(Context* c, long long* cells, long long* voxels) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    long long* h = (long long*)c->host_mailbox;
    B2_CUDA_OK(c, cudaMemcpyAsync(h, m->counters, sizeof(long long) * 2, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (cells) *cells = h[0];
    if (voxels) *voxels = h[1];
    return B2_OK;
}

int map_export(Context* c, float4* out_pts, int* out_keys, double* out_sums, long long max_out, long long* n_out) {
    MapState* m = c->map;
    if (!m) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    long long voxels = 0;
    B2_TRY(map_counts(c, nullptr, &voxels));
    if (voxels > max_out) return set_error(c, B2_ERR_INVALID, "map_export: map holds %lld voxels, buffer %lld", voxels, max_out);
    if (n_out) *n_out = voxels;
    if (voxels == 0) return B2_OK;
    const int n = (int)voxels;
    B2_WS(c, keysA, unsigned long long, WS_KEYS_A, sizeof(unsigned long long) * (size_t)n);
    B2_WS(c, keysB, unsigned long long, WS_KEYS_B, sizeof(unsigned long long) * (size_t)n);
    B2_WS(c, valsA, unsigned, WS_VALS_A, sizeof(unsigned) * (size_t)n);
    B2_WS(c, valsB, unsigned, WS_VALS_B, sizeof(unsigned) * (size_t)n);
    B2_WS(c, pos, unsigned, WS_EXPORT_KEYS, sizeof(unsigned) * (size_t)n);
    B2_WS(c, ctl, SortCtl, WS_CTL, sizeof(SortCtl));
    B2_CUDA_OK(c, cudaMemsetAsync(ctl, 0, sizeof(SortCtl), c->stream));
    if (m->brick == 1)
        brick_export_collect_kernel<<<div_up((long long)m->pool_bricks * kBrickCells, kBlock), kBlock, 0, c->stream>>>(
            m->cent, m->bkeys, m->counters, m->pool_bricks, keysA, pos, ctl, (long long)n);
    else
        map_export_collect_kernel<<<div_up(m->capacity, kBlock), kBlock, 0, c->stream>>>(
            m->entries, m->capacity, m->lid, m->brick, m->cap_slots, keysA, pos, ctl, (long long)n);
    B2_LAUNCH_CHECK(c);
    map_export_ctl_kernel<<<1, 1, 0, c->stream>>>(ctl, (long long)n);
    B2_LAUNCH_CHECK(c);
    B2_TRY(radix_sort_pairs(c, keysA, keysB, valsA, valsB, ctl, n, 8));
    gather_pos_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(keysA, keysB, valsA, valsB, pos, ctl);
    B2_LAUNCH_CHECK(c);
    map_export_write_kernel<<<div_up(n, kBlock), kBlock, 0, c->stream>>>(keysA, keysB, valsA, valsB, ctl, m->cent,
                                                                        m->acc, out_pts, out_keys, out_sums);
    B2_LAUNCH_CHECK(c);
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return B2_OK;
}

}  // namespace b2
