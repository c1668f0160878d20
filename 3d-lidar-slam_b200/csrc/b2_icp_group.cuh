// b2_icp_group.cuh — the 4-lane-group ICP step kernel (dense resident map; single pairs on pair / slab grids),
// the CTA / grid folds shared by all step kernels.  Included by b2_icp.cu inside namespace b2 { namespace {.
#pragma once
// ---------------------------------------------------------------------------
// the hot kernels
// ---------------------------------------------------------------------------
// Work decomposition: a warp serves EIGHT queries at a time, four lanes each, so that a 21.7k-point
// source is one round over the 148 x 160 groups of the grid (one 640-thread CTA per SM).  Lane g of a
// group probes stencil cells {g, g+4, ..., g+24 (<27)} in batches of four hash probes that are issued
// back to back (memory-level parallelism 4 per lane, 128 per warp) before any of them is consumed,
// and the group's arg-min is a 2-step shuffle butterfly.  The group leader (lane g == 0) owns the query:
// it loads and transforms the point, and adds the query's contribution to its private column of
// the shared-memory accumulators, so no per-query warp-wide reduction is needed.
//
// Two drivers share that body:
//   icp_iter_kernel     one launch per A.3 step, gridDim.y = pairs (batches; last CTA of a pair solves)
//   icp_persist_kernel  one cooperative launch per registration of a single pair: the CTAs stay
//                       resident, keep their query points in registers, and meet once per step at a
//                       ticket barrier whose last arriver reduces, solves and publishes T — no
//                       kernel boundary, no relaunch, immediate exit on convergence.
#ifdef B2_TAIL_TIMING
__device__ unsigned long long g_tail_ts[8 + 2 * 1024];
__device__ unsigned long long g_pts[300 * 8];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define B2_TS(slot) do { if (threadIdx.x == 0) g_tail_ts[slot] = gtime(); } while (0)
#define B2_TS_BLOCK(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_tail_ts[8 + 2 * blockIdx.x + (k)] = gtime(); } while (0)
#else
#define B2_TS(slot) do {} while (0)
#define B2_TS_BLOCK(k) do {} while (0)
#endif
constexpr int kGroup = 4;
constexpr int kCellsPerLane = (27 + kGroup - 1) / kGroup;  // 7
constexpr int kProbeBatch = 4;  // probes of a lane in flight together
constexpr int kLeaders = kThreads / kGroup;  // accumulator columns per CTA

struct SearchArgs {
    const CellEntry* entries;
    const float4* gpts;
    const float4* tgt_normals;
    unsigned mask;
    int tgt_base;
    unsigned pair;
    double r2;
    int* corr_out;
    double* d2_out;
    int search_only;
    const unsigned char* walk;  // sweep kernels: walk order table in shared memory
};

// One round: the warp's four groups each register one query (p = the leader's source point, valid
// per group) against the grid and add its normal-equation terms to column `col` of sAccL.
template <int MODE, int KIND>
__device__ __forceinline__ void search_round(const SearchArgs& a, const double* __restrict__ sT,
                                             const double* __restrict__ sPivot, const GridMeta& meta, float4 p,
                                             bool valid, int q, double (*sAccL)[kLeaders], int col) {
    const int lane = threadIdx.x & 31;
    const int g = lane & (kGroup - 1);
    const int lead = lane & ~(kGroup - 1);
    // ---- leader: transform the query, find its cell ---------------------------------------------
    double Px = 0.0, Py = 0.0, Pz = 0.0;
    int cx = 0, cy = 0, cz = 0;
    if (g == 0 && valid) {
        const D3 P = xform_strict(sT, (double)p.x, (double)p.y, (double)p.z);
        Px = P.x; Py = P.y; Pz = P.z;
        cx = voxel_coord(Px, meta.origin[0], meta.cell);
        cy = voxel_coord(Py, meta.origin[1], meta.cell);
        cz = voxel_coord(Pz, meta.origin[2], meta.cell);
        if (KIND == 1 && meta.brick > 1) {
            cx = floor_div(cx, meta.brick);
            cy = floor_div(cy, meta.brick);
            cz = floor_div(cz, meta.brick);
        }
        // map grids: one range test per query instead of one per probe (a query outside the 2^21-cell
        // key space cannot have a neighbour); poisoning cx makes every lane of the group skip the probes
        if (KIND) {
            const int lim = (1 << 20) - 2;
            if (cx < -lim || cx > lim || cy < -lim || cy > lim || cz < -lim || cz > lim) cx = INT_MIN;
        }
    }
    Px = __shfl_sync(kFull, Px, lead);
    Py = __shfl_sync(kFull, Py, lead);
    Pz = __shfl_sync(kFull, Pz, lead);
    cx = __shfl_sync(kFull, cx, lead);
    cy = __shfl_sync(kFull, cy, lead);
    cz = __shfl_sync(kFull, cz, lead);

    // ---- the group's 27 stencil cells: 4 (3 for g >= 3) per lane, probes issued together ----------
    double best = INFINITY;
    int best_idx = 0x7fffffff;
    float bx = 0.f, by = 0.f, bz = 0.f;
    // pair / slab grids: float32 screening state of this lane and the cells it resolved
    const float Pxf = (float)Px, Pyf = (float)Py, Pzf = (float)Pz;
    float bf = INFINITY, sf = INFINITY;
    unsigned bpos = 0;
    unsigned coff[kCellsPerLane + kProbeBatch];
    int ccnt[kCellsPerLane + kProbeBatch];
#pragma unroll
    for (int k = 0; k < kCellsPerLane + kProbeBatch; ++k) {
        coff[k] = 0;
        ccnt[k] = 0;
    }
    if (valid && cx != INT_MIN) {
#pragma unroll
        for (int k0 = 0; k0 < kCellsPerLane; k0 += kProbeBatch) {
            unsigned long long key[kProbeBatch];
            unsigned hpos[kProbeBatch];
            bool live[kProbeBatch];
#pragma unroll
            for (int k = 0; k < kProbeBatch; ++k) {
                const int c = g + kGroup * (k0 + k);
                const int c3 = (c * 171) >> 9;  // c / 3 for c < 128
                const int c9 = (c * 57) >> 9;   // c / 9
                const int nx = cx + c9 - 1, ny = cy + (c3 - 3 * c9) - 1, nz = cz + (c - 3 * c3) - 1;
                const bool in_stencil = (k0 + k) < kCellsPerLane && c < 27;
                if (KIND) {
                    live[k] = in_stencil;
                    key[k] = map_cell_key(nx, ny, nz);
                } else {
                    live[k] = in_stencil && nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536;
                    key[k] = pair_cell_key(a.pair, nx, ny, nz);
                }
                hpos[k] = cell_hash(key[k]) & a.mask;
            }
            {
                uint4 raw[kProbeBatch];
#pragma unroll
                for (int k = 0; k < kProbeBatch; ++k)
                    if (live[k]) raw[k] = __ldg(reinterpret_cast<const uint4*>(a.entries + hpos[k]));
#pragma unroll
                for (int k = 0; k < kProbeBatch; ++k) {
                    if (!live[k]) continue;
                    unsigned h = hpos[k];
                    unsigned long long kk = ((unsigned long long)raw[k].y << 32) | raw[k].x;
                    while (kk != key[k] && kk != kEmptyKey) {
                        h = (h + 1) & a.mask;
                        raw[k] = __ldg(reinterpret_cast<const uint4*>(a.entries + h));
                        kk = ((unsigned long long)raw[k].y << 32) | raw[k].x;
                    }
                    if (kk != key[k]) continue;
                    const unsigned off = raw[k].z;
                    const int cnt = (int)raw[k].w;
                    coff[k0 + k] = off;
                    ccnt[k0 + k] = cnt;
                    // float32 screening pass over the cell's points (uniform, no fp64): keep this lane's
                    // best and second-best float32 squared distance
#pragma unroll 4
                    for (int j = 0; j < cnt; ++j) {
                        const float4 t = __ldg(a.gpts + off + j);
                        const float fx = Pxf - t.x, fy = Pyf - t.y, fz = Pzf - t.z;
                        const float d = fx * fx + fy * fy + fz * fz;
                        if (d < bf) {
                            sf = bf;
                            bf = d;
                            bpos = off + j;
                        } else if (d < sf) {
                            sf = d;
                        }
                    }
                }
            }
        }
    }
    {
        // Exact phase.  Let M bound |d2_f32 - d2| for every candidate (rounding of P to float32 plus the
        // float32 arithmetic): every candidate's d2_f32 >= d2* - M, where d2* is the true minimum, so a
        // candidate with d2_f32 > gmin + 2M has d2 > d2* and cannot be the nearest neighbour.  Only the
        // contenders (normally one per group) are evaluated in fp64, in the strict order of the oracle.
        float gmin = bf;
#pragma unroll
        for (int o = 1; o < kGroup; o <<= 1) gmin = fminf(gmin, __shfl_xor_sync(kFull, gmin, o));
        if (valid && gmin < INFINITY) {
            const float e_abs = (fmaxf(fmaxf(fabsf(Pxf), fabsf(Pyf)), fabsf(Pzf)) + 4.0f * (float)meta.cell *
                                                                                         (float)meta.brick + 1e-3f) * 1.2e-7f;
            const float M = 3.5f * e_abs * sqrtf(gmin) + 3.0f * e_abs * e_abs + 1e-6f * gmin;
            const float thr = (gmin + 2.0f * M) * 1.000002f + 1e-30f;
            if (bf <= thr) {
                if (sf > thr) {  // one contender in this lane
                    const float4 t = __ldg(a.gpts + bpos);
                    best = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                    best_idx = __float_as_int(t.w);
                    bx = t.x; by = t.y; bz = t.z;
                } else {  // several within the margin (near-ties): exact pass over this lane's cells
#pragma unroll
                    for (int k = 0; k < kCellsPerLane; ++k) {
                        for (int j = 0; j < ccnt[k]; ++j) {
                            const float4 t = __ldg(a.gpts + coff[k] + j);
                            const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                            const int idx = __float_as_int(t.w);
                            if (d2 < best || (d2 == best && idx < best_idx)) {
                                best = d2;
                                best_idx = idx;
                                bx = t.x; by = t.y; bz = t.z;
                            }
                        }
                    }
                }
            }
        }
    }
    // ---- arg-min over the lanes of the group: (d2, idx) lexicographic -----------------------------
    double gbest = best;
    int gidx = best_idx;
#pragma unroll
    for (int o = 1; o < kGroup; o <<= 1) {
        const double od = __shfl_xor_sync(kFull, gbest, o);
        const int oi = __shfl_xor_sync(kFull, gidx, o);
        if (od < gbest || (od == gbest && oi < gidx)) {
            gbest = od;
            gidx = oi;
        }
    }
    const bool found = valid && gbest < a.r2;
    const unsigned winners = __ballot_sync(kFull, found && best == gbest && best_idx == gidx);
    const int wl = lead + __ffs((winners >> lead) & ((1u << kGroup) - 1u)) - 1;  // meaningful only when found
    const float Qxf = __shfl_sync(kFull, bx, wl & 31);
    const float Qyf = __shfl_sync(kFull, by, wl & 31);
    const float Qzf = __shfl_sync(kFull, bz, wl & 31);
    if (g != 0 || !valid) return;
    // ---- leader: outputs + contribution of this correspondence --------------------------------------
    if (a.corr_out) a.corr_out[q] = found ? gidx : -1;
    if (a.d2_out) a.d2_out[q] = found ? gbest : 0.0;
    if (a.search_only || !found) return;
    const double Qx = Qxf, Qy = Qyf, Qz = Qzf;
    if (MODE == 0) {
        const double px = Px - sPivot[0], py = Py - sPivot[1], pz = Pz - sPivot[2];
        const double qx = Qx - sPivot[0], qy = Qy - sPivot[1], qz = Qz - sPivot[2];
        sAccL[0][col] += px;       sAccL[1][col] += py;       sAccL[2][col] += pz;
        sAccL[3][col] += qx;       sAccL[4][col] += qy;       sAccL[5][col] += qz;
        sAccL[6][col] += qx * px;  sAccL[7][col] += qx * py;  sAccL[8][col] += qx * pz;
        sAccL[9][col] += qy * px;  sAccL[10][col] += qy * py; sAccL[11][col] += qy * pz;
        sAccL[12][col] += qz * px; sAccL[13][col] += qz * py; sAccL[14][col] += qz * pz;
        sAccL[15][col] += gbest;
        sAccL[16][col] += 1.0;
    } else {
        const float4 nq = __ldg(a.tgt_normals + a.tgt_base + gidx);
        const double nx = nq.x, ny = nq.y, nz = nq.z;
        double J[6];
        J[0] = Py * nz - Pz * ny;
        J[1] = Pz * nx - Px * nz;
        J[2] = Px * ny - Py * nx;
        J[3] = nx; J[4] = ny; J[5] = nz;
        const double r = (Px - Qx) * nx + (Py - Qy) * ny + (Pz - Qz) * nz;
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) sAccL[t++][col] += J[i] * J[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) sAccL[21 + i][col] += J[i] * r;
        sAccL[27][col] += gbest;
        sAccL[28][col] += 1.0;
    }
}

// CTA-wide fixed-order fold of the leader columns into one partial per term (warp w: terms w, w+16).
template <int NT, int COLS>
__device__ __forceinline__ void cta_fold(const double* __restrict__ acc, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int t = w; t < NT; t += (int)(blockDim.x >> 5)) {
        double s = 0.0;
#pragma unroll
        for (int c0 = 0; c0 < COLS; c0 += 32) s += acc[t * COLS + c0 + lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        if (lane == 0) out[t] = s;
    }
}

// Fixed-order sum of the per-CTA partials (all loads of a thread in flight before the first add).
template <int NT>
__device__ __forceinline__ void grid_fold(const double* __restrict__ partial, int n_cta, double (*sRed)[kTerms],
                                          double* __restrict__ sSum) {
    constexpr int kParts = kThreads / 32;
    constexpr int kMaxPer = (kCtasPerSm * kSMs + kParts - 1) / kParts;
    const int j = threadIdx.x & 31, part = threadIdx.x >> 5;
    double s = 0.0;
    if (j < NT && part < kParts) {  // (a sweep CTA may have more warps than kParts: the extra ones idle)
        if (n_cta <= kCtasPerSm * kSMs) {
            double v[kMaxPer];
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) {
                const int b = part + k * kParts;
                v[k] = b < n_cta ? __ldcg(partial + (size_t)b * kTerms + j) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) s += v[k];
        } else {
            for (int b = part; b < n_cta; b += kParts) s += __ldcg(partial + (size_t)b * kTerms + j);
        }
    }
    if (part < kParts) sRed[part][j] = s;
    __syncthreads();
    if (threadIdx.x < kTerms) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kParts; ++k) t += sRed[k][threadIdx.x];
        sSum[threadIdx.x] = t;
    }
    __syncthreads();
}

template <int MODE, int KIND>  // KIND 0: pair grid, 1: resident map with slabs (the dense map has b2_icp_dense.cuh)
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
icp_iter_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs, const int* __restrict__ ns_dev,
                int ns_max, const CellEntry* __restrict__ entries, unsigned mask,
                const float4* __restrict__ gpts, const GridMeta* __restrict__ metas, const float4* __restrict__ tgt_normals,
                const int* __restrict__ tgt_offs, IcpParams prm, IcpState* __restrict__ states,
                double* __restrict__ partial, int* __restrict__ tickets, IcpResultDev* __restrict__ results,
                int* __restrict__ corr_out, int* __restrict__ ndone, double* __restrict__ d2_out,
                int search_only, unsigned long long loop_handle, int spread_partial) {
    constexpr int NT = MODE == 0 ? 17 : 29;  // live accumulator terms
    const int pair = blockIdx.y;
    IcpState* st = states + pair;

    __shared__ IcpState sState;  // snapshot of the pair's state: every later read is on-chip
    __shared__ GridMeta sMeta;
    __shared__ double sAccL[NT][kLeaders];
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ int sLast;
    static_assert(sizeof(IcpState) % 8 == 0 && sizeof(GridMeta) % 8 == 0, "copied as 8-byte words");

    B2_TS_BLOCK(0);
    // Programmatic dependent launch: let the next step's grid be scheduled as soon as our CTAs retire
    // (all but the solving CTA leave right after the search), and do everything that does not depend
    // on the previous step — grid metadata, accumulator reset, this group's query point — before
    // waiting for the previous grid's writes (the pair state) to land.
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] =
            reinterpret_cast<const double*>(metas + (KIND ? 0 : pair))[threadIdx.x - 32];
    for (int i = threadIdx.x; i < NT * kLeaders; i += kThreads) (&sAccL[0][0])[i] = 0.0;
    const int lane = threadIdx.x & 31;
    const int col = threadIdx.x / kGroup;
    const int groups_total = gridDim.x * kLeaders;
    const int warp_first = blockIdx.x * kLeaders + col - lane / kGroup;
    float4 p_first = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!src_offs) {  // single pair: prefetch the first query (the buffer holds ns_max rows; validity is decided later)
        const int q = warp_first + lane / kGroup;
        if ((lane & (kGroup - 1)) == 0 && q < ns_max) p_first = __ldg(src + q);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // all remaining prologue loads are independent of each other: one memory round trip, not a chain
    if (threadIdx.x < sizeof(IcpState) / 8)
        reinterpret_cast<double*>(&sState)[threadIdx.x] = reinterpret_cast<const double*>(st)[threadIdx.x];
    int sb, se;
    if (src_offs) {
        sb = src_offs[pair];
        se = src_offs[pair + 1];
    } else {
        sb = 0;
        se = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    }
    __syncthreads();
    if (sState.done) {  // (as the body of a graph WHILE node: make sure the loop ends)
        if (loop_handle && threadIdx.x == 0 && blockIdx.x == 0)
            cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, 0u);
        return;
    }
    const double* sT = sState.T;
    const double* sPivot = sState.pivot;

    SearchArgs a;
    a.entries = entries; a.gpts = gpts; a.tgt_normals = tgt_normals; a.mask = mask;
    a.tgt_base = tgt_offs ? tgt_offs[pair] : 0;
    a.pair = (unsigned)pair;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = d2_out; a.search_only = search_only;

    // Queries are dealt to groups round-robin over (round, CTA, group): a CTA works on 160 consecutive queries per
    // round (neighbours in space: its warps share L1 lines).  When the source needs more than one round, the LAST,
    // partial round is dealt in 8-query warp chunks round-robin over the CTAs instead, so that every SM gets an
    // equal share of it rather than the first CTAs a full extra round (config 1: 29.8 k queries = 1.26 rounds).
    const int full_rounds = (se - sb) / groups_total;
    const bool spread = full_rounds >= 1 || spread_partial;
    bool first = !src_offs;
    const int q_end = spread ? sb + full_rounds * groups_total : se;
    for (int q0 = sb + warp_first; q0 < q_end; q0 += groups_total) {  // warp-uniform trip count
        const int q = q0 + lane / kGroup;
        const bool valid = q < q_end;
        float4 p = p_first;
        if (!first && (lane & (kGroup - 1)) == 0 && valid) p = __ldg(src + q);
        first = false;
        search_round<MODE, KIND>(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
    }
    if (spread) {
        const int q0 = q_end + ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * (32 / kGroup);
        if (q0 < se) {  // (at most one chunk per warp: the partial round holds fewer than gridDim.x * 20 chunks)
            const int q = q0 + lane / kGroup;
            const bool valid = q < se;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((lane & (kGroup - 1)) == 0 && valid) p = __ldg(src + q);
            search_round<MODE, KIND>(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
        }
    }
    if (search_only) return;

    __syncthreads();
    B2_TS_BLOCK(1);
    cta_fold<NT, kLeaders>(&sAccL[0][0], partial + ((size_t)pair * gridDim.x + blockIdx.x) * kTerms);
    __syncthreads();
    if (threadIdx.x == 0) {
        // one gpu-scope fence by the signalling thread: the CTA barrier above made every warp's
        // partial visible to it, and the fence is cumulative
        __threadfence();
        int t = atomicAdd(&tickets[pair], 1);
        sLast = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!sLast) return;
    B2_TS(0);
    grid_fold<NT>(partial + (size_t)pair * gridDim.x * kTerms, (int)gridDim.x, sRed, sSum);
    B2_TS(1);
    if (threadIdx.x == 0) {
        tickets[pair] = 0;
        icp_step<MODE>(sSum, se - sb, prm, sState, st, results + pair, ndone);
        // body of a CUDA-graph WHILE node (single pair): run another step unless the registration is done
        if (loop_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, st->done ? 0u : 1u);
    }
    B2_TS(2);
}
