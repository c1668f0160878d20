// b2_icp.cu — the ICP iteration: correspondence search fused with the normal-equation reduction and the
// on-device SE(3) solve, and its host-side drivers.
//
// Replaces open3d::pipelines::registration::RegistrationICP (reference call site
// icp_processor.py:103-113).  One launch of a step kernel == one "Eval + ComputeTransformation" step of that loop:
//
//   per query:  P  = T * p                  (fp64, un-fused, fixed order — A.7)
//               j* = argmin_j |P - q_j|^2 over the 27-cell stencil, strict d^2 < r^2, ties -> lowest index
//                                           (KDTreeFlann::SearchHybrid(P, r, 1), A.4)
//               the point-to-point (17 sums: Sp, Sq, Sqp^T, Sd2, n) or point-to-plane (29 sums: upper JtJ, Jtr,
//               Sd2, n) terms of the correspondence go to a private accumulator column in shared memory
//   per CTA:    fixed-order fold of the columns -> one 32-double partial
//   last CTA of the pair (atomic ticket): fixed-order fold over the CTAs, then one thread evaluates fitness / rmse,
//               the convergence test of A.3, the Umeyama rotation (A.5) or the 6x6 solve + Rz*Ry*Rx (A.6) and
//               composes T <- dT * T.  Nothing returns to the host inside the loop; pairs that have converged turn
//               later launches into no-ops.
//
//   b2_icp_solve.cuh    the per-step solve
//   b2_icp_group.cuh    icp_iter_kernel: four lanes per query, probes in parallel (dense map, single pairs)
//   b2_icp_sweep.cuh    icp_sweep_kernel: one thread per query, pruned walk (batches on pair / slab grids)
//   b2_icp_dense.cuh    icp_dense_kernel: the streaming path (dense resident map): redundant fold + solve per CTA,
//                       no ticket / last-CTA tail, search from a speculative pass run beside the solve; by default ONE
//                       persistent launch per registration whose CTAs exchange tagged partial sums (b2_set_dense_driver)
//   b2_icp_persist.cuh  persistent driver — experiment, only built with -DB2_EXPERIMENTS (B2_EXPERIMENTS=1 build)
//   this file           launchers, the driver choice, the chained-launch / WHILE-node loop (icp_run)
//
// Algorithmic bytes of one launch: 16*Ns (queries) + 16*sum_q cand(q) (candidate points) +
// 16*27*Ns (cell records) + 256*CTAs.
#include <climits>
#include <cstdlib>

#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 640;   // 20 warps, one CTA per SM (<= 96 registers per thread)
constexpr int kCtasPerSm = 1;
constexpr int kTerms = 32;
constexpr unsigned kFull = 0xffffffffu;

#include "b2_icp_solve.cuh"
#include "b2_icp_group.cuh"
#include "b2_icp_sweep.cuh"
#include "b2_icp_dense.cuh"
#ifdef B2_EXPERIMENTS
#include "b2_icp_persist.cuh"  // opt-in persistent driver (measured slower inside the captured scan step)
#endif

__global__ void icp_init_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs,
                                const int* __restrict__ ns_dev, int ns_max, const double* __restrict__ init,
                                int init_stride, IcpState* __restrict__ states, int* __restrict__ tickets,
                                int* __restrict__ ndone, int n_pairs) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair == 0) *ndone = 0;
    if (pair >= n_pairs) return;
    IcpState s;
    for (int i = 0; i < 16; ++i) s.T[i] = init ? init[(size_t)pair * init_stride + i] : ((i % 5 == 0) ? 1.0 : 0.0);
    s.fitness = s.rmse = s.prev_fitness = s.prev_rmse = 0.0;
    int sb = src_offs ? src_offs[pair] : 0;
    int se = src_offs ? src_offs[pair + 1] : (ns_dev ? min(*ns_dev, ns_max) : ns_max);
    if (se > sb) {
        float4 p = src[sb + (se - sb) / 2];
        D3 P = xform_strict(s.T, (double)p.x, (double)p.y, (double)p.z);
        s.pivot[0] = P.x; s.pivot[1] = P.y; s.pivot[2] = P.z;
    } else {
        s.pivot[0] = s.pivot[1] = s.pivot[2] = 0.0;
    }
    s.iter = 0;
    s.done = 0;
    s.n_corr = 0;
    s.has_prev = 0;
    states[pair] = s;
    tickets[pair] = 0;
}

template <int MODE, int IS_MAP>
void launch_iter(Context* c, dim3 grid, const float4* src, const int* src_offs, const int* ns_dev, int ns_max,
                 const GridView& g, const float4* nrm, const int* tgt_offs, const IcpParams& prm, IcpState* st,
                 double* partial, int* tickets, IcpResultDev* res, int* corr, int* ndone, double* d2_out,
                 int search_only, cudaStream_t stream, unsigned long long loop_handle) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("B2_NO_PDL") != nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    static const int spread_partial = getenv("B2_SPREAD_PARTIAL") ? atoi(getenv("B2_SPREAD_PARTIAL")) : 0;  // experiment
    cudaLaunchKernelEx(&cfg, icp_iter_kernel<MODE, IS_MAP>, src, src_offs, ns_dev, ns_max,
                       (const CellEntry*)g.entries, g.mask, (const float4*)g.pts, (const GridMeta*)g.meta, nrm, tgt_offs, prm, st, partial, tickets, res, corr, ndone, d2_out,
                       search_only, loop_handle, spread_partial);
}

template <bool PERSIST>
void launch_dense(Context* c, dim3 grid, const float4* src, const int* ns_dev, int ns_max, const GridView& g,
                  const IcpParams& prm, DenseState* states, long long* seqw, double* partial, IcpResultDev* res, int* corr,
                  int* ndone, double* d2_out, int step, cudaStream_t stream, unsigned long long loop_handle) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kDenseThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("B2_NO_PDL") != nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = (no_pdl || PERSIST) ? 0 : 1;  // (the persistent variant starts after the init kernel has completed)
    cudaLaunchKernelEx(&cfg, icp_dense_kernel<PERSIST>, src, ns_dev, ns_max, (const BrickEntry*)g.btab, g.mask,
                       (const float4*)g.pts, (const GridMeta*)g.meta, prm, states, seqw, partial, res, corr, ndone, d2_out, step, loop_handle, c->dev_status);
}

template <int MODE, int KIND, int L>
void launch_sweep(Context* c, dim3 grid, const float4* src, const int* src_offs, const int* ns_dev, int ns_max,
                  const GridView& g, const float4* nrm, const int* tgt_offs, const IcpParams& prm, IcpState* st,
                  double* partial, int* tickets, IcpResultDev* res, int* corr, int* ndone, double* d2_out,
                  int search_only) {
    constexpr size_t smem = SweepCfg<MODE>::kSmem;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(SweepCfg<MODE>::kT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("B2_NO_PDL") != nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    cudaLaunchKernelEx(&cfg, icp_sweep_kernel<MODE, KIND, L>, src, src_offs, ns_dev, ns_max, (const CellEntry*)g.entries,
                       g.mask, (const float4*)g.pts, (const GridMeta*)g.meta, nrm, tgt_offs, prm, st, partial, tickets,
                       res, corr, ndone, d2_out, search_only);
}

// Candidate census of the 27-cell stencil (the data-dependent term of the byte model); kind as in GridView.
__global__ void __launch_bounds__(256)
stencil_count_kernel(const float4* __restrict__ src, int ns, const double* __restrict__ T,
                     const CellEntry* __restrict__ entries, const BrickEntry* __restrict__ btab, unsigned mask,
                     const float4* __restrict__ cells,
                     const GridMeta* __restrict__ meta, int kind, unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (q >= ns) return;
    const float4 p = __ldg(src + q);
    D3 P{(double)p.x, (double)p.y, (double)p.z};
    if (T) P = xform_strict(T, P.x, P.y, P.z);
    const int brick = meta->brick;
    int cx = voxel_coord(P.x, meta->origin[0], meta->cell);
    int cy = voxel_coord(P.y, meta->origin[1], meta->cell);
    int cz = voxel_coord(P.z, meta->origin[2], meta->cell);
    if (kind == 1 && brick > 1) {
        cx = floor_div(cx, brick);
        cy = floor_div(cy, brick);
        cz = floor_div(cz, brick);
    }
    unsigned cnt = 0;
    if (lane < 27) {
        const int nx = cx + lane / 9 - 1, ny = cy + (lane / 3) % 3 - 1, nz = cz + lane % 3 - 1;
        if (kind == 0) {
            if (nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536) {
                const unsigned long long key = pair_cell_key(0u, nx, ny, nz);
                const CellEntry e = grid_lookup(entries, mask, key);
                if (e.key == key) cnt = e.cnt;
            }
        } else {
            const unsigned long long key = map_cell_key(nx, ny, nz);
            if (kind == 2) {
                const unsigned bi = brick_lookup(btab, mask, map_cell_key(nx >> 3, ny >> 3, nz >> 3));
                if (bi != kNoBrick)
                    cnt = __float_as_int(__ldg(cells + (size_t)bi * kBrickCells + brick_local(nx, ny, nz)).w) != 0 ? 1u : 0u;
            } else {
                const CellEntry e = grid_lookup(entries, mask, key);
                if (e.key == key) cnt = e.cnt;
            }
        }
    }
    cnt = __reduce_add_sync(kFull, cnt);
    if (lane == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

}  // namespace

int map_stencil_count(Context* c, const float4* src, int ns, const double* T_dev, const GridView& g,
                      unsigned long long* total_dev) {
    stencil_count_kernel<<<div_up((long long)ns * 32, 256), 256, 0, c->stream>>>(src, ns, T_dev, g.entries, g.btab,
                                                                                  g.mask, g.pts, g.meta, g.is_map, total_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

#ifdef B2_EXPERIMENTS
namespace {

template <int MODE, int KIND>
int launch_persist(Context* c, const float4* src, const int* ns_dev, int ns_max, const GridView& g,
                   const float4* nrm, const IcpParams& prm, IcpState* st, double* partial, PersistSync* psync,
                   IcpResultDev* res, int* corr, int* ndone) {
    static int max_ctas = 0;  // co-resident CTAs of this instantiation (cooperative launch limit)
    if (max_ctas == 0) {
        int per_sm = 0, sms = 0;
        B2_CUDA_OK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_persist_kernel<MODE, KIND>, kThreads, 0));
        B2_CUDA_OK(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        max_ctas = per_sm * sms;
        if (max_ctas < 1) return set_error(c, B2_ERR_CUDA, "persistent ICP kernel cannot be made resident");
    }
    int grid = div_up(ns_max, kLeaders);  // one query per 8-lane group when it fits
    if (grid > max_ctas) grid = max_ctas;
    if (grid > kCtasPerSm * kSMs) grid = kCtasPerSm * kSMs;
    if (grid < 1) grid = 1;
    const CellEntry* entries = g.entries;
    unsigned mask = g.mask;
    const float4* gpts = g.pts;
    const GridMeta* meta = g.meta;
    IcpParams prm_v = prm;
    void* args[] = {(void*)&src, (void*)&ns_dev, (void*)&ns_max, (void*)&entries, (void*)&mask, (void*)&gpts,
                    (void*)&meta, (void*)&nrm, (void*)&prm_v, (void*)&st, (void*)&partial,
                    (void*)&psync, (void*)&res, (void*)&corr, (void*)&ndone};
    B2_CUDA_OK(c, cudaLaunchCooperativeKernel((const void*)icp_persist_kernel<MODE, KIND>, dim3(grid), dim3(kThreads),
                                              args, 0, c->stream));
    c->launches++;
    return B2_OK;
}

}  // namespace
#endif  // B2_EXPERIMENTS

// The sweep kernels keep one accumulator column per thread in dynamic shared memory (> 48 KB): opt in once
// per context (device attribute, not a stream operation — done at creation so it never lands in a capture).
int icp_configure(Context* c) {
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<0>::kSmem));
#ifdef B2_EXPERIMENTS
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<0, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<0>::kSmem));
#endif
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<0>::kSmem));
#ifdef B2_EXPERIMENTS
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<0, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<0>::kSmem));
#endif
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<1>::kSmem));
#ifdef B2_EXPERIMENTS
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<1, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<1>::kSmem));
#endif
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<1>::kSmem));
#ifdef B2_EXPERIMENTS
    B2_CUDA_OK(c, cudaFuncSetAttribute(icp_sweep_kernel<1, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<1>::kSmem));
#endif
    return B2_OK;
}

constexpr int kChainSteps = 40;  // captured scan step: steps launched back to back before the WHILE node takes over
constexpr long long kSweepMinQueries = 2LL * kCtasPerSm * kSMs * kThreads;  // ~190 k queries in flight

int icp_grid_x(int ns_max, int n_pairs, int kind) {
    const int pairs = n_pairs > 0 ? n_pairs : 1;
    int cap = (kCtasPerSm * kSMs) / pairs;
    if (cap < 2) cap = 2;
    if (pairs == 1) cap = kCtasPerSm * kSMs;
    int gx;
    if (kind == 2) {
        // dense map, single pair: one query per 4-lane group when it fits (152 groups per CTA, one CTA per SM)
        gx = div_up(ns_max / pairs + 1, kDenseLeaders);
    } else if (pairs == 1) {
        gx = div_up(ns_max, 32);  // one thread per query: spread the 32-query blocks over every SM
    } else {
        gx = div_up(ns_max / pairs + 1, 640);  // batches fill the machine through gridDim.y
    }
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    return gx;
}

int icp_run(Context* c, const float4* src, int ns_max, const int* src_offsets, const int* ns_dev, int n_pairs,
            const GridView& grid, const float4* tgt_normals, const int* tgt_offsets, const double* init_dev,
            const IcpParams& prm, IcpResultDev* results_dev, int* corr_out, double* d2_out, int search_only,
            int check_every) {
    if (n_pairs <= 0 || ns_max <= 0) return set_error(c, B2_ERR_INVALID, "icp: empty source");
    if (prm.mode == 1 && !tgt_normals && !search_only)
        return set_error(c, B2_ERR_INVALID, "point-to-plane ICP requires target normals");
    if (!(prm.max_corr > 0.0)) return set_error(c, B2_ERR_INVALID, "max_correspondence_distance must be > 0");
    // Pair / slab grids have two drivers: the 4-lane-group kernel (parallel probes: best when the queries do
    // not fill the machine and latency is the cost) and the thread-per-query sweep (pruned walk: best when
    // they do).  b2_set_icp_driver forces one (parity tests run both; results are identical).
    bool sweep = grid.is_map != 2 && (long long)ns_max >= kSweepMinQueries;
    int sweep_lanes = 1;  // lanes per query of the sweep kernel: 4 when a small problem has to fill the machine
    if (c->icp_driver != 0 && grid.is_map != 2) {
        sweep = c->icp_driver >= 2;
        sweep_lanes = c->icp_driver == 3 ? 4 : 1;
    }
    int gx = icp_grid_x(ns_max, n_pairs, sweep ? grid.is_map : 2);
    if (sweep && sweep_lanes == 4 && n_pairs == 1) {  // one 8-query block per warp before any warp gets a second one
        gx = div_up(ns_max, 8);
        if (gx > kCtasPerSm * kSMs) gx = kCtasPerSm * kSMs;
    }
    // the streaming path (one scan against the dense map, point-to-point) has its own step kernel
    const bool dense = grid.is_map == 2;
    if (dense && (prm.mode != 0 || n_pairs != 1 || src_offsets || search_only))
        return set_error(c, B2_ERR_INVALID, "the dense resident map registers one point-to-point scan at a time");
    B2_WS(c, states, IcpState, WS_ICP_STATE, sizeof(IcpState) * (size_t)n_pairs);
    // (+ 8 KB: the persistent dense variant keeps 2 x 148 x 17 tagged 16-byte entries here)
    B2_WS(c, partial, double, WS_ICP_PARTIAL, sizeof(double) * kTerms * (size_t)(gx > kCtasPerSm * kSMs ? gx : kCtasPerSm * kSMs) * n_pairs * 2 + 8192);
    B2_WS(c, tickets, int, WS_ICP_TICKET, sizeof(int) * ((size_t)n_pairs + 1));
    B2_WS(c, dstates, DenseState, WS_ICP_DENSE, sizeof(DenseState) * 2 + 2 * sizeof(long long));  // + seqw, grid-barrier word
    long long* seqw = reinterpret_cast<long long*>(dstates + 2);
    int* ndone = tickets + n_pairs;
    if (dense)
        icp_dense_init_kernel<<<1, 1, 0, c->stream>>>(src, ns_dev, ns_max, init_dev, dstates, seqw, ndone);
    else
        icp_init_kernel<<<div_up(n_pairs, 128), 128, 0, c->stream>>>(src, src_offsets, ns_dev, ns_max, init_dev, 16,
                                                                     states, tickets, ndone, n_pairs);
    B2_LAUNCH_CHECK(c);
    dim3 g(gx, n_pairs);
    // the dense-map registration as ONE persistent launch (b2_set_dense_driver; every CTA of the grid must be resident)
    const bool dense_persist = dense && c->dense_driver == B2_DENSE_DRIVER_PERSISTENT && gx <= kCtasPerSm * kSMs &&
                               live_handles(c->device) == 1 &&  // (see live_handles: two resident grids could starve each other)
                               prm.max_iter + 2 < 1024;          // (the partials' tags keep ten bits for the step)
    const int launches = search_only ? 1 : (dense_persist ? 1 : prm.max_iter + 1 + (dense ? 1 : 0));
#ifdef B2_EXPERIMENTS
    // (opt-in: on B200 the software grid barrier + solve of a step costs more than a kernel boundary)
    static const bool persist = getenv("B2_PERSIST") != nullptr;
    if (persist && grid.is_map != 2 && n_pairs == 1 && !src_offsets && !search_only && !d2_out && !c->capturing) {
        // single registration: one cooperative launch runs the whole A.3 loop
        B2_WS(c, psync, PersistSync, WS_ICP_SYNC, sizeof(PersistSync));
        B2_CUDA_OK(c, cudaMemsetAsync(psync, 0, sizeof(PersistSync), c->stream));
        int rc = B2_OK;
        if (prm.mode == 0) {
            if (grid.is_map == 1) rc = launch_persist<0, 1>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
            else rc = launch_persist<0, 0>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
        } else {
            if (grid.is_map == 1) rc = launch_persist<1, 1>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
            else rc = launch_persist<1, 0>(c, src, ns_dev, ns_max, grid, tgt_normals, prm, states, partial, psync, results_dev, corr_out, ndone);
        }
#ifdef B2_TAIL_TIMING
        if (getenv("B2_TIMING")) {
            cudaStreamSynchronize(c->stream);
            static unsigned long long h[300 * 8];
            cudaMemcpyFromSymbol(h, g_pts, sizeof(h));
            int nb = div_up(ns_max, kLeaders); if (nb > 296) nb = 296;
            unsigned long long t0 = ~0ull; for (int b = 0; b < nb; ++b) if (h[b * 8] < t0) t0 = h[b * 8];
            unsigned long long mx[7] = {0}, mn[7]; for (int k = 0; k < 7; ++k) mn[k] = ~0ull;
            for (int b = 0; b < nb; ++b) for (int k = 0; k < 7; ++k) { unsigned long long v = h[b * 8 + k]; if (!v) continue; v -= t0; if (v > mx[k]) mx[k] = v; if (v < mn[k]) mn[k] = v; }
            fprintf(stderr, "[b2 persist step3 ns] start %llu..%llu | main done %llu..%llu | folded %llu..%llu | ticketed %llu..%llu | grid-fold %llu | step %llu | released %llu..%llu\n",
                    mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], mn[3], mx[3], mx[4], mx[5], mn[6], mx[6]);
        }
#endif
        return rc;
    }
#endif  // B2_EXPERIMENTS
    int* host_done = (int*)c->host_mailbox + 512;  // scratch word in the pinned mailbox
    // B2_TIMING=1: bracket every launch with CUDA events and print the per-launch device times
    static const bool timing_env = getenv("B2_TIMING") != nullptr;
    const bool timing = timing_env && !c->capturing;
    std::vector<cudaEvent_t> evs;
    if (timing) {
        evs.resize(launches + 1);
        for (auto& e : evs) cudaEventCreate(&e);
        cudaEventRecord(evs[0], c->stream);
    }
    // Inside a captured scan step with a long iteration budget (the reference default is 500) only the first
    // kChainSteps steps are chained launches (programmatic dependent launch: the cheapest per step); the rest
    // of the A.3 loop is a CUDA-graph WHILE node whose body is ONE step kernel.  The solving CTA clears the
    // loop condition when the registration is done, so a replay runs the steps it needs instead of
    // max_iteration + 1 launches of which most return at once (500: 985 -> ~495 us per scan).
    // B2_ICP_WHILE=0 disables the loop node, =1 uses it for every step.
    static const char* while_env = getenv("B2_ICP_WHILE");
    const int while_mode = while_env ? atoi(while_env) : -1;
    const bool can_while = c->capturing && n_pairs == 1 && !sweep && !search_only && !corr_out && !d2_out && while_mode != 0 && !dense_persist;
    const int chained = !can_while ? launches : (while_mode == 1 ? 0 : (launches > kChainSteps ? kChainSteps : launches));
    auto launch_one = [&](cudaStream_t launch_stream, unsigned long long loop_handle, int step) {
        if (dense_persist) {
            launch_dense<true>(c, g, src, ns_dev, ns_max, grid, prm, dstates, seqw, partial, results_dev, corr_out, ndone,
                               d2_out, prm.max_iter + 2, launch_stream, 0ull);
            return;
        }
        if (dense) {
            launch_dense<false>(c, g, src, ns_dev, ns_max, grid, prm, dstates, seqw, partial, results_dev, corr_out, ndone,
                                d2_out, step, launch_stream, loop_handle);
            return;
        }
#define B2_ICP_LAUNCH(M, K)                                                                                   \
    launch_iter<M, K>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states, partial, \
                      tickets, results_dev, corr_out, ndone, d2_out, search_only, launch_stream, loop_handle)
#ifdef B2_EXPERIMENTS
#define B2_SWEEP4_BRANCH(M, K)                                                                                 \
        if (sweep_lanes == 4)                                                                                  \
            launch_sweep<M, K, 4>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states, \
                                  partial, tickets, results_dev, corr_out, ndone, d2_out, search_only);       \
        else
#else
#define B2_SWEEP4_BRANCH(M, K)
#endif
#define B2_ICP_SWEEP(M, K)                                                                                     \
    do {                                                                                                       \
        B2_SWEEP4_BRANCH(M, K)                                                                                 \
            launch_sweep<M, K, 1>(c, g, src, src_offsets, ns_dev, ns_max, grid, tgt_normals, tgt_offsets, prm, states, \
                                  partial, tickets, results_dev, corr_out, ndone, d2_out, search_only);       \
    } while (0)
        if (prm.mode == 0) {
            if (!sweep && grid.is_map == 1) B2_ICP_LAUNCH(0, 1);
            else if (!sweep) B2_ICP_LAUNCH(0, 0);
            else if (grid.is_map == 1) B2_ICP_SWEEP(0, 1);
            else B2_ICP_SWEEP(0, 0);
        } else {
            if (!sweep && grid.is_map == 1) B2_ICP_LAUNCH(1, 1);
            else if (!sweep) B2_ICP_LAUNCH(1, 0);
            else if (grid.is_map == 1) B2_ICP_SWEEP(1, 1);
            else B2_ICP_SWEEP(1, 0);
        }
#undef B2_ICP_LAUNCH
#undef B2_ICP_SWEEP
#undef B2_SWEEP4_BRANCH
    };
    for (int it = 0; it < chained; ++it) {
        launch_one(c->stream, 0ull, it);
        B2_LAUNCH_CHECK(c);
        if (timing) cudaEventRecord(evs[it + 1], c->stream);
        if (check_every > 0 && !search_only && !timing && (it + 1) % check_every == 0 && it + 1 < launches) {
            B2_CUDA_OK(c, cudaMemcpyAsync(host_done, ndone, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
            if (*host_done >= n_pairs) break;
        }
    }
    if (chained < launches) {  // the remaining steps: a WHILE node after the chained launches
        cudaStreamCaptureStatus cs;
        cudaGraph_t parent = nullptr;
        const cudaGraphNode_t* deps = nullptr;
        size_t n_deps = 0;
        B2_CUDA_OK(c, cudaStreamGetCaptureInfo(c->stream, &cs, nullptr, &parent, &deps, &n_deps));
        cudaGraphConditionalHandle handle;
        B2_CUDA_OK(c, cudaGraphConditionalHandleCreate(&handle, parent, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t loop_node;
        B2_CUDA_OK(c, cudaGraphAddNode(&loop_node, parent, deps, n_deps, &np));
        // the body is captured through the copy stream (capture only: nothing runs on that stream)
        B2_CUDA_OK(c, cudaStreamBeginCaptureToGraph(c->copy_stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                                    cudaStreamCaptureModeThreadLocal));
        launch_one(c->copy_stream, (unsigned long long)handle, -1);
        cudaGraph_t body = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->copy_stream, &body);
        B2_LAUNCH_CHECK(c);
        B2_CUDA_OK(c, e);
        B2_CUDA_OK(c, cudaStreamUpdateCaptureDependencies(c->stream, &loop_node, 1, cudaStreamSetCaptureDependencies));
    }
    if (timing) {
        cudaStreamSynchronize(c->stream);
        fprintf(stderr, "[b2 timing] icp launches (us):");
        for (int it = 0; it < launches; ++it) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, evs[it], evs[it + 1]);
            fprintf(stderr, " %.1f", ms * 1e3f);
        }
        fprintf(stderr, "\n");
#ifdef B2_TAIL_TIMING
        if (dense) {
            static unsigned long long h[64 * 148 * 16];
            static unsigned mis[64];
            cudaMemcpyFromSymbol(h, g_dense_ts, sizeof(h));
            cudaMemcpyFromSymbol(mis, g_dense_mis, sizeof(mis));
            fprintf(stderr, "[b2 dense] queries on the full path per launch:");
            for (int jj = 0; jj < 64 && jj < launches; ++jj) fprintf(stderr, " %u", mis[jj]);
            fprintf(stderr, "\n");
            const char* names[12] = {"entry", "waited", "folded", "solved", "searched", "exit", "w0:arrived", "w0:spec-issued", "w0:T-seen", "w0:cell", "w0:reloaded", "w0:finished"};
            for (int jj = 2; jj < 8 && jj < launches; ++jj) {
                unsigned long long t0 = ~0ull;
                for (int b = 0; b < gx; ++b) if (h[(jj * 148 + b) * 16] < t0) t0 = h[(jj * 148 + b) * 16];
                unsigned long long prev_exit = 0;
                for (int b = 0; b < gx; ++b) if (h[((jj - 1) * 148 + b) * 16 + 5] > prev_exit) prev_exit = h[((jj - 1) * 148 + b) * 16 + 5];
                fprintf(stderr, "[b2 dense launch %d] first entry %+lld ns after last exit of the previous launch |", jj, (long long)t0 - (long long)prev_exit);
                fprintf(stderr, " [cell changed for %u queries]", mis[jj]);
                for (int k = 0; k < 12; ++k) {
                    unsigned long long mn = ~0ull, mx = 0, sum = 0;
                    for (int b = 0; b < gx; ++b) { unsigned long long v = h[(jj * 148 + b) * 16 + k] - t0; if (v < mn) mn = v; if (v > mx) mx = v; sum += v; }
                    fprintf(stderr, " %s %llu/%llu/%llu", names[k], mn, sum / gx, mx);
                }
                fprintf(stderr, "  (min/mean/max ns over CTAs)\n");
            }
        } else {
            static unsigned long long h[8 + 2 * 1024];
            cudaMemcpyFromSymbol(h, g_tail_ts, sizeof(h));
            unsigned long long t0 = ~0ull, tmain_min = ~0ull, tmain_max = 0, tstart_max = 0;
            for (int b = 0; b < gx && b < 1024; ++b) {
                if (h[8 + 2 * b] < t0) t0 = h[8 + 2 * b];
                if (h[8 + 2 * b] > tstart_max) tstart_max = h[8 + 2 * b];
                if (h[8 + 2 * b + 1] < tmain_min) tmain_min = h[8 + 2 * b + 1];
                if (h[8 + 2 * b + 1] > tmain_max) tmain_max = h[8 + 2 * b + 1];
            }
            fprintf(stderr, "[b2 blocks start-minus-t0 ns]");
            for (int b = 0; b < gx && b < 1024; ++b) fprintf(stderr, " %llu", h[8 + 2 * b] - t0);
            fprintf(stderr, "\n");
            fprintf(stderr, "[b2 blocks main-done-minus-start ns]");
            for (int b = 0; b < gx && b < 1024; ++b) fprintf(stderr, " %llu", h[8 + 2 * b + 1] - h[8 + 2 * b]);
            fprintf(stderr, "\n");
            fprintf(stderr, "[b2 tail] last launch (ns from first CTA start): last CTA start %llu | main done min %llu max %llu | "
                            "last-CTA: ticket won %llu, reduced %llu, step done %llu | (sweep) fold start %llu, folded %llu, fenced %llu\n",
                    tstart_max - t0, tmain_min - t0, tmain_max - t0, h[0] - t0, h[1] - t0, h[2] - t0, h[3] - t0, h[4] - t0, h[5] - t0);
        }
#endif
        for (auto& e : evs) cudaEventDestroy(e);
    }
    return B2_OK;
}

#ifdef B2_TAIL_TIMING
// developer builds only: the globaltimer stamps of the dense step kernel, [64 launches][148 CTAs][16 slots]
int debug_dense_stamps(unsigned long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, g_dense_ts, sizeof(unsigned long long) * 64 * 148 * 16) == cudaSuccess ? 0 : -2;
}
int debug_dense_full_path(unsigned* out_host) {  // [64 launches]: queries that took the full path; cleared by the read
    if (cudaMemcpyFromSymbol(out_host, g_dense_mis, sizeof(unsigned) * 64) != cudaSuccess) return -2;
    static const unsigned zeros[64] = {0};
    return cudaMemcpyToSymbol(g_dense_mis, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -2;
}
#endif

}  // namespace b2

#ifdef B2_TAIL_TIMING
extern "C" int b2_debug_dense_stamps(unsigned long long* out_host) { return b2::debug_dense_stamps(out_host); }
extern "C" int b2_debug_dense_full_path(unsigned* out_host64) { return b2::debug_dense_full_path(out_host64); }
#endif
