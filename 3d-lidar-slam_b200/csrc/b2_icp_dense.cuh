// b2_icp_dense.cuh — the ICP step kernel of the streaming path (one scan against the dense resident map).
// Included by b2_icp.cu inside namespace b2 { namespace {.
#pragma once
// One registration is a chain of launches of ONE kernel, and the cost of a step is its chain of dependent
// latencies, not its bytes.  Round 1 ended a step with: CTA fold -> fence -> atomic ticket -> the LAST CTA folds
// the 148 partials -> one thread solves -> state to global -> kernel boundary -> every CTA reloads the state
// (13.9 us per step, half of it that tail, 147 SMs idle meanwhile).  Here the tail is gone:
//
//   launch j:  every CTA loads the 148 partials of launch j-1 and the state (one memory round trip, all loads in
//              flight together), folds them in a fixed order and evaluates the A.3 step REDUNDANTLY — same inputs,
//              same instruction sequence, same T_j on every SM — then searches with T_j and leaves its own partial.
//
// No ticket, no fence, no last-CTA election, no serial section that the other SMs wait for across the chip; the
// partials and the state ping-pong between two slots by launch parity, so a launch never overwrites what a
// slower CTA of the same launch still has to read.  CTA 0 alone publishes the state / result records.
// A registration of I steps is I + 2 launches: launch 0 only searches, launch I + 1 only folds and finishes.
#ifdef B2_TAIL_TIMING
__device__ unsigned long long g_dense_ts[64 * 148 * 8];
#define B2_DTS(k) do { if (threadIdx.x == 0 && j >= 0 && j < 64) g_dense_ts[((int)j * 148 + blockIdx.x) * 8 + (k)] = gtime(); } while (0)
#else
#define B2_DTS(k) do {} while (0)
#endif
struct DenseState {
    IcpState st;
    long long seq;  // launches applied so far (WHILE-node bodies, whose parity is not baked in, read it first)
};

__global__ void icp_dense_init_kernel(const float4* __restrict__ src, const int* __restrict__ ns_dev, int ns_max,
                                      const double* __restrict__ init, DenseState* __restrict__ states,
                                      long long* __restrict__ seqw, int* __restrict__ ndone) {
    // launch 0 reads slot 1 (parity of "launch -1")
    DenseState s;
    for (int i = 0; i < 16; ++i) s.st.T[i] = init ? init[i] : ((i % 5 == 0) ? 1.0 : 0.0);
    s.st.fitness = s.st.rmse = s.st.prev_fitness = s.st.prev_rmse = 0.0;
    const int n = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    if (n > 0) {
        const float4 p = src[n / 2];
        const D3 P = xform_strict(s.st.T, (double)p.x, (double)p.y, (double)p.z);
        s.st.pivot[0] = P.x; s.st.pivot[1] = P.y; s.st.pivot[2] = P.z;
    } else {
        s.st.pivot[0] = s.st.pivot[1] = s.st.pivot[2] = 0.0;
    }
    s.st.iter = 0; s.st.done = 0; s.st.n_corr = 0; s.st.has_prev = 0;
    s.seq = 0;
    states[1] = s;
    states[0] = s;
    *seqw = 0;
    *ndone = 0;
}

// The A.3 step on the CTA's private copy of the state (point-to-point).  Returns true when the registration is
// finished (the result record is written by CTA 0 only).
__device__ __noinline__ bool icp_step_local(const double* S, int ns, const IcpParams prm, IcpState& s,
                                            IcpResultDev* res, int* ndone, bool publish) {
    const double n = S[16];
    const double sd2 = S[15];
    const double fitness = ns > 0 ? n / (double)ns : 0.0;
    const double rmse = n > 0 ? sqrt(sd2 / n) : 0.0;
    bool done = false;
    if (s.has_prev && fabs(s.prev_fitness - fitness) < prm.rel_fitness && fabs(s.prev_rmse - rmse) < prm.rel_rmse)
        done = true;
    if (s.iter >= prm.max_iter) done = true;
    s.fitness = fitness;
    s.rmse = rmse;
    s.n_corr = (int)n;
    if (done) {
        s.done = 1;
        if (publish) {
#pragma unroll
            for (int k = 0; k < 12; ++k) res->T[k] = s.T[k];
            res->T[12] = 0.0; res->T[13] = 0.0; res->T[14] = 0.0; res->T[15] = 1.0;
            res->fitness = fitness;
            res->rmse = rmse;
            res->iterations = s.iter;
            res->n_corr = (int)n;
            atomicAdd(ndone, 1);
        }
        return true;
    }
    double U[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (n > 0) {
        const double inv = 1.0 / n;
        double mp0 = S[0] * inv, mp1 = S[1] * inv, mp2 = S[2] * inv;
        double mq0 = S[3] * inv, mq1 = S[4] * inv, mq2 = S[5] * inv;
        double C[9], R[9];
        C[0] = S[6] * inv - mq0 * mp0;  C[1] = S[7] * inv - mq0 * mp1;  C[2] = S[8] * inv - mq0 * mp2;
        C[3] = S[9] * inv - mq1 * mp0;  C[4] = S[10] * inv - mq1 * mp1; C[5] = S[11] * inv - mq1 * mp2;
        C[6] = S[12] * inv - mq2 * mp0; C[7] = S[13] * inv - mq2 * mp1; C[8] = S[14] * inv - mq2 * mp2;
        if (!procrustes_newton(C, R) && !polar_rotation(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8], R))
            rotation_from_covariance(C, R);
        const double p0 = s.pivot[0], p1 = s.pivot[1], p2 = s.pivot[2];
        mp0 += p0; mp1 += p1; mp2 += p2;
        mq0 += p0; mq1 += p1; mq2 += p2;
        U[0] = R[0]; U[1] = R[1]; U[2] = R[2];  U[3] = mq0 - (R[0] * mp0 + R[1] * mp1 + R[2] * mp2);
        U[4] = R[3]; U[5] = R[4]; U[6] = R[5];  U[7] = mq1 - (R[3] * mp0 + R[4] * mp1 + R[5] * mp2);
        U[8] = R[6]; U[9] = R[7]; U[10] = R[8]; U[11] = mq2 - (R[6] * mp0 + R[7] * mp1 + R[8] * mp2);
        s.pivot[0] = mq0; s.pivot[1] = mq1; s.pivot[2] = mq2;
    }
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = s.T[k];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double u0 = U[4 * r], u1 = U[4 * r + 1], u2 = U[4 * r + 2], u3 = U[4 * r + 3];
        s.T[4 * r + 0] = u0 * T[0] + u1 * T[4] + u2 * T[8];
        s.T[4 * r + 1] = u0 * T[1] + u1 * T[5] + u2 * T[9];
        s.T[4 * r + 2] = u0 * T[2] + u1 * T[6] + u2 * T[10];
        s.T[4 * r + 3] = u0 * T[3] + u1 * T[7] + u2 * T[11] + u3;
    }
    s.iter = s.iter + 1;
    s.prev_fitness = fitness;
    s.prev_rmse = rmse;
    s.has_prev = 1;
    return false;
}

// ---- correspondence search against the brick-major map -------------------------------------------------
// A warp serves eight queries per round, four lanes each (21.8 k queries = one round over 148 x 160 groups).
// Round 1's search was bound by instruction issue (2.9 M warp instructions per step: every lane hashed and
// probed its own 7 stencil cells, and the group leader alone transformed the point and accumulated the 17
// sums while three lanes idled).  Here the work of a query is spread evenly over its four lanes:
//   * lanes 0-2 each compute ONE axis of P = T p and of the query's cell (one IEEE division each);
//   * the 27-cell stencil lies in at most 2 x 2 x 2 bricks: lane g looks up the two corner bricks (y, z) = (g>>1,
//     g&1) — 2 hash probes per lane instead of 7, shared by shuffles;
//   * cell c = g + 4k of the stencil is then a directly indexed 16-byte load (all 7 of a lane in flight at once),
//     an emptiness test on the count in w and an un-fused fp64 distance — no key compare, no probe chain;
//   * arg-min over the group is a 2-step butterfly on (d2, stencil cell number): ties go to the lexicographically
//     smallest voxel key, i.e. the lowest index of the canonical (key-ordered) export, as in the oracle;
//   * the 17 normal-equation terms are added by all four lanes (5 / 4 / 4 / 4 terms) to the group's column.
struct DenseArgs {
    const BrickEntry* btab;
    unsigned bmask;
    const float4* cells;
    double r2;
    int* corr_out;
    double* d2_out;
};

__device__ __forceinline__ void dense_search_round(const DenseArgs& a, const double* __restrict__ sT,
                                                   const double* __restrict__ sPivot, const GridMeta& meta, float4 p,
                                                   bool valid, int q, double (*sAccL)[kLeaders], int col) {
    const int lane = threadIdx.x & 31;
    const int g = lane & (kGroup - 1);
    const int lead = lane & ~(kGroup - 1);
    // ---- P = T p and the query's cell: axis g on lane g ------------------------------------------------
    const float pxf = __shfl_sync(kFull, p.x, lead), pyf = __shfl_sync(kFull, p.y, lead), pzf = __shfl_sync(kFull, p.z, lead);
    valid = __shfl_sync(kFull, (int)valid, lead) != 0;
    double Pa = 0.0;
    int ca = 0;
    if (g < 3) {
        const double* row = sT + 4 * g;
        Pa = dadd(dadd(dadd(dmul(row[0], (double)pxf), dmul(row[1], (double)pyf)), dmul(row[2], (double)pzf)), row[3]);
        ca = voxel_coord(Pa, meta.origin[g], meta.cell);
    }
    const double Px = __shfl_sync(kFull, Pa, lead), Py = __shfl_sync(kFull, Pa, lead + 1), Pz = __shfl_sync(kFull, Pa, lead + 2);
    const int cx = __shfl_sync(kFull, ca, lead), cy = __shfl_sync(kFull, ca, lead + 1), cz = __shfl_sync(kFull, ca, lead + 2);
    const int lim = (1 << 20) - 10;
    const bool in_range = cx > -lim && cx < lim && cy > -lim && cy < lim && cz > -lim && cz < lim;
    const bool live = valid && in_range;  // (a query outside the key space has no neighbour)

    // ---- the (up to) eight bricks under the stencil ------------------------------------------------------
    const int bx0 = (cx - 1) >> 3, by0 = (cy - 1) >> 3, bz0 = (cz - 1) >> 3;
    const int bx1 = (cx + 1) >> 3;
    unsigned idx0 = kNoBrick, idx1 = kNoBrick;
    if (live) {
        const int by = (g & 2) ? (cy + 1) >> 3 : by0, bz = (g & 1) ? (cz + 1) >> 3 : bz0;
        idx0 = brick_lookup(a.btab, a.bmask, map_cell_key(bx0, by, bz));
        idx1 = bx1 == bx0 ? idx0 : brick_lookup(a.btab, a.bmask, map_cell_key(bx1, by, bz));
    }
    // ---- this lane's cells: all loads in flight before the first distance ----------------------------------
    float4 t[kCellsPerLane];
    bool have[kCellsPerLane];
#pragma unroll
    for (int k = 0; k < kCellsPerLane; ++k) {
        const int c = g + kGroup * k;
        const int c3 = (c * 171) >> 9;  // c / 3 for c < 128
        const int c9 = (c * 57) >> 9;   // c / 9
        const int nx = cx + c9 - 1, ny = cy + (c3 - 3 * c9) - 1, nz = cz + (c - 3 * c3) - 1;
        const int srcl = lead + ((((ny >> 3) != by0) ? 2 : 0) | (((nz >> 3) != bz0) ? 1 : 0));
        const unsigned i0 = __shfl_sync(kFull, idx0, srcl), i1 = __shfl_sync(kFull, idx1, srcl);
        const unsigned bi = (nx >> 3) != bx0 ? i1 : i0;
        have[k] = c < 27 && bi != kNoBrick;
        if (have[k]) t[k] = __ldg(a.cells + (size_t)bi * kBrickCells + brick_local(nx, ny, nz));
    }
    double best = INFINITY;
    int best_c = 0x7fffffff;
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int k = 0; k < kCellsPerLane; ++k) {
        if (!have[k] || __float_as_int(t[k].w) == 0) continue;
        const double d2 = dist2_strict(Px, Py, Pz, (double)t[k].x, (double)t[k].y, (double)t[k].z);
        if (d2 < best) {  // (cells of a lane come in increasing stencil order: ties keep the earlier one)
            best = d2;
            best_c = g + kGroup * k;
            bx = t[k].x; by = t[k].y; bz = t[k].z;
        }
    }
    // ---- arg-min over the group: (d2, stencil cell) lexicographic ------------------------------------------
    double gbest = best;
    int gc = best_c;
#pragma unroll
    for (int o = 1; o < kGroup; o <<= 1) {
        const double od = __shfl_xor_sync(kFull, gbest, o);
        const int oc = __shfl_xor_sync(kFull, gc, o);
        if (od < gbest || (od == gbest && oc < gc)) {
            gbest = od;
            gc = oc;
        }
    }
    const bool found = valid && gbest < a.r2;
    const int wl = lead + (gc & (kGroup - 1));  // the lane that holds stencil cell gc (meaningful when found)
    const float Qxf = __shfl_sync(kFull, bx, wl), Qyf = __shfl_sync(kFull, by, wl), Qzf = __shfl_sync(kFull, bz, wl);
    if (!valid) return;
    if (g == 0) {
        if (a.corr_out) a.corr_out[q] = found ? gc : -1;  // (stencil cell number of the neighbour: diagnostic only)
        if (a.d2_out) a.d2_out[q] = found ? gbest : 0.0;
    }
    if (!found) return;
    // ---- the 17 terms, spread over the four lanes (column `col` belongs to this group) ------------------------
    const double px = Px - sPivot[0], py = Py - sPivot[1], pz = Pz - sPivot[2];
    if (g == 0) {
        sAccL[0][col] += px;
        sAccL[1][col] += py;
        sAccL[2][col] += pz;
        sAccL[15][col] += gbest;
        sAccL[16][col] += 1.0;
    } else {
        const int ax = g - 1;
        const double qa = (double)(ax == 0 ? Qxf : (ax == 1 ? Qyf : Qzf)) - sPivot[ax];
        sAccL[3 + ax][col] += qa;
        sAccL[6 + 3 * ax][col] += qa * px;
        sAccL[7 + 3 * ax][col] += qa * py;
        sAccL[8 + 3 * ax][col] += qa * pz;
    }
}

// step >= 0: index of this launch in the chain (parity baked in by the host); step < 0: body of a CUDA-graph
// WHILE node — the launch index is read from *seqw first (one more round trip, only in the tail of long budgets).
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
icp_dense_kernel(const float4* __restrict__ src, const int* __restrict__ ns_dev, int ns_max,
                 const BrickEntry* __restrict__ btab, unsigned bmask, const float4* __restrict__ cells,
                 const GridMeta* __restrict__ metas, IcpParams prm,
                 DenseState* __restrict__ states, long long* __restrict__ seqw, double* __restrict__ partial,
                 IcpResultDev* __restrict__ results, int* __restrict__ corr_out, int* __restrict__ ndone,
                 double* __restrict__ d2_out, int step, unsigned long long loop_handle) {
    constexpr int NT = 17;
    __shared__ DenseState sDS;
    __shared__ GridMeta sMeta;
    __shared__ double sAccL[NT][kLeaders];
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ long long sSeq;
    static_assert(sizeof(DenseState) % 8 == 0 && sizeof(GridMeta) % 8 == 0, "copied as 8-byte words");

    long long j = step;
    B2_DTS(0);
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] = reinterpret_cast<const double*>(metas)[threadIdx.x - 32];
    for (int i = threadIdx.x; i < NT * kLeaders; i += kThreads) (&sAccL[0][0])[i] = 0.0;
    const int lane = threadIdx.x & 31;
    const int col = threadIdx.x / kGroup;
    const int groups_total = gridDim.x * kLeaders;
    const int warp_first = blockIdx.x * kLeaders + col - lane / kGroup;
    float4 p_first = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const int q = warp_first + lane / kGroup;
        if ((lane & (kGroup - 1)) == 0 && q < ns_max) p_first = __ldg(src + q);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    B2_DTS(1);
    if (step < 0) {
        if (threadIdx.x == 0) sSeq = __ldcg(seqw);
        __syncthreads();
        j = sSeq;
    }
    const int rd = (int)((j + 1) & 1), wr = (int)(j & 1);
    const int n_cta = (int)gridDim.x;
    // one round trip: the state record and this thread's share of the partials of launch j-1
    if (threadIdx.x < sizeof(DenseState) / 8)
        reinterpret_cast<double*>(&sDS)[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(states + rd) + threadIdx.x);
    const int ns = ns_dev ? min(__ldg(ns_dev), ns_max) : ns_max;
    if (j > 0) {
        grid_fold<NT>(partial + (size_t)rd * kCtasPerSm * kSMs * kTerms, n_cta, sRed, sSum);  // (two CTA barriers inside)
    } else {
        __syncthreads();
    }
    B2_DTS(2);
    if (sDS.st.done) {  // (a finished registration turns the remaining launches into no-ops)
        if (loop_handle && threadIdx.x == 0 && blockIdx.x == 0)
            cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, 0u);
        return;
    }
    if (threadIdx.x == 0) {
        const bool publish = blockIdx.x == 0;
        bool done = false;
        if (j > 0) done = icp_step_local(sSum, ns, prm, sDS.st, results, ndone, publish);
        sDS.seq = j + 1;
        if (publish) {
            states[wr] = sDS;
            *seqw = j + 1;
            if (loop_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, done ? 0u : 1u);
        }
    }
    __syncthreads();
    B2_DTS(3);
    if (sDS.st.done) return;
    const double* sT = sDS.st.T;
    const double* sPivot = sDS.st.pivot;

    DenseArgs a;
    a.btab = btab; a.bmask = bmask; a.cells = cells;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = d2_out;

    // same dealing of queries as icp_iter_kernel: 160 consecutive queries per CTA and round, the partial last
    // round of a multi-round source in 8-query warp chunks over all CTAs
    const int full_rounds = ns / groups_total;
    const bool spread = full_rounds >= 1;
    bool first = true;
    const int q_end = spread ? full_rounds * groups_total : ns;
    for (int q0 = warp_first; q0 < q_end; q0 += groups_total) {
        const int q = q0 + lane / kGroup;
        const bool valid = q < q_end;
        float4 p = p_first;
        if (!first && (lane & (kGroup - 1)) == 0 && valid) p = __ldg(src + q);
        first = false;
        dense_search_round(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
    }
    if (spread) {
        const int q0 = q_end + ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * (32 / kGroup);
        if (q0 < ns) {
            const int q = q0 + lane / kGroup;
            const bool valid = q < ns;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((lane & (kGroup - 1)) == 0 && valid) p = __ldg(src + q);
            dense_search_round(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
        }
    }
    __syncthreads();
    B2_DTS(4);
    cta_fold<NT, kLeaders>(&sAccL[0][0], partial + ((size_t)wr * kCtasPerSm * kSMs + blockIdx.x) * kTerms);
    B2_DTS(5);
}
