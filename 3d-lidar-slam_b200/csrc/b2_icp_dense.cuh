// b2_icp_dense.cuh — the ICP step kernel of the streaming path (one scan against the dense resident map).
// Included by b2_icp.cu inside namespace b2 { namespace {.
#pragma once
// One registration is a chain of launches of ONE kernel, and the cost of a step is its chain of dependent
// latencies, not its bytes.  Round 1 ended a step with: CTA fold -> fence -> atomic ticket -> the LAST CTA folds
// the 148 partials -> one thread solves -> state to global -> kernel boundary -> every CTA reloads the state
// (13.9 us per step, half of it that tail, 147 SMs idle meanwhile).  Here:
//
//   * no tail: every CTA loads the 148 partials of launch j-1 and the state, folds them in a fixed order and
//     evaluates the A.3 step REDUNDANTLY — same inputs, same instruction sequence, same T_j on every SM.  No
//     ticket, no fence, no last-CTA election.  Partials and state ping-pong between two slots by launch parity, so
//     a launch never overwrites what a slower CTA of the same launch still has to read; CTA 0 alone publishes the
//     state / result records.  A registration of I steps is I + 2 launches: launch 0 only searches, launch I + 1
//     only folds and finishes.
//   * the fold + solve (one thread's worth of dependent fp64, ~2 us) runs on a 20th CONTROL WARP while the 19
//     search warps run the search SPECULATIVELY with the previous transform T_{j-1}: an ICP update moves a query
//     by far less than a 5 cm cell after the first steps, so the cell it lands in, the bricks, the 27 stencil cells
//     AND the ranking of those cells by distance are almost always what T_j will give.  Each lane fetches its nine
//     cells, screens them in float32 and keeps its two nearest cells plus a lower bound for the rest (SpecKept).
//     When T_j arrives (named barrier) the query has moved by a measured delta; if the bound minus delta still
//     exceeds the exact distance of the nearest kept cell (and the query has not left the fetched box), the kept
//     cells decide the search: one exact d2 per lane.  Anything else takes the full path around the new cell.
//     Probes, cell loads and the screening are off the critical path.
//   * distances are screened in float32 (full-rate pipes, no fp32->fp64 conversions) and only the contenders are
//     evaluated in un-fused fp64 in the oracle's order, so correspondences, d2 and ties are bit-identical to an
//     all-fp64 search (tests/test_gpu_parity.py::test_dense_map_correspondences_*).
#ifdef B2_TAIL_TIMING
__device__ unsigned long long g_dense_ts[64 * 148 * 16];
__device__ unsigned g_dense_mis[64];
#define B2_DTS(k) do { if ((threadIdx.x == 0 || ((k) == 2 || (k) == 3) && threadIdx.x == kDenseSearchThreads) && j >= 0 && j < 64) g_dense_ts[((int)j * 148 + blockIdx.x) * 16 + (k)] = gtime(); } while (0)
// latest warp of the CTA to pass a point (slots 12..15)
#define B2_DTS_MAX(k) do { if ((threadIdx.x & 31) == 0 && j >= 0 && j < 64) atomicMax(&g_dense_ts[((int)j * 148 + blockIdx.x) * 16 + (k)], gtime()); } while (0)
#else
#define B2_DTS(k) do {} while (0)
#define B2_DTS_MAX(k) do {} while (0)
#endif
// 640 threads = 96 registers each: 19 search warps + the control warp (fold + solve).  (A 21st warp would put
// six warps on one SM sub-partition: 6 x 32 x 96 registers > its 16,384.)  A query is served by THREE lanes (one
// z-plane of the stencil each), ten queries per warp, lanes 30 and 31 idle: 190 queries per CTA and round, i.e.
// 28,120 per round over 148 SMs.  With four lanes per query (152 per CTA, 22,496 per round) a quarter of the
// bench scans (21-24 k points after the 2 cm down-sample) overflowed into a second round on the first CTAs — a
// second, unspeculated round trip to the map on the critical path of every step of those scans (+3 us).
constexpr int kDenseThreads = kThreads;
constexpr int kDenseSearchThreads = kThreads - 32;
constexpr int kDenseLanes = 3;                                   // lanes per query
constexpr int kDenseQueriesPerWarp = 10;
constexpr int kDenseLeaders = (kDenseSearchThreads / 32) * kDenseQueriesPerWarp;  // 190 queries per CTA and round
constexpr int kDenseCols = 192;                                  // accumulator columns (padded to whole warps)
struct DenseState {
    IcpState st;
    long long seq;  // launches applied so far (WHILE-node bodies, whose parity is not baked in, read it first)
};

__global__ void icp_dense_init_kernel(const float4* __restrict__ src, const int* __restrict__ ns_dev, int ns_max,
                                      const double* __restrict__ init, DenseState* __restrict__ states,
                                      long long* __restrict__ seqw, int* __restrict__ ndone) {
    // persistent variant: a fresh epoch per registration, so that a partial left by an earlier registration can never
    // pass for one of this one (the partials carry (epoch, step) tags instead of being fenced by a barrier)
    reinterpret_cast<unsigned*>(seqw + 1)[0] += 1u;
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
// A warp serves ten queries per round, three lanes each (lanes 30 / 31 idle):
//   * lane g computes axis g of P = T p and of the query's cell (one IEEE division each);
//   * lane g owns the z-plane dz = g - 1 of the stencil: its 9 cells (dx, dy) have COMPILE-TIME offsets, lie in at
//     most 2 x 2 bricks (the lane's own 1-4 hash probes, issued together), and each is a directly indexed 16-byte
//     load, all 9 in flight at once;
//   * arg-min over the group: every lane reads the three lanes' (d2, stencil cell number) and keeps the
//     lexicographic minimum: ties go to the smallest voxel key, i.e. the lowest index of the canonical (key-ordered)
//     export, as in the oracle;
//   * the 17 normal-equation terms live in REGISTERS of the three lanes (6 / 6 / 5 terms) across rounds and are
//     written to the group's shared-memory column once.
struct DenseArgs {
    const BrickEntry* btab;
    unsigned bmask;
    const float4* cells;
    double r2;
    int* corr_out;
    double* d2_out;
};

struct PlaneCells {   // one lane's z-plane of a query's stencil
    float4 t[9];      // cells (dx, dy) = (k / 3 - 1, k % 3 - 1); w == 0: empty or no brick
    unsigned pos[9];  // position in the brick pool (kNoBrick: none)
    int cx, cy, cz;   // the query's cell they were fetched for
};

// axis g (< 3) of P = T p and of its cell; T row as four doubles
__device__ __forceinline__ void dense_axis(const double* __restrict__ row, float4 p, double origin, double cell,
                                           double& Pa, int& ca) {
    Pa = dadd(dadd(dadd(dmul(row[0], (double)p.x), dmul(row[1], (double)p.y)), dmul(row[2], (double)p.z)), row[3]);
    ca = voxel_coord(Pa, origin, cell);
}

// Fetch the nine cells of z-plane nz = cz + g - 1 around (cx, cy): four brick probes issued together, then nine
// direct loads.  No warp-collective operation inside (callable under divergence).
__device__ __forceinline__ void dense_load_plane(const DenseArgs& a, int cx, int cy, int cz, int g, bool live,
                                                 PlaneCells& pc) {
    pc.cx = cx; pc.cy = cy; pc.cz = cz;
    const int nz = cz + g - 1;
    const int bx0 = (cx - 1) >> 3, by0 = (cy - 1) >> 3, bzl = nz >> 3;
    const bool xc0 = (cx >> 3) != bx0, xc1 = ((cx + 1) >> 3) != bx0;
    const bool yc0 = (cy >> 3) != by0, yc1 = ((cy + 1) >> 3) != by0;
    unsigned b[4] = {kNoBrick, kNoBrick, kNoBrick, kNoBrick};  // (x, y) corner bricks 00, 10, 01, 11
    if (live) {
        const bool need[4] = {true, xc1, yc1, xc1 && yc1};
        unsigned long long key[4];
        unsigned h[4];
        uint4 raw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            key[i] = map_cell_key(bx0 + (i & 1), by0 + (i >> 1), bzl);
            h[i] = cell_hash(key[i]) & a.bmask;
            raw[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, kNoBrick, 0u);
            if (need[i]) raw[i] = __ldg(reinterpret_cast<const uint4*>(a.btab + h[i]));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned long long k = ((unsigned long long)raw[i].y << 32) | raw[i].x;
            while (k != key[i] && k != kEmptyKey) {  // collision chain (the table is a quarter full at most)
                h[i] = (h[i] + 1) & a.bmask;
                raw[i] = __ldg(reinterpret_cast<const uint4*>(a.btab + h[i]));
                k = ((unsigned long long)raw[i].y << 32) | raw[i].x;
            }
            b[i] = k == key[i] ? raw[i].z : kNoBrick;
        }
        if (!xc1) b[1] = b[0];
        if (!yc1) b[2] = b[0];
        if (!(xc1 && yc1)) b[3] = xc1 ? b[1] : b[2];
    }
#pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
        const bool xs = ix == 0 ? false : (ix == 1 ? xc0 : xc1);
        const unsigned r0 = xs ? b[1] : b[0], r1 = xs ? b[3] : b[2];
        const unsigned xo = ((unsigned)((cx + ix - 1) & 7) << 6) | (unsigned)(nz & 7);
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const bool ys = iy == 0 ? false : (iy == 1 ? yc0 : yc1);
            const unsigned bi = ys ? r1 : r0;
            const int k = 3 * ix + iy;
            pc.pos[k] = bi == kNoBrick ? kNoBrick : bi * (unsigned)kBrickCells + (xo | ((unsigned)((cy + iy - 1) & 7) << 3));
            pc.t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pc.pos[k] != kNoBrick) pc.t[k] = __ldg(a.cells + pc.pos[k]);
        }
    }
}

// What survives the barrier from the SPECULATIVE float32 screening (done with T_{j-1} in the shadow of the solve):
// the lane's two nearest cells and a lower bound for the distance of every cell that was not kept.
struct SpecKept {
    float x1, y1, z1, x2, y2, z2;  // centroids of the nearest / second-nearest cell of this lane's z-plane
    unsigned pos1, pos2;           // their positions in the brick pool
    int c12;                       // stencil cell numbers 9 * ix + 3 * iy of the two (c1 | c2 << 8; 0xff: none)
    float s2, s3;                  // sqrt(d~) (1 - 4e-6) - 2 e_abs of the second-nearest / of the third: with the
                                   // query's motion subtracted, lower bounds of their exact distances under T_j
    float e_abs;                   // per-coordinate float32 error bound of the screening
};

__device__ __forceinline__ float sqrt_approx(float x) {  // (relative error 2^-23: inside every margin below)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float dense_e_abs(float Pxf, float Pyf, float Pzf, double cellw) {
    return (fmaxf(fmaxf(fabsf(Pxf), fabsf(Pyf)), fabsf(Pzf)) + 4.0f * (float)cellw + 1e-3f) * 1.2e-7f;
}

// The three smallest of the lane's nine float32 squared distances by a min / max insertion network on integer keys
// (distance bits with the low four mantissa bits replaced by the cell's index k: non-negative floats order like
// integers; the truncation only LOWERS a distance by <= 2^-19 relative, which the bounds below allow for), then
// the two winners are re-read from L1 (their lines were fetched a moment ago) instead of being carried through
// the network.
__device__ __forceinline__ void dense_screen_top2(const DenseArgs& a, const PlaneCells& pc, float Pxf, float Pyf, float Pzf,
                                                  double cellw, SpecKept& kp) {
    int ka = 0x7fffffff, kb = 0x7fffffff, kc = 0x7fffffff;  // smallest, second, third key
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float fx = Pxf - pc.t[k].x, fy = Pyf - pc.t[k].y, fz = Pzf - pc.t[k].z;
        float d = fx * fx + fy * fy + fz * fz;
        if (__float_as_int(pc.t[k].w) == 0) d = INFINITY;
        const int key = (__float_as_int(d) & ~15) | k;
        kc = min(kc, max(kb, key));
        kb = min(kb, max(ka, key));
        ka = min(ka, key);
    }
    const bool has1 = ka < 0x7f800000, has2 = kb < 0x7f800000;
    const int k1 = ka & 15, k2 = kb & 15;
    unsigned p1 = 0u, p2 = 0u;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == k1) p1 = pc.pos[k];
        if (k == k2) p2 = pc.pos[k];
    }
    float4 t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1;
    if (has1) t1 = __ldg(a.cells + p1);
    if (has2) t2 = __ldg(a.cells + p2);
    kp.x1 = t1.x; kp.y1 = t1.y; kp.z1 = t1.z; kp.pos1 = p1;
    kp.x2 = t2.x; kp.y2 = t2.y; kp.z2 = t2.z; kp.pos2 = p2;
    const int c1 = has1 ? 9 * (k1 / 3) + 3 * (k1 % 3) : 0xff, c2 = has2 ? 9 * (k2 / 3) + 3 * (k2 % 3) : 0xff;
    kp.c12 = c1 | (c2 << 8);
    kp.e_abs = dense_e_abs(Pxf, Pyf, Pzf, cellw);
    const float d2 = __int_as_float(kb & ~15), d3 = __int_as_float(kc & ~15);  // (INFINITY stays INFINITY)
    kp.s2 = sqrt_approx(d2) * 0.999996f - 2.0f * kp.e_abs;
    kp.s3 = sqrt_approx(d3) * 0.999996f - 2.0f * kp.e_abs;
}

// One lane's nearest cell among the nine of a freshly fetched z-plane (the full path): float32 screening, then the
// contenders in exact fp64.  Warp-collective (all 32 lanes enter; lanes 30 / 31 and lanes of groups that do not take
// part come in with empty planes).
__device__ __forceinline__ void dense_lane_best(double cellw, double Px, double Py, double Pz, const PlaneCells& pc,
                                                int g, int lead, double& best, int& best_c, unsigned& best_pos,
                                                float& bx, float& by, float& bz) {
    // ---- float32 screening of this lane's nine cells: best and second-best squared distance ------------------
    const float Pxf = (float)Px, Pyf = (float)Py, Pzf = (float)Pz;
    float bf = INFINITY, sf = INFINITY;
    bx = by = bz = 0.f;
    unsigned bpos = 0;
    int bc = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float fx = Pxf - pc.t[k].x, fy = Pyf - pc.t[k].y, fz = Pzf - pc.t[k].z;
        float d = fx * fx + fy * fy + fz * fz;
        if (__float_as_int(pc.t[k].w) == 0) d = INFINITY;
        if (d < bf) {
            sf = bf;
            bf = d;
            bx = pc.t[k].x; by = pc.t[k].y; bz = pc.t[k].z;
            bpos = pc.pos[k];
            bc = 9 * (k / 3) + 3 * (k % 3);
        } else if (d < sf) {
            sf = d;
        }
    }
    const float gmin = fminf(fminf(__shfl_sync(kFull, bf, lead), __shfl_sync(kFull, bf, lead + 1)),
                             __shfl_sync(kFull, bf, lead + 2));
    // ---- exact phase: let M bound |d2_f32 - d2| for every candidate (rounding of P to float32 plus the float32
    // arithmetic).  Every candidate's d2_f32 >= d2* - M where d2* is the true minimum, so a candidate with
    // d2_f32 > gmin + 2M has d2 > d2* and cannot be the nearest neighbour; only the contenders (normally one per
    // group) are evaluated in fp64, in the strict order of the oracle.
    best = INFINITY;
    best_c = 0x7fffffff;
    best_pos = 0;
    if (gmin < INFINITY) {
        const float e_abs = dense_e_abs(Pxf, Pyf, Pzf, cellw);
        const float M = 3.5f * e_abs * sqrtf(gmin) + 3.0f * e_abs * e_abs + 1e-6f * gmin;
        const float thr = (gmin + 2.0f * M) * 1.000002f + 1e-30f;
        if (bf <= thr) {
            if (sf > thr) {  // one contender in this lane
                best = dist2_strict(Px, Py, Pz, (double)bx, (double)by, (double)bz);
                best_c = bc + g;
                best_pos = bpos;
            } else {  // several within the margin (near-ties): exact pass over the lane's cells
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (__float_as_int(pc.t[k].w) == 0) continue;
                    const double d2 = dist2_strict(Px, Py, Pz, (double)pc.t[k].x, (double)pc.t[k].y, (double)pc.t[k].z);
                    if (d2 < best) {  // (increasing stencil order: ties keep the earlier cell)
                        best = d2;
                        best_c = 9 * (k / 3) + 3 * (k % 3) + g;
                        best_pos = pc.pos[k];
                        bx = pc.t[k].x; by = pc.t[k].y; bz = pc.t[k].z;
                    }
                }
            }
        }
    }
}

// Arg-min over the group's three lanes + the correspondence's contribution to the sums.
// Warp-collective (all 32 lanes enter; `valid` is uniform over a group; g = lane's z-plane 0..2, lead = the
// group's first lane; lanes 30 / 31 come in with g >= 3, valid == false and best == INFINITY).
__device__ __forceinline__ void dense_group_tail(const DenseArgs& a, const double* __restrict__ sPivot, double Px,
                                                 double Py, double Pz, double best, int best_c, unsigned best_pos,
                                                 float bx, float by, float bz, bool valid, int q, int g, int lead,
                                                 double* __restrict__ colp, bool first) {
    // (d2, stencil cell) lexicographic.  Normally exactly one lane of the group holds a contender: its values are
    // the group's (one ballot, no comparison of doubles across lanes); otherwise (near-ties across z-planes, or no
    // candidate at all) every lane reads the three lanes' (d2, cell) and keeps the lexicographic minimum.
    const unsigned has = (__ballot_sync(kFull, best < INFINITY) >> lead) & 7u;
    double gbest;
    int gc;
    unsigned gpos;
    if (__all_sync(kFull, __popc(has) == 1)) {  // (warp-uniform: the shuffles below name all 32 lanes)
        const int wlane = lead + (__ffs(has) - 1);
        gbest = __shfl_sync(kFull, best, wlane);
        gc = __shfl_sync(kFull, best_c, wlane);
        gpos = __shfl_sync(kFull, best_pos, wlane);
    } else {
        gbest = __shfl_sync(kFull, best, lead);
        gc = __shfl_sync(kFull, best_c, lead);
        gpos = __shfl_sync(kFull, best_pos, lead);
#pragma unroll
        for (int o = 1; o < kDenseLanes; ++o) {
            const double od = __shfl_sync(kFull, best, lead + o);
            const int oc = __shfl_sync(kFull, best_c, lead + o);
            const unsigned op = __shfl_sync(kFull, best_pos, lead + o);
            if (od < gbest || (od == gbest && oc < gc)) {
                gbest = od;
                gc = oc;
                gpos = op;
            }
        }
    }
    const bool found = valid && gbest < a.r2;
    if (g == 0 && valid) {
        if (a.corr_out) a.corr_out[q] = found ? (int)gpos : -1;  // (cell position in the brick pool: diagnostic only)
        if (a.d2_out) a.d2_out[q] = found ? gbest : 0.0;
    }
    // the neighbour's coordinates come from the lane that owns stencil cell gc (its z-plane: gc % 3)
    const int wl = lead + (found ? gc % 3 : 0);
    float4 Q;
    Q.x = __shfl_sync(kFull, bx, wl); Q.y = __shfl_sync(kFull, by, wl); Q.z = __shfl_sync(kFull, bz, wl);
    if (g >= kDenseLanes) return;
    // ---- the 17 terms, spread over the three lanes: lane g holds axis g of q (q_g, q_g p^T) and p_g; lane 0 also
    // the squared distance, lane 1 the count; they go straight to the group's shared-memory column (rows by lane
    // role; the first round stores, a later round — sources beyond 190 queries per CTA — adds) ------------------
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0;
    if (found) {
        const double px = Px - sPivot[0], py = Py - sPivot[1], pz = Pz - sPivot[2];
        const double qa = (double)(g == 0 ? Q.x : (g == 1 ? Q.y : Q.z)) - sPivot[g];
        t0 = qa; t1 = qa * px; t2 = qa * py; t3 = qa * pz;
        t4 = g == 0 ? px : (g == 1 ? py : pz);
        t5 = g == 0 ? gbest : (g == 1 ? 1.0 : 0.0);
    } else if (!first) {
        return;
    }
    double* r0 = colp + (3 + g) * kDenseCols;        // q_g
    double* r1 = colp + (6 + 3 * g) * kDenseCols;    // q_g p^T (three rows)
    double* r4 = colp + g * kDenseCols;              // p_g
    double* r5 = colp + (15 + g) * kDenseCols;       // d2 (lane 0) / count (lane 1)
    if (first) {
        *r0 = t0; r1[0] = t1; r1[kDenseCols] = t2; r1[2 * kDenseCols] = t3; *r4 = t4;
        if (g < 2) *r5 = t5;
    } else {
        *r0 += t0; r1[0] += t1; r1[kDenseCols] += t2; r1[2 * kDenseCols] += t3; *r4 += t4;
        if (g < 2) *r5 += t5;
    }
}

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Tagged partials of the persistent variant: every fp64 term travels as a 16-byte entry {lo, tag, hi, tag}.  Each
// 8-byte half validates itself, so an entry is complete as soon as both tags read (epoch, step) — no fence, no
// arrival counter and no second round trip between "everybody has arrived" and "now load the data": the CTAs poll the
// data itself, and the last partial to arrive IS the grid barrier.
// (vector accesses are performed element by element and a naturally aligned 64-bit element is single-copy atomic: the
// elements are the two {value half, tag} words)
__device__ __forceinline__ void st_tagged(uint4* p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ uint4 ld_tagged(const uint4* p) {  // -> {lo, tag, hi, tag}
    unsigned long long w0, w1;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    return make_uint4((unsigned)w0, (unsigned)(w0 >> 32), (unsigned)w1, (unsigned)(w1 >> 32));
}

// PERSIST == false — one launch per step.  step >= 0: index of this launch in the chain (parity baked in by the host);
// step < 0: body of a CUDA-graph WHILE node — the launch index is read from *seqw first (one more round trip, only in
// the tail of long budgets).
// PERSIST == true — ONE launch per registration: the steps are trips of a loop, `step` is their maximum number, and
// the kernel boundary between two steps is replaced by a grid barrier (arrival counter in global memory; one CTA per
// SM, all co-resident) — and the barrier is implicit: the only thing CTAs exchange is their partial sums, which
// travel as self-validating tagged entries (st_tagged / ld_tagged), so a CTA simply polls the 148 partials it needs.
// The state record then never leaves shared memory, the query point, its doubles and the map metadata stay in
// registers, and the launches that would only find "done" do not exist.
template <bool PERSIST>
__global__ void __launch_bounds__(kDenseThreads, 1)
icp_dense_kernel(const float4* __restrict__ src, const int* __restrict__ ns_dev, int ns_max,
                 const BrickEntry* __restrict__ btab, unsigned bmask, const float4* __restrict__ cells,
                 const GridMeta* __restrict__ metas, IcpParams prm,
                 DenseState* __restrict__ states, long long* __restrict__ seqw, double* __restrict__ partial,
                 IcpResultDev* __restrict__ results, int* __restrict__ corr_out, int* __restrict__ ndone,
                 double* __restrict__ d2_out, int step, unsigned long long loop_handle, int* __restrict__ status) {
    constexpr int NT = 17;
    constexpr int kParts = kDenseSearchThreads / 32;  // 20
    __shared__ DenseState sDS;
    __shared__ double sAccL[NT][kDenseCols];
    __shared__ double sRed[kParts][kTerms];
    static_assert(sizeof(DenseState) % 8 == 0 && sizeof(GridMeta) % 8 == 0, "copied as 8-byte words");

    long long j = PERSIST ? 0 : step;
    B2_DTS(0);
    if (!PERSIST) asm volatile("griddepcontrol.launch_dependents;");
    const bool control = threadIdx.x >= kDenseSearchThreads;
    const int lane = threadIdx.x & 31;
    const int grp = min(lane / kDenseLanes, kDenseQueriesPerWarp - 1);  // (lanes 30, 31: g = 3, 4 -> idle)
    const int g = lane - kDenseLanes * grp;
    const int lead = kDenseLanes * grp;
    const int col = (threadIdx.x >> 5) * kDenseQueriesPerWarp + grp;  // (search threads)
    // The source size is final before the first launch of the chain starts (the init kernel ahead of it completes
    // first), so it can be read ahead of the grid dependency.  The queries are dealt evenly: every CTA takes
    // ceil(ns / CTAs) of them, in whole warps of ten (148 each for a 21.8 k-point source: 15 of the 19 search warps),
    // and only a source beyond 190 per CTA makes a second round.
    const int n_cta = (int)gridDim.x;
    const int ns = ns_dev ? min(__ldcg(ns_dev), ns_max) : ns_max;
    const int share = min(kDenseLeaders, ((ns + n_cta - 1) / n_cta + kDenseQueriesPerWarp - 1) / kDenseQueriesPerWarp *
                                             kDenseQueriesPerWarp);
    const int groups_total = n_cta * share;
    const bool has_query = col < share && g < kDenseLanes;
    const int q_first = blockIdx.x * share + col;  // this group's query of the first round
    // everything that does not depend on the previous launch: grid metadata, this group's query point
    const double m_cell = __ldg(&metas->cell);                              // voxel edge (= cell edge: brick == 1)
    const double m_org = __ldg(&metas->origin[lane < 30 ? lane % 3 : 0]);   // this lane's axis of the map origin
    float4 p_first = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!control && has_query && q_first < ns) p_first = __ldg(src + q_first);  // (one address per group)
    if (!PERSIST) asm volatile("griddepcontrol.wait;" ::: "memory");
    B2_DTS(1);
    if (!PERSIST && step < 0) j = __ldcg(seqw);

    DenseArgs a;
    a.btab = btab; a.bmask = bmask; a.cells = cells;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = d2_out;
    const unsigned epoch_tag = PERSIST ? __ldcg(reinterpret_cast<const unsigned*>(seqw + 1)) << 10 : 0u;  // (steps < 1024)
    if (PERSIST) {  // the state record stays in shared memory for the whole registration
        if (control && lane < (int)(sizeof(DenseState) / 8))
            reinterpret_cast<double*>(&sDS)[lane] = __ldcg(reinterpret_cast<const double*>(states + 1) + lane);
        __syncthreads();
    }
  for (;;) {  // (PERSIST: one trip per step; otherwise a single trip)
    const int rd = (int)((j + 1) & 1), wr = (int)(j & 1);
    const DenseState* prev = states + rd;
    if (control) {
        // ---- control warp: state record, fold of the previous launch's partials, the A.3 step -----------------
        if (!PERSIST && lane < (int)(sizeof(DenseState) / 8))
            reinterpret_cast<double*>(&sDS)[lane] = __ldcg(reinterpret_cast<const double*>(prev) + lane);
        if (lane < kDenseCols - kDenseLeaders)  // (the fold reads columns in chunks of 32: zero the padding columns)
            for (int t = 0; t < NT; ++t) sAccL[t][kDenseLeaders + lane] = 0.0;
        __syncwarp();
        named_sync(1, kDenseThreads);  // the search warps have left their 20 x 17 partial sums in sRed
        bool done = sDS.st.done != 0;
        if (!done && j > 0) {
            double ssum = 0.0;
            if (lane < NT) {
#pragma unroll
                for (int k = 0; k < kParts; ++k) ssum += sRed[k][lane];
            }
            __shared__ double sSum[kTerms];
            if (lane < NT) sSum[lane] = ssum;
            __syncwarp();
            B2_DTS(2);
            if (lane == 0) done = icp_step_local(sSum, ns, prm, sDS.st, results, ndone, blockIdx.x == 0);
        }
        if (lane == 0) {
            sDS.seq = j + 1;
            if (!PERSIST && blockIdx.x == 0) {
                states[wr] = sDS;
                *seqw = j + 1;
                if (loop_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, sDS.st.done ? 0u : 1u);
            }
        }
        __syncwarp();
        B2_DTS(3);
        named_sync(2, kDenseThreads);  // T_j (or `done`) is in sDS
        if (sDS.st.done) return;
    } else {
    // ---- search warps ----------------------------------------------------------------------------------------------
    // (a) this thread's share of the previous launch's partials: part = warp, term = lane
        double s = 0.0;
        if (PERSIST) {
            if (j > 0 && lane < NT) {
                constexpr int kMaxPer = (kCtasPerSm * kSMs + kParts - 1) / kParts;
                const uint4* pp = reinterpret_cast<const uint4*>(partial) + (size_t)rd * kCtasPerSm * kSMs * NT + lane;
                const int part = threadIdx.x >> 5;
                const unsigned tag = epoch_tag + (unsigned)j;  // written at the end of trip j - 1
                double v[kMaxPer];
                unsigned pending = 0;
#pragma unroll
                for (int k = 0; k < kMaxPer; ++k) {
                    v[k] = 0.0;
                    if (part + k * kParts < n_cta) pending |= 1u << k;
                }
                // (every warp polls the entries it needs: a cheaper hint phase in front — one entry per row — and a
                // back-off between the rounds were both measured slower: what counts is the number of round trips
                // between the last partial landing and the fold, not the 6 MB per round the polling moves)
                for (unsigned spins = 0; pending; ++spins) {
#pragma unroll
                    for (int k = 0; k < kMaxPer; ++k) {
                        if (!(pending >> k & 1u)) continue;
                        const uint4 e = ld_tagged(pp + (size_t)(part + k * kParts) * NT);
                        if (e.y == tag && e.w == tag) {
                            v[k] = __longlong_as_double((long long)(((unsigned long long)e.z << 32) | e.x));
                            pending &= ~(1u << k);
                        }
                    }
                    if (spins > (1u << 20)) {  // (seconds: some CTA of the grid is not resident — never with one CTA per SM)
                        atomicCAS(status, 0, B2_ERR_STATE);
                        break;
                    }
                }
#pragma unroll
                for (int k = 0; k < kMaxPer; ++k) s += v[k];
            }
        } else if (j > 0 && lane < NT) {
            constexpr int kMaxPer = (kCtasPerSm * kSMs + kParts - 1) / kParts;
            const double* pp = partial + (size_t)rd * kCtasPerSm * kSMs * kTerms + lane;
            const int part = threadIdx.x >> 5;
            double v[kMaxPer];
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) {
                const int b = part + k * kParts;
                v[k] = b < n_cta ? __ldcg(pp + (size_t)b * kTerms) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) s += v[k];
        }
        // (b) meanwhile: the previous transform for the speculative fetch — ONE request per warp (the record is a
        // single line that 148 x 19 warps read at the same moment), rows handed out by shuffles
        // (PERSIST: the state record of the previous trip is in shared memory; the control warp rewrites it only
        // after this warp has arrived at barrier 1, and a finished registration never gets here)
        const double tv = lane < 12 ? (PERSIST ? sDS.st.T[lane] : __ldg(prev->st.T + lane)) : 0.0;
        const int prev_done = PERSIST ? 0 : __ldg(&prev->st.done);
        double Trow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) Trow[i] = __shfl_sync(kFull, tv, (4 * g + i) & 15);
        if (lane < NT) sRed[threadIdx.x >> 5][lane] = s;
        __threadfence_block();
        named_arrive(1, kDenseThreads);
        B2_DTS(6);
        B2_DTS_MAX(14);

        // (c) speculative fetch AND float32 screening for the first round's query with T_{j-1}: what crosses the
        // barrier is the lane's two nearest cells + a bound for the others (SpecKept), not the nine cells
        SpecKept kp;
        const int warp_q0 = q_first - grp;
        const bool valid0 = q_first < ns && has_query;
        double lo = 1.0, hi = 0.0;  // bounds of the predicted cell along this lane's axis (empty: nothing predicted)
        double Pa_old = 0.0;        // this lane's axis of T_{j-1} p
        double pdx = (double)p_first.x, pdy = (double)p_first.y, pdz = (double)p_first.z;  // (converted in the shadow)
        bool spec_live;
        {
            PlaneCells pc;
            int cs = 0;
            if (g < 3) {
                dense_axis(Trow, p_first, m_org, m_cell, Pa_old, cs);
                lo = dadd(dadd(m_org, dmul((double)cs, m_cell)), 1e-9);
                hi = dsub(dadd(dadd(m_org, dmul((double)cs, m_cell)), m_cell), 1e-9);
            }
            const int cx = __shfl_sync(kFull, cs, lead), cy = __shfl_sync(kFull, cs, lead + 1), cz = __shfl_sync(kFull, cs, lead + 2);
            const int lim = (1 << 20) - 10;
            const bool in_range = cx > -lim && cx < lim && cy > -lim && cy < lim && cz > -lim && cz < lim;
            spec_live = valid0 && in_range && !prev_done;
            dense_load_plane(a, cx, cy, cz, g, spec_live, pc);
            B2_DTS(7);
            const float paf = (float)Pa_old;
            const float Pxf = __shfl_sync(kFull, paf, lead), Pyf = __shfl_sync(kFull, paf, lead + 1), Pzf = __shfl_sync(kFull, paf, lead + 2);
            // (consumes the loads: the wait sits here, in the shadow of the solve, not behind the barrier)
            dense_screen_top2(a, pc, Pxf, Pyf, Pzf, m_cell, kp);
        }
        B2_DTS_MAX(12);
        named_sync(2, kDenseThreads);  // T_j is in sDS
        B2_DTS(8);
        if (sDS.st.done) return;
        const double* sT = sDS.st.T;
        const double* sPivot = sDS.st.pivot;
        const float r_up = (float)prm.max_corr * 1.000001f;
        double* colp = &sAccL[0][0] + col;
        const bool warp_has_queries = (int)(threadIdx.x >> 5) * kDenseQueriesPerWarp < share;
        for (int q0 = warp_has_queries ? warp_q0 : ns; q0 < ns; q0 += groups_total) {  // (warp-uniform trip count)
            const int q = q0 + grp;
            const bool valid = q < ns && has_query;
            const bool first = q0 == warp_q0;
            if (!first && valid) {
                const float4 pq = __ldg(src + q);
                pdx = (double)pq.x; pdy = (double)pq.y; pdz = (double)pq.z;
            }
            double Pa = 0.0;
            if (g < 3) {
                const double* row = sT + 4 * g;
                Pa = dadd(dadd(dadd(dmul(row[0], pdx), dmul(row[1], pdy)), dmul(row[2], pdz)), row[3]);
            }
            const double Px = __shfl_sync(kFull, Pa, lead), Py = __shfl_sync(kFull, Pa, lead + 1), Pz = __shfl_sync(kFull, Pa, lead + 2);
            double best = INFINITY;
            int best_c = 0x7fffffff;
            unsigned best_pos = 0;
            float bx = 0.f, by = 0.f, bz = 0.f;
            unsigned redo = kFull;  // lanes whose group has to take the full path (fetch around the new cell)
            if (first) {
                // ---- does the speculative screening stand under T_j?  The query moved by delta = |T_j p - T_{j-1} p|,
                // so a cell at float32-screened distance D~ is now at least LB(D~) = D~ (1 - 4e-6) - 2 e_abs - delta
                // away.  With rho = the exact distance of the nearest KEPT cell (capped at the gate r), the kept cells
                // decide the search when (i) LB(third) > rho: no cell that was dropped can win, tie or match, and
                // (ii) the ball of radius rho around the query lies inside the fetched 3 x 3 x 3 box (always true while
                // the query stays in its cell: cell edge >= r).  The second-nearest is evaluated exactly only if
                // LB(second) <= rho.  Anything else takes the full path.
                const float fd = (float)dsub(Pa, Pa_old);
                const float ddx = __shfl_sync(kFull, fd, lead), ddy = __shfl_sync(kFull, fd, lead + 1), ddz = __shfl_sync(kFull, fd, lead + 2);
                const float delta = (fabsf(ddx) + fabsf(ddy) + fabsf(ddz)) * 1.00001f + 1e-12f;  // (>= the Euclidean norm)
                const int c1 = kp.c12 & 0xff, c2 = kp.c12 >> 8;
                if (c1 != 0xff) {
                    best = dist2_strict(Px, Py, Pz, (double)kp.x1, (double)kp.y1, (double)kp.z1);
                    best_c = c1 + g;
                    best_pos = kp.pos1;
                    bx = kp.x1; by = kp.y1; bz = kp.z1;
                }
                const double gm = fmin(fmin(__shfl_sync(kFull, best, lead), __shfl_sync(kFull, best, lead + 1)), __shfl_sync(kFull, best, lead + 2));
                const float rho = fminf(sqrt_approx((float)gm) * 1.000001f + 1e-30f, r_up);
                bool ok = kp.s3 - delta > rho;
                if (g < 3 && !(Pa > lo && Pa < hi)) {  // left the cell along this axis: distance to the face of the fetched box
                    const double m = fmin(dsub(Pa, dsub(lo, m_cell)), dsub(dadd(hi, m_cell), Pa));
                    ok = ok && ((float)m * 0.999999f - 4.0f * kp.e_abs > rho);
                }
                redo = __ballot_sync(kFull, valid && g < 3 && (!spec_live || !ok));
                // a lane that cannot hold the minimum is no contender (the tail then takes its one-ballot path)
                if (best > gm) best = INFINITY;
                if (c2 != 0xff && kp.s2 - delta <= rho) {  // the second-nearest may be nearer (or tie) under T_j
                    const double m2 = dist2_strict(Px, Py, Pz, (double)kp.x2, (double)kp.y2, (double)kp.z2);
                    if (m2 < best || (m2 == best && c2 + g < best_c)) {
                        best = m2; best_c = c2 + g; best_pos = kp.pos2;
                        bx = kp.x2; by = kp.y2; bz = kp.z2;
                    }
                }
            }
            B2_DTS(9);
            if (redo) {  // (warp-uniform) the full path for the groups that need it
                const bool gredo = ((redo >> lead) & 7u) != 0 && g < 3;
                int ca = 0;
                if (g < 3) ca = voxel_coord(Pa, m_org, m_cell);
                const int cx = __shfl_sync(kFull, ca, lead), cy = __shfl_sync(kFull, ca, lead + 1), cz = __shfl_sync(kFull, ca, lead + 2);
                const int lim = (1 << 20) - 10;
                const bool in_range = cx > -lim && cx < lim && cy > -lim && cy < lim && cz > -lim && cz < lim;
#ifdef B2_TAIL_TIMING
                if (g == 0 && valid && gredo && first && j >= 0 && j < 64) atomicAdd(&g_dense_mis[j], 1u);
#endif
                PlaneCells pc;
                dense_load_plane(a, cx, cy, cz, g, gredo && valid && in_range, pc);
                double sb;
                int sc;
                unsigned sp;
                float sx, sy, sz;
                dense_lane_best(m_cell, Px, Py, Pz, pc, g, lead, sb, sc, sp, sx, sy, sz);
                if (gredo) {
                    best = sb; best_c = sc; best_pos = sp;
                    bx = sx; by = sy; bz = sz;
                }
            }
            B2_DTS(10);
            B2_DTS_MAX(13);
            dense_group_tail(a, sPivot, Px, Py, Pz, best, best_c, best_pos, bx, by, bz, valid, q, g, lead, colp, first);
            B2_DTS(11);
        }
        // a warp without queries (or a source smaller than the grid) leaves zeros in its columns
        if (g < kDenseLanes && !(warp_has_queries && warp_q0 < ns)) {
            colp[(3 + g) * kDenseCols] = 0.0;
            colp[(6 + 3 * g) * kDenseCols] = 0.0; colp[(7 + 3 * g) * kDenseCols] = 0.0; colp[(8 + 3 * g) * kDenseCols] = 0.0;
            colp[g * kDenseCols] = 0.0;
            if (g < 2) colp[(15 + g) * kDenseCols] = 0.0;
        }
    }
    __syncthreads();  // (the search warps' columns are complete; the control warp has no part in the fold)
    B2_DTS(4);
    if (!PERSIST) {
        cta_fold<NT, kDenseCols>(&sAccL[0][0], partial + ((size_t)wr * kCtasPerSm * kSMs + blockIdx.x) * kTerms);
        B2_DTS(5);
        return;
    }
    {   // the same fold, the partial published as tagged entries (tag of trip j: read by everybody in trip j + 1)
        uint4* out = reinterpret_cast<uint4*>(partial) + ((size_t)wr * kCtasPerSm * kSMs + blockIdx.x) * NT;
        const int w = threadIdx.x >> 5;
        for (int t = w; t < NT; t += kDenseThreads / 32) {
            double acc = 0.0;
#pragma unroll
            for (int c0 = 0; c0 < kDenseCols; c0 += 32) acc += sAccL[t][c0 + lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
            if (lane == 0) st_tagged(out + t, acc, epoch_tag + (unsigned)(j + 1));
        }
    }
    B2_DTS(5);
    __syncthreads();  // (the columns are rewritten by the next trip)
    ++j;
    if (j >= step) return;
  }
}
