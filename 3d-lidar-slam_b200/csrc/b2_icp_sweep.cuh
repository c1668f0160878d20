// b2_icp_sweep.cuh — the thread-per-query ICP step kernel for batches on pair / slab grids: pruned nearest-first
// stencil walk, float32 screening + exact fp64 contenders, neighbour-cell tasks pooled over the warp.
// Included by b2_icp.cu inside namespace b2 { namespace {.
#pragma once
// ---- pair / slab grids: one thread per query ------------------------------------------------------
// A pair or slab grid holds several points per cell (~6 at a 2 cm target under a 5 cm cell), so a query
// meets ~57 candidates under the full stencil and the walk — not memory latency — is the cost.  Here a
// thread owns a query and walks the stencil nearest-first with a bound: its own cell, then the 6 face,
// 12 edge and 8 corner neighbours, skipping every cell whose box (squared distance from the query to
// the box, shrunk by a safety margin) is farther than the best candidate so far or than the radius.
// After convergence ~3 of 27 cells and ~15 of 57 candidates survive.  Candidates are screened in
// float32; only the contenders — those within the float32 error bound of the best — are evaluated in
// fp64 in the oracle's strict order, so the result is the exact nearest neighbour of the full stencil,
// ties included.  Queries of a warp are consecutive in voxel-key order, i.e. spatial neighbours: their
// probes and candidate reads hit the same L1 lines.
// Walk order: 0 the query's cell, 1-6 faces, 7-18 edges, 19-26 corners, as a 6-bit offset code
// (dx+1) | (dy+1) << 2 | (dz+1) << 4.
__host__ __device__ constexpr int offset_code(int dx, int dy, int dz) {
    return (dx + 1) | ((dy + 1) << 2) | ((dz + 1) << 4);
}
__host__ __device__ constexpr int walk_code(int b) {
    if (b == 0) return offset_code(0, 0, 0);
    if (b < 7) {
        const int f = b - 1, s = (f & 1) ? 1 : -1;
        return (f >> 1) == 0 ? offset_code(s, 0, 0) : ((f >> 1) == 1 ? offset_code(0, s, 0) : offset_code(0, 0, s));
    }
    if (b < 19) {
        const int e = b - 7, s1 = (e & 2) ? 1 : -1, s2 = (e & 1) ? 1 : -1;
        return (e >> 2) == 0 ? offset_code(s1, s2, 0) : ((e >> 2) == 1 ? offset_code(s1, 0, s2) : offset_code(0, s1, s2));
    }
    const int c = b - 19;
    return offset_code((c & 4) ? 1 : -1, (c & 2) ? 1 : -1, (c & 1) ? 1 : -1);
}

#ifndef B2_SWEEP_T0
#define B2_SWEEP_T0 1024  // p2p: 64 registers per thread, 32 warps per SM (19 % faster than 640 threads at 96)
#define B2_SWEEP_T1 768   // p2l: 29 accumulator terms per thread limit the CTA through shared memory
#endif
template <int MODE>
struct SweepCfg {  // threads per sweep CTA (one CTA per SM): the register budget per thread is 64 K / threads
    static constexpr int kT = MODE == 0 ? B2_SWEEP_T0 : B2_SWEEP_T1;
    static constexpr int kNT = MODE == 0 ? 17 : 29;
    static constexpr size_t kSmem = sizeof(double) * kNT * kT;  // one accumulator column per thread
};

// L = lanes of a warp per query: 1 for batches (32 queries per warp round); 4 when one pair has to fill the
// machine (8 queries per round: the lanes of a group share the screening of the query's own cell, and the
// three that hold no query are extra hands for the pooled neighbour tasks).
template <int MODE, int KIND, int L>
__device__ __forceinline__ void search_one(const SearchArgs& a, const double* __restrict__ sT,
                                           const double* __restrict__ sPivot, const GridMeta& meta, float4 p, bool valid,
                                           int q, double* __restrict__ acc, unsigned long long* __restrict__ wbest,
                                           unsigned* __restrict__ wsecond) {
    // (warp-collective: all 32 lanes enter; `valid` marks the lanes that hold a query — with L > 1 only the
    // first lane of a group can)
    const int lane = threadIdx.x & 31;
    const D3 P = xform_strict(sT, (double)p.x, (double)p.y, (double)p.z);
    const double Px = P.x, Py = P.y, Pz = P.z;
    int cx = voxel_coord(Px, meta.origin[0], meta.cell);
    int cy = voxel_coord(Py, meta.origin[1], meta.cell);
    int cz = voxel_coord(Pz, meta.origin[2], meta.cell);
    if (KIND == 1 && meta.brick > 1) {
        cx = floor_div(cx, meta.brick);
        cy = floor_div(cy, meta.brick);
        cz = floor_div(cz, meta.brick);
    }
    bool searching = valid;
    if (KIND) {  // a query outside the 2^21-cell key space of a map cannot have a neighbour
        const int lim = (1 << 20) - 2;
        searching = searching && !(cx < -lim || cx > lim || cy < -lim || cy > lim || cz < -lim || cz > lim);
    }
    const double Ed = KIND == 1 ? meta.cell * (double)meta.brick : meta.cell;
    const float E = (float)Ed;
    const float ax = (float)(Px - (meta.origin[0] + (double)cx * Ed));  // offset inside the cell, in [0, E]
    const float ay = (float)(Py - (meta.origin[1] + (double)cy * Ed));
    const float az = (float)(Pz - (meta.origin[2] + (double)cz * Ed));
    const float Pxf = (float)Px, Pyf = (float)Py, Pzf = (float)Pz;
    const float slack = 1e-5f * E;  // >> the float32 rounding of ax and E - ax
    const float e_abs = (fmaxf(fmaxf(fabsf(Pxf), fabsf(Pyf)), fabsf(Pzf)) + 4.0f * E + 1e-3f) * 1.2e-7f;
    auto margin = [&](float d2f) -> float {  // bound of |d2_f32 - d2| for a candidate at float32 distance^2 d2f
        return 3.5f * e_abs * sqrtf(d2f) + 3.0f * e_abs * e_abs + 1e-6f * d2f;
    };
    // keys are additive in the cell offsets while no field overflows: always for map keys (range-checked
    // above), and for pair keys when the query's cell is not on the border of the 16-bit coordinate range
    const bool interior = KIND || (cx >= 1 && cx <= 65534 && cy >= 1 && cy <= 65534 && cz >= 1 && cz <= 65534);
    const unsigned long long kbase = KIND ? map_cell_key(cx - 1, cy - 1, cz - 1) : pair_cell_key(a.pair, cx - 1, cy - 1, cz - 1);
    auto key_delta = [](int code) -> unsigned long long {
        constexpr int sx = KIND ? 42 : 32, sy = KIND ? 21 : 16;
        return ((unsigned long long)(code & 3) << sx) + ((unsigned long long)((code >> 2) & 3) << sy) +
               (unsigned long long)(code >> 4);
    };
    auto key_of = [&](int code, bool& ok) -> unsigned long long {
        if (interior) return kbase + key_delta(code);
        const int nx = cx + (code & 3) - 1, ny = cy + ((code >> 2) & 3) - 1, nz = cz + (code >> 4) - 1;
        ok = nx >= 0 && nx < 65536 && ny >= 0 && ny < 65536 && nz >= 0 && nz < 65536;
        return pair_cell_key(a.pair, nx, ny, nz);
    };
    auto probe = [&](unsigned long long key, unsigned& off, int& cnt) -> bool {  // is the cell occupied?
        unsigned h = cell_hash(key) & a.mask;
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(a.entries + h));
        unsigned long long kk = ((unsigned long long)raw.y << 32) | raw.x;
        while (kk != key && kk != kEmptyKey) {  // collision chain
            h = (h + 1) & a.mask;
            raw = __ldg(reinterpret_cast<const uint4*>(a.entries + h));
            kk = ((unsigned long long)raw.y << 32) | raw.x;
        }
        off = raw.z;
        cnt = (int)raw.w;
        return kk == key;
    };

    float bf = INFINITY, sf = INFINITY;  // best / second-best float32 squared distance
    unsigned bpos = 0;
    float ub = (float)a.r2 * 1.00001f;  // nothing at or beyond the radius can be a correspondence
    // float32 screening of the points [off, off + cnt) for a query at (x, y, z): running best two
    auto screen_cell = [&](unsigned off, int cnt, float x, float y, float z, float& b1, float& b2, unsigned& bp,
                           int first, int stride) {
        const float4* cp = a.gpts + off;
        for (int j = first; j < cnt; j += stride) {
            const float4 t = __ldg(cp + j);
            const float fx = x - t.x, fy = y - t.y, fz = z - t.z;
            const float d = fx * fx + fy * fy + fz * fz;
            bp = d < b1 ? off + j : bp;
            b2 = fminf(b2, fmaxf(b1, d));  // second best of what has been seen
            b1 = fminf(b1, d);
        }
    };
    unsigned m = 0;  // this query's list: bit b = walk position b can still hold a closer point
    float xm = 0.f, xp = 0.f, ym = 0.f, yp = 0.f, zm = 0.f, zp = 0.f;
    // merge (d2, position) of one screened share into the slot of its query: an atomic min keeps the best,
    // whatever loses goes to the second best (non-negative floats order like their bit patterns)
    auto merge_into = [&](int slot, float l1, float l2, unsigned lp) {
        const unsigned long long mine = ((unsigned long long)__float_as_uint(l1) << 32) | lp;
        const unsigned long long old = atomicMin(wbest + slot, mine);
        const unsigned loser = mine < old ? (unsigned)(old >> 32) : __float_as_uint(l1);
        atomicMin(wsecond + slot, min(loser, __float_as_uint(l2)));
    };
    // the query's own cell first: it usually holds the neighbour and gives the bound for the rest
    if (L == 1) {
        if (searching) {
            bool ok = true;
            const unsigned long long key = key_of(offset_code(0, 0, 0), ok);
            unsigned off;
            int cnt;
            if (ok && probe(key, off, cnt)) screen_cell(off, cnt, Pxf, Pyf, Pzf, bf, sf, bpos, 0, 1);
        }
    } else {
        // the L lanes of the group screen the cell together: lane s takes points s, s + L, ...
        const int lead = lane & ~(L - 1);
        bool ok = searching;
        unsigned long long key = 0;
        if (searching) key = key_of(offset_code(0, 0, 0), ok);
        ok = __shfl_sync(kFull, ok, lead);
        key = __shfl_sync(kFull, key, lead);
        const float gx = __shfl_sync(kFull, Pxf, lead), gy = __shfl_sync(kFull, Pyf, lead), gz = __shfl_sync(kFull, Pzf, lead);
        wbest[lane] = 0x7f80000000000000ull;  // (+inf, position 0)
        wsecond[lane] = 0x7f800000u;
        __syncwarp();
        unsigned off;
        int cnt;
        if (ok && probe(key, off, cnt)) {
            float l1 = INFINITY, l2 = INFINITY;
            unsigned lp = 0;
            screen_cell(off, cnt, gx, gy, gz, l1, l2, lp, lane - lead, L);
            if (l1 < INFINITY) merge_into(lead, l1, l2, lp);
        }
        __syncwarp();
        if (searching) {
            const unsigned long long w = wbest[lane];
            bf = __uint_as_float((unsigned)(w >> 32));
            bpos = (unsigned)w;
            sf = __uint_as_float(wsecond[lane]);
        }
        __syncwarp();
    }
    if (searching) {
        if (bf < INFINITY) ub = fminf(ub, (bf + 2.0f * margin(bf)) * 1.00001f);  // true d2 <= bf + M
        // conservative (too small) squared distance from the query to the slab of cells at -1 / +1 per axis
        auto sq = [](float v) { return v * v * 0.9999f; };
        xm = sq(fmaxf(ax - slack, 0.f)); xp = sq(fmaxf(E - ax - slack, 0.f));
        ym = sq(fmaxf(ay - slack, 0.f)); yp = sq(fmaxf(E - ay - slack, 0.f));
        zm = sq(fmaxf(az - slack, 0.f)); zp = sq(fmaxf(E - az - slack, 0.f));
        // (a face neighbour qualifies when its slab does; an edge or corner neighbour only if all its faces do)
        const bool fxm = xm <= ub, fxp = xp <= ub, fym = ym <= ub, fyp = yp <= ub, fzm = zm <= ub, fzp = zp <= ub;
        m = (fxm ? 1u << 1 : 0u) | (fxp ? 1u << 2 : 0u) | (fym ? 1u << 3 : 0u) | (fyp ? 1u << 4 : 0u) |
            (fzm ? 1u << 5 : 0u) | (fzp ? 1u << 6 : 0u);
        if (((fxm || fxp) ? 1 : 0) + ((fym || fyp) ? 1 : 0) + ((fzm || fzp) ? 1 : 0) >= 2) {
#pragma unroll
            for (int b = 7; b < 27; ++b) {
                const int code = walk_code(b), dx = (code & 3) - 1, dy = ((code >> 2) & 3) - 1, dz = (code >> 4) - 1;
                const float lb = (dx < 0 ? xm : (dx > 0 ? xp : 0.f)) + (dy < 0 ? ym : (dy > 0 ? yp : 0.f)) +
                                 (dz < 0 ? zm : (dz > 0 ? zp : 0.f));
                if (lb <= ub) m |= 1u << b;
            }
        }
    }
    // ---- the neighbour cells, pooled over the warp ---------------------------------------------------
    // List lengths are very uneven (most queries need 1-3 neighbours, a query without a close point needs
    // all 26), and a warp that walks per lane waits for its longest list.  So the (query, cell) pairs of the
    // whole warp form one pool of tasks, 32 at a time: a lane takes a task, fetches the owner's key base and
    // coordinates by shuffle, probes the cell, screens its points and merges its best two into the owner's
    // slot in shared memory (an atomic min on (d2, position) keeps the best; whatever loses goes to the
    // second best).  The float32 arithmetic is the owner's, so the result does not depend on who computed it.
    const unsigned mp = interior ? m : 0u;  // (border cells of the 16-bit pair key space: walked by the owner below)
    const int c_own = __popc(mp);
    int pre = c_own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, pre, o);
        if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(kFull, pre, 31);
    pre -= c_own;  // exclusive prefix: first task of this lane
    wbest[lane] = ((unsigned long long)__float_as_uint(bf) << 32) | bpos;
    wsecond[lane] = __float_as_uint(sf);
    __syncwarp();
    for (int g0 = 0; g0 < total; g0 += 32) {  // warp-uniform trip count
        const int g = g0 + lane;
        int o = 0;  // owner: the last lane whose first task is <= g
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const int cand = o + step;
            const int pv = __shfl_sync(kFull, pre, cand & 31);
            if (cand < 32 && pv <= g) o = cand;
        }
        const int opre = __shfl_sync(kFull, pre, o);
        const unsigned om = __shfl_sync(kFull, mp, o);
        const unsigned long long okb = __shfl_sync(kFull, kbase, o);
        const float ox = __shfl_sync(kFull, Pxf, o), oy = __shfl_sync(kFull, Pyf, o), oz = __shfl_sync(kFull, Pzf, o);
        if (g < total) {
            const int b = __fns(om, 0, g - opre + 1);  // the (g - opre)-th cell of the owner's list
            unsigned off;
            int cnt;
            if (probe(okb + key_delta(a.walk[b]), off, cnt)) {
                float l1 = INFINITY, l2 = INFINITY;
                unsigned lp = 0;
                screen_cell(off, cnt, ox, oy, oz, l1, l2, lp, 0, 1);
                if (l1 < INFINITY) merge_into(o, l1, l2, lp);
            }
        }
    }
    __syncwarp();
    if (mp) {
        const unsigned long long w = wbest[lane];
        bf = __uint_as_float((unsigned)(w >> 32));
        bpos = (unsigned)w;
        sf = __uint_as_float(wsecond[lane]);
    }
    __syncwarp();  // the slots are reused by the warp's next round
    if (!interior && m) {  // rare: per-lane walk with range-checked keys
        unsigned mm = m;
        while (mm) {
            const int b = __ffs(mm) - 1;
            mm &= mm - 1;
            bool ok = true;
            const unsigned long long key = key_of(a.walk[b], ok);
            unsigned off;
            int cnt;
            if (ok && probe(key, off, cnt)) screen_cell(off, cnt, Pxf, Pyf, Pzf, bf, sf, bpos, 0, 1);
        }
    }
    if (!valid) return;
    // Exact phase.  Let M bound |d2_f32 - d2| for every candidate (rounding of P to float32 plus the float32
    // arithmetic): every candidate's d2_f32 >= d2* - M, where d2* is the true minimum, so a candidate with
    // d2_f32 > bf + 2M has d2 > d2* and cannot be the nearest neighbour.
    double best = INFINITY;
    int best_idx = 0x7fffffff;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (bf < INFINITY) {
        const float thr = (bf + 2.0f * margin(bf)) * 1.000002f + 1e-30f;
        if (sf > thr) {  // a single contender
            const float4 t = __ldg(a.gpts + bpos);
            best = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
            best_idx = __float_as_int(t.w);
            qx = t.x; qy = t.y; qz = t.z;
        } else {  // near-ties: exact (d2, idx) arg-min over the whole stencil
#pragma unroll 1
            for (int b = 0; b < 27; ++b) {
                bool ok = true;
                const unsigned long long key = key_of(a.walk[b], ok);
                if (!ok) continue;
                const CellEntry e = grid_lookup(a.entries, a.mask, key);
                if (e.key != key) continue;
                for (unsigned j = 0; j < e.cnt; ++j) {
                    const float4 t = __ldg(a.gpts + e.off + j);
                    const double d2 = dist2_strict(Px, Py, Pz, (double)t.x, (double)t.y, (double)t.z);
                    const int idx = __float_as_int(t.w);
                    if (d2 < best || (d2 == best && idx < best_idx)) {
                        best = d2;
                        best_idx = idx;
                        qx = t.x; qy = t.y; qz = t.z;
                    }
                }
            }
        }
    }
    const bool found = best < a.r2;
    if (a.corr_out) a.corr_out[q] = found ? best_idx : -1;
    if (a.d2_out) a.d2_out[q] = found ? best : 0.0;
    if (a.search_only || !found) return;
    const double Qx = qx, Qy = qy, Qz = qz;
    constexpr int S = SweepCfg<MODE>::kT;  // acc[t * S]: term t of this thread's column
    if (MODE == 0) {
        const double px = Px - sPivot[0], py = Py - sPivot[1], pz = Pz - sPivot[2];
        const double ux = Qx - sPivot[0], uy = Qy - sPivot[1], uz = Qz - sPivot[2];
        acc[0 * S] += px;       acc[1 * S] += py;       acc[2 * S] += pz;
        acc[3 * S] += ux;       acc[4 * S] += uy;       acc[5 * S] += uz;
        acc[6 * S] += ux * px;  acc[7 * S] += ux * py;  acc[8 * S] += ux * pz;
        acc[9 * S] += uy * px;  acc[10 * S] += uy * py; acc[11 * S] += uy * pz;
        acc[12 * S] += uz * px; acc[13 * S] += uz * py; acc[14 * S] += uz * pz;
        acc[15 * S] += best;
        acc[16 * S] += 1.0;
    } else {
        const float4 nq = __ldg(a.tgt_normals + a.tgt_base + best_idx);
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
            for (int j = i; j < 6; ++j) acc[(t++) * S] += J[i] * J[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[(21 + i) * S] += J[i] * r;
        acc[27 * S] += best;
        acc[28 * S] += 1.0;
    }
}

template <int MODE, int KIND, int L>  // KIND 0: pair grid, 1: resident map with slabs; L: lanes per query
__global__ void __launch_bounds__(SweepCfg<MODE>::kT, 1)
icp_sweep_kernel(const float4* __restrict__ src, const int* __restrict__ src_offs, const int* __restrict__ ns_dev,
                 int ns_max, const CellEntry* __restrict__ entries, unsigned mask, const float4* __restrict__ gpts,
                 const GridMeta* __restrict__ metas, const float4* __restrict__ tgt_normals,
                 const int* __restrict__ tgt_offs, IcpParams prm, IcpState* __restrict__ states,
                 double* __restrict__ partial, int* __restrict__ tickets, IcpResultDev* __restrict__ results,
                 int* __restrict__ corr_out, int* __restrict__ ndone, double* __restrict__ d2_out, int search_only) {
    constexpr int NT = SweepCfg<MODE>::kNT, kT = SweepCfg<MODE>::kT;
    const int pair = blockIdx.y;
    IcpState* st = states + pair;
    extern __shared__ double sAcc[];  // [NT][kT]: one accumulator column per thread
    __shared__ unsigned char sWalk[32];
    __shared__ unsigned long long sBest[SweepCfg<MODE>::kT];  // per warp: 32 slots of (best d2 bits, position)
    __shared__ unsigned sSecond[SweepCfg<MODE>::kT];          //           and of the second-best d2 bits
    __shared__ IcpState sState;
    __shared__ GridMeta sMeta;
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ int sLast;

    B2_TS_BLOCK(0);
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] =
            reinterpret_cast<const double*>(metas + (KIND ? 0 : pair))[threadIdx.x - 32];
#pragma unroll
    for (int t = 0; t < NT; ++t) sAcc[t * kT + threadIdx.x] = 0.0;
    if (threadIdx.x < 27) sWalk[threadIdx.x] = (unsigned char)walk_code(threadIdx.x);
    asm volatile("griddepcontrol.wait;" ::: "memory");
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
    if (sState.done) return;

    SearchArgs a;
    a.walk = sWalk;
    a.entries = entries; a.gpts = gpts; a.tgt_normals = tgt_normals; a.mask = mask;
    a.tgt_base = tgt_offs ? tgt_offs[pair] : 0;
    a.pair = (unsigned)pair;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = d2_out; a.search_only = search_only;

    // blocks of Q queries are dealt to the warps round-robin over (round, warp, CTA), so that a source smaller
    // than the grid still spreads over every SM
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (kT / 32);
    constexpr int Q = 32 / L;  // queries of a warp round
    for (int blk = (threadIdx.x >> 5) * gridDim.x + blockIdx.x; sb + blk * Q < se; blk += warps_total) {
        const int q = sb + blk * Q + lane / L;
        const bool valid = (lane % L) == 0 && q < se;
        const float4 p = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        search_one<MODE, KIND, L>(a, sState.T, sState.pivot, sMeta, p, valid, q, sAcc + threadIdx.x,
                                  sBest + (threadIdx.x & ~31), sSecond + (threadIdx.x & ~31));
    }
    if (search_only) return;

    __syncthreads();
    B2_TS_BLOCK(1);
    B2_TS(3);
    cta_fold<NT, kT>(sAcc, partial + ((size_t)pair * gridDim.x + blockIdx.x) * kTerms);
    __syncthreads();
    B2_TS(4);
    if (threadIdx.x == 0) {
        __threadfence();
        B2_TS(5);
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
    }
    B2_TS(2);
}
