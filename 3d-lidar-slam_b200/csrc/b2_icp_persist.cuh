// b2_icp_persist.cuh — opt-in persistent single-pair driver (B2_PERSIST=1, eager calls only; DESIGN.md section 4,
// 'tried and rejected').  Included by b2_icp.cu inside namespace b2 { namespace {.
#pragma once
// ---- persistent single-pair driver ------------------------------------------------------------------
struct PersistSync {
    unsigned int ticket;  // cumulative CTA arrivals
    unsigned int epoch;   // steps published so far
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kPersistRounds = 2;  // query points a group keeps in registers across steps

template <int MODE, int KIND>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
icp_persist_kernel(const float4* __restrict__ src, const int* __restrict__ ns_dev, int ns_max,
                   const CellEntry* __restrict__ entries, unsigned mask, const float4* __restrict__ gpts,
                   const GridMeta* __restrict__ metas,
                   const float4* __restrict__ tgt_normals, IcpParams prm, IcpState* __restrict__ st,
                   double* __restrict__ partial, PersistSync* __restrict__ sync, IcpResultDev* __restrict__ results,
                   int* __restrict__ corr_out, int* __restrict__ ndone) {
    constexpr int NT = MODE == 0 ? 17 : 29;
    __shared__ IcpState sState;  // every CTA keeps a snapshot of the pair's state: the solver never re-reads it
    __shared__ GridMeta sMeta;
    __shared__ double sAccL[NT][kLeaders];
    __shared__ double sRed[kThreads / 32][kTerms];
    __shared__ double sSum[kTerms];
    __shared__ int sFlag;  // 1: this CTA was the last arriver of the step
    const double* sT = sState.T;
    const double* sPivot = sState.pivot;

    const int se = ns_dev ? min(*ns_dev, ns_max) : ns_max;
    if (threadIdx.x < sizeof(IcpState) / 8)
        reinterpret_cast<double*>(&sState)[threadIdx.x] = reinterpret_cast<const double*>(st)[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + sizeof(GridMeta) / 8)
        reinterpret_cast<double*>(&sMeta)[threadIdx.x - 32] = reinterpret_cast<const double*>(metas)[threadIdx.x - 32];

    SearchArgs a;
    a.entries = entries; a.gpts = gpts; a.tgt_normals = tgt_normals; a.mask = mask;
    a.tgt_base = 0;
    a.pair = 0u;
    a.r2 = dmul(prm.max_corr, prm.max_corr);
    a.corr_out = corr_out; a.d2_out = nullptr; a.search_only = 0;

    const int lane = threadIdx.x & 31;
    const int col = threadIdx.x / kGroup;
    const bool leader = (lane & (kGroup - 1)) == 0;
    const int groups_total = gridDim.x * kLeaders;
    const int warp_first = blockIdx.x * kLeaders + col - lane / kGroup;
    // the queries of a group do not change between steps: keep the first rounds in registers
    float4 pq[kPersistRounds];
#pragma unroll
    for (int r = 0; r < kPersistRounds; ++r) {
        const int q = warp_first + r * groups_total + lane / kGroup;
        pq[r] = (leader && q < se) ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const unsigned n_cta = gridDim.x;

    for (unsigned step = 0;; ++step) {
#ifdef B2_TAIL_TIMING
#define B2_PTS(k) do { if (step == 3 && threadIdx.x == 0 && blockIdx.x < 300) g_pts[blockIdx.x * 8 + (k)] = gtime(); } while (0)
#else
#define B2_PTS(k) do {} while (0)
#endif
        B2_PTS(0);
        for (int i = threadIdx.x; i < NT * kLeaders; i += kThreads) (&sAccL[0][0])[i] = 0.0;
        __syncthreads();  // sT / sPivot / sMeta of this step and the zeroed accumulators are visible
#pragma unroll
        for (int r = 0; r < kPersistRounds; ++r) {
            const int q0 = warp_first + r * groups_total;
            if (q0 < se) {
                const int q = q0 + lane / kGroup;
                search_round<MODE, KIND>(a, sT, sPivot, sMeta, pq[r], q < se, q, sAccL, col);
            }
        }
        for (int q0 = warp_first + kPersistRounds * groups_total; q0 < se; q0 += groups_total) {
            const int q = q0 + lane / kGroup;
            const bool valid = q < se;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (leader && valid) p = __ldg(src + q);
            search_round<MODE, KIND>(a, sT, sPivot, sMeta, p, valid, q, sAccL, col);
        }
        __syncthreads();
        B2_PTS(1);
        cta_fold<NT, kLeaders>(&sAccL[0][0], partial + (size_t)blockIdx.x * kTerms);
        __syncthreads();
        B2_PTS(2);
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned t = atomicAdd(&sync->ticket, 1u);
            sFlag = (t == (step + 1) * n_cta - 1);
        }
        __syncthreads();
        B2_PTS(3);
        if (sFlag) {  // last arriver of this step: the ticket ordered every CTA's partial before this point
            grid_fold<NT>(partial, (int)n_cta, sRed, sSum);
            B2_PTS(4);
            if (threadIdx.x == 0) {
                icp_step<MODE>(sSum, se, prm, sState, st, results, ndone);
                B2_PTS(5);
                __threadfence();
                st_release_u32(&sync->epoch, step + 1);
            }
        } else if (threadIdx.x == 0) {
            while (ld_acquire_u32(&sync->epoch) < step + 1) __nanosleep(20);
        }
        __syncthreads();
        B2_PTS(6);
        // pick up the published state (written by another CTA: read it from L2)
        if (threadIdx.x < sizeof(IcpState) / 8)
            reinterpret_cast<double*>(&sState)[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(st) + threadIdx.x);
        __syncthreads();
        if (sState.done) break;
    }
}
