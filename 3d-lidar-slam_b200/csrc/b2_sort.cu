// b2_sort.cu — stable LSD radix sort of (u64 key, u32 value) pairs.
//
// Used by voxel down-sampling (A.1), uniform-grid construction and map export.
// 8-bit digits; per pass: block histograms -> one-block exclusive scan ->
// stable scatter (warp match-any ranking).  The element count and the number
// of passes live in device memory (SortCtl), so a pipeline never has to read
// them back: passes >= ctl->npasses return immediately.  Pass p reads buffer
// (p&1 ? B : A) and writes the other one; on pass 0 a null value input means
// "value = element index".
//
// Grid sizing: at most 4 CTAs per SM (592) each owning a contiguous run of
// 2048-element chunks, so the histogram table stays small enough for a
// single-CTA scan while large inputs still fill the 148 SMs.
#include "b2_common.cuh"

namespace b2 {

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kChunk = kThreads * kItems;  // 2048
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBlocks = kSMs * 4;

__device__ __forceinline__ void block_range(int n, int& begin, int& end) {
    int nchunks = (n + kChunk - 1) / kChunk;
    int cpb = (nchunks + gridDim.x - 1) / gridDim.x;
    long long b = (long long)blockIdx.x * cpb * kChunk;
    long long e = b + (long long)cpb * kChunk;
    begin = (int)(b < n ? b : n);
    end = (int)(e < n ? e : n);
}

__global__ void __launch_bounds__(kThreads)
sort_hist_kernel(const unsigned long long* __restrict__ keysA, const unsigned long long* __restrict__ keysB,
                 const SortCtl* __restrict__ ctl, int pass, unsigned* __restrict__ hist) {
    pdl_enter();
    if (pass >= ctl->npasses) return;
    const unsigned long long* in = (pass & 1) ? keysB : keysA;
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    int begin, end;
    block_range(ctl->n, begin, end);
    const int shift = 8 * pass;
    for (int i = begin + threadIdx.x; i < end; i += kThreads) {
        unsigned d = (unsigned)(in[i] >> shift) & 255u;
        atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * gridDim.x + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan of `total` counters in place, one CTA of 1024 threads (`total` is a multiple of 256 and the
// buffer is 16-byte aligned).  The counters are walked in chunks of 8192: every thread loads two uint4 that are
// adjacent to its neighbours' (coalesced), scans its eight values in registers, and the CTA scans the
// per-thread sums; a running carry links the chunks.
__global__ void __launch_bounds__(1024)
sort_scan_kernel(unsigned* __restrict__ hist, int total, const SortCtl* __restrict__ ctl, int pass) {
    pdl_enter();
    if (pass >= ctl->npasses) return;
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    uint4* h4 = reinterpret_cast<uint4*>(hist);
    const int total4 = total >> 2;
    for (int base = 0; base < total4; base += 2048) {
        // thread t owns uint4 #(base + 2t) and #(base + 2t + 1): eight consecutive counters
        const int i0 = base + 2 * (int)threadIdx.x;
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (i0 < total4) a = h4[i0];
        if (i0 + 1 < total4) b = h4[i0 + 1];
        const unsigned s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
        unsigned incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned carry = carry_s;  // (written two barriers ago; rewritten only after the next one)
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0) {
            unsigned v = warp_sums[lane];
            unsigned inc2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += t;
            }
            warp_sums[lane] = inc2 - v;
            if (lane == 31) carry_s = carry + inc2;  // read by everybody after the next barrier
        }
        __syncthreads();
        unsigned run = carry + warp_sums[w] + incl - s;
        uint4 oa, ob;
        oa.x = run; run += a.x;
        oa.y = run; run += a.y;
        oa.z = run; run += a.z;
        oa.w = run; run += a.w;
        ob.x = run; run += b.x;
        ob.y = run; run += b.y;
        ob.z = run; run += b.z;
        ob.w = run;
        if (i0 < total4) h4[i0] = oa;
        if (i0 + 1 < total4) h4[i0 + 1] = ob;
        __syncthreads();  // warp_sums / carry_s are rewritten by the next chunk
    }
}

__global__ void __launch_bounds__(kThreads)
sort_scatter_kernel(unsigned long long* keysA, unsigned long long* keysB, unsigned* valsA, unsigned* valsB,
                    const SortCtl* __restrict__ ctl, int pass, const unsigned* __restrict__ hist,
                    int vals_implicit, int fused_scan) {
    pdl_enter();
    if (pass >= ctl->npasses) return;
    const bool odd = pass & 1;
    const unsigned long long* __restrict__ kin = odd ? keysB : keysA;
    const unsigned* __restrict__ vin = odd ? valsB : valsA;
    unsigned long long* __restrict__ kout = odd ? keysA : keysB;
    unsigned* __restrict__ vout = odd ? valsA : valsB;
    const bool iota = (pass == 0) && vals_implicit;

    __shared__ unsigned running[256];
    __shared__ unsigned digit_base[256];
    __shared__ unsigned wcount[kWarps][256];

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (!fused_scan) {
        running[threadIdx.x] = hist[threadIdx.x * gridDim.x + blockIdx.x];  // already scanned
    } else {
        // small grids: every CTA derives its own bases from the raw histogram table (256 x gridDim.x
        // counters) instead of paying a separate one-CTA scan launch per pass
        unsigned below = 0, total = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) {
            const unsigned v = hist[threadIdx.x * gridDim.x + b];
            total += v;
            if (b < blockIdx.x) below += v;
        }
        // exclusive scan of the 256 digit totals
        unsigned incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) digit_base[w] = incl;
        __syncthreads();
        unsigned woff = 0;
#pragma unroll
        for (int k = 0; k < kWarps; ++k)
            if (k < w) woff += digit_base[k];
        __syncthreads();
        running[threadIdx.x] = woff + incl - total + below;
    }
    int begin, end;
    const int n = ctl->n;
    block_range(n, begin, end);
    const int shift = 8 * pass;

    for (int base = begin; base < end; base += kChunk) {
#pragma unroll
        for (int k = 0; k < kWarps; ++k) wcount[k][threadIdx.x] = 0;
        __syncthreads();
        unsigned long long key[kItems];
        unsigned lrank[kItems];
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = base + w * (32 * kItems) + r * 32 + lane;
            const bool valid = i < end;
            key[r] = valid ? kin[i] : 0ull;
            const unsigned d = (unsigned)(key[r] >> shift) & 255u;
            const unsigned active = __ballot_sync(0xffffffffu, valid);
            lrank[r] = 0;
            if (valid) {
                const unsigned peers = __match_any_sync(active, d);
                const int leader = __ffs(peers) - 1;
                unsigned old = 0;
                if (lane == leader) {
                    old = wcount[w][d];
                    wcount[w][d] = old + __popc(peers);
                }
                old = __shfl_sync(peers, old, leader);
                lrank[r] = old + __popc(peers & lt_mask);
            }
            __syncwarp();
        }
        __syncthreads();
        {
            unsigned acc = 0;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) {
                unsigned t = wcount[k][threadIdx.x];
                wcount[k][threadIdx.x] = acc;
                acc += t;
            }
            unsigned run = running[threadIdx.x];
            digit_base[threadIdx.x] = run;
            running[threadIdx.x] = run + acc;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const int i = base + w * (32 * kItems) + r * 32 + lane;
            if (i < end) {
                const unsigned d = (unsigned)(key[r] >> shift) & 255u;
                const unsigned dst = digit_base[d] + wcount[w][d] + lrank[r];
                kout[dst] = key[r];
                vout[dst] = iota ? (unsigned)i : vin[i];
            }
        }
        __syncthreads();
    }
}

}  // namespace

int radix_sort_pairs(Context* c, unsigned long long* keysA, unsigned long long* keysB, unsigned* valsA,
                     unsigned* valsB, const SortCtl* ctl, int n_max, int max_passes) {
    if (n_max <= 0) return B2_OK;
    int nblocks = div_up(n_max, kChunk);
    if (nblocks > kMaxBlocks) nblocks = kMaxBlocks;
    B2_WS(c, hist, unsigned, WS_HIST, sizeof(unsigned) * 256 * (size_t)kMaxBlocks);
    const int fused_scan = nblocks <= 64;
    for (int p = 0; p < max_passes; ++p) {
        pdl_launch(sort_hist_kernel, dim3(nblocks), dim3(kThreads), c->stream, (const unsigned long long*)keysA,
                   (const unsigned long long*)keysB, ctl, p, hist);
        B2_LAUNCH_CHECK(c);
        if (!fused_scan) {
            pdl_launch(sort_scan_kernel, dim3(1), dim3(1024), c->stream, hist, 256 * nblocks, ctl, p);
            B2_LAUNCH_CHECK(c);
        }
        pdl_launch(sort_scatter_kernel, dim3(nblocks), dim3(kThreads), c->stream, keysA, keysB, valsA, valsB, ctl, p,
                   (const unsigned*)hist, 1, fused_scan);
        B2_LAUNCH_CHECK(c);
    }
    return B2_OK;
}

}  // namespace b2
