// b2_api.cu — the C ABI (include/b2slam.h) over the kernels of this directory.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <functional>

#include "b2_common.cuh"

namespace b2 {

int set_error(Context* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->last_error = buf;
    return code;
}

void* ws_get(Context* c, int slot, size_t bytes) {
    DevBuf& b = c->ws[slot];
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return b.p;
    if (c->capturing) {
        set_error(c, B2_ERR_CUDA, "workspace slot %d would have to grow during graph capture", slot);
        return nullptr;
    }
    c->ws_generation++;
    size_t want = bytes + bytes / 4;  // headroom so slowly growing inputs do not reallocate every call
    want = (want + 255) & ~(size_t)255;
    if (b.p) {
        cudaStreamSynchronize(c->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        set_error(c, B2_ERR_CUDA, "cudaMalloc(%zu bytes) for workspace slot %d failed: %s", want, slot,
                  cudaGetErrorString(e));
        b.p = nullptr;
        return nullptr;
    }
    b.cap = want;
    return b.p;
}

namespace {

constexpr int kBlock = 256;

__global__ void pack_xyz_kernel(const float* __restrict__ xyz, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
}

__global__ void unpack_xyz_kernel(const float4* __restrict__ in, int n, float* __restrict__ xyz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    xyz[3 * i] = p.x;
    xyz[3 * i + 1] = p.y;
    xyz[3 * i + 2] = p.z;
}

// generic_point_cloud_processor.py:93-108
__global__ void project_kernel(const float* __restrict__ rec, int n, double fx, double fy, double cx, double cy,
                               int numpy2, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float xp = rec[3 * i], yp = rec[3 * i + 1], z = rec[3 * i + 2];
    float x, y;
    if (numpy2) {
        x = __fdiv_rn(__fmul_rn(__fsub_rn(xp, (float)cx), z), (float)fx);
        y = __fdiv_rn(__fmul_rn(__fsub_rn(yp, (float)cy), z), (float)fy);
    } else {
        x = (float)ddiv(dmul(dsub((double)xp, cx), (double)z), fx);
        y = (float)ddiv(dmul(dsub((double)yp, cy), (double)z), fy);
    }
    out[i] = make_float4(x, y, z, 0.f);
}

// sensor_msgs_py.point_cloud2.read_points(msg, ("x","y","z"), skip_nans=True) (input_handler.py:65) for the
// three float32 fields of a PointCloud2 byte buffer: one record per thread, byte-wise field assembly when a
// field is not 4-byte aligned, optional byte swap, NaN flags for the order-preserving compaction that
// follows, and (optionally) the pixel -> metric projection of generic_point_cloud_processor.py:93-108 fused in.
template <bool ALIGNED>
__global__ void ingest_kernel(const unsigned char* __restrict__ data, int n, int point_step, int off_x, int off_y,
                              int off_z, int big_endian, int skip_nans, int project, double fx, double fy, double cx,
                              double cy, int numpy2, float4* __restrict__ out, unsigned char* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* rec = data + (size_t)i * point_step;
    auto field = [&](int off) -> float {
        unsigned v;
        if (ALIGNED) {
            v = *reinterpret_cast<const unsigned*>(rec + off);
        } else {
            v = (unsigned)rec[off] | ((unsigned)rec[off + 1] << 8) | ((unsigned)rec[off + 2] << 16) |
                ((unsigned)rec[off + 3] << 24);
        }
        if (big_endian) v = __byte_perm(v, 0, 0x0123);
        return __uint_as_float(v);
    };
    const float xp = field(off_x), yp = field(off_y), z = field(off_z);
    if (keep) keep[i] = (skip_nans && (isnan(xp) || isnan(yp) || isnan(z))) ? 0 : 1;
    float x = xp, y = yp;
    if (project) {
        if (numpy2) {
            x = __fdiv_rn(__fmul_rn(__fsub_rn(xp, (float)cx), z), (float)fx);
            y = __fdiv_rn(__fmul_rn(__fsub_rn(yp, (float)cy), z), (float)fy);
        } else {
            x = (float)ddiv(dmul(dsub((double)xp, cx), (double)z), fx);
            y = (float)ddiv(dmul(dsub((double)yp, cy), (double)z), fy);
        }
    }
    out[i] = make_float4(x, y, z, 0.f);
}

__global__ void transform_kernel(const float4* __restrict__ in, int n, const double* __restrict__ T,
                                 float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    const D3 q = xform_strict(T, (double)p.x, (double)p.y, (double)p.z);
    out[i] = make_float4((float)q.x, (float)q.y, (float)q.z, p.w);
}

// Commit of a scan step: the registered pose becomes the prior of the next scan — unless the step failed on
// the device (sticky status word: a sort beyond its pass budget, a full map), in which case neither the pose
// nor the map (the merge kernels test the same word) is touched and the host can retry the scan.
__global__ void copy_pose_kernel(const IcpResultDev* __restrict__ res, double* __restrict__ pose,
                                 const int* __restrict__ dev_status) {
    if (*dev_status != 0) return;
    if (threadIdx.x < 16) pose[threadIdx.x] = res->T[threadIdx.x];
}

int check_T(Context* c, const double* T, const char* what) {
    if (!T) return B2_OK;
    for (int i = 0; i < 16; ++i)
        if (!std::isfinite(T[i])) return set_error(c, B2_ERR_INVALID, "%s: transform has non-finite entries", what);
    if (T[12] != 0.0 || T[13] != 0.0 || T[14] != 0.0 || T[15] != 1.0)
        return set_error(c, B2_ERR_INVALID, "%s: last row of the transform must be (0,0,0,1)", what);
    return B2_OK;
}

// Upload 16 doubles to the small device scratch area (slot k of WS_MISC).
int upload_T(Context* c, const double* T, int k, double** dev) {
    B2_WS(c, misc, double, WS_MISC, sizeof(double) * 16 * 16);
    double* d = misc + 16 * k;
    B2_CUDA_OK(c, cudaMemcpyAsync(d, T, sizeof(double) * 16, cudaMemcpyHostToDevice, c->stream));
    *dev = d;
    return B2_OK;
}

// Read and clear the sticky device status word (synchronises).
int check_status(Context* c) {
    int* h = (int*)c->host_mailbox + 640;
    B2_CUDA_OK(c, cudaMemcpyAsync(h, c->dev_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    const int st = *h;
    if (st == 0) return B2_OK;
    B2_CUDA_OK(c, cudaMemsetAsync(c->dev_status, 0, sizeof(int), c->stream));
    switch (st) {
        case B2_ERR_KEY_RANGE:
            return set_error(c, st, "voxel / cell index out of range (int32 voxel index, packed key width or "
                                    "radix pass budget exceeded)");
        case B2_ERR_CAPACITY:
            return set_error(c, st, "resident map is full: raise capacity_cells in b2_map_init");
        default:
            return set_error(c, st, "device-side failure %d", st);
    }
}

IcpParams to_internal(const b2_icp_params* p) {
    IcpParams q;
    q.max_corr = p->max_correspondence_distance;
    q.max_iter = p->max_iteration;
    q.rel_fitness = p->relative_fitness;
    q.rel_rmse = p->relative_rmse;
    q.mode = p->mode;
    return q;
}

int check_params(Context* c, const b2_icp_params* p) {
    if (!p) return set_error(c, B2_ERR_INVALID, "icp params are NULL");
    if (!(p->max_correspondence_distance > 0.0)) return set_error(c, B2_ERR_INVALID, "max_correspondence_distance must be > 0");
    if (p->max_iteration < 0) return set_error(c, B2_ERR_INVALID, "max_iteration must be >= 0");
    if (p->mode != B2_ICP_POINT_TO_POINT && p->mode != B2_ICP_POINT_TO_PLANE)
        return set_error(c, B2_ERR_INVALID, "unknown estimation mode %d", p->mode);
    return B2_OK;
}

void fill_result(const IcpResultDev* r, b2_icp_result* out) {
    std::memcpy(out->transformation, r->T, sizeof(double) * 16);
    out->fitness = r->fitness;
    out->inlier_rmse = r->rmse;
    out->iterations = r->iterations;
    out->n_correspondences = r->n_corr;
}

struct HintGuard {  // widen the radix pass budget for the duration of a call
    Context* c;
    int saved;
    HintGuard(Context* ctx, int h) : c(ctx), saved(ctx->sort_passes_hint) {
        if (h > c->sort_passes_hint) c->sort_passes_hint = h;
    }
    ~HintGuard() { c->sort_passes_hint = saved; }
};

}  // namespace
}  // namespace b2

namespace b2 {
namespace {
std::atomic<int> g_live_handles[64];
std::atomic<unsigned long long> g_handles_epoch{0};
}  // namespace
void handle_born(int device) { if (device >= 0 && device < 64) g_live_handles[device]++; g_handles_epoch++; }
void handle_gone(int device) { if (device >= 0 && device < 64) g_live_handles[device]--; g_handles_epoch++; }
int live_handles(int device) { return device >= 0 && device < 64 ? g_live_handles[device].load() : 2; }
unsigned long long handles_epoch() { return g_handles_epoch.load(); }
}  // namespace b2

using namespace b2;

static_assert(sizeof(b2_icp_result) == sizeof(IcpResultDev), "result layouts must match");

#define CTX(h)                               \
    Context* c = reinterpret_cast<Context*>(h); \
    if (!c) return B2_ERR_INVALID;           \
    cudaSetDevice(c->device)

// A sort whose key range does not fit the current radix pass budget reports KEY_RANGE once;
// retry the whole call with the full 8-pass budget before giving up.
#define B2_RETRY_WIDE(c, body)                                          \
    do {                                                                \
        int _rc = (body);                                               \
        if (_rc == B2_ERR_KEY_RANGE && (c)->sort_passes_hint < 8) {     \
            (c)->sort_passes_hint = 8;                                  \
            _rc = (body);                                               \
        }                                                               \
        return _rc;                                                     \
    } while (0)

extern "C" {

const char* b2_version(void) {
#ifdef B2_EXPERIMENTS
    return "b2slam 0.2 (sm_100a) +experiments";
#else
    return "b2slam 0.2 (sm_100a)";
#endif
}

int b2_create(int device, b2_handle* out) {
    if (!out) return B2_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return B2_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return B2_ERR_CUDA;
    Context* c = new Context();
    c->scan_path = getenv("B2_STREAM_SORTED") ? B2_SCAN_PATH_SORTED : B2_SCAN_PATH_HASH;
    c->dense_driver = (getenv("B2_DENSE_PERSIST") && atoi(getenv("B2_DENSE_PERSIST")) == 0) ? B2_DENSE_DRIVER_CHAINED
                                                                                              : B2_DENSE_DRIVER_PERSISTENT;
    c->device = device;
    // the main stream (ICP chain, fusion) outranks the front stream (uploads, down-sample of the NEXT scan), whose
    // CTAs only fill what the ICP chain leaves free
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
        delete c;
        return B2_ERR_CUDA;
    }
    c->own_stream = true;
    if (cudaMalloc(&c->dev_status, sizeof(int)) != cudaSuccess ||
        cudaMemset(c->dev_status, 0, sizeof(int)) != cudaSuccess ||
        cudaHostAlloc(&c->host_mailbox, 4096, cudaHostAllocDefault) != cudaSuccess) {
        delete c;
        return B2_ERR_CUDA;
    }
    bool aux_ok = cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
                  cudaHostAlloc(&c->result_ring, sizeof(IcpResultDev) * kResultRing, cudaHostAllocDefault) == cudaSuccess &&
                  cudaHostAlloc((void**)&c->count_ring, sizeof(int) * kResultRing, cudaHostAllocDefault) == cudaSuccess;
    for (int i = 0; i < 4 && aux_ok; ++i)
        aux_ok = cudaEventCreateWithFlags(&c->ev_block[i], cudaEventDisableTiming) == cudaSuccess;
    aux_ok = aux_ok && cudaEventCreateWithFlags(&c->ev_dev_ready, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && aux_ok; ++i)
        aux_ok = cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
    c->ring_first = new unsigned char[kResultRing]();
    if (!aux_ok || icp_configure(c) != B2_OK) {
        delete[] c->ring_first;
        if (c->result_ring) cudaFreeHost(c->result_ring);
        if (c->count_ring) cudaFreeHost(c->count_ring);
        for (int i = 0; i < 4; ++i)
            if (c->ev_block[i]) cudaEventDestroy(c->ev_block[i]);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        cudaFree(c->dev_status);
        cudaFreeHost(c->host_mailbox);
        cudaStreamDestroy(c->stream);
        delete c;
        return B2_ERR_CUDA;
    }
    handle_born(device);
    *out = reinterpret_cast<b2_handle>(c);
    return B2_OK;
}

int b2_destroy(b2_handle h) {
    CTX(h);
    cudaStreamSynchronize(c->stream);
    handle_gone(c->device);
    for (int i = 0; i < 2; ++i) {
        if (c->front_graph[i].exec) cudaGraphExecDestroy(c->front_graph[i].exec);
        if (c->main_graph[i].exec) cudaGraphExecDestroy(c->main_graph[i].exec);
    }
    map_free(c);
    for (auto& b : c->ws)
        if (b.p) cudaFree(b.p);
    cudaFree(c->dev_status);
    cudaFreeHost(c->host_mailbox);
    cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c->ev_copied[i]);
        cudaEventDestroy(c->ev_consumed[i]);
    }
    cudaStreamDestroy(c->copy_stream);
    if (c->ev_dev_ready) cudaEventDestroy(c->ev_dev_ready);
    for (int i = 0; i < Context::kHostStages; ++i) {
        if (c->host_stage[i]) cudaFreeHost(c->host_stage[i]);
        if (c->ev_host_stage[i]) cudaEventDestroy(c->ev_host_stage[i]);
    }
    cudaFreeHost(c->result_ring);
    cudaFreeHost(c->count_ring);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(c->ev_block[i]);
    delete[] c->ring_first;
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return B2_OK;
}

const char* b2_last_error(b2_handle h) {
    Context* c = reinterpret_cast<Context*>(h);
    return c ? c->last_error.c_str() : "invalid handle";
}

int b2_set_stream(b2_handle h, void* cuda_stream) {
    CTX(h);
    cudaStreamSynchronize(c->stream);
    if (cuda_stream) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        c->stream = (cudaStream_t)cuda_stream;
        c->own_stream = false;
    } else if (!c->own_stream) {
        B2_CUDA_OK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return B2_OK;
}

int b2_synchronize(b2_handle h) {
    CTX(h);
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return check_status(c);
}

int b2_set_icp_driver(b2_handle h, int driver) {
    CTX(h);
#ifdef B2_EXPERIMENTS
    const int last = B2_ICP_DRIVER_SWEEP4;
#else
    const int last = B2_ICP_DRIVER_SWEEP;
#endif
    if (driver < B2_ICP_DRIVER_AUTO || driver > last)
        return set_error(c, B2_ERR_INVALID, "b2_set_icp_driver: unknown driver %d%s", driver,
                         driver == B2_ICP_DRIVER_SWEEP4 ? " (SWEEP4 is an experiment: build with B2_EXPERIMENTS=1)" : "");
    c->icp_driver = driver;
    return B2_OK;
}

int b2_set_scan_path(b2_handle h, int path) {
    CTX(h);
    if (path != B2_SCAN_PATH_HASH && path != B2_SCAN_PATH_SORTED)
        return set_error(c, B2_ERR_INVALID, "b2_set_scan_path: unknown path %d", path);
    if (c->scan_path != path) {
        c->scan_path = path;
        c->ws_generation++;  // the captured scan step holds the other path's kernels
    }
    return B2_OK;
}

int b2_set_dense_driver(b2_handle h, int driver) {
    CTX(h);
    if (driver != B2_DENSE_DRIVER_PERSISTENT && driver != B2_DENSE_DRIVER_CHAINED)
        return set_error(c, B2_ERR_INVALID, "b2_set_dense_driver: unknown driver %d", driver);
    if (c->dense_driver != driver) {
        c->dense_driver = driver;
        c->ws_generation++;  // the captured scan step holds the other driver's launches
    }
    return B2_OK;
}

unsigned long long b2_launch_count(b2_handle h) {
    Context* c = reinterpret_cast<Context*>(h);
    return c ? c->launches : 0;
}

int b2_pack_xyz(b2_handle h, const float* xyz, int n, int src_on_host, float* out_xyzw_dev) {
    CTX(h);
    if (n < 0 || (n > 0 && (!xyz || !out_xyzw_dev))) return set_error(c, B2_ERR_INVALID, "b2_pack_xyz: bad arguments");
    if (n == 0) return B2_OK;
    const float* src = xyz;
    if (src_on_host) {
        B2_WS(c, stage, float, WS_STAGE_XYZ, sizeof(float) * 3 * (size_t)n);
        B2_CUDA_OK(c, cudaMemcpyAsync(stage, xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        src = stage;
    }
    pack_xyz_kernel<<<div_up(n, kBlock), kBlock, 0, c->stream>>>(src, n, (float4*)out_xyzw_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

int b2_unpack_xyz(b2_handle h, const float* xyzw_dev, int n, float* out_xyz, int dst_on_host) {
    CTX(h);
    if (n < 0 || (n > 0 && (!xyzw_dev || !out_xyz))) return set_error(c, B2_ERR_INVALID, "b2_unpack_xyz: bad arguments");
    if (n == 0) return B2_OK;
    float* dst = out_xyz;
    if (dst_on_host) {
        B2_WS(c, stage, float, WS_STAGE_XYZ, sizeof(float) * 3 * (size_t)n);
        dst = stage;
    }
    unpack_xyz_kernel<<<div_up(n, kBlock), kBlock, 0, c->stream>>>((const float4*)xyzw_dev, n, dst);
    B2_LAUNCH_CHECK(c);
    if (dst_on_host) {
        B2_CUDA_OK(c, cudaMemcpyAsync(out_xyz, dst, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    }
    return B2_OK;
}

int b2_project_pixels(b2_handle h, const float* records_host, int n, double fx, double fy, double cx, double cy,
                      int numpy2_semantics, float* out_xyzw_dev) {
    CTX(h);
    if (n < 0 || (n > 0 && (!records_host || !out_xyzw_dev))) return set_error(c, B2_ERR_INVALID, "b2_project_pixels: bad arguments");
    if (n == 0) return B2_OK;
    B2_WS(c, stage, float, WS_STAGE_XYZ, sizeof(float) * 3 * (size_t)n);
    B2_CUDA_OK(c, cudaMemcpyAsync(stage, records_host, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    project_kernel<<<div_up(n, kBlock), kBlock, 0, c->stream>>>(stage, n, fx, fy, cx, cy, numpy2_semantics,
                                                                (float4*)out_xyzw_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

int b2_ingest_pointcloud2(b2_handle h, const void* data, int data_on_host, int n_points, int point_step, int off_x,
                          int off_y, int off_z, int is_bigendian, int skip_nans, const double* intrinsics,
                          int numpy2_semantics, float* out_xyzw_dev, int* n_out) {
    CTX(h);
    if (n_out) *n_out = 0;
    if (n_points < 0 || (n_points > 0 && (!data || !out_xyzw_dev)))
        return set_error(c, B2_ERR_INVALID, "b2_ingest_pointcloud2: bad arguments");
    const int offs[3] = {off_x, off_y, off_z};
    for (int k = 0; k < 3; ++k)
        if (offs[k] < 0 || point_step < 12 || offs[k] + 4 > point_step)
            return set_error(c, B2_ERR_INVALID, "b2_ingest_pointcloud2: field offset %d does not fit point_step %d",
                             offs[k], point_step);
    if (n_points == 0) return B2_OK;
    const size_t bytes = (size_t)n_points * (size_t)point_step;
    const unsigned char* raw = (const unsigned char*)data;
    if (data_on_host) {
        B2_WS(c, stage, unsigned char, WS_INGEST_RAW, bytes);
        B2_CUDA_OK(c, cudaMemcpyAsync(stage, data, bytes, cudaMemcpyHostToDevice, c->stream));
        raw = stage;
    }
    float4* dst = (float4*)out_xyzw_dev;
    unsigned char* keep = nullptr;
    if (skip_nans) {
        B2_WS(c, tmp, float4, WS_INGEST_TMP, sizeof(float4) * (size_t)n_points);
        B2_WS(c, kp, unsigned char, WS_INGEST_KEEP, (size_t)n_points);
        dst = tmp;
        keep = kp;
    }
    const double fx = intrinsics ? intrinsics[0] : 1.0, fy = intrinsics ? intrinsics[1] : 1.0;
    const double cx = intrinsics ? intrinsics[2] : 0.0, cy = intrinsics ? intrinsics[3] : 0.0;
    const bool aligned = point_step % 4 == 0 && off_x % 4 == 0 && off_y % 4 == 0 && off_z % 4 == 0 &&
                         ((uintptr_t)raw % 4 == 0);
    if (aligned)
        ingest_kernel<true><<<div_up(n_points, kBlock), kBlock, 0, c->stream>>>(
            raw, n_points, point_step, off_x, off_y, off_z, is_bigendian, skip_nans, intrinsics ? 1 : 0, fx, fy, cx, cy,
            numpy2_semantics, dst, keep);
    else
        ingest_kernel<false><<<div_up(n_points, kBlock), kBlock, 0, c->stream>>>(
            raw, n_points, point_step, off_x, off_y, off_z, is_bigendian, skip_nans, intrinsics ? 1 : 0, fx, fy, cx, cy,
            numpy2_semantics, dst, keep);
    B2_LAUNCH_CHECK(c);
    if (!skip_nans) {
        if (n_out) *n_out = n_points;
        return B2_OK;
    }
    B2_WS(c, cnt, int, WS_COUNTS, sizeof(int) * 4);
    B2_TRY(select_points(c, dst, n_points, keep, (float4*)out_xyzw_dev, cnt));
    int* hn = (int*)c->host_mailbox + 724;
    B2_CUDA_OK(c, cudaMemcpyAsync(hn, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    if (n_out) *n_out = *hn;
    return B2_OK;
}

static int voxel_downsample_impl(Context* c, const float* pts_dev, int n, double voxel, const double* origin,
                                 const double* T, float* out_pts_dev, int* out_keys_dev, int* n_out) {
    double* T_dev = nullptr;
    if (T) B2_TRY(upload_T(c, T, 0, &T_dev));
    DownsampleOut out{(float4*)out_pts_dev, out_keys_dev, nullptr, nullptr};
    VoxelSortOut vs;
    B2_TRY(voxel_downsample(c, (const float4*)pts_dev, n, nullptr, nullptr, 1, voxel, origin ? 1 : 0, origin, T_dev,
                            out, &vs));
    int* hn = (int*)c->host_mailbox + 700;
    B2_CUDA_OK(c, cudaMemcpyAsync(hn, &vs.ctl->nseg, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    if (n_out) *n_out = *hn;
    return B2_OK;
}

int b2_voxel_downsample(b2_handle h, const float* pts_dev, int n, double voxel, const double* origin,
                        const double* T, float* out_pts_dev, int* out_keys_dev, int* n_out) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !out_pts_dev))) return set_error(c, B2_ERR_INVALID, "b2_voxel_downsample: bad arguments");
    if (!(voxel > 0.0)) return set_error(c, B2_ERR_INVALID, "voxel size must be > 0 (got %g)", voxel);
    B2_TRY(check_T(c, T, "b2_voxel_downsample"));
    if (n == 0) {
        if (n_out) *n_out = 0;
        return B2_OK;
    }
    B2_RETRY_WIDE(c, voxel_downsample_impl(c, pts_dev, n, voxel, origin, T, out_pts_dev, out_keys_dev, n_out));
}

int b2_scan_downsample(b2_handle h, const float* pts_dev, int n, double voxel, float* out_pts_dev, int* n_out) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !out_pts_dev)) || !n_out) return set_error(c, B2_ERR_INVALID, "b2_scan_downsample: bad arguments");
    if (!(voxel > 0.0)) return set_error(c, B2_ERR_INVALID, "voxel size must be > 0 (got %g)", voxel);
    *n_out = 0;
    if (n == 0) return B2_OK;
    if (!stream_applicable(n))
        return set_error(c, B2_ERR_INVALID, "b2_scan_downsample: %d points exceed the single-scan limit (use b2_voxel_downsample)", n);
    // (the scan tables are shared with the down-samples of deferred scans still in flight on the front stream)
    B2_CUDA_OK(c, cudaStreamSynchronize(c->copy_stream));
    B2_TRY(stream_begin_pre(c, n, 0));
    B2_WS(c, cnt, int, WS_COUNTS, sizeof(int) * 4);
    B2_TRY(stream_downsample(c, (const float4*)pts_dev, n, nullptr, voxel, (float4*)out_pts_dev, cnt, 0));
    B2_TRY(stream_propagate_status(c, n, 0));
    int* hn = (int*)c->host_mailbox;
    B2_CUDA_OK(c, cudaMemcpyAsync(hn, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    *n_out = *hn;
    return B2_OK;
}

int b2_estimate_normals(b2_handle h, const float* pts_dev, int n, double radius, int max_nn, float* out_normals_dev) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !out_normals_dev))) return set_error(c, B2_ERR_INVALID, "b2_estimate_normals: bad arguments");
    if (n == 0) return B2_OK;
    auto body = [&]() -> int {
        B2_TRY(estimate_normals(c, (const float4*)pts_dev, n, nullptr, nullptr, 1, radius, max_nn, (float4*)out_normals_dev));
        return check_status(c);
    };
    B2_RETRY_WIDE(c, body());
}

int b2_nn_search(b2_handle h, const float* src_dev, int ns, const float* tgt_dev, int nt, const double* T,
                 double radius, int* out_idx_dev, double* out_d2_dev) {
    CTX(h);
    if (ns < 0 || nt < 0 || (ns > 0 && (!src_dev || !out_idx_dev)) || (nt > 0 && !tgt_dev))
        return set_error(c, B2_ERR_INVALID, "b2_nn_search: bad arguments");
    if (!(radius > 0.0)) return set_error(c, B2_ERR_INVALID, "radius must be > 0");
    B2_TRY(check_T(c, T, "b2_nn_search"));
    if (ns == 0) return B2_OK;
    if (nt == 0) {
        B2_CUDA_OK(c, cudaMemsetAsync(out_idx_dev, 0xFF, sizeof(int) * (size_t)ns, c->stream));
        if (out_d2_dev) B2_CUDA_OK(c, cudaMemsetAsync(out_d2_dev, 0, sizeof(double) * (size_t)ns, c->stream));
        return B2_OK;
    }
    auto body = [&]() -> int {
        double* T_dev = nullptr;
        if (T) B2_TRY(upload_T(c, T, 1, &T_dev));
        GridView g;
        B2_TRY(grid_build(c, (const float4*)tgt_dev, nt, nullptr, nullptr, 1, radius, WS_GRID_ENTRIES, &g));
        IcpParams prm{radius, 0, 0.0, 0.0, 0};
        B2_WS(c, res, IcpResultDev, WS_BATCH_D, sizeof(IcpResultDev));
        B2_TRY(icp_run(c, (const float4*)src_dev, ns, nullptr, nullptr, 1, g, nullptr, nullptr, T_dev, prm, res,
                       out_idx_dev, out_d2_dev, /*search_only=*/1, 0));
        return check_status(c);
    };
    B2_RETRY_WIDE(c, body());
}

int b2_icp(b2_handle h, const float* src_dev, int ns, const float* tgt_dev, int nt, const float* tgt_normals_dev,
           const double* init, const b2_icp_params* params, b2_icp_result* out, int* out_corr_dev) {
    CTX(h);
    B2_TRY(check_params(c, params));
    if (ns <= 0 || nt <= 0 || !src_dev || !tgt_dev || !out) return set_error(c, B2_ERR_INVALID, "b2_icp: empty cloud or NULL argument");
    if (params->mode == B2_ICP_POINT_TO_PLANE && !tgt_normals_dev)
        return set_error(c, B2_ERR_INVALID, "point-to-plane ICP requires target normals");
    B2_TRY(check_T(c, init, "b2_icp"));
    auto body = [&]() -> int {
        double* init_dev = nullptr;
        if (init) B2_TRY(upload_T(c, init, 2, &init_dev));
        GridView g;
        B2_TRY(grid_build(c, (const float4*)tgt_dev, nt, nullptr, nullptr, 1, params->max_correspondence_distance,
                          WS_GRID_ENTRIES, &g));
        B2_WS(c, res, IcpResultDev, WS_BATCH_D, sizeof(IcpResultDev));
        B2_TRY(icp_run(c, (const float4*)src_dev, ns, nullptr, nullptr, 1, g, (const float4*)tgt_normals_dev, nullptr,
                       init_dev, to_internal(params), res, out_corr_dev, nullptr, 0, /*check_every=*/16));
        IcpResultDev* hr = (IcpResultDev*)((char*)c->host_mailbox + 1024);
        B2_CUDA_OK(c, cudaMemcpyAsync(hr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
        B2_TRY(check_status(c));
        fill_result(hr, out);
        return B2_OK;
    };
    B2_RETRY_WIDE(c, body());
}

int b2_transform(b2_handle h, const float* pts_dev, int n, const double* T, float* out_pts_dev) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !out_pts_dev)) || !T) return set_error(c, B2_ERR_INVALID, "b2_transform: bad arguments");
    B2_TRY(check_T(c, T, "b2_transform"));
    if (n == 0) return B2_OK;
    double* T_dev = nullptr;
    B2_TRY(upload_T(c, T, 3, &T_dev));
    transform_kernel<<<div_up(n, kBlock), kBlock, 0, c->stream>>>((const float4*)pts_dev, n, T_dev, (float4*)out_pts_dev);
    B2_LAUNCH_CHECK(c);
    return B2_OK;
}

static int count_kept(Context* c, const unsigned char* keep, int n, int* n_kept) {
    if (!n_kept) return check_status(c);
    B2_WS(c, tmp, float4, WS_TMP_PTS2, sizeof(float4) * 4);
    B2_WS(c, cnt, int, WS_COUNTS, sizeof(int) * 4);
    // count through the compaction's own prefix sum (no output written: n_out only)
    B2_WS(c, scratch, float4, WS_TMP_PTS, sizeof(float4) * (size_t)n);
    (void)tmp;
    B2_TRY(select_points(c, scratch, n, keep, scratch, cnt));
    int* hn = (int*)c->host_mailbox + 720;
    B2_CUDA_OK(c, cudaMemcpyAsync(hn, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    *n_kept = *hn;
    return B2_OK;
}

int b2_radius_outlier_mask(b2_handle h, const float* pts_dev, int n, int nb_points, double radius,
                           unsigned char* keep_dev, int* n_kept) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !keep_dev))) return set_error(c, B2_ERR_INVALID, "b2_radius_outlier_mask: bad arguments");
    if (n == 0) { if (n_kept) *n_kept = 0; return B2_OK; }
    auto body = [&]() -> int {
        B2_TRY(radius_outlier_mask(c, (const float4*)pts_dev, n, nb_points, radius, keep_dev));
        return count_kept(c, keep_dev, n, n_kept);
    };
    B2_RETRY_WIDE(c, body());
}

int b2_statistical_outlier_mask(b2_handle h, const float* pts_dev, int n, int nb_neighbors, double std_ratio,
                                double cell, unsigned char* keep_dev, int* n_kept) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !keep_dev))) return set_error(c, B2_ERR_INVALID, "b2_statistical_outlier_mask: bad arguments");
    if (n == 0) { if (n_kept) *n_kept = 0; return B2_OK; }
    auto body = [&]() -> int {
        // one coarse voxelisation gives both the bounding box (how many grid levels the k-NN sweep
        // needs) and the density the finest cell is sized from: occupancy of 5 cm voxels -> points
        // per m^2 of a surface-like cloud -> 9 c^2 rho ~ 6 k points in a 27-cell stencil
        B2_WS(c, tmp, float4, WS_TMP_PTS, sizeof(float4) * (size_t)n);
        DownsampleOut dso{tmp, nullptr, nullptr, nullptr};
        VoxelSortOut vs;
        B2_TRY(voxel_downsample(c, (const float4*)pts_dev, n, nullptr, nullptr, 1, 0.05, 0, nullptr, nullptr, dso, &vs));
        int* hn = (int*)c->host_mailbox + 724;
        VoxelPrep* hp = (VoxelPrep*)((char*)c->host_mailbox + 3072);
        B2_CUDA_OK(c, cudaMemcpyAsync(hn, &vs.ctl->nseg, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA_OK(c, cudaMemcpyAsync(hp, vs.prep, sizeof(VoxelPrep), cudaMemcpyDeviceToHost, c->stream));
        B2_TRY(check_status(c));
        double extent = 0.0;
        for (int d = 0; d < 3; ++d) extent = std::fmax(extent, hp->maxb[d] - hp->minb[d]);
        extent = extent * 1.001 + 1e-6;
        double use_cell = cell;
        if (!(use_cell > 0.0)) {
            const double per_voxel = (double)n / (double)(*hn > 0 ? *hn : 1);
            const double rho = per_voxel / (0.05 * 0.05);
            use_cell = std::sqrt(6.0 * (double)nb_neighbors / (9.0 * rho));
            if (use_cell < 0.002) use_cell = 0.002;
            if (use_cell > 0.5) use_cell = 0.5;
        }
        B2_TRY(statistical_outlier_mask(c, (const float4*)pts_dev, n, nb_neighbors, std_ratio, use_cell, extent, keep_dev));
        return count_kept(c, keep_dev, n, n_kept);
    };
    B2_RETRY_WIDE(c, body());
}

int b2_select_points(b2_handle h, const float* pts_dev, int n, const unsigned char* keep_dev, float* out_pts_dev,
                     int* n_out) {
    CTX(h);
    if (n < 0 || (n > 0 && (!pts_dev || !keep_dev || !out_pts_dev))) return set_error(c, B2_ERR_INVALID, "b2_select_points: bad arguments");
    if (n == 0) { if (n_out) *n_out = 0; return B2_OK; }
    B2_WS(c, cnt, int, WS_COUNTS, sizeof(int) * 4);
    B2_TRY(select_points(c, (const float4*)pts_dev, n, keep_dev, (float4*)out_pts_dev, cnt));
    int* hn = (int*)c->host_mailbox + 720;
    B2_CUDA_OK(c, cudaMemcpyAsync(hn, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    if (n_out) *n_out = *hn;
    return B2_OK;
}

int b2_map_init(b2_handle h, double voxel, const double* origin, double max_corr, long long capacity_cells) {
    CTX(h);
    return map_init(c, voxel, origin, max_corr, capacity_cells);
}

int b2_map_reset(b2_handle h) {
    CTX(h);
    return map_reset(c);
}

int b2_map_fuse(b2_handle h, const float* pts_dev, int n, const double* T) {
    CTX(h);
    if (n < 0 || (n > 0 && !pts_dev)) return set_error(c, B2_ERR_INVALID, "b2_map_fuse: bad arguments");
    B2_TRY(check_T(c, T, "b2_map_fuse"));
    if (n == 0) return B2_OK;
    auto body = [&]() -> int {
        double* T_dev = nullptr;
        if (T) B2_TRY(upload_T(c, T, 4, &T_dev));
        B2_TRY(map_fuse(c, (const float4*)pts_dev, n, nullptr, T_dev));
        return check_status(c);
    };
    // a KEY_RANGE failure happens before the merge kernel touches the map, so the retry is safe
    B2_RETRY_WIDE(c, body());
}

int b2_map_size(b2_handle h, long long* n_voxels) {
    CTX(h);
    return map_counts(c, nullptr, n_voxels);
}

int b2_map_export(b2_handle h, float* out_pts_dev, int* out_keys_dev, long long max_out, long long* n_out) {
    CTX(h);
    if (!out_pts_dev || max_out < 0) return set_error(c, B2_ERR_INVALID, "b2_map_export: bad arguments");
    return map_export(c, (float4*)out_pts_dev, out_keys_dev, nullptr, max_out, n_out);
}

int b2_map_export_records(b2_handle h, int* out_keys_dev, double* out_sums_dev, long long max_out, long long* n_out) {
    CTX(h);
    if (!out_keys_dev || !out_sums_dev || max_out < 0)
        return set_error(c, B2_ERR_INVALID, "b2_map_export_records: bad arguments");
    return map_export(c, nullptr, out_keys_dev, out_sums_dev, max_out, n_out);
}

int b2_map_merge_records(b2_handle h, const int* keys_dev, const double* sums_dev, long long n) {
    CTX(h);
    if (n < 0 || n > 0x7fffffffLL || (n > 0 && (!keys_dev || !sums_dev)))
        return set_error(c, B2_ERR_INVALID, "b2_map_merge_records: bad arguments");
    if (n == 0) return B2_OK;
    B2_TRY(map_merge_records(c, keys_dev, sums_dev, (int)n));
    return check_status(c);
}

int b2_map_icp(b2_handle h, const float* src_dev, int ns, const double* init, const b2_icp_params* params,
               b2_icp_result* out) {
    CTX(h);
    B2_TRY(check_params(c, params));
    if (ns <= 0 || !src_dev || !out) return set_error(c, B2_ERR_INVALID, "b2_map_icp: empty cloud or NULL argument");
    if (params->mode != B2_ICP_POINT_TO_POINT) return set_error(c, B2_ERR_INVALID, "the resident map holds no normals: point-to-point only");
    B2_TRY(check_T(c, init, "b2_map_icp"));
    GridView g;
    B2_TRY(map_grid_view(c, &g));
    if (params->max_correspondence_distance > map_max_corr(c) * (1.0 + 1e-12))
        return set_error(c, B2_ERR_INVALID, "max_correspondence_distance %g exceeds the map's cell edge %g",
                         params->max_correspondence_distance, map_max_corr(c));
    double* init_dev = nullptr;
    if (init) B2_TRY(upload_T(c, init, 5, &init_dev));
    IcpResultDev* res = map_result_dev(c);
    B2_TRY(icp_run(c, (const float4*)src_dev, ns, nullptr, nullptr, 1, g, nullptr, nullptr, init_dev,
                   to_internal(params), res, nullptr, nullptr, 0, 16));
    IcpResultDev* hr = (IcpResultDev*)((char*)c->host_mailbox + 1024);
    B2_CUDA_OK(c, cudaMemcpyAsync(hr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    fill_result(hr, out);
    return B2_OK;
}

// matched centroid of every query from the brick-pool positions the dense step kernel reports
__global__ void gather_matches_kernel(const int* __restrict__ pos, int n, const float4* __restrict__ cells,
                                      float4* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int p = pos[q];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p >= 0) {
        v = __ldg(cells + p);
        v.w = 1.0f;
    }
    out[q] = v;
}

int b2_map_icp_corr(b2_handle h, const float* src_dev, int ns, const double* init, const b2_icp_params* params,
                    b2_icp_result* out, double* d2_out_dev, float* match_out_dev) {
    CTX(h);
    B2_TRY(check_params(c, params));
    if (ns <= 0 || !src_dev || !out || !d2_out_dev || !match_out_dev)
        return set_error(c, B2_ERR_INVALID, "b2_map_icp_corr: empty cloud or NULL argument");
    if (params->mode != B2_ICP_POINT_TO_POINT) return set_error(c, B2_ERR_INVALID, "the resident map holds no normals: point-to-point only");
    B2_TRY(check_T(c, init, "b2_map_icp_corr"));
    GridView g;
    B2_TRY(map_grid_view(c, &g));
    if (g.is_map != 2)
        return set_error(c, B2_ERR_INVALID, "b2_map_icp_corr: per-query matches are reported for the brick-major map (voxel edge >= gate)");
    if (params->max_correspondence_distance > map_max_corr(c) * (1.0 + 1e-12))
        return set_error(c, B2_ERR_INVALID, "max_correspondence_distance %g exceeds the map's cell edge %g",
                         params->max_correspondence_distance, map_max_corr(c));
    double* init_dev = nullptr;
    if (init) B2_TRY(upload_T(c, init, 5, &init_dev));
    B2_WS(c, corr, int, WS_CORR, sizeof(int) * (size_t)ns);
    IcpResultDev* res = map_result_dev(c);
    B2_TRY(icp_run(c, (const float4*)src_dev, ns, nullptr, nullptr, 1, g, nullptr, nullptr, init_dev,
                   to_internal(params), res, corr, d2_out_dev, 0, 0));
    gather_matches_kernel<<<div_up(ns, 256), 256, 0, c->stream>>>(corr, ns, (const float4*)g.pts, (float4*)match_out_dev);
    B2_LAUNCH_CHECK(c);
    IcpResultDev* hr = (IcpResultDev*)((char*)c->host_mailbox + 1024);
    B2_CUDA_OK(c, cudaMemcpyAsync(hr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
    B2_TRY(check_status(c));
    fill_result(hr, out);
    return B2_OK;
}

int b2_stencil_candidates(b2_handle h, const float* src_dev, int ns, const float* tgt_dev, int nt, const double* T,
                          double cell, long long* n_candidates) {
    CTX(h);
    if (ns <= 0 || nt <= 0 || !src_dev || !tgt_dev || !n_candidates || !(cell > 0.0))
        return set_error(c, B2_ERR_INVALID, "b2_stencil_candidates: bad arguments");
    B2_TRY(check_T(c, T, "b2_stencil_candidates"));
    auto body = [&]() -> int {
        double* T_dev = nullptr;
        if (T) B2_TRY(upload_T(c, T, 7, &T_dev));
        GridView g;
        B2_TRY(grid_build(c, (const float4*)tgt_dev, nt, nullptr, nullptr, 1, cell, WS_GRID_ENTRIES, &g));
        B2_WS(c, cnt, unsigned long long, WS_COUNTS, sizeof(unsigned long long) * 2);
        B2_CUDA_OK(c, cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 2, c->stream));
        B2_TRY(map_stencil_count(c, (const float4*)src_dev, ns, T_dev, g, cnt));
        unsigned long long* hc = (unsigned long long*)((char*)c->host_mailbox + 2048);
        B2_CUDA_OK(c, cudaMemcpyAsync(hc, cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        B2_TRY(check_status(c));
        *n_candidates = (long long)*hc;
        return B2_OK;
    };
    B2_RETRY_WIDE(c, body());
}

int b2_slam_set_pose(b2_handle h, const double* T) {
    CTX(h);
    if (!T) return set_error(c, B2_ERR_INVALID, "b2_slam_set_pose: NULL transform");
    B2_TRY(check_T(c, T, "b2_slam_set_pose"));
    double* pose = map_pose_dev(c);
    if (!pose) return set_error(c, B2_ERR_STATE, "map not initialised (call b2_map_init)");
    B2_CUDA_OK(c, cudaMemcpyAsync(pose, T, sizeof(double) * 16, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_slam_get_pose(b2_handle h, double* T) {
    CTX(h);
    double* pose = map_pose_dev(c);
    if (!pose || !T) return set_error(c, B2_ERR_STATE, "map not initialised or NULL output");
    B2_CUDA_OK(c, cudaMemcpyAsync(T, pose, sizeof(double) * 16, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return B2_OK;
}

// After a failed step the host's view of the map ("has anything been fused yet?") is re-read from the device.
static void slam_resync_after_failure(Context* c) {
    long long voxels = 0;
    if (map_counts(c, nullptr, &voxels) == B2_OK) map_set_voxels_bound(c, voxels);
}

// ---- the scan step -----------------------------------------------------------------------------------------------
// Two halves on two streams, double-buffered by scan parity `par`:
//   front (c->copy_stream): upload + pack of the scan into scan[par], its live point count into n_dev[par], and —
//          sort-free path, map not empty — its 2 cm down-sample into sdown[par] / cnt[par].  None of this depends on
//          the map or on the previous pose, so with a caller that runs ahead (out == NULL) the front half of scan k+1
//          runs while scan k is still being registered: its kernels are small (128-thread CTAs fit next to the
//          one 640-thread ICP CTA of an SM) and live in the shadow of the ICP chain.
//   main  (c->stream): ICP of sdown[par] against the resident map seeded with the pose kept on the device, fusion of
//          scan[par] with the registered pose, pose commit.
// ev_copied[par]: front -> main (the buffers of `par` are filled); ev_consumed[par]: main -> front (they may be
// reused, two scans later).  A failed down-sample leaves its reason in the parity's status word; the fusion of the
// same scan hands it to the sticky status word, so a failure of scan k+1 cannot be blamed on scan k.
struct StreamSwap {  // the internal launchers use c->stream: the front half runs them on the front stream
    Context* c;
    cudaStream_t saved;
    StreamSwap(Context* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};

struct ScanBuffers {
    float4* scan;   // [cap] the raw scan, packed
    int* n_dev;     // live point count
    float4* sdown;  // [cap] down-sampled source
    int* cnt;       // its size
};

static int scan_buffers(Context* c, int cap, int par, ScanBuffers* sb) {
    B2_WS(c, scan, float4, par ? WS_SCAN_PTS2 : WS_SCAN_PTS, sizeof(float4) * (size_t)cap);
    B2_WS(c, counts, int, WS_SCAN_COUNT, sizeof(int) * 16);
    B2_WS(c, sdown, float4, par ? WS_SRC_DOWN2 : WS_SRC_DOWN, sizeof(float4) * (size_t)cap);
    sb->scan = scan;
    sb->n_dev = counts + 4 * par;
    sb->cnt = counts + 8 + 4 * par;
    sb->sdown = sdown;
    return B2_OK;
}

static bool scan_fast(Context* c, int n) { return c->scan_path == B2_SCAN_PATH_HASH && stream_applicable(n); }

// front half after the scan sits in scan[par] (launches on c->stream = the front stream while it runs)
static int slam_front_body(Context* c, const ScanBuffers& sb, int n, double ds_voxel, int par) {
    B2_TRY(stream_begin_pre(c, n, par));
    B2_TRY(stream_begin_main(c, n, par));  // (the fusion's zero-fill, taken off the main stream's critical path)
    B2_TRY(stream_downsample(c, sb.scan, n, sb.n_dev, ds_voxel, sb.sdown, sb.cnt, par));
    return B2_OK;
}

// main half: ICP against the resident map (pose chained on the device) -> fuse -> pose commit
static int slam_main_body(Context* c, const ScanBuffers& sb, int n, double ds_voxel, const b2_icp_params* params,
                          bool map_empty, int par, bool front_downsampled) {
    // n is the capacity the kernels are launched for; the live point count sits in *n_dev
    double* pose = map_pose_dev(c);
    IcpResultDev* res = map_result_dev(c);
    const bool fast = scan_fast(c, n);
    if (fast && !front_downsampled) B2_TRY(stream_begin_main(c, n, par));
    const int pre_par = front_downsampled ? par : -1;
    if (!map_empty) {
        const float4* src = sb.scan;
        const int* ns_dev = sb.n_dev;
        if (ds_voxel > 0.0) {
            if (!front_downsampled) {
                DownsampleOut dso{sb.sdown, nullptr, nullptr, nullptr};
                VoxelSortOut vs;
                B2_TRY(voxel_downsample(c, sb.scan, n, sb.n_dev, nullptr, 1, ds_voxel, 0, nullptr, nullptr, dso, &vs));
                // keep the voxel count where later sorts cannot overwrite it
                B2_CUDA_OK(c, cudaMemcpyAsync(sb.cnt, &vs.ctl->nseg, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
            }
            src = sb.sdown;
            ns_dev = sb.cnt;
        }
        GridView g;
        B2_TRY(map_grid_view(c, &g));
        B2_TRY(icp_run(c, src, n, nullptr, ns_dev, 1, g, nullptr, nullptr, pose, to_internal(params), res, nullptr,
                       nullptr, 0, 0));
        // fuse with the registered pose straight from the result record; the pose prior is committed last
        bool committed = false;
        if (fast) B2_TRY(map_fuse_stream(c, sb.scan, n, sb.n_dev, res->T, par, pre_par, res, &committed));
        else B2_TRY(map_fuse(c, sb.scan, n, sb.n_dev, res->T));
        if (!committed) {
            copy_pose_kernel<<<1, 32, 0, c->stream>>>(res, pose, c->dev_status);
            B2_LAUNCH_CHECK(c);
        }
        return B2_OK;
    }
    if (fast) B2_TRY(map_fuse_stream(c, sb.scan, n, sb.n_dev, pose, par, -1, nullptr, nullptr));
    else B2_TRY(map_fuse(c, sb.scan, n, sb.n_dev, pose));
    return B2_OK;
}

// Each half is ~10-40 small kernels; replaying it as a CUDA graph removes the per-launch CPU cost and
// most of the inter-kernel gap.  A graph is keyed by everything baked into its nodes (sizes,
// parameters, workspace/map addresses, scan path); the first call with a new key runs eagerly (it sizes the
// workspaces), the second captures, later calls replay.  `body` launches on c->stream.
static int run_graphed(Context* c, ScanGraph& sg, int n, double ds_voxel, const void* scan,
                       const b2_icp_params* params, const std::function<int()>& body) {
    static const bool no_graph = getenv("B2_NO_GRAPH") != nullptr;
    const bool same = sg.n == n && sg.ds_voxel == ds_voxel && sg.scan == scan && sg.passes_hint == c->sort_passes_hint &&
                      std::memcmp(&sg.params, params, sizeof(*params)) == 0 && sg.generation == c->ws_generation &&
                      sg.handles_epoch == handles_epoch();
    if (no_graph) return body();
    if (same && sg.exec) {
        B2_CUDA_OK(c, cudaGraphLaunch(sg.exec, c->stream));
        c->launches += sg.kernel_nodes;
        return B2_OK;
    }
    if (!same) {  // new shape: eager run, remember the key
        if (sg.exec) cudaGraphExecDestroy(sg.exec);
        sg.exec = nullptr;
        B2_TRY(body());
        sg.n = n; sg.ds_voxel = ds_voxel; sg.scan = scan; sg.params = *params; sg.generation = c->ws_generation;
        sg.handles_epoch = handles_epoch();
        sg.passes_hint = c->sort_passes_hint;
        return B2_OK;
    }
    // second call with this key: capture
    const unsigned long long launches_before = c->launches;
    B2_CUDA_OK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    int rc = body();
    c->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != B2_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        sg.n = -1;  // do not try again with this key until it changes
        c->launches = launches_before;
        if (rc != B2_OK && rc != B2_ERR_CUDA) return rc;
        return body();
    }
    e = cudaGraphInstantiate(&sg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        sg.exec = nullptr;
        sg.n = -1;
        c->launches = launches_before;
        return body();
    }
    sg.kernel_nodes = (int)(c->launches - launches_before);
    c->launches = launches_before + sg.kernel_nodes;
    B2_CUDA_OK(c, cudaGraphLaunch(sg.exec, c->stream));
    return B2_OK;
}

// proj != null: xyz_host holds (x_px, y_px, depth) records and proj = {fx, fy, cx, cy, numpy2 != 0}: the front half
// back-projects them (generic_point_cloud_processor.py:93-108) while it packs.
static int slam_scan_impl(Context* c, const float* xyz_host, const float4* scan_dev, int n, double ds_voxel,
                          const b2_icp_params* params, b2_icp_result* out, bool map_empty, const double* proj) {
    // The scan lands in one of two fixed device buffers, kernels are launched for a capacity bucket and
    // read the live point count from device memory: the step's graphs are replayed for every scan size
    // of the bucket (real scans drop a varying number of invalid pixels).
    const int cap = (n + 8191) / 8192 * 8192;
    const int par = (int)(c->scan_seq & 1u);
    ScanBuffers sb;
    B2_TRY(scan_buffers(c, cap, par, &sb));
    const bool front_ds = scan_fast(c, cap) && !map_empty && ds_voxel > 0.0;
    b2_icp_params front_key{};  // (the front half does not depend on the ICP parameters)
    // ---- front half ------------------------------------------------------------------------------------------
    {
        // the buffers of this parity were last read by the main half of the scan before the previous one
        if (c->consumed_valid[par]) B2_CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[par], 0));
        static const bool no_overlap = getenv("B2_NO_OVERLAP") != nullptr;  // A/B switch: front half after the previous scan
        if (no_overlap && c->consumed_valid[par ^ 1])
            B2_CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[par ^ 1], 0));
        {   // the live count goes through a pinned ring slot of its own: the copy below is asynchronous and a caller
            // that passes out == NULL runs many scans ahead of the device
            const unsigned slot = c->scan_seq % kResultRing;
            constexpr unsigned kPerBlock = kResultRing / 4;
            if (slot % kPerBlock == 0) {
                const int blk = (int)(slot / kPerBlock);
                // entering block blk: its previous use (kResultRing scans ago) must have run.  Everything up to the
                // start of the block after it was covered by that block's event.
                const int guard = (blk + 1) & 3;
                if (c->block_valid[guard]) B2_CUDA_OK(c, cudaEventSynchronize(c->ev_block[guard]));
                B2_CUDA_OK(c, cudaEventRecord(c->ev_block[blk], c->copy_stream));
                c->block_valid[blk] = true;
            }
            c->scan_seq++;
            int* hn = c->count_ring + slot;
            *hn = n;
            B2_CUDA_OK(c, cudaMemcpyAsync(sb.n_dev, hn, sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
        }
        if (scan_dev) {
            // A device scan may have been produced by a library call on the main stream since the last scan step
            // (b2_pack_xyz, b2_project_pixels, b2_ingest_pointcloud2 ...): then the front stream has to wait for it.
            // A scan that was ready before the previous step was submitted needs no such wait and overlaps it.
            if (c->launches != c->launches_after_scan) {
                B2_CUDA_OK(c, cudaEventRecord(c->ev_dev_ready, c->stream));
                B2_CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_dev_ready, 0));
            }
            B2_CUDA_OK(c, cudaMemcpyAsync(sb.scan, scan_dev, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, c->copy_stream));
        } else {
            B2_WS(c, stage0, float, WS_STAGE_XYZ, sizeof(float) * 3 * (size_t)cap);
            B2_WS(c, stage1, float, WS_STAGE_XYZ2, sizeof(float) * 3 * (size_t)cap);
            float* stage = par ? stage1 : stage0;
            // A pageable source (the ordinary NumPy array a ROS callback holds) is first copied into a pinned ring
            // slot of the library: the device copy is then truly asynchronous, overlaps the previous scan's
            // registration, and the caller's array may be released as soon as the call returns.
            const float* h2d_src = xyz_host;
            cudaPointerAttributes pa;
            const bool pageable = cudaPointerGetAttributes(&pa, xyz_host) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered;
            cudaGetLastError();
            const size_t bytes = sizeof(float) * 3 * (size_t)n;
            if (pageable) {
                const int hs = (int)(c->host_stage_seq++ % Context::kHostStages);
                if (c->host_stage_bytes[hs] < bytes) {
                    if (c->host_stage[hs]) {
                        B2_CUDA_OK(c, cudaEventSynchronize(c->ev_host_stage[hs]));
                        cudaFreeHost(c->host_stage[hs]);
                        c->host_stage[hs] = nullptr;
                    }
                    const size_t want = sizeof(float) * 3 * (size_t)cap;
                    B2_CUDA_OK(c, cudaHostAlloc(&c->host_stage[hs], want, cudaHostAllocDefault));
                    c->host_stage_bytes[hs] = want;
                    if (!c->ev_host_stage[hs]) B2_CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_host_stage[hs], cudaEventDisableTiming));
                } else {
                    B2_CUDA_OK(c, cudaEventSynchronize(c->ev_host_stage[hs]));  // its previous upload has run
                }
                std::memcpy(c->host_stage[hs], xyz_host, bytes);
                h2d_src = (const float*)c->host_stage[hs];
                B2_CUDA_OK(c, cudaMemcpyAsync(stage, h2d_src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
                B2_CUDA_OK(c, cudaEventRecord(c->ev_host_stage[hs], c->copy_stream));
            } else {
                B2_CUDA_OK(c, cudaMemcpyAsync(stage, h2d_src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
            }
            if (proj)
                project_kernel<<<div_up(n, kBlock), kBlock, 0, c->copy_stream>>>(stage, n, proj[0], proj[1], proj[2],
                                                                               proj[3], proj[4] != 0.0, sb.scan);
            else
                pack_xyz_kernel<<<div_up(n, kBlock), kBlock, 0, c->copy_stream>>>(stage, n, sb.scan);
            B2_LAUNCH_CHECK(c);
        }
        if (front_ds) {
            StreamSwap swap(c, c->copy_stream);
            B2_TRY(run_graphed(c, c->front_graph[par], cap, ds_voxel, sb.scan, &front_key,
                               [&]() { return slam_front_body(c, sb, cap, ds_voxel, par); }));
        }
        B2_CUDA_OK(c, cudaEventRecord(c->ev_copied[par], c->copy_stream));
    }
    // ---- main half -------------------------------------------------------------------------------------------
    B2_CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->ev_copied[par], 0));
    if (map_empty) B2_TRY(slam_main_body(c, sb, cap, ds_voxel, params, true, par, false));
    else B2_TRY(run_graphed(c, c->main_graph[par], cap, ds_voxel, sb.scan, params,
                            [&]() { return slam_main_body(c, sb, cap, ds_voxel, params, false, par, front_ds); }));
    double* pose = map_pose_dev(c);
    IcpResultDev* res = map_result_dev(c);
    {   // every scan leaves its result in the pinned ring (asynchronous copy): b2_slam_results reads it later
        const int slot = (int)(c->scans_total % kResultRing);
        IcpResultDev* rr = (IcpResultDev*)c->result_ring + slot;
        c->ring_first[slot] = map_empty ? 1 : 0;
        if (map_empty) B2_CUDA_OK(c, cudaMemcpyAsync(rr->T, pose, sizeof(double) * 16, cudaMemcpyDeviceToHost, c->stream));
        else B2_CUDA_OK(c, cudaMemcpyAsync(rr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
        c->scans_total++;
    }
    B2_CUDA_OK(c, cudaEventRecord(c->ev_consumed[par], c->stream));
    c->consumed_valid[par] = true;
    c->launches_after_scan = c->launches;
    if (out) {
        IcpResultDev* hr = (IcpResultDev*)((char*)c->host_mailbox + 1024);
        if (map_empty) {
            B2_CUDA_OK(c, cudaMemcpyAsync(hr->T, pose, sizeof(double) * 16, cudaMemcpyDeviceToHost, c->stream));
            hr->fitness = 0.0; hr->rmse = 0.0; hr->iterations = 0; hr->n_corr = 0;
        } else {
            B2_CUDA_OK(c, cudaMemcpyAsync(hr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
        }
        int rc = check_status(c);
        if (rc == B2_ERR_KEY_RANGE && c->sort_passes_hint < 8 && !scan_fast(c, cap)) {
            // the scan's voxel extent needs more radix passes than the captured step launches: nothing was
            // committed (pose and map are gated on the status word), so run the step again with the full
            // budget; the graph is keyed on the budget and is re-captured for the following scans
            c->sort_passes_hint = 8;
            IcpResultDev* rr = (IcpResultDev*)c->result_ring + (int)((c->scans_total - 1) % kResultRing);
            if (map_empty) {
                map_set_voxels_bound(c, 0);
                B2_TRY(slam_main_body(c, sb, cap, ds_voxel, params, true, par, false));
                B2_CUDA_OK(c, cudaMemcpyAsync(hr->T, pose, sizeof(double) * 16, cudaMemcpyDeviceToHost, c->stream));
                B2_CUDA_OK(c, cudaMemcpyAsync(rr->T, pose, sizeof(double) * 16, cudaMemcpyDeviceToHost, c->stream));
            } else {
                B2_TRY(run_graphed(c, c->main_graph[par], cap, ds_voxel, sb.scan, params,
                                   [&]() { return slam_main_body(c, sb, cap, ds_voxel, params, false, par, false); }));
                B2_CUDA_OK(c, cudaMemcpyAsync(hr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
                B2_CUDA_OK(c, cudaMemcpyAsync(rr, res, sizeof(IcpResultDev), cudaMemcpyDeviceToHost, c->stream));
            }
            B2_CUDA_OK(c, cudaEventRecord(c->ev_consumed[par], c->stream));
            rc = check_status(c);
        }
        if (rc != B2_OK) {
            slam_resync_after_failure(c);
            return rc;
        }
        fill_result(hr, out);
    }
    return B2_OK;
}

static int slam_scan_entry(Context* c, const float* xyz_host, const float4* scan_dev, int n, double ds_voxel,
                           const b2_icp_params* params, b2_icp_result* out, const double* proj = nullptr) {
    B2_TRY(check_params(c, params));
    if (n <= 0 || (!xyz_host && !scan_dev)) return set_error(c, B2_ERR_INVALID, "b2_slam_scan: empty scan");
    if (params->mode != B2_ICP_POINT_TO_POINT) return set_error(c, B2_ERR_INVALID, "the resident map holds no normals: point-to-point only");
    GridView g;
    B2_TRY(map_grid_view(c, &g));
    if (params->max_correspondence_distance > map_max_corr(c) * (1.0 + 1e-12))
        return set_error(c, B2_ERR_INVALID, "max_correspondence_distance %g exceeds the map's cell edge %g",
                         params->max_correspondence_distance, map_max_corr(c));
    // "first scan becomes the map" (icp_processor.py:53-56): decided on the host from the fuse history
    const bool map_empty = map_is_empty_host(c);
    return slam_scan_impl(c, xyz_host, scan_dev, n, ds_voxel, params, out, map_empty, proj);
}

int b2_slam_scan(b2_handle h, const float* xyz_host, int n, double ds_voxel, const b2_icp_params* params,
                 b2_icp_result* out) {
    CTX(h);
    return slam_scan_entry(c, xyz_host, nullptr, n, ds_voxel, params, out);
}

int b2_slam_scan_records(b2_handle h, const float* records_host, int n, double fx, double fy, double cx, double cy,
                         int numpy2_semantics, double ds_voxel, const b2_icp_params* params, b2_icp_result* out) {
    CTX(h);
    if (!(fx != 0.0) || !(fy != 0.0) || !std::isfinite(fx) || !std::isfinite(fy) || !std::isfinite(cx) || !std::isfinite(cy))
        return set_error(c, B2_ERR_INVALID, "b2_slam_scan_records: intrinsics must be finite, fx and fy non-zero");
    const double proj[5] = {fx, fy, cx, cy, numpy2_semantics ? 1.0 : 0.0};
    return slam_scan_entry(c, records_host, nullptr, n, ds_voxel, params, out, proj);
}

int b2_slam_results(b2_handle h, int count, b2_icp_result* out) {
    CTX(h);
    if (count < 0 || (count > 0 && !out)) return set_error(c, B2_ERR_INVALID, "b2_slam_results: bad arguments");
    if ((unsigned long long)count > c->scans_total || count > kResultRing)
        return set_error(c, B2_ERR_INVALID, "b2_slam_results: %d results asked, %llu scans registered, ring of %d",
                         count, c->scans_total, kResultRing);
    {   // waits for the stream: every pending result has landed in the ring
        const int rc = check_status(c);
        if (rc != B2_OK) {
            // a deferred scan failed on the device: from that scan on nothing was committed (pose and map are
            // gated on the status word).  Widen the sort budget so that resubmitting the scans succeeds.
            if (rc == B2_ERR_KEY_RANGE) c->sort_passes_hint = 8;
            slam_resync_after_failure(c);
            c->last_error += "; scans submitted since the failing one were not applied, resubmit them";
            return rc;
        }
    }
    for (int i = 0; i < count; ++i) {
        const int slot = (int)((c->scans_total - (unsigned long long)count + i) % kResultRing);
        IcpResultDev r = ((const IcpResultDev*)c->result_ring)[slot];
        if (c->ring_first[slot]) { r.fitness = 0.0; r.rmse = 0.0; r.iterations = 0; r.n_corr = 0; }
        fill_result(&r, out + i);
    }
    return B2_OK;
}

int b2_slam_scan_dev(b2_handle h, const float* scan_xyzw_dev, int n, double ds_voxel, const b2_icp_params* params,
                     b2_icp_result* out) {
    CTX(h);
    return slam_scan_entry(c, nullptr, (const float4*)scan_xyzw_dev, n, ds_voxel, params, out);
}

int b2_map_stencil_candidates(b2_handle h, const float* src_dev, int ns, const double* T, long long* n_candidates) {
    CTX(h);
    if (ns <= 0 || !src_dev || !n_candidates) return set_error(c, B2_ERR_INVALID, "b2_map_stencil_candidates: bad arguments");
    B2_TRY(check_T(c, T, "b2_map_stencil_candidates"));
    GridView g;
    B2_TRY(map_grid_view(c, &g));
    double* T_dev = nullptr;
    if (T) B2_TRY(upload_T(c, T, 6, &T_dev));
    B2_WS(c, cnt, unsigned long long, WS_COUNTS, sizeof(unsigned long long) * 2);
    B2_CUDA_OK(c, cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 2, c->stream));
    B2_TRY(map_stencil_count(c, (const float4*)src_dev, ns, T_dev, g, cnt));
    unsigned long long* hc = (unsigned long long*)((char*)c->host_mailbox + 2048);
    B2_CUDA_OK(c, cudaMemcpyAsync(hc, cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    *n_candidates = (long long)*hc;
    return B2_OK;
}

static int icp_batch_impl(Context* c, const float* src_dev, const int* src_offsets_dev, int src_total,
                          const float* tgt_dev, const int* tgt_offsets_dev, int tgt_total, int n_pairs, double ds_voxel,
                          double normal_radius, int normal_max_nn, const double* init_dev, const b2_icp_params* params,
                          b2_icp_result* results_dev) {
    B2_TRY(check_params(c, params));
    if (n_pairs <= 0 || src_total <= 0 || tgt_total <= 0 || !src_dev || !tgt_dev || !src_offsets_dev ||
        !tgt_offsets_dev || !results_dev)
        return set_error(c, B2_ERR_INVALID, "b2_icp_batch: empty batch or NULL argument");
    if (n_pairs > 32768) return set_error(c, B2_ERR_INVALID, "b2_icp_batch: at most 32768 pairs per call (got %d)", n_pairs);
    const bool p2l = params->mode == B2_ICP_POINT_TO_PLANE;
    if (p2l && (!(normal_radius > 0.0) || normal_max_nn < 1))
        return set_error(c, B2_ERR_INVALID, "point-to-plane batch: normal_radius > 0 and normal_max_nn >= 1 required "
                                            "(use b2_icp_batch_p2l)");
    HintGuard guard(c, 8);
    const float4* src = (const float4*)src_dev;
    const float4* tgt = (const float4*)tgt_dev;
    const int* soffs = src_offsets_dev;
    const int* toffs = tgt_offsets_dev;
    if (ds_voxel > 0.0) {
        B2_WS(c, sd, float4, WS_BATCH_A, sizeof(float4) * (size_t)src_total);
        B2_WS(c, so, int, WS_OFFS_A, sizeof(int) * ((size_t)n_pairs + 1));
        DownsampleOut o1{sd, nullptr, nullptr, so};
        B2_TRY(voxel_downsample(c, src, src_total, nullptr, src_offsets_dev, n_pairs, ds_voxel, 0, nullptr, nullptr, o1, nullptr));
        B2_WS(c, td, float4, WS_BATCH_B, sizeof(float4) * (size_t)tgt_total);
        B2_WS(c, to, int, WS_OFFS_B, sizeof(int) * ((size_t)n_pairs + 1));
        DownsampleOut o2{td, nullptr, nullptr, to};
        B2_TRY(voxel_downsample(c, tgt, tgt_total, nullptr, tgt_offsets_dev, n_pairs, ds_voxel, 0, nullptr, nullptr, o2, nullptr));
        src = sd; tgt = td; soffs = so; toffs = to;
    }
    const float4* normals = nullptr;
    if (p2l) {  // icp_processor.py:94-99 on every (down-sampled) target of the batch
        B2_WS(c, nrm, float4, WS_BATCH_C, sizeof(float4) * (size_t)tgt_total);
        B2_TRY(estimate_normals(c, tgt, tgt_total, nullptr, toffs, n_pairs, normal_radius, normal_max_nn, nrm));
        normals = nrm;
    }
    GridView g;
    B2_TRY(grid_build(c, tgt, tgt_total, nullptr, toffs, n_pairs, params->max_correspondence_distance, WS_GRID_ENTRIES, &g));
    B2_TRY(icp_run(c, src, src_total, soffs, nullptr, n_pairs, g, normals, toffs, init_dev, to_internal(params),
                   (IcpResultDev*)results_dev, nullptr, nullptr, 0, 0));
    return B2_OK;
}

int b2_icp_batch(b2_handle h, const float* src_dev, const int* src_offsets_dev, int src_total, const float* tgt_dev,
                 const int* tgt_offsets_dev, int tgt_total, int n_pairs, double ds_voxel, const double* init_dev,
                 const b2_icp_params* params, b2_icp_result* results_dev) {
    CTX(h);
    if (params && params->mode != B2_ICP_POINT_TO_POINT)
        return set_error(c, B2_ERR_INVALID, "b2_icp_batch: point-to-point only (b2_icp_batch_p2l estimates the normals)");
    return icp_batch_impl(c, src_dev, src_offsets_dev, src_total, tgt_dev, tgt_offsets_dev, tgt_total, n_pairs, ds_voxel,
                          0.0, 0, init_dev, params, results_dev);
}

int b2_icp_batch_p2l(b2_handle h, const float* src_dev, const int* src_offsets_dev, int src_total, const float* tgt_dev,
                     const int* tgt_offsets_dev, int tgt_total, int n_pairs, double ds_voxel, double normal_radius,
                     int normal_max_nn, const double* init_dev, const b2_icp_params* params, b2_icp_result* results_dev) {
    CTX(h);
    if (params && params->mode != B2_ICP_POINT_TO_PLANE)
        return set_error(c, B2_ERR_INVALID, "b2_icp_batch_p2l: params->mode must be B2_ICP_POINT_TO_PLANE");
    return icp_batch_impl(c, src_dev, src_offsets_dev, src_total, tgt_dev, tgt_offsets_dev, tgt_total, n_pairs, ds_voxel,
                          normal_radius, normal_max_nn, init_dev, params, results_dev);
}

int b2_estimate_normals_batch(b2_handle h, const float* pts_dev, const int* offsets_dev, int total, int n_clouds,
                              double radius, int max_nn, float* out_normals_dev) {
    CTX(h);
    if (total < 0 || n_clouds <= 0 || n_clouds > 32768 || (total > 0 && (!pts_dev || !offsets_dev || !out_normals_dev)))
        return set_error(c, B2_ERR_INVALID, "b2_estimate_normals_batch: bad arguments");
    if (total == 0) return B2_OK;
    HintGuard guard(c, 8);
    auto body = [&]() -> int {
        B2_TRY(estimate_normals(c, (const float4*)pts_dev, total, nullptr, offsets_dev, n_clouds, radius, max_nn,
                                (float4*)out_normals_dev));
        return check_status(c);
    };
    B2_RETRY_WIDE(c, body());
}

}  // extern "C"
