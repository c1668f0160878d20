/* b2slam.h — C ABI of libb2slam.so, the B200 (sm_100a) scan-registration and
 * map-fusion hot path.
 *
 * The reference (MatthewKazan/3D-lidar-slam) has no FFI of its own: its plugin
 * `ICPProcessor` (src/slam/scripts/point_cloud_processors/icp_processor.py:18-160)
 * reaches its arithmetic through Open3D's pybind11 module.  Each entry point
 * below is what a binding for that path binds instead; the Open3D call and the
 * reference line it replaces are cited per function.  Plain pointers and sizes
 * only — no torch / numpy types cross this boundary.
 *
 * Conventions
 *   - every function returns B2_OK (0) or a negative B2_ERR_* code;
 *     b2_last_error(h) returns the message of the last failure on that handle.
 *   - `*_dev` pointers are caller-owned CUDA device pointers on the handle's
 *     device.  Point clouds are float32 [n,4] rows (x, y, z, w): 16-byte rows so
 *     one point is one 128-bit load; w is ignored on input unless stated.
 *   - 4x4 transforms are row-major float64[16] in HOST memory, last row must be
 *     (0,0,0,1).
 *   - all work is enqueued on the handle's stream (b2_set_stream); functions
 *     that return data-dependent scalars to the host synchronise that stream.
 *   - there is no CPU fallback anywhere: without a CUDA device b2_create fails.
 */
#ifndef B2SLAM_H_
#define B2SLAM_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_context_s* b2_handle;

enum {
    B2_OK = 0,
    B2_ERR_INVALID = -1,   /* bad argument */
    B2_ERR_CUDA = -2,      /* CUDA runtime failure (message has the call) */
    B2_ERR_CAPACITY = -3,  /* resident map / hash table full */
    B2_ERR_KEY_RANGE = -4, /* voxel index outside int32 / packed key range */
    B2_ERR_STATE = -5      /* call order violated (e.g. map not initialised) */
};

enum { B2_ICP_POINT_TO_POINT = 0, B2_ICP_POINT_TO_PLANE = 1 };

/* o3d.pipelines.registration.ICPConvergenceCriteria + max_correspondence_distance
 * + estimation method (icp_processor.py:103-113). */
typedef struct b2_icp_params {
    double max_correspondence_distance;
    int max_iteration;
    int mode; /* B2_ICP_POINT_TO_POINT | B2_ICP_POINT_TO_PLANE */
    double relative_fitness;
    double relative_rmse;
} b2_icp_params;

/* o3d RegistrationResult: transformation, fitness, inlier_rmse (+ iterations run). */
typedef struct b2_icp_result {
    double transformation[16];
    double fitness;
    double inlier_rmse;
    int iterations;
    int n_correspondences;
} b2_icp_result;

/* ---- lifecycle ---------------------------------------------------------- */
const char* b2_version(void);
int b2_create(int device, b2_handle* out);
int b2_destroy(b2_handle h);
const char* b2_last_error(b2_handle h);
int b2_set_stream(b2_handle h, void* cuda_stream); /* cudaStream_t; NULL = handle-owned stream */
int b2_synchronize(b2_handle h);
/* number of kernels this handle has launched so far (bench.py "gpu_launches") */
unsigned long long b2_launch_count(b2_handle h);
/* Which kernel runs the ICP step on pair / slab grids (results are identical; tests and tuning):
 * B2_ICP_DRIVER_AUTO picks by problem size, _GROUP = 4 lanes per query (latency), _SWEEP = one thread
 * per query with a pruned stencil walk (throughput).  Dense resident maps always use the group kernel. */
#define B2_ICP_DRIVER_AUTO 0
#define B2_ICP_DRIVER_GROUP 1
#define B2_ICP_DRIVER_SWEEP 2
#define B2_ICP_DRIVER_SWEEP4 3 /* experiment (measured slower), only in B2_EXPERIMENTS=1 builds: four lanes per query */
int b2_set_icp_driver(b2_handle h, int driver);
/* How b2_slam_scan / b2_slam_scan_dev voxelise a scan (both the 2 cm down-sample in front of ICP and the
 * fusion into the map; voxel keys, occupancy and counts are identical, centroids agree to ~1e-15 relative):
 * B2_SCAN_PATH_HASH (default) = one scratch hash per scan filled with atomics, no sort (b2_stream.cu: 8 launches);
 * B2_SCAN_PATH_SORTED = the radix-sort pipeline of b2_voxel_downsample (28 launches; in-order fp64 sums, the
 * down-sampled source comes out in voxel-key order).  The environment variable B2_STREAM_SORTED=1 selects
 * SORTED at b2_create. */
#define B2_SCAN_PATH_HASH 0
#define B2_SCAN_PATH_SORTED 1
int b2_set_scan_path(b2_handle h, int path);
/* How a registration against the brick-major resident map (voxel edge >= gate) is launched; results are identical bit
 * for bit (the same kernel body, tests run both):
 * B2_DENSE_DRIVER_PERSISTENT (default) = ONE launch per registration: one CTA per SM stays resident, the ICP steps
 *   are trips of a loop, and the CTAs exchange their partial sums as self-validating tagged records they poll for
 *   (no kernel boundary, no launches that only find "done");
 * B2_DENSE_DRIVER_CHAINED = one launch per ICP step, chained with programmatic dependent launch (+ a CUDA-graph
 *   WHILE node beyond 40 steps).  The environment variable B2_DENSE_PERSIST=0 selects CHAINED at b2_create. */
#define B2_DENSE_DRIVER_PERSISTENT 0
#define B2_DENSE_DRIVER_CHAINED 1
int b2_set_dense_driver(b2_handle h, int driver);

/* ---- layout helpers ----------------------------------------------------- */
/* float32 [n,3] (host when src_on_host != 0, else device) -> float32 [n,4] device rows.
 * Replaces o3d.utility.Vector3dVector(points) (icp_processor.py:60,65). */
int b2_pack_xyz(b2_handle h, const float* xyz, int n, int src_on_host, float* out_xyzw_dev);
/* float32 [n,4] device rows -> float32 [n,3] (host or device).  Replaces
 * np.asarray(pcd.points) (icp_processor.py:74). */
int b2_unpack_xyz(b2_handle h, const float* xyzw_dev, int n, float* out_xyz, int dst_on_host);
/* ProcessPointClouds.project_pixel_to_3d (generic_point_cloud_processor.py:93-108):
 * records (x_px, y_px, depth_m) float32 [n,3] on the host -> metric float32 [n,4] device rows.
 * numpy2_semantics != 0: float32 arithmetic (NEP 50); else float64 arithmetic rounded once. */
int b2_project_pixels(b2_handle h, const float* records_host, int n, double fx, double fy, double cx, double cy,
                      int numpy2_semantics, float* out_xyzw_dev);
/* sensor_msgs_py.point_cloud2.read_points(msg, field_names=("x","y","z"), skip_nans=True)
 * (input_handler.py:65-66) on the raw PointCloud2 payload: n_points records of point_step bytes, the three
 * float32 fields at byte offsets off_x/off_y/off_z (ARDepthViewController.swift:136-150 sends 12-byte
 * records, offsets 0/4/8, little endian).  Records with a NaN in any of the three fields are dropped
 * (skip_nans != 0), order preserved.  intrinsics != NULL ({fx, fy, cx, cy}): the pixel -> metric projection
 * of b2_project_pixels is applied to the kept records in the same pass.  data may be host or device memory.
 * out_xyzw_dev: float32 [n_points,4] device rows (first *n_out valid).  SURVEY.md section 8 row f3. */
int b2_ingest_pointcloud2(b2_handle h, const void* data, int data_on_host, int n_points, int point_step, int off_x,
                          int off_y, int off_z, int is_bigendian, int skip_nans, const double* intrinsics,
                          int numpy2_semantics, float* out_xyzw_dev, int* n_out);

/* ---- stateless primitives ----------------------------------------------- */
/* PointCloud.voxel_down_sample(voxel) (icp_processor.py:90-91,132-133).
 * origin == NULL: Open3D's own origin (min_bound - voxel/2).  T != NULL: points are
 * transformed first (fuses icp_processor.py:73).  Output order is canonical
 * (lexicographic voxel key); out_pts_dev [n,4] (w = bit pattern of the int32 point
 * count of the voxel), out_keys_dev int32 [n,3] or NULL, *n_out = voxel count. */
int b2_voxel_downsample(b2_handle h, const float* pts_dev, int n, double voxel, const double* origin,
                        const double* T, float* out_pts_dev, int* out_keys_dev, int* n_out);
/* The same down-sample for ONE scan (n <= 262,144 points) without a sort — the pass b2_slam_scan runs in front
 * of ICP (icp_processor.py:90): Open3D's origin, centroids bit-identical to b2_voxel_downsample's, emitted in the
 * order of each voxel's first point in the input (pixel order) instead of voxel-key order. */
int b2_scan_downsample(b2_handle h, const float* pts_dev, int n, double voxel, float* out_pts_dev, int* n_out);

/* PointCloud.estimate_normals(KDTreeSearchParamHybrid(radius, max_nn))
 * (icp_processor.py:94-99).  out_normals_dev float32 [n,4] (w = 0).
 * Normal sign is unspecified (no orientation step), as in Open3D. */
int b2_estimate_normals(b2_handle h, const float* pts_dev, int n, double radius, int max_nn,
                        float* out_normals_dev);

/* One correspondence pass: KDTreeFlann.SearchHybrid(T*p, radius, 1) per source point
 * (the Eval step of registration_icp).  out_idx_dev int32 [ns]
 * (-1 = no neighbour with d^2 < radius^2), out_d2_dev float64 [ns] or NULL. */
int b2_nn_search(b2_handle h, const float* src_dev, int ns, const float* tgt_dev, int nt, const double* T,
                 double radius, int* out_idx_dev, double* out_d2_dev);

/* o3d.pipelines.registration.registration_icp(source, target, max_corr, init, estimation,
 * criteria) (icp_processor.py:103-113).  tgt_normals_dev is required for
 * point-to-plane.  out_corr_dev int32 [ns] or NULL receives the final correspondence map. */
int b2_icp(b2_handle h, const float* src_dev, int ns, const float* tgt_dev, int nt,
           const float* tgt_normals_dev, const double* init, const b2_icp_params* params,
           b2_icp_result* out, int* out_corr_dev);

/* PointCloud.transform(T) (icp_processor.py:73). In-place allowed. */
int b2_transform(b2_handle h, const float* pts_dev, int n, const double* T, float* out_pts_dev);

/* ---- map maintenance (ICPProcessor.do_stuff, icp_processor.py:121-148) ---- */
/* PointCloud.remove_radius_outlier(nb_points, radius) (icp_processor.py:139-140):
 * keep_dev uint8 [n] = 1 where more than nb_points points (the point itself included) lie within
 * radius (d^2 < r^2); *n_kept may be NULL. */
int b2_radius_outlier_mask(b2_handle h, const float* pts_dev, int n, int nb_points, double radius,
                           unsigned char* keep_dev, int* n_kept);
/* PointCloud.remove_statistical_outlier(nb_neighbors, std_ratio) (icp_processor.py:136-137,145-146;
 * SURVEY A.8): exact k-nearest mean distance per point, keep where 0 < mean < mu + std_ratio * sigma.
 * cell <= 0: the search cell is sized from the cloud's density. */
int b2_statistical_outlier_mask(b2_handle h, const float* pts_dev, int n, int nb_neighbors, double std_ratio,
                                double cell, unsigned char* keep_dev, int* n_kept);
/* PointCloud.select_by_index of the kept points, order preserved: out_pts_dev float32 [n,4]. */
int b2_select_points(b2_handle h, const float* pts_dev, int n, const unsigned char* keep_dev, float* out_pts_dev,
                     int* n_out);

/* ---- resident global map ------------------------------------------------ */
/* The reference keeps the map as a host ndarray that it re-voxelises, re-indexes and
 * re-uploads for every scan (icp_processor.py:64-65,91,103).  Here the map is a
 * GPU-resident voxel hash: (sum xyz, n) per voxel of edge `voxel` anchored at
 * `origin`, grouped in cells of edge >= max_corr so it doubles as the uniform grid
 * of the correspondence search.  capacity_cells: hash slots (power of two is
 * enforced by rounding up); the map reports B2_ERR_CAPACITY beyond 70% load. */
int b2_map_init(b2_handle h, double voxel, const double* origin, double max_corr, long long capacity_cells);
int b2_map_reset(b2_handle h); /* ProcessPointClouds.reset (generic_point_cloud_processor.py:128-138) */
/* point_cloud.transform(T) + (point_cloud + map) + voxel merge (icp_processor.py:73-74,132-133). */
int b2_map_fuse(b2_handle h, const float* pts_dev, int n, const double* T);
int b2_map_size(b2_handle h, long long* n_voxels);
/* Voxel centroids in canonical key order: out_pts_dev float32 [max_out,4] (w = count bits),
 * out_keys_dev int32 [max_out,3] or NULL. */
int b2_map_export(b2_handle h, float* out_pts_dev, int* out_keys_dev, long long max_out, long long* n_out);
/* The map as mergeable records, canonical key order: out_keys_dev int32 [max_out,3] voxel keys and
 * out_sums_dev float64 [max_out,4] = (sum x, sum y, sum z, point count) of every voxel — what ranks exchange
 * when partial maps are gathered.  b2_map_merge_records adds such records (unique
 * keys within one call) to the resident map: per voxel the sums and counts add up, so merging the records of
 * several partial maps equals voxel_down_sample of the concatenated points (icp_processor.py:74,132-133). */
int b2_map_export_records(b2_handle h, int* out_keys_dev, double* out_sums_dev, long long max_out, long long* n_out);
int b2_map_merge_records(b2_handle h, const int* keys_dev, const double* sums_dev, long long n);
/* registration_icp with the resident map as the target (icp_processor.py:68-69,103-113). */
int b2_map_icp(b2_handle h, const float* src_dev, int ns, const double* init, const b2_icp_params* params,
               b2_icp_result* out);

/* The same registration, also reporting the correspondence set of the LAST evaluated transform — what Open3D's
 * RegistrationResult.correspondence_set holds (icp_processor.py:103-113): d2_out_dev float64 [ns]
 * squared distance of every source point to its nearest map voxel within the gate (0 where there is none),
 * match_out_dev float32 [ns,4] that voxel's centroid (w = 1) or zeros.  When the iteration budget ends the
 * registration, that transform is the result's.  Brick-major maps (voxel edge >= gate) only. */
int b2_map_icp_corr(b2_handle h, const float* src_dev, int ns, const double* init, const b2_icp_params* params,
                    b2_icp_result* out, double* d2_out_dev, float* match_out_dev);

/* ---- fused per-scan step (ICPProcessor.construct_global_map, icp_processor.py:43-75) ---- */
/* xyz_host: float32 [n,3] scan in the sensor frame (the plugin-boundary format).
 * Runs: upload -> voxel_down_sample(ds_voxel) -> ICP against the resident map seeded with
 * the previous pose (kept on the device) -> transform + fuse the RAW scan.  When the map
 * is empty the scan is fused with the current pose and no ICP runs (icp_processor.py:53-56).
 * out may be NULL (no host read-back, fully asynchronous).  xyz_host may be pinned (cudaHostAlloc /
 * cudaHostRegister: uploaded in place) or ordinary pageable memory (copied into a pinned staging ring of the
 * library first, so the caller's buffer may be released as soon as the call returns). */
int b2_slam_scan(b2_handle h, const float* xyz_host, int n, double ds_voxel, const b2_icp_params* params,
                 b2_icp_result* out);
/* ProcessPointClouds.process (generic_point_cloud_processor.py:111-126) in one call: the same step fed with the
 * wire-format records (x_px, y_px, depth_m) float32 [n,3] of one depth image; the pixel -> metric projection of
 * b2_project_pixels (generic_point_cloud_processor.py:93-108; numpy2_semantics as there) runs on the device as
 * part of the upload, on the library's front stream, so it overlaps the registration of the previous scan. */
int b2_slam_scan_records(b2_handle h, const float* records_host, int n, double fx, double fy, double cx, double cy,
                         int numpy2_semantics, double ds_voxel, const b2_icp_params* params, b2_icp_result* out);
/* Same step for a scan that already lives on the device as float32 [n,4] rows (offline replay of
 * device-resident batches): no upload, otherwise identical.  The rows are copied on the library's front stream,
 * which is ordered after every library call made on this handle so far; a scan written by the CALLER'S OWN kernels
 * must be complete (or ordered before this call by b2_synchronize / an event wait on the handle's stream plus any
 * library call) when the call is made. */
int b2_slam_scan_dev(b2_handle h, const float* scan_xyzw_dev, int n, double ds_voxel, const b2_icp_params* params,
                     b2_icp_result* out);
/* Deferred read-back for callers that submit scans faster than they need the poses: pass out == NULL to
 * b2_slam_scan / b2_slam_scan_dev (the call then only enqueues work; a PINNED host scan is read in place and must stay
 * valid until the next synchronising call, a pageable one is copied before the call returns) and collect the results of the last `count` scans here, oldest
 * first (one stream synchronisation for the whole batch; at most 1024 results are kept). */
int b2_slam_results(b2_handle h, int count, b2_icp_result* out);
/* Number of resident-map points inside the 27-cell stencils of the transformed source points —
 * the data-dependent term of the search's algorithmic-byte model. */
int b2_map_stencil_candidates(b2_handle h, const float* src_dev, int ns, const double* T, long long* n_candidates);
/* The same census for an arbitrary target cloud gridded with cell edge `cell` (pair mode). */
int b2_stencil_candidates(b2_handle h, const float* src_dev, int ns, const float* tgt_dev, int nt, const double* T,
                          double cell, long long* n_candidates);
int b2_slam_set_pose(b2_handle h, const double* T);
int b2_slam_get_pose(b2_handle h, double* T);

/* ---- batched independent pairs (loop-closure candidates / offline replay) ----
 * The batch calls are ASYNCHRONOUS: they enqueue the whole batch on the handle's stream and return; the
 * records in results_dev are complete, and device-side failures (B2_ERR_KEY_RANGE from a grid or voxel sort,
 * ...) are reported, by the next b2_synchronize(h) — call it before reading results_dev from another stream. */
/* src_dev/tgt_dev: concatenated float32 [sum,4]; *_offsets_dev: int32 [n_pairs+1] (device).
 * ds_voxel > 0: both clouds of every pair are voxel-down-sampled first (icp_processor.py:90-91).
 * init_dev: float64 [n_pairs,16] device or NULL (identity).  results_dev: device array of
 * n_pairs b2_icp_result records (gathered across ranks by the caller, e.g. with NCCL). */
int b2_icp_batch(b2_handle h, const float* src_dev, const int* src_offsets_dev, int src_total,
                 const float* tgt_dev, const int* tgt_offsets_dev, int tgt_total, int n_pairs, double ds_voxel,
                 const double* init_dev, const b2_icp_params* params, b2_icp_result* results_dev);
/* Point-to-plane batch (BASELINE config 3 over many pairs): the normals of every (down-sampled) target are
 * estimated first — estimate_normals(KDTreeSearchParamHybrid(normal_radius, normal_max_nn)), icp_processor.py:94-99
 * uses radius = 2 voxels, max_nn = 20 — then the batch runs with params->mode = B2_ICP_POINT_TO_PLANE. */
int b2_icp_batch_p2l(b2_handle h, const float* src_dev, const int* src_offsets_dev, int src_total, const float* tgt_dev,
                     const int* tgt_offsets_dev, int tgt_total, int n_pairs, double ds_voxel, double normal_radius,
                     int normal_max_nn, const double* init_dev, const b2_icp_params* params, b2_icp_result* results_dev);
/* b2_estimate_normals over a batch of clouds (concatenated rows, int32 [n_clouds+1] device offsets): neighbours are
 * searched inside a point's own cloud only. */
int b2_estimate_normals_batch(b2_handle h, const float* pts_dev, const int* offsets_dev, int total, int n_clouds,
                              double radius, int max_nn, float* out_normals_dev);

#ifdef __cplusplus
}
#endif
#endif /* B2SLAM_H_ */
