/* dmf_b200.h -- C ABI of libdmf_b200.so: the RayTracingEngine hot path of REXJJ/depth-map-fusion-utils as
 * hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI layer: the boundary it exposes is the public C++ surface of three headers
 * (include/Camera.hpp, include/Volume.hpp, include/RayTracingEngine.hpp).  The drop-in headers in
 * depth-map-fusion-utils_b200/dropin/ keep those class/method signatures and forward to the entry points
 * below; each entry point cites the reference interface it replaces.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; dmf_last_error() gives the message
 *     (the reference has no error convention at all: include/RayTracingEngine.hpp prints and continues).
 *   - poses are row-major 3x4 float camera->world affines, i.e. rows 0..2 of Eigen::Affine3f::matrix().
 *   - voxel ids are VoxelVolume::getHashId values: (x<<40)^(y<<20)^z  (include/Volume.hpp:143-148).
 *   - "occupied order" is the order of VoxelVolume::occupied_cells_ (first insertion, Volume.hpp:216).
 *   - one host thread per context; a context owns one CUDA device, its streams and all device buffers.
 *   - the *_dev entry points enqueue on the caller's stream and return; calls on one context share scratch buffers, so the
 *     library orders each call after the previous one on that context (an event wait when the stream changes).  Host-buffer
 *     calls (dmf_forward, dmf_reverse, ...) run on the context's own streams and synchronise; do not overlap them with *_dev
 *     work of the same context that is still in flight on another stream.
 *   - there is NO CPU fallback: every compute entry point fails if no sm_100 device is usable.
 */
#ifndef DMF_B200_H
#define DMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmf_ctx dmf_ctx;

#define DMF_NO_VOXEL 0xFFFFFFFFFFFFFFFFull

/* forward routines of include/RayTracingEngine.hpp */
enum {
    DMF_MODE_POINTS      = 0, /* rayTraceAndGetPoints      :447-494 */
    DMF_MODE_GOOD_POINTS = 1, /* rayTraceAndGetGoodPoints  :377-445 */
    DMF_MODE_CLASSIFY    = 2, /* rayTraceAndClassify       :311-375 */
    DMF_MODE_MARK        = 3, /* rayTrace                  :268-309 */
    DMF_MODE_MINIMUM     = 4  /* rayTraceAndGetMinimum     :229-264 */
};

/* occupancy format probed by the march */
enum {
    DMF_GRID_BIT  = 0, /* linear bit grid, [x][y][z] z fastest, uint32 words (1/8 byte per voxel)  */
    DMF_GRID_BYTE = 1, /* one byte per voxel, [x][y][z] z fastest like voxels_[x][y][z] (Volume.hpp:126): Chebyshev distance bytes  */
    DMF_GRID_AUTO = 2  /* the library picks: the bit grid for the first forward call on a volume, the distance bytes once they exist
                          (the reverse march builds them) or from the second call on (results are identical either way)          */
};

/* ---- lifetime --------------------------------------------------------------------------------- */
int  dmf_create(dmf_ctx** out, int device);              /* new: process-global state the by-value RayTracingEngine cannot own */
void dmf_destroy(dmf_ctx* ctx);
const char* dmf_last_error(void);
int  dmf_device_count(void);                             /* CUDA devices visible; 0 => compute calls fail */
int  dmf_version(void);

/* pinned host memory for the host-buffer entry points (plain malloc'd buffers work too, slower) */
void* dmf_host_alloc(size_t bytes);
void  dmf_host_free(void* p);

/* ---- Camera(K, height, width)                                        include/Camera.hpp:23 ---- */
int dmf_set_camera(dmf_ctx* ctx, const float K[9], int height, int width);

/* ---- VoxelVolume ------------------------------------------------------------------------------ */
/* Upload a volume that was built on the host by the (drop-in) VoxelVolume: replaces the per-probe
 * voxels_[x][y][z] reads of RayTracingEngine.hpp (:62,:97,:154,:194,:255,:298,:344,:414,:479,:517).
 * bounds = xmin,xmax,ymin,ymax,zmin,zmax; delta = xdelta_,ydelta_,zdelta_; dim = xdim_,ydim_,zdim_
 * (Volume.hpp:55-59).  occupied_ids = occupied_cells_; normal_offsets (n_occ+1) / normals (xyz) are the
 * per-voxel Voxel::normals lists in occupied order (Volume.hpp:32); both may be NULL (no normals). */
int dmf_upload_volume(dmf_ctx* ctx, const double bounds[6], const double delta[3], const int dim[3],
                      const uint64_t* occupied_ids, size_t n_occ,
                      const uint32_t* normal_offsets, const float* normals);

/* Build the volume inside the library from a point cloud: setDimensions + setVolumeSize + constructVolume
 * + integratePointCloud(cloud, normals) (Volume.hpp:89-128,199-228), then upload.  normals may be NULL
 * (the xyz-only overload, Volume.hpp:172-197). */
int dmf_volume_from_points(dmf_ctx* ctx, const double bounds[6], const int dims[3],
                           const float* xyz, const float* normals, size_t n_points);

/* The same with integratePointCloud itself on the GPU (first-insertion order of occupied_cells_ and point order of every
 * voxel's normal list preserved); for large clouds.  Host pointers in, identical volume out. */
int dmf_volume_from_points_gpu(dmf_ctx* ctx, const double bounds[6], const int dims[3],
                               const float* xyz, const float* normals, size_t n_points);

/* dims[3], deltas[3], n_occ, n_normals of the uploaded volume (any pointer may be NULL) */
int dmf_volume_info(dmf_ctx* ctx, int dims[3], double deltas[3], double* voxel_size, size_t* n_occ, size_t* n_normals);
int dmf_volume_get_occupied(dmf_ctx* ctx, uint64_t* ids /* n_occ */);
/* CSR of the per-voxel normal lists in occupied order: offsets[n_occ+1], normals[3*n_normals] */
int dmf_volume_get_normals(dmf_ctx* ctx, uint32_t* offsets, float* normals);

/* Voxel::view / Voxel::good (Volume.hpp:33-34) live on the device between calls. */
int dmf_clear_marks(dmf_ctx* ctx);
/* n = entries in the caller's arrays; must equal the uploaded volume's n_occ (fails otherwise: a mismatch means the caller's
 * VoxelVolume is not the one mirrored on the device).  n == 0 is a no-op. */
int dmf_download_marks(dmf_ctx* ctx, int32_t* view /* n */, uint8_t* good /* n */, size_t n);
int dmf_upload_marks(dmf_ctx* ctx, const int32_t* view /* n */, const uint8_t* good /* n */, size_t n);

/* ---- forward per-pixel march ------------------------------------------------------------------ */
typedef struct {
    int mode;        /* DMF_MODE_*                                                                    */
    int zdelta;      /* z-plane stride in mm (reference defaults: 10; 1 for MINIMUM)                   */
    int sparse;      /* pixel stride 5 (10 for MINIMUM) instead of 1                                   */
    int view_id0;    /* CLASSIFY: `view` argument of pose 0; pose i uses view_id0 + i                  */
    int grid_format; /* DMF_GRID_*                                                                     */
    int flags;       /* DMF_FWD_*                                                                      */
} dmf_forward_params;

/* By default the march skips probes that a macro-cell distance field proves to be in-bounds misses (results, id
 * lists and the samples/inbounds counters are identical either way; DMF_CNT_SKIPPED says how many were skipped).
 * DMF_FWD_NO_SKIP evaluates every probe (the brute-force kernel; used to validate the skipping one). */
#define DMF_FWD_NO_SKIP 1
/* DMF_GRID_BYTE only: use the previous generation of the skipping march (every probed sample evaluated exactly, two
 * per iteration) instead of the line-first one.  Same results; kept as the A/B baseline for profiles/. */
#define DMF_FWD_TWO_PROBE 2

/* Carve mode -- "occupied/free voxel marking".  NOT in the reference (it never records free space); defined by and pinned to
 * oracle/dmf_oracle.hpp only (PixelOut::observed): every sample that passes validPoints (Volume.hpp:230-233) and is visited
 * by its ray up to and including the ray's first hit -- exactly the samples DMF_CNT_INBOUNDS counts -- sets the bit of the
 * voxel getVoxel (Volume.hpp:150-156) puts it in, in a per-context "observed" bit grid that accumulates over views and calls
 * until dmf_clear_observed / a new volume.  observed & ~occupied = voxels seen free; observed & occupied = voxels hit.
 * Every sample has to be located, so none is skipped: with DMF_GRID_BYTE the line-first kernel finds the hit and then places
 * the samples on the line (exact evaluation only near voxel faces and the volume boundary); DMF_FWD_NO_SKIP or DMF_GRID_BIT
 * select the brute-force march that evaluates every sample the reference's way.  Not available in MINIMUM mode. */
#define DMF_FWD_CARVE 4
/* Do not update the DMF_CNT_* probe counters in this call.  They are instrumentation the reference does not have (the parity
 * tests use them to show that skipping never changes which samples are visited); a production sweep saves the warp reductions
 * and atomics.  Results are unaffected. */
#define DMF_FWD_NO_COUNTERS 8

/* Per-view outputs; any pointer may be NULL.  For the *_dev entry point these are device pointers. */
typedef struct {
    int32_t*  depth_mm;    /* [n_views][H][W]    z_depth (mm) of each pixel's first occupied sample, -1 = none/not cast */
    float*    points;      /* [n_views][H][W][3] world-space sample point of that first hit (simulated depth cloud), 0 if none */
    uint64_t* hit_voxel;   /* [n_views][H][W]    voxel id of the first hit, DMF_NO_VOXEL if none           */
    uint64_t* visibility;  /* [n_views][vis_words] bit i <=> occupied_cells_[i] is in the view's returned id list */
    int32_t*  found_any;   /* [n_views]          .first of the returned pair (any pixel hit)               */
    int32_t*  min_depth;   /* [n_views]          rayTraceAndGetMinimum result (-1 = none); MINIMUM mode    */
    uint64_t* ids;         /* returned id lists, discovery order (z_depth, r, c), views concatenated       */
    int64_t*  ids_offsets; /* [n_views+1]        view v owns ids[ids_offsets[v] .. ids_offsets[v+1])       */
    size_t    ids_capacity;/* entries available in ids; fails (no partial write) if too small             */
    uint16_t* depth_u16;   /* [n_views][H][W]    the same first-hit z_depth as 16 bits (z_depth < 1000 always), 0xFFFF = none/not cast:
                              half the device->host bytes of depth_mm for consumers that only need the depth map */
} dmf_forward_out;

size_t dmf_visibility_words(dmf_ctx* ctx);   /* ceil(n_occ/64) */

/* The observed bit grid of carve mode: uint32 words over the PADDED index space [0,dim_x] x [0,dim_y] x [0,dim_z]
 * (z fastest, like voxels_[x][y][z]): bit index = (x*(dim_y+1) + y)*(dim_z+1) + z, bit i of word w = index 32w+i. */
size_t dmf_observed_words(dmf_ctx* ctx);                       /* number of uint32 words (a multiple of 8)                  */
int dmf_clear_observed(dmf_ctx* ctx);
int dmf_download_observed(dmf_ctx* ctx, uint32_t* words /* dmf_observed_words */);
int dmf_observed_dev(dmf_ctx* ctx, void** d_words);           /* device pointer (for OR-reducing the grids of several GPUs) */
int dmf_observed_counts(dmf_ctx* ctx, uint64_t out[3]);       /* observed, observed & occupied (hit), observed & ~occupied (free) */

/* Host-buffer call: copies poses H2D, casts all views, copies the requested outputs D2H, synchronises.
 * Replaces n_views consecutive calls of one forward routine (RayTracingEngine.hpp:229,268,311,377,447). */
int dmf_forward(dmf_ctx* ctx, const dmf_forward_params* p, const float* poses, int n_views, const dmf_forward_out* out);

/* Device-buffer call: poses and outputs already in HBM (ids / ids_offsets not supported here); enqueues on
 * `stream` (a cudaStream_t, NULL = the context's stream) and returns without synchronising. */
int dmf_forward_dev(dmf_ctx* ctx, const dmf_forward_params* p, const float* d_poses, int n_views,
                    const dmf_forward_out* d_out, void* stream);

/* ---- reverse per-voxel march ------------------------------------------------------------------ */
typedef struct {
    uint64_t* visibility;  /* [n_views][vis_words] bit i <=> occupied_cells_[i] emitted (good_points)      */
    uint64_t* unoccluded;  /* [n_views][vis_words] bit i <=> occupied_cells_[i] not occluded (view=1 when viz) */
    int32_t*  found_any;   /* [n_views]                                                                   */
    uint64_t* ids;         /* emitted centroid hashes in emission order, views concatenated               */
    int64_t*  ids_offsets; /* [n_views+1]                                                                 */
    size_t    ids_capacity;
} dmf_reverse_out;

/* fast = 1: reverseRayTraceFast (:136-226); fast = 0: reverseRayTrace (:45-134).  viz != 0 also updates
 * Voxel::view / Voxel::good on the device (dmf_download_marks). */
/* Grid the reverse march probes: DMF_GRID_BYTE (default; distance bytes, skips provably empty steps) or DMF_GRID_BIT
 * (every step evaluated).  Results and counters are identical. */
int dmf_set_reverse_format(dmf_ctx* ctx, int grid_format);
int dmf_reverse(dmf_ctx* ctx, int fast, int viz, const float* poses, int n_views, const dmf_reverse_out* out);
int dmf_reverse_dev(dmf_ctx* ctx, int fast, int viz, const float* d_poses, int n_views, const dmf_reverse_out* d_out, void* stream);

/* ---- batched stand-off search: Algorithms::optimizeCameraPosition(volume, engine, res, Affine3f camera)
 *      (Algorithms.hpp:394-421; repositionCamera/moveCamera :170-188) for n cameras at once ------------------------- */
/* Binary search of the stand-off in [low0, high0] mm (reference: 300, 600) on the number of ids reverseRayTrace returns
 * at the two ends; every search step evaluates all still-active cameras in one batched reverse march.  mid_out[n] gets
 * the final `mid`, poses_out[n][12] the repositioned cameras; either may be NULL. */
int dmf_optimize_standoff(dmf_ctx* ctx, const float* poses, int n, unsigned low0, unsigned high0, uint32_t* mid_out, float* poses_out);

/* ---- segment collision: willCollide(volume, a, b) (tests/CameraPathGen.cpp:128-156) for n segments --------------- */
/* a, b: [n][3] end points; guard_coords != 0 reproduces the validCoords guard of CameraPathGen.cpp:147, 0 the unguarded
 * copies (CameraMotionTSP.cpp:236-261, CameraMotionPlanner.cpp:246-271).  out[i] = 1 if the 1 mm march from a to b meets
 * an occupied voxel. */
int dmf_segments_collide(dmf_ctx* ctx, const float* a, const float* b, int n, int guard_coords, uint8_t* out);

/* ---- z-buffer splat: rayTraceVolume (:498-564) ------------------------------------------------ */
/* depth receives the H*W z-buffer (mm, -1 = empty; not returned by the reference, exposed for parity);
 * Voxel::view is set to 1 on the device for voxels whose depth equals the buffer. */
int dmf_zbuffer(dmf_ctx* ctx, const float pose[12], int32_t* depth /* H*W or NULL */, int64_t* n_splat /* or NULL */);

/* ---- greedy set cover over visibility bitsets: Algorithms::greedySetCover (Algorithms.hpp:38-86) */
/* bitsets: [n_sets][words] host buffer (bit i = element i).  selected receives the chosen set indices in
 * selection order (capacity n_sets); *n_selected their count. */
int dmf_greedy_set_cover(dmf_ctx* ctx, const uint64_t* bitsets, int n_sets, size_t words, int32_t* selected, int* n_selected);
int dmf_greedy_set_cover_dev(dmf_ctx* ctx, const uint64_t* d_bitsets, int n_sets, size_t words, int32_t* selected, int* n_selected);

/* ---- bitwise OR of per-rank bitsets (the "seen" map combine step of a sharded sweep) ------------ */
/* d_dst[w] |= OR over r of d_src[r*words + w];  device pointers, enqueued on stream */
int dmf_or_reduce_dev(dmf_ctx* ctx, uint64_t* d_dst, const uint64_t* d_src, int n_src, size_t words, void* stream);

/* ---- multi-GPU: the candidate-view sweep sharded over the GPUs of one box ---------------------------------------------
 * Replaces the per-view loops of the reference's drivers (tests/SetCover.cpp:218-240, tests/CameraMotionPlanner.cpp:334-356,
 * tests/CameraPathGen.cpp:158-180: one reverseRayTraceFast / rayTraceAndGetPoints call per candidate view, one CPU thread).
 * Views are dealt round-robin (GPU r marches views r, r+N, ...), the volume is replicated on every GPU, and every GPU ends up
 * with every view's visibility row: each GPU's kernels store its finished rows straight into the peers' gathered buffers
 * through peer-mapped pointers over NVLink (no collective call; rows are disjoint by view), or -- where the GPUs cannot address
 * each other, or with DMF_COMM_EXCHANGE=nccl -- the rows travel through ncclAllGather.  NCCL is loaded with dlopen when a group is formed. */
typedef struct dmf_comm dmf_comm;
#define DMF_UNIQUE_ID_BYTES 128
/* how the visibility rows reach the peers: peer-mapped 16-byte stores over NVLink issued by one small push kernel right behind
 * the march (default), the same stores issued from the march kernels' own epilogue (DMF_COMM_EXCHANGE=epilogue: a ticket per
 * view finds the last block; measured slower, kept as the A/B), or ncclAllGather (DMF_COMM_EXCHANGE=nccl, and the fallback
 * where the GPUs cannot address each other) */
enum { DMF_EXCHANGE_NONE = 0, DMF_EXCHANGE_P2P_PUSH = 1, DMF_EXCHANGE_NCCL = 2, DMF_EXCHANGE_P2P_EPILOGUE = 3 };
enum { DMF_SWEEP_ROWS_OWN = 0, DMF_SWEEP_ROWS_ALL = 1 };

/* ONE process drives n_gpus devices (0 .. n_gpus-1; n_gpus <= 0: all visible, at most 8): creates a context per device,
 * enables peer access, forms the NCCL communicators (ncclCommInitAll).  The shape the reference's single-threaded drivers need. */
int dmf_comm_init_all(dmf_comm** out, int n_gpus);
/* One process per GPU (torchrun / MPI): rank 0 makes an id (dmf_comm_unique_id: ncclGetUniqueId), shares it out of band, every
 * rank calls dmf_comm_init_rank with its own context (ncclCommInitRank; peer buffers are mapped with CUDA IPC). */
int dmf_comm_unique_id(void* id /* DMF_UNIQUE_ID_BYTES */);
int dmf_comm_init_rank(dmf_comm** out, dmf_ctx* ctx, const void* unique_id, int rank, int world);
void dmf_comm_destroy(dmf_comm* comm);
int dmf_comm_info(dmf_comm* comm, int* world, int* n_local, int* first_rank, int* exchange /* DMF_EXCHANGE_* */);
dmf_ctx* dmf_comm_ctx(dmf_comm* comm, int local_index);           /* the context of local member i (0 .. n_local-1) */
int dmf_comm_set_camera(dmf_comm* comm, const float K[9], int height, int width);
/* The volume uploaded on rank `root` goes to every other GPU, GPU to GPU: only the occupied id list and the normals travel
 * (cudaMemcpyPeer / ncclBroadcast), each GPU rebuilds its march structures on the device.  Collective. */
int dmf_comm_replicate_volume(dmf_comm* comm, int root);
int dmf_comm_synchronize(dmf_comm* comm);

/* Host outputs of a sweep, in VIEW order; any pointer may be NULL. */
typedef struct {
    uint64_t* visibility;   /* [n_views][dmf_visibility_words]                                                        */
    int32_t*  found_any;    /* [n_views]                                                                              */
    int       rows_to_host; /* one process per GPU: DMF_SWEEP_ROWS_OWN copies back only the rows this rank marched (the others
                               are left untouched), DMF_SWEEP_ROWS_ALL the whole gathered array.  A single-process group always
                               fills the whole array: every GPU sends its own rows over its own PCIe link.                  */
} dmf_sweep_out;

/* n_views consecutive calls of one forward routine (POINTS, GOOD_POINTS or CLASSIFY; RayTracingEngine.hpp:311,377,447), resp. of
 * reverseRayTraceFast (:136-226), sharded over the group.  poses = the WHOLE list [n_views][12] (every rank passes the same
 * list).  Collective; synchronises when `out` is given.  Afterwards every member holds all rows (dmf_sweep_gathered_dev) and
 * dmf_sweep_set_cover runs Algorithms::greedySetCover (Algorithms.hpp:38-86) over them.  After a CLASSIFY sweep call
 * dmf_comm_fuse_marks; with DMF_FWD_CARVE call dmf_comm_fuse_observed. */
int dmf_sweep_forward(dmf_comm* comm, const dmf_forward_params* p, const float* poses, int n_views, const dmf_sweep_out* out);
int dmf_sweep_reverse(dmf_comm* comm, int fast, const float* poses, int n_views, const dmf_sweep_out* out);
/* The same with each local member's poses already on its device (d_poses[i][j] = global view first_rank + i + j * world) and
 * nothing copied back: enqueues on streams[i] (NULL array / entry = the member's own stream) and returns; the last thing
 * enqueued on each stream is the wait for the peers' rows. */
/* d_out (NULL, or one entry per local member, each may be NULL): device buffers for the per-pixel outputs of the member's own
 * views -- depth_mm / depth_u16 / points / hit_voxel, [views of the member][H][W] -- ; the other fields are ignored (visibility
 * and found_any live in the gathered rows). */
int dmf_sweep_forward_dev(dmf_comm* comm, const dmf_forward_params* p, const float* const* d_poses, int n_views,
                          const dmf_forward_out* const* d_out, void* const* streams);
int dmf_sweep_reverse_dev(dmf_comm* comm, int fast, const float* const* d_poses, int n_views, void* const* streams);
/* the gathered rows of the last sweep on local member i: [n_views] rows of row_words uint64 (vis_words of visibility, then one
 * word holding found_any, then padding) */
int dmf_sweep_gathered_dev(dmf_comm* comm, int local_index, uint64_t** d_rows, size_t* row_words, size_t* vis_words, int* n_views);
int dmf_sweep_set_cover(dmf_comm* comm, int32_t* selected /* capacity n_views */, int* n_selected);

/* Carve mode over the group: bitwise OR of the members' observed grids, as a reduce-scatter + all-gather over peer memory
 * (every GPU reduces 1/N of the words from all peers and pushes the result back); afterwards every grid is the union. */
int dmf_comm_fuse_observed(dmf_comm* comm);
/* After a sharded CLASSIFY sweep: Voxel::view = view_id0 + the smallest global view index that hit the voxel (first-wins in call
 * order, RayTracingEngine.hpp:354-355) where it was still 0 -- a min-reduce over the GPUs --, Voxel::good = OR (:356-370). */
int dmf_comm_fuse_marks(dmf_comm* comm, int view_id0);

/* ---- host utility (no GPU needed) ------------------------------------------------------------- */
/* The reference's good-point test  degree(acos(n.v)) in [0,90]  (CommonUtilities.hpp:17, RayTracingEngine.hpp:211-212)
 * depends on the HOST libm's float acos.  The library bisects it once; the kernels then test dot_min <= d <= 1.
 * out[0] = dot_min, out[1..2] = [band_lo, band_hi) where the host acosf was seen non-monotonic (empty if lo >= hi). */
int dmf_host_angle_test(float out[3]);

/* GPU self-test: exhaustive comparison (all 2^32 float inputs) of the kernels' division-free `x / 1000.0f` against the
 * IEEE division the reference performs (RayTracingEngine.hpp:82,173).  mismatches[0] = the two-step form the kernels use
 * (must be 0), mismatches[1] = the one-step form (informational). */
int dmf_selftest_div1000(dmf_ctx* ctx, uint64_t mismatches[5]);

/* ---- counters ---------------------------------------------------------------------------------- */
enum {
    DMF_CNT_SAMPLES = 0,   /* probes evaluated (pixel,z_depth) / (voxel,step)                              */
    DMF_CNT_INBOUNDS = 1,  /* probes that passed validPoints, up to and including the first hit           */
    DMF_CNT_HITS = 2,      /* forward: rays that hit; reverse: unoccluded voxels                          */
    DMF_CNT_EXACT_DIV = 3, /* probes whose voxel index needed the exact-division path (quotient near an integer) */
    DMF_CNT_OOB = 4,       /* reads the reference would perform out of bounds (treated as empty)          */
    DMF_CNT_ACOS_TIES = 5, /* good-point tests whose dot product fell in the host-libm acosf ambiguity band */
    DMF_CNT_LAUNCHES = 6,  /* kernels launched by this context                                           */
    DMF_CNT_RUNAWAY = 7,   /* reverse marches stopped by the step cap                                     */
    DMF_CNT_F64_PATH = 8,  /* probes whose float index filter was inconclusive and were redone in double           */
    DMF_CNT_SKIPPED = 9,   /* probes proven empty by the macro-cell traversal without being evaluated              */
    DMF_CNT_BOUNDS = 10,   /* checked build only (-DDMF_CHECKED): computed grid/table indices that were out of range; must be 0 */
    DMF_CNT_COUNT = 12
};
int dmf_counters(dmf_ctx* ctx, uint64_t out[DMF_CNT_COUNT]);  /* cumulative; synchronises */
int dmf_reset_counters(dmf_ctx* ctx);

/* duration in ms of the kernels of the most recent *_dev/host call on this context, measured with CUDA
 * events on the launching stream (synchronises) */
int dmf_last_kernel_ms(dmf_ctx* ctx, float* ms);
/* same, for the dominant march kernel(s) alone (the k_forward* / k_reverse launches of the last call or pass) */
int dmf_last_hot_kernel_ms(dmf_ctx* ctx, float* ms);
/* One-off cost of preparing the uploaded volume for the march, measured with CUDA events: build_ms = bit grid + rank directory
 * + macro-cell clearance + centroid hashes from the occupied id list (all on the device, dmf_volume.cuh); bytes_ms = the
 * Chebyshev distance bytes of DMF_GRID_BYTE (dmf_distance.cuh; -1 until that format has been used or dmf_prepare_grid called).
 * Replaces constructVolume + the per-probe pointer grid of the reference (Volume.hpp:119-128). */
int dmf_volume_prepare_ms(dmf_ctx* ctx, float* build_ms, float* bytes_ms);
/* build the structures of `grid_format` now instead of on first use (so that a timed first call does not include them) */
int dmf_prepare_grid(dmf_ctx* ctx, int grid_format);
int dmf_synchronize(dmf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DMF_B200_H */
