/* ls3d.h — C ABI of libls3d_b200.so, the B200 (sm_100a) implementation of LiveScan3D's per-frame
 * point-cloud hot path.  Plain C: pointers and sizes only, no C++/torch types.
 *
 * Part 1 keeps the reference's own exports (what LiveScanServer P/Invokes from NativeUtils.dll), same
 * names, argument meaning, ownership and "never throws, returns nothing" error behaviour; each entry
 * cites the reference declaration it replaces (paths relative to the LiveScan3D tree).
 * Part 2 adds flat-array entries for the stages that only have a C++ signature in the reference
 * (filter) or no single entry (the whole frame pipeline), plus the error side channel.
 * Part 3 is the device-resident API (device pointers + a cudaStream_t passed as void*): what bench.py's
 * HBM-resident timing and the multi-GPU host (one process per GPU, torch.distributed for collectives)
 * drive.  Part 4 covers the data formats either side of the path (client frame blob, frames dump, PLY, transfer frame).
 * Nothing here ever falls back to a CPU implementation: without a CUDA device every compute
 * entry fails and ls3d_last_error() says why.
 */
#ifndef LS3D_H
#define LS3D_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference data types (bit-identical layouts) -------------------------------------------------- */

/* include/NativeUtils/icp.h:15-18 (and include/LiveScanClient/utils.h:44-61): 12 bytes, no padding */
typedef struct Point3f { float X, Y, Z; } Point3f;

/* include/LiveScanClient/utils.h:105-111: the colour type filter() carries, 4 bytes */
typedef struct RGB { unsigned char rgbBlue, rgbGreen, rgbRed, rgbReserved; } RGB;

/* include/NativeUtils/depthprocessing.h:29-33 (C# twin LiveScanServer/Utils.cs:314-333): 16 bytes */
typedef struct VertexC4ubV3f { unsigned char R, G, B, A; float X, Y, Z; } VertexC4ubV3f;

/* include/NativeUtils/depthprocessing.h:42-48 (C# LiveScanServer/Utils.cs:335-342) */
typedef struct Mesh {
	int nVertices;
	VertexC4ubV3f *vertices;   /* callee-allocated; released by deleteMesh */
	int nTriangles;
	int *triangles;            /* callee-allocated (never NULL after a call); released by deleteMesh */
} Mesh;

/* ---- Part 1: the reference's exports -------------------------------------------------------------- */

/* Replaces include/NativeUtils/icp.h:65 (src/NativeUtils/icp.cpp:75-177); C# binding
 * LiveScanServer/MainWindowForm.cs:42-43.  Point-to-point ICP of verts2 onto verts1: verts2 is transformed
 * in place, R (row-major 3x3) and t are accumulated into (R = R*Rk, t += T*R^T per iteration).
 * Returns 1.0f like the reference (its error value is never computed, icp.cpp:85,176). */
float ICP(Point3f *verts1, Point3f *verts2, int nVerts1, int nVerts2, float *R, float *t, int maxIter);

/* Replaces include/NativeUtils/depthprocessing.h:103-105 (src/NativeUtils/depthprocessing.cpp:1631-1657);
 * C# binding LiveScanServer/KinectServer.cs:41-44.  One sensor (depth_map_index) of the packed frame:
 * depth -> camera space -> +t -> R* -> strict bbox cull -> row-major compaction -> Mesh (no triangles).
 * depth_maps: packed little-endian u16 images; depth_colors: packed RGB triples; intr_params: 7 floats per
 * map (cx,cy,fx,fy,r2,r4,r6); wtransform_params: 12 floats per map (t[3], R row-major). */
void generateVerticesFromDepthMap(unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int depth_map_index);

/* Replaces include/NativeUtils/depthprocessing.h:108-110 (src/NativeUtils/depthprocessing.cpp:1715-1792);
 * C# binding LiveScanServer/KinectServer.cs:35-38.  All sensors -> one merged Mesh in sensor order.
 * Result = the reference's (bcolor_transfer, bgenerate_triangles) = (false, false) branch: vertices plus the depth-grid
 * triangles generateTriangles always produces (depthprocessing.cpp:1786, meshGenerator.cpp:76-181), indices rebased per
 * sensor as formMesh does (:1611-1626).  The two flags only add colour correction and the multi-view vertex merge in
 * the reference (:1757-1778); both are outside this path (read as their low byte: the C# side marshals 4-byte BOOLs,
 * KinectServer.cs:36-38).  When either flag is set the call still succeeds and returns the (false, false) mesh, and
 * ls3d_last_error() carries a NON-FATAL text starting with "note:" that says so (the server's default is
 * bGenerateTriangles = true, KinectSettings.cs:50).  Host schedule as for ls3d_frame_pipeline (chunks of sensors, read-back overlapped with
 * the upload, one CUDA graph when the inputs are page-locked); Mesh.vertices / Mesh.triangles are page-locked blocks sized for the worst case. */
void generateMeshFromDepthMaps(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh, int bcolor_transfer,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int bgenerate_triangles);

/* Replaces include/NativeUtils/depthprocessing.h:111 (src/NativeUtils/depthprocessing.cpp:1794-1815, per sensor :191-261);
 * C# binding LiveScanServer/KinectServer.cs:51-53.  Radial-distortion correction of every sensor's depth and colour map,
 * IN PLACE in the caller's packed buffers: forward warp by 1 - r2*r - r4*r^2 - r6*r^3 (the last source pixel in raster order
 * wins a destination), then the reference's in-place raster-order hole fill (a zero pixel with more than 4 mutually
 * consistent non-zero neighbours becomes their integer mean).  Bit-identical to the reference.  The buffers are written only
 * after every kernel has finished and the device status has been read back clean: a failure before that leaves them untouched
 * and ls3d_last_error() says why (only a failing device-to-host transfer of the finished result can leave them undefined, and
 * the error text then says so). */
void depthMapAndColorSetRadialCorrection(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights, float *intr_params);

/* Replace src/NativeUtils/depthprocessing.cpp:1818-1835 (C# KinectServer.cs:56-60).  deleteMesh releases the
 * two arrays and leaves the struct itself alone, as the reference does. */
Mesh *createMesh(void);
void deleteMesh(Mesh *mesh);

/* ---- Part 2: flat-array entries for the C++-only stages ------------------------------------------ */

/* Replaces `std::unordered_map<int,int> filter(std::vector<Point3f>&, std::vector<RGB>&, int k, float maxDist)`
 * (include/LiveScanClient/filter.h:64, src/LiveScanClient/filter.cpp:36-81).  Removes point i iff the squared
 * distance to its k-th nearest neighbour (itself included) exceeds (float)pow(maxDist,2); verts and colors are
 * compacted in place in their original order; old_to_new[i] (may be NULL) receives the new index or -1.
 * k<=0 or maxDist<=0: nothing is touched, old_to_new is filled with -2 (the reference returns a map holding only
 * its {-1:-1} sentinel) and n is returned.  Returns the surviving count, or -1 on error. */
int ls3d_filter(Point3f *verts, RGB *colors, int n, int k, float maxDist, int *old_to_new);

/* The whole per-frame path in one call: per sensor map -> world transform -> cull (createVertices,
 * depthprocessing.cpp:122-187) -> neighbour-count filter per sensor (filter.cpp:36-81; skipped when
 * filter_k<=0 or filter_maxDist<=0) -> merge in sensor order (formMesh, depthprocessing.cpp:1578-1629).
 * per_map_counts (may be NULL) receives each sensor's surviving vertex count.  Returns total vertices or -1.
 *
 * Host schedule (organized filter path): the sensors are processed in chunks on four streams — upload, count, merge, read-back —
 * so a chunk's records travel to the host while the next chunk's depth is still arriving; Mesh.vertices is a page-locked
 * block the device stores into directly (freed by deleteMesh as usual).  If depth_maps / depth_colors are themselves
 * page-locked (cudaHostAlloc / cudaHostRegister'ed by the caller) the colours are never uploaded — the merge kernel reads the
 * surviving pixels' colour bytes out of the caller's buffer — and the whole schedule is replayed as one CUDA graph.
 * Pageable inputs take the same schedule with plain stream launches.  Tuning / diagnosis (environment, read once):
 * LS3D_E2E_MODE 2|1|0 (pull colours | upload colours | unpipelined), LS3D_E2E_CHUNKS (default 4), LS3D_E2E_GRAPH 1|0,
 * LS3D_E2E_TRACE n (print the device timeline of call n+8 to stderr). */
int ls3d_frame_pipeline(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ,
	int filter_k, float filter_maxDist, int *per_map_counts);

/* Replaces `void KinectCapture::filterFlyingPixels(int neighbourhoodSize, float thr, int maxNonFittingNeighbours)`
 * (include/LiveScanClient/kinectCapture.h:39, src/LiveScanClient/kinectCapture.cpp:132-174; called per acquired frame at :199) on one
 * width x height u16 depth image, in place: a pixel is zeroed when more than half of the (2k+1)^2-1 pixels around it differ from
 * it by more than thr.  maxNonFittingNeighbours is accepted and ignored, as in the reference (:150 overwrites it).  Returns 1 / -1. */
int ls3d_filter_flying_pixels(unsigned short *depth, int width, int height, int neighbourhoodSize, float thr, int maxNonFittingNeighbours);

/* FindClosestPointForEach (src/NativeUtils/icp.cpp:18-32) as a flat entry, for stage-level parity tests:
 * index (into verts1) and squared distance of the nearest verts1 point for every verts2 point.  Returns 0/-1. */
int ls3d_find_closest(const Point3f *verts1, int nVerts1, const Point3f *verts2, int nVerts2,
	unsigned long long *indices, float *distances);

/* Per-iteration trace of ICP (same fields as the oracle's), filled when a trace buffer is installed. */
typedef struct Ls3dIcpTrace {
	int n_matched;      /* one-to-one matches before rejection (icp.cpp:95-126) */
	int n_accepted;     /* after RejectOutlierMatches (icp.cpp:56-73) */
	float sigma;        /* GetStandardDeviation of squared distances (icp.cpp:34-54) */
	float T[3];         /* tempT (icp.cpp:141) */
	float Rk[9];        /* tempR (icp.cpp:153-163) */
} Ls3dIcpTrace;

/* ICP with a per-iteration trace (trace: maxIter entries, may be NULL).  Same contract as ICP(). */
float ls3d_icp_trace(Point3f *verts1, Point3f *verts2, int nVerts1, int nVerts2, float *R, float *t, int maxIter,
	Ls3dIcpTrace *trace);

/* Last error of the calling thread's most recent failing call ("" if none).  The reference has no error
 * reporting at all (every export returns void or a constant); this is the side channel that replaces the
 * C++ exceptions it would let escape into the CLR (nanoflann.h:904 on an empty target, OpenCV asserts). */
const char *ls3d_last_error(void);

/* Library/device identification: "ls3d-b200 <version> sm_100a; device: <name> (cc X.Y)" or the failure text. */
const char *ls3d_version(void);

/* Device self-check of arithmetic shortcuts that are exhaustive facts rather than identities (d / 1000.0f for every u16 d in three
 * instructions).  Returns the number of failures (0 = good) or -1 without a usable device. */
int ls3d_selftest(void);

/* Number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches). */
long long ls3d_launch_count(void);
void ls3d_reset_launch_count(void);

/* ---- Part 3: device-resident API ---------------------------------------------------------------- */

typedef struct Ls3dFrame Ls3dFrame;   /* persistent per-rig state: descriptors, scratch, voxel hash, outputs */
typedef struct Ls3dIcp Ls3dIcp;       /* persistent ICP state: target grid, dedupe slots, partial sums */

/* A frame context for n_maps sensors of the given sizes on the current CUDA device. */
Ls3dFrame *ls3d_frame_create(int n_maps, const int *widths, const int *heights);
void ls3d_frame_destroy(Ls3dFrame *f);

/* Host-side parameters (intrinsics 7/map, world transforms 12/map, bounds, filter settings): copied to the
 * device descriptors asynchronously on `stream`.  Returns 0/-1. */
int ls3d_frame_set_params(Ls3dFrame *f, const float *intr_params, const float *wtransform_params,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ,
	int filter_k, float filter_maxDist, void *stream);

/* Enqueue the whole path on `stream` for device-resident packed inputs (same layouts as the host API).
 * first_map/n_run select a contiguous sensor range (n_run<=0: all).  No host synchronisation.  The merged
 * cloud is left in ls3d_frame_vertices() (16-byte VertexC4ubV3f records), its length in device memory at
 * ls3d_frame_count_ptr().  Returns the number of kernels enqueued, or -1. */
int ls3d_frame_run(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run, void *stream);

const void *ls3d_frame_vertices(Ls3dFrame *f);        /* device pointer: merged, filtered cloud */
const void *ls3d_frame_culled_vertices(Ls3dFrame *f); /* device pointer: mapped+culled cloud before the filter */
const int *ls3d_frame_count_ptr(Ls3dFrame *f);        /* device int[5]: {n_final, n_culled, error_flags, n_kept, n_triangles} */
const int *ls3d_frame_sensor_starts(Ls3dFrame *f);    /* device int[n_maps+1]: start of each sensor in the merged cloud */
const int *ls3d_frame_culled_starts(Ls3dFrame *f);    /* device int[n_maps+1]: same for the culled cloud */
const int *ls3d_frame_old_to_new(Ls3dFrame *f);       /* device int[n_culled]: culled index -> merged index or -1 */
const unsigned char *ls3d_frame_keep_mask(Ls3dFrame *f); /* device u8[sum w*h]: 1 where the pixel's vertex survived cull + filter in the last ORGANIZED run
                                                           (the reference's filter mask in pixel order); NULL after any other run */
const int *ls3d_frame_depth_to_vertex(Ls3dFrame *f);  /* device int[sum w*h]: pixel -> index of its vertex in the culled cloud of the run (all sensors), or -1;
                                                         produced from the next run on (createVertices' depth_to_vertices_map + formMesh's rebasing) */

/* Triangle stage (generateTrianglesGradients, meshGenerator.cpp:76-181): when enabled, every UNFILTERED run also leaves the
 * triangles of all sensors (3 ints each: indices into the merged cloud, sensor order, raster order inside a sensor) in
 * ls3d_frame_triangles(), their number in ls3d_frame_count_ptr()[4] and per-sensor starts in ls3d_frame_triangle_starts(). */
void ls3d_frame_enable_triangles(Ls3dFrame *f, int on);
const int *ls3d_frame_triangles(Ls3dFrame *f);        /* device int[3 * n_triangles] */
const int *ls3d_frame_triangle_starts(Ls3dFrame *f);  /* device int[n_maps+1] */

/* How the neighbour count enumerates candidates (results are identical; only speed differs):
 *   0 auto      : organized (pixel-window) count when every sensor's pose and intrinsics admit its bound and the window of a point
 *                 1 m away fits the kernel's shared-memory halo (8 px), else voxel hash
 *   1 voxel hash: always the voxel-hash path (also what ls3d_filter uses: it has no image to exploit)
 *   2 organized : insist on the pixel-window path; ls3d_frame_run fails if it is not applicable.
 * On the organized path the unfiltered culled cloud is never materialised: ls3d_frame_culled_vertices and
 * ls3d_frame_old_to_new return NULL after such a run. */
int ls3d_frame_set_filter_mode(Ls3dFrame *f, int mode);

/* The same switch for the host-buffer entry points (ls3d_frame_pipeline); process-wide, default 0. */
int ls3d_set_default_filter_mode(int mode);

/* Optional per-stage timing for the roofline report: with timing on, ls3d_frame_run brackets its kernels with CUDA
 * events on the run's stream; ls3d_frame_stage_ms waits for the last run and returns, in milliseconds (0 = stage
 * did not run): [0] map/cull/compact [1] hash clear [2] voxel insert [3] cell ranges+scatter [4] voxel-hash neighbour
 * count [5] survivor compaction [6] organized neighbour count [7] whole run [8] triangles.  Returns 0/-1. */
void ls3d_frame_enable_timing(Ls3dFrame *f, int on);
int ls3d_frame_stage_ms(Ls3dFrame *f, float out[9]);

/* Same path, writing the merged cloud to a caller-provided device buffer at record offset read from
 * d_dst_offset (device int, may be NULL = 0): the multi-GPU merge writes peer-mapped memory through this. */
int ls3d_frame_run_to(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run,
	void *d_dst_vertices, const int *d_dst_offset, void *stream);

/* Multi-GPU merge: as ls3d_frame_run_to, but the final compaction stores every surviving record to n_peers
 * destination buffers (peer-mapped device pointers, one per rank; host array of n_peers <= 8 pointers) at the
 * same record offset — compaction and all-gather in one kernel over NVLink peer stores. */
int ls3d_frame_run_peers(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run,
	int n_peers, void *const *peer_dst_vertices, const int *d_dst_offset, void *stream);

/* The same in two stream-ordered stages, for ranks that must exchange their survivor counts in between:
 *   ls3d_frame_run_count   : everything up to and including the neighbour count; ls3d_frame_count_ptr()[3] then
 *                            holds this rank's survivor count (n_kept) although nothing has been compacted yet
 *   ls3d_frame_merge_peers : the final compaction, storing to every peer buffer at *d_dst_offset. */
int ls3d_frame_run_count(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run, void *stream);
int ls3d_frame_merge_peers(Ls3dFrame *f, int first_map, int n_run, int n_peers, void *const *peer_dst_vertices, const int *d_dst_offset, void *stream);

/* The exchange around that pair without any collective library: counts and completion flags travel as NVLink peer stores.
 * Every rank owns one Ls3dFrameSync block in device memory, zero-initialised once and mapped by all ranks (CUDA IPC);
 * peer_sync[r] is rank r's block as seen from this process.
 *   ls3d_frame_publish_counts : after ls3d_frame_run_count — stores this rank's survivor count into every rank's block, waits for
 *                               all ranks' counts of this frame and leaves this rank's exclusive prefix in its own block's
 *                               `offset` (pass &block->offset as ls3d_frame_merge_peers' d_dst_offset) and the sum in `total`.
 *                               f == NULL publishes 0 (a rank that owns no sensor).
 *   ls3d_frame_wait_peers     : after ls3d_frame_merge_peers — raises this rank's "delivered" word on every rank and waits for
 *                               everyone's: afterwards this rank's merged buffer holds the whole cloud.
 * Both are single-block kernels on `stream` (no host synchronisation; capturable in a CUDA graph). */
typedef struct Ls3dFrameSync {
	unsigned epoch; int offset; int total; int err;
	unsigned cnt_tag[8]; int cnt_val[8]; unsigned done_tag[8];
	unsigned pad[36];
} Ls3dFrameSync;                               /* 256 bytes */
int ls3d_frame_publish_counts(Ls3dFrame *f, int rank, int world, void *const *peer_sync, void *stream);
int ls3d_frame_wait_peers(int rank, int world, void *const *peer_sync, void *stream);

/* The two pre-passes on device-resident buffers, enqueued on `stream` without synchronisation: radial correction of a packed
 * frame in place (same layouts as ls3d_frame_run's inputs, so it chains straight into it), flying-pixel filter from d_in to a
 * distinct d_out.  Return the number of kernels enqueued or -1. */
int ls3d_radial_correction_device(int n_maps, void *d_depth_maps, void *d_depth_colors, const int *widths, const int *heights, const float *intr_params, void *stream);
int ls3d_filter_flying_pixels_device(const void *d_in, void *d_out, int width, int height, int neighbourhoodSize, float thr, void *stream);

/* Plain device allocations that can be shared between the ranks of one node (one process per GPU) through CUDA
 * IPC: export a 64-byte handle, open it in another process to get a peer-mapped pointer (NVLink loads/stores). */
void *ls3d_dev_alloc(unsigned long long bytes);
void ls3d_dev_free(void *p);
int ls3d_ipc_export(void *p, unsigned char handle[64]);
void *ls3d_ipc_open(const unsigned char handle[64]);
void ls3d_ipc_close(void *p);

/* ICP context sized for at most n1_max target and n2_max source points on the current device. */
Ls3dIcp *ls3d_icp_create(int n1_max, int n2_max);
void ls3d_icp_destroy(Ls3dIcp *c);

/* Build the target grid from device points (Point3f layout) — once per ICP call (the reference rebuilds its
 * kd-tree every iteration although verts1 never changes, icp.cpp:21-23). */
int ls3d_icp_set_target(Ls3dIcp *c, const void *d_verts1, int n1, void *stream);
/* Source cloud (device, Point3f layout, transformed in place).  [i_begin,i_end) is the slice whose nearest
 * neighbours THIS rank searches (0,n2 on one GPU); R/t accumulators start from R0 (9 floats), t0 (3). */
int ls3d_icp_set_source(Ls3dIcp *c, void *d_verts2, int n2, int i_begin, int i_end, const float *R0, const float *t0, void *stream);

/* One iteration = two stream-ordered stages:
 *   ls3d_icp_match  : apply the previous iteration's (T,Rk) to verts2, exact NN search of the local source slice (warp packets,
 *                     then a block-wide stage for the few packets that need 60+ node visits), one-to-one dedupe by 64-bit
 *                     atomicMin into the dedupe slots (ls3d_icp_slots: int64[n1]; with peers set, into the OWNER rank's slots
 *                     over NVLink)
 *   ls3d_icp_reduce : mean / sigma of the matched squared distances in the reference's two-pass form (icp.cpp:34-54), the
 *                     2.5 sigma gate, the 16 correspondence sums, the 3x3 Kabsch/SVD -> the (T, Rk) the next match applies, and
 *                     the R,t accumulation of icp.cpp:167-168 — ONE launch with in-kernel grid barriers; every slot is reset.
 *                     With peers set the ranks exchange their partial sums inside the kernel (peer stores + flags): no
 *                     collective library call anywhere in the loop.  This pair is what ls3d_icp_run and ICP() enqueue.
 * ls3d_icp_finish applies the last (T,Rk), leaving R,t in ls3d_icp_Rt (device f32[12]: R[9] then t[3]). */
int ls3d_icp_match(Ls3dIcp *c, void *stream);
int ls3d_icp_reduce(Ls3dIcp *c, void *stream);
/* Sharded ICP over `world` <= 8 GPUs of one node, one process each: replicated clouds, source slices searched per rank, dedupe
 * slots and reduction chunks owned per target range.  slots/part/flag[r] = rank r's ls3d_icp_slots / ls3d_icp_red_part /
 * ls3d_icp_red_flag as mapped into this process (CUDA IPC); entry `rank` must be the context's own.  world <= 1 switches it off. */
int ls3d_icp_set_peers(Ls3dIcp *c, int world, int rank, void *const *slots, void *const *part, void *const *flag);
double *ls3d_icp_red_part(Ls3dIcp *c);      /* device f64[3][296][16] */
unsigned *ls3d_icp_red_flag(Ls3dIcp *c);    /* device u32[64] */
int ls3d_icp_finish(Ls3dIcp *c, void *stream);
/* All maxIter iterations on one GPU, no host round trips (captured once per shape as a CUDA graph). */
int ls3d_icp_run(Ls3dIcp *c, int maxIter, void *stream);

/* Tuning aid: per-source-point work statistics of the last match stage (3 u32 each: octree child steps, candidate points
 * scanned, resume level + 1 with 0 = finished in the first kernel); pass NULL to switch off. */
void ls3d_icp_set_debug(Ls3dIcp *c, void *d_stats);

long long *ls3d_icp_slots(Ls3dIcp *c);      /* device int64[n1] */
float *ls3d_icp_Rt(Ls3dIcp *c);             /* device f32[12] */
const int *ls3d_icp_nn_index(Ls3dIcp *c);   /* device int[n2]: last NN index per source point (-1 outside the slice) */
const float *ls3d_icp_nn_dist(Ls3dIcp *c);  /* device f32[n2] */
Ls3dIcpTrace *ls3d_icp_trace_buf(Ls3dIcp *c); /* device trace[64]: entry i filled by iteration i (i < 64) */
const int *ls3d_icp_status(Ls3dIcp *c);     /* device int[4]: {iterations applied, error flags, 0, 0} */

/* ---- Part 4: the data formats either side of the path ------------------------------ */

/* The client's frame blob: SerializeFrame (src/LiveScanClient/liveScanClient.cpp:185-290) -> KinectSocket.ReceiveFrame
 * (LiveScanServer/KinectSocket.cs:211-304).  16-byte header {int payload_bytes, int compressed, int width, int height}, then the
 * payload — u16 depth[w*h], RGB u8[3*w*h], int n_bodies, per body {u8 tracked, int n_joints, n_joints x {int type, int state,
 * float x, y, z, float cx, cy}} — as is or as one zstd frame (compressed == 1). */
typedef struct Ls3dClientFrameInfo {
	int payload_bytes, compressed, width, height;   /* the header fields */
	int n_bodies;                                   /* filled by ls3d_client_frame_unpack (-1 after ls3d_client_frame_header) */
	long long raw_bytes;                            /* payload size after decompression (ditto) */
} Ls3dClientFrameInfo;
/* Header only.  0 / -1 (payload_bytes <= 0 is the client's "no more stored frames" marker, KinectSocket.cs:231-235). */
int ls3d_client_frame_header(const unsigned char *blob, long long blob_bytes, Ls3dClientFrameInfo *info);
/* Decompresses if needed and copies depth (2*w*h bytes) and colours (3*w*h) to where the caller wants them — typically sensor
 * i's slot of the packed arrays the path takes (KinectServer.CopyLatestFrames, KinectServer.cs:404-500) — and the body records
 * (count included) to bodies_out.  Any output may be NULL.  Returns the byte length of the body records or -1. */
int ls3d_client_frame_unpack(const unsigned char *blob, long long blob_bytes, unsigned char *depth_out, unsigned char *colors_out,
	unsigned char *bodies_out, long long bodies_cap, Ls3dClientFrameInfo *info);
/* The writer: depth + per-depth-pixel colours (+ body records, NULL = none) -> blob; compression_level > 0 compresses the payload
 * with zstd at that level (the client's default is 2, liveScanClient.cpp:60-61).  out == NULL returns a sufficient capacity.
 * Returns the blob length or -1. */
long long ls3d_client_frame_pack(const unsigned char *depth, const unsigned char *colors, int width, int height,
	const unsigned char *bodies, long long bodies_bytes, int compression_level, unsigned char *out, long long out_cap);

/* NativeUtils' frames dump: storeAllFramesInformation / loadAllFramesInformation (src/NativeUtils/depthprocessing.cpp:1316-1385), the
 * file generateMeshFromDepthMaps reads instead of its arguments when built with LOAD_FRAMES_INFORMATION (:16,:1726-1730): exactly the
 * argument list of the path's entry points.  Loaded images sit in page-locked memory.  0 / -1. */
typedef struct Ls3dFramesInfo {
	int n_maps;
	unsigned char *depth_maps, *depth_colors;
	int *widths, *heights;
	float *intr_params, *wtransform_params;
} Ls3dFramesInfo;
int ls3d_frames_info_store(const char *filename, int n_maps, const unsigned char *depth_maps, const unsigned char *depth_colors,
	const int *widths, const int *heights, const float *intr_params, const float *wtransform_params);
int ls3d_frames_info_load(const char *filename, Ls3dFramesInfo *out);
void ls3d_frames_info_free(Ls3dFramesInfo *info);

/* Binary PLY as Utils.saveToPly writes it (LiveScanServer/Utils.cs:222-293; n_triangles < 0: the vertex-only overload :173-220):
 * text header, 15 bytes per vertex (float x, y, z, uchar r, g, b), 13 per face (uchar 3, int a, b, c).  The first header line ends
 * in "\r\n" (StreamWriter.WriteLine on Windows), the others in "\n", as in the files the server writes.  The ASCII variant
 * (off by default, KinectSettings.cs:48) is .NET number formatting and is not provided.
 * ls3d_ply_binary_size: bytes of the whole file.  ls3d_write_ply_binary: host arrays -> file image in out (the body is packed on
 * the device); returns its length or -1.  ls3d_pack_ply_body_device: the body only (15*n + 13*nt bytes), device to device
 * (d_out 16-byte aligned), enqueued on `stream`. */
long long ls3d_ply_binary_size(int n_vertices, int n_triangles);
long long ls3d_write_ply_binary(const VertexC4ubV3f *vertices, int n_vertices, const int *triangles, int n_triangles, unsigned char *out, long long out_cap);
int ls3d_pack_ply_body_device(const void *d_vertices, int n_vertices, const int *d_triangles, int n_triangles, void *d_out, void *stream);
/* ASCII PLY, Utils.saveToPly with binary = false (Utils.cs:204-214, :276-289), byte for byte including its quirks: "\r\n" after a
 * header line that already ends in "\n", a blank before every vertex line's end, face lines "3 " + the three indices written with
 * nothing between them.  Floats as Single.ToString(InvariantCulture) of .NET Framework 4.5 (7 significant digits, general format).
 * Host text codec (no device work).  out_cap must be at least ls3d_ply_ascii_bound(); returns the file's length or -1.
 * n_triangles < 0: the vertex-only overload. */
long long ls3d_ply_ascii_bound(int n_vertices, int n_triangles);
long long ls3d_write_ply_ascii(const VertexC4ubV3f *vertices, int n_vertices, const int *triangles, int n_triangles, unsigned char *out, long long out_cap);

/* The mesh frame TransferServer streams to viewers (HoloLens / Unity): formVerticesChunks / formMeshChunks
 * (LiveScanServer/TransferServer.cs:179-271) split the mesh into chunks of at most 64 997 vertices — with triangles, a vertex is
 * re-emitted in every chunk that uses it and the index list becomes chunk-local — and TransferSocket.SendFrame
 * (LiveScanServer/TransferSocket.cs:50-105) sends int n_vertices, int n_triangles, int n_chunks, int chunk_vertices[n_chunks],
 * int chunk_triangles[n_chunks], float xyz[3*n_vertices], uchar rgb[3*n_vertices], int triangles[3*n_triangles].
 * ls3d_write_transfer_frame: host mesh -> that byte stream (out == NULL: returns the exact length); chunking and re-packing run
 * on the device.  ls3d_transfer_chunks_device: the same for a device-resident mesh (e.g. ls3d_frame_vertices / _triangles): chunk
 * sizes to the host arrays, the body (xyz | rgb | triangles) stays on the device at *d_body (library memory, valid until the
 * next call); returns the number of chunks.  ls3d_pack_transfer_body_device: only the re-packing. */
long long ls3d_transfer_frame_size(int n_vertices, int n_triangles, int n_chunks);
int ls3d_set_transfer_chunk_limit(int limit);   /* test hook: the chunk size limit (default 65000 - 3, TransferServer.cs:181,205) */
long long ls3d_write_transfer_frame(const VertexC4ubV3f *vertices, int n_vertices, const int *triangles, int n_triangles, unsigned char *out, long long out_cap);
int ls3d_transfer_chunks_device(const void *d_vertices, int n_vertices, const int *d_triangles, int n_triangles,
	int *chunk_vertices, int *chunk_triangles, int chunk_cap, int *n_vertices_out, const void **d_body, void *stream);
int ls3d_pack_transfer_body_device(const void *d_vertices, int n_vertices, const int *d_triangles, int n_triangles, void *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LS3D_H */
