// ls3d_oracle — CPU restatement of the LiveScan3D per-frame point-cloud hot path.
//
// THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it, and only as the checker (or the timed CPU baseline) — the product
// (livescan3d_b200/csrc, libls3d_b200.so) never links, imports or falls back to anything in oracle/.
//
// Every function restates, in its own words, the reference function cited next to it (paths relative to
// /root/reference).  It is pinned, in tests/test_oracle_vs_ref.py, against the reference's own sources
// compiled in place into oracle/_ref/ (see oracle/Makefile) and against the golden vectors under
// tests/golden/ that were generated from that build.  The reference ships no golden vectors of its own
// for this path, and the OpenCV 3.2 arithmetic inside ICP (icp.cpp:138-168) is not in
// the reference tree at all: for that boundary parity is UNPINNED by fixtures and rests on the stated
// tolerances (R 1e-5, t 1e-4 m).
//
// Third-party algorithms restated here:
//   * nanoflann 1.1.9 (vendored by the reference as include/nanoflann.h:71) — kd-tree build with leaf size
//     10 and the "middle split" rule, exact k-NN search with its fp32 bound bookkeeping and first-visited
//     tie rule.
//   * OpenCV 3.2.0 core (include/opencv2/core/version.hpp:53-55; binaries absent) — reduce(AVG), gemm and
//     3x3 SVD semantics for CV_32F, as far as icp.cpp uses them.
//
// Build: `make -C oracle oracle`  (g++ -O2 -fopenmp -ffp-contract=off; no FMA contraction anywhere, so the
// fp32 evaluation order below is the evaluation order of the reference built the same way).

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

struct P3 { float X, Y, Z; };                       // icp.h:15-18 / utils.h:44-61 (12 B, no padding)
struct Vtx { unsigned char R, G, B, A; float X, Y, Z; };   // VertexC4ubV3f, depthprocessing.h:29-33 (16 B)

// ---------------------------------------------------------------------------------------------------
// nanoflann restatement
// ---------------------------------------------------------------------------------------------------
struct Interval { float low, high; };
struct BBox { Interval d[3]; };

struct KdNode {
	int child1, child2;      // -1,-1 => leaf
	int left, right;         // leaf: range in vind
	int divfeat;             // inner: split axis
	float divlow, divhigh;   // inner: max of left subtree / min of right subtree on divfeat
};

struct KdTree {
	const P3 *pts = nullptr;
	size_t n = 0;
	std::vector<size_t> vind;
	std::vector<KdNode> nodes;
	int root = -1;
	BBox root_bbox;
	static const size_t leaf_max = 10;   // KDTreeSingleIndexAdaptorParams default, nanoflann.h:711

	inline float get(size_t idx, int dim) const { return dim == 0 ? pts[idx].X : (dim == 1 ? pts[idx].Y : pts[idx].Z); }

	// PointCloud::kdtree_distance, icp.h:40-47 / filter.h:38-45 — THE parity-defining metric.
	inline float dist(const float *q, size_t idx) const {
		const float d0 = q[0] - pts[idx].X;
		const float d1 = q[1] - pts[idx].Y;
		const float d2 = q[2] - pts[idx].Z;
		return d0 * d0 + d1 * d1 + d2 * d2;
	}

	// buildIndex, nanoflann.h:859-867 (+ computeBoundingBox :1000-1022)
	void build(const P3 *p, size_t count) {
		pts = p; n = count;
		vind.resize(n);
		for (size_t i = 0; i < n; i++) vind[i] = i;
		nodes.clear(); root = -1;
		if (n == 0) return;
		for (int i = 0; i < 3; i++) root_bbox.d[i].low = root_bbox.d[i].high = get(0, i);
		for (size_t k = 1; k < n; k++)
			for (int i = 0; i < 3; i++) {
				if (get(k, i) < root_bbox.d[i].low) root_bbox.d[i].low = get(k, i);
				if (get(k, i) > root_bbox.d[i].high) root_bbox.d[i].high = get(k, i);
			}
		nodes.reserve(n / 4 + 16);
		root = divide(0, n, root_bbox);
	}

	// computeMinMax, nanoflann.h:1084-1094
	void min_max(const size_t *ind, size_t count, int element, float &mn, float &mx) const {
		mn = mx = get(ind[0], element);
		for (size_t i = 1; i < count; i++) {
			float v = get(ind[i], element);
			if (v < mn) mn = v;
			if (v > mx) mx = v;
		}
	}

	// planeSplit, nanoflann.h:1146-1174
	void plane_split(size_t *ind, size_t count, int cutfeat, float cutval, size_t &lim1, size_t &lim2) const {
		size_t left = 0, right = count - 1;
		for (;;) {
			while (left <= right && get(ind[left], cutfeat) < cutval) ++left;
			while (right && left <= right && get(ind[right], cutfeat) >= cutval) --right;
			if (left > right || !right) break;
			std::swap(ind[left], ind[right]);
			++left; --right;
		}
		lim1 = left;
		right = count - 1;
		for (;;) {
			while (left <= right && get(ind[left], cutfeat) <= cutval) ++left;
			while (right && left <= right && get(ind[right], cutfeat) > cutval) --right;
			if (left > right || !right) break;
			std::swap(ind[left], ind[right]);
			++left; --right;
		}
		lim2 = left;
	}

	// middleSplit_, nanoflann.h:1096-1135.  Note the 1.1.9 quirk kept on purpose: inside the axis loop the
	// spread is measured on the CURRENT cutfeat, not on axis i (:1111).
	void middle_split(size_t *ind, size_t count, size_t &index, int &cutfeat, float &cutval, const BBox &bbox) const {
		const float EPS = 0.00001f;
		float max_span = bbox.d[0].high - bbox.d[0].low;
		for (int i = 1; i < 3; i++) {
			float span = bbox.d[i].high - bbox.d[i].low;
			if (span > max_span) max_span = span;
		}
		float max_spread = -1;
		cutfeat = 0;
		for (int i = 0; i < 3; i++) {
			float span = bbox.d[i].high - bbox.d[i].low;
			if (span > (1 - EPS) * max_span) {
				float mn, mx;
				min_max(ind, count, cutfeat, mn, mx);
				float spread = mx - mn;
				if (spread > max_spread) { cutfeat = i; max_spread = spread; }
			}
		}
		float split_val = (bbox.d[cutfeat].low + bbox.d[cutfeat].high) / 2;
		float mn, mx;
		min_max(ind, count, cutfeat, mn, mx);
		if (split_val < mn) cutval = mn;
		else if (split_val > mx) cutval = mx;
		else cutval = split_val;
		size_t lim1, lim2;
		plane_split(ind, count, cutfeat, cutval, lim1, lim2);
		if (lim1 > count / 2) index = lim1;
		else if (lim2 < count / 2) index = lim2;
		else index = count / 2;
	}

	// divideTree, nanoflann.h:1034-1082
	int divide(size_t left, size_t right, BBox &bbox) {
		int id = (int)nodes.size();
		nodes.push_back(KdNode());
		if ((right - left) <= leaf_max) {
			KdNode nd; nd.child1 = nd.child2 = -1; nd.left = (int)left; nd.right = (int)right; nd.divfeat = 0; nd.divlow = nd.divhigh = 0;
			for (int i = 0; i < 3; i++) bbox.d[i].low = bbox.d[i].high = get(vind[left], i);
			for (size_t k = left + 1; k < right; k++)
				for (int i = 0; i < 3; i++) {
					if (bbox.d[i].low > get(vind[k], i)) bbox.d[i].low = get(vind[k], i);
					if (bbox.d[i].high < get(vind[k], i)) bbox.d[i].high = get(vind[k], i);
				}
			nodes[id] = nd;
		} else {
			size_t idx; int cutfeat; float cutval;
			middle_split(&vind[0] + left, right - left, idx, cutfeat, cutval, bbox);
			BBox lb = bbox; lb.d[cutfeat].high = cutval;
			int c1 = divide(left, left + idx, lb);
			BBox rb = bbox; rb.d[cutfeat].low = cutval;
			int c2 = divide(left + idx, right, rb);
			KdNode nd; nd.child1 = c1; nd.child2 = c2; nd.left = nd.right = 0; nd.divfeat = cutfeat;
			nd.divlow = lb.d[cutfeat].high; nd.divhigh = rb.d[cutfeat].low;
			for (int i = 0; i < 3; i++) {
				bbox.d[i].low = std::min(lb.d[i].low, rb.d[i].low);
				bbox.d[i].high = std::max(lb.d[i].high, rb.d[i].high);
			}
			nodes[id] = nd;
		}
		return id;
	}
};

// KNNResultSet, nanoflann.h:76-134 (ties: the earlier-visited point stays in front)
struct KnnSet {
	size_t *indices; float *dists; size_t capacity, count;
	void init(size_t *i, float *d, size_t cap) { indices = i; dists = d; capacity = cap; count = 0; dists[capacity - 1] = FLT_MAX; }
	inline void add(float dist, size_t index) {
		size_t i;
		for (i = count; i > 0; --i) {
			if (dists[i - 1] > dist) {
				if (i < capacity) { dists[i] = dists[i - 1]; indices[i] = indices[i - 1]; }
			} else break;
		}
		if (i < capacity) { dists[i] = dist; indices[i] = index; }
		if (count < capacity) count++;
	}
	inline float worst() const { return dists[capacity - 1]; }
};

// searchLevel, nanoflann.h:1200-1247 (epsError == 1)
static void search_level(const KdTree &t, KnnSet &rs, const float *vec, int node, float mindistsq, float dists[3]) {
	const KdNode &nd = t.nodes[node];
	if (nd.child1 < 0 && nd.child2 < 0) {
		float worst_dist = rs.worst();
		for (int i = nd.left; i < nd.right; ++i) {
			const size_t index = t.vind[i];
			float d = t.dist(vec, index);
			if (d < worst_dist) rs.add(d, index);
		}
		return;
	}
	int idx = nd.divfeat;
	float val = vec[idx];
	float diff1 = val - nd.divlow;
	float diff2 = val - nd.divhigh;
	int best, other; float cut_dist;
	if ((diff1 + diff2) < 0) { best = nd.child1; other = nd.child2; cut_dist = (val - nd.divhigh) * (val - nd.divhigh); }
	else { best = nd.child2; other = nd.child1; cut_dist = (val - nd.divlow) * (val - nd.divlow); }
	search_level(t, rs, vec, best, mindistsq, dists);
	float dst = dists[idx];
	mindistsq = mindistsq + cut_dist - dst;
	dists[idx] = cut_dist;
	if (mindistsq * 1.0f <= rs.worst()) search_level(t, rs, vec, other, mindistsq, dists);
	dists[idx] = dst;
}

// findNeighbors + computeInitialDistances, nanoflann.h:899-911, 1176-1194
static void find_neighbors(const KdTree &t, KnnSet &rs, const float *vec) {
	float dists[3] = {0, 0, 0};
	float distsq = 0;
	for (int i = 0; i < 3; i++) {
		if (vec[i] < t.root_bbox.d[i].low) { dists[i] = (vec[i] - t.root_bbox.d[i].low) * (vec[i] - t.root_bbox.d[i].low); distsq += dists[i]; }
		if (vec[i] > t.root_bbox.d[i].high) { dists[i] = (vec[i] - t.root_bbox.d[i].high) * (vec[i] - t.root_bbox.d[i].high); distsq += dists[i]; }
	}
	search_level(t, rs, vec, t.root, distsq, dists);
}

// ---------------------------------------------------------------------------------------------------
// OpenCV-3.2 semantics used by ICP (see header comment) — same restatement as oracle/ref_shim/opencv\cv.h
// ---------------------------------------------------------------------------------------------------
static void svd3(const float M[9], float U[9], float Vt[9]) {
	const int n = 3;
	float At[3][3], V[3][3]; double W[3];
	for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { At[i][j] = M[j * 3 + i]; V[i][j] = (i == j) ? 1.0f : 0.0f; }
	for (int i = 0; i < n; i++) { double sd = 0; for (int k = 0; k < n; k++) sd += (double)At[i][k] * At[i][k]; W[i] = sd; }
	const float eps = 1.1920929e-07f * 2;
	for (int iter = 0; iter < 30; iter++) {
		bool changed = false;
		for (int i = 0; i < n - 1; i++) for (int j = i + 1; j < n; j++) {
			float *Ai = At[i], *Aj = At[j];
			double a = W[i], p = 0, b = W[j];
			for (int k = 0; k < n; k++) p += (double)Ai[k] * Aj[k];
			if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
			p *= 2;
			double beta = a - b, gamma = hypot(p, beta);
			float c, s;
			if (beta < 0) { double delta = (gamma - beta) * 0.5; s = (float)std::sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
			else { c = (float)std::sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
			a = b = 0;
			for (int k = 0; k < n; k++) { float t0 = c * Ai[k] + s * Aj[k]; float t1 = -s * Ai[k] + c * Aj[k]; Ai[k] = t0; Aj[k] = t1; a += (double)t0 * t0; b += (double)t1 * t1; }
			W[i] = a; W[j] = b;
			changed = true;
			float *Vi = V[i], *Vj = V[j];
			for (int k = 0; k < n; k++) { float t0 = c * Vi[k] + s * Vj[k]; float t1 = -s * Vi[k] + c * Vj[k]; Vi[k] = t0; Vj[k] = t1; }
		}
		if (!changed) break;
	}
	for (int i = 0; i < n; i++) { double sd = 0; for (int k = 0; k < n; k++) sd += (double)At[i][k] * At[i][k]; W[i] = std::sqrt(sd); }
	for (int i = 0; i < n - 1; i++) {
		int j = i; for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
		if (i != j) { std::swap(W[i], W[j]); for (int k = 0; k < n; k++) { std::swap(At[i][k], At[j][k]); std::swap(V[i][k], V[j][k]); } }
	}
	for (int i = 0; i < n; i++) {
		float inv = W[i] > 0 ? (float)(1.0 / W[i]) : 0.0f;
		for (int k = 0; k < n; k++) { U[k * 3 + i] = At[i][k] * inv; Vt[i * 3 + k] = V[i][k]; }
	}
}

static inline void mul33_f32(const float A[9], const float B[9], float D[9]) {   // gemm small-matrix path
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			D[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}

static inline float det33_f32(const float *m) {
	return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------
// Vertex generation
// ---------------------------------------------------------------------------------------------------

// createVertices, src/NativeUtils/depthprocessing.cpp:122-187 (+ RotatePoint :109-120).
// Outputs: xyz[3*n], rgb[3*n], depth_to_vertices[w*h] (-1 = none), vertices_to_depth[n].  Returns n.
int orc_create_vertices(const unsigned short *depth_map, const unsigned char *depth_colors, int w, int h,
	const float *intr7, const float *wt12, const float *bounds6,
	float *xyz, unsigned char *rgb, int *depth_to_vertices, int *vertices_to_depth)
{
	const float cx = intr7[0], cy = intr7[1], fx = intr7[2], fy = intr7[3];
	const float *t = wt12, *R = wt12 + 3;   // WorldTranformation(float*): t[3] then R row-major, depthprocessing.h:56-63
	const float minX = bounds6[0], minY = bounds6[1], minZ = bounds6[2], maxX = bounds6[3], maxY = bounds6[4], maxZ = bounds6[5];
	int n = 0;
	for (int y = 0; y < h; y++) {
		const unsigned short *row = depth_map + (size_t)y * w;
		for (int x = 0; x < w; x++) {
			int pos = x + y * w;
			if (depth_to_vertices) depth_to_vertices[pos] = -1;
			if (row[x] == 0) continue;
			float val = row[x];
			float Z = val / 1000.0f;
			float X = (x - cx) / fx;
			float Y = (cy - y) / fy;
			X = X * Z;
			Y = Y * Z;
			X += t[0]; Y += t[1]; Z += t[2];
			float rx = X * R[0] + Y * R[1] + Z * R[2];
			float ry = X * R[3] + Y * R[4] + Z * R[5];
			float rz = X * R[6] + Y * R[7] + Z * R[8];
			if (rx < minX || rx > maxX || ry < minY || ry > maxY || rz < minZ || rz > maxZ) continue;
			if (depth_to_vertices) depth_to_vertices[pos] = n;
			if (vertices_to_depth) vertices_to_depth[n] = pos;
			xyz[3 * n] = rx; xyz[3 * n + 1] = ry; xyz[3 * n + 2] = rz;
			rgb[3 * n] = depth_colors[pos * 3]; rgb[3 * n + 1] = depth_colors[pos * 3 + 1]; rgb[3 * n + 2] = depth_colors[pos * 3 + 2];
			n++;
		}
	}
	return n;
}

// generateVerticesFromDepthMaps + formMesh (depthprocessing.cpp:708-733, 1578-1629) for the vertex part of
// generateMeshFromDepthMaps / generateVerticesFromDepthMap (:1631-1657, :1715-1792): sensors are processed
// independently and concatenated in sensor order into 16-byte VertexC4ubV3f records with A = 255.
// map_index = -1 => all sensors; otherwise only that sensor (byte offsets as at :1646-1650).
// out_vertices must hold sum(w*h) records.  per_map_counts may be NULL.  Returns the total vertex count.
int orc_generate_mesh(int n_maps, const unsigned char *depth_maps, const unsigned char *depth_colors,
	const int *widths, const int *heights, const float *intr_params, const float *wtransform_params,
	const float *bounds6, int map_index, void *out_vertices, int *per_map_counts)
{
	Vtx *out = (Vtx*)out_vertices;
	size_t depth_pos = 0, colors_pos = 0;
	int total = 0;
	for (int i = 0; i < n_maps; i++) {
		size_t npx = (size_t)widths[i] * heights[i];
		if (map_index == -1 || map_index == i) {
			std::vector<float> xyz(3 * npx);
			std::vector<unsigned char> rgb(3 * npx);
			int n = orc_create_vertices((const unsigned short*)(depth_maps + depth_pos), depth_colors + colors_pos, widths[i], heights[i],
				intr_params + 7 * i, wtransform_params + 12 * i, bounds6, xyz.data(), rgb.data(), nullptr, nullptr);
			for (int j = 0; j < n; j++) {
				Vtx &v = out[total + j];
				v.R = rgb[3 * j]; v.G = rgb[3 * j + 1]; v.B = rgb[3 * j + 2]; v.A = 255;
				v.X = xyz[3 * j]; v.Y = xyz[3 * j + 1]; v.Z = xyz[3 * j + 2];
			}
			if (per_map_counts) per_map_counts[i] = n;
			total += n;
		} else if (per_map_counts) per_map_counts[i] = 0;
		depth_pos += npx * 2;
		colors_pos += npx * 3;
	}
	return total;
}

// ---------------------------------------------------------------------------------------------------
// Triangle generation
// ---------------------------------------------------------------------------------------------------

// MeshGenerator::checkTriangleConstraints, src/NativeUtils/meshGenerator.cpp:14-62, on pixel indices a, b, c of one
// depth image.  All three depths must be non-zero; every edge (a->b, b->c, c->a) must either be flat
// (|dv| < thr) or continue the depth gradient of the pixel one step beyond either end of the edge.  The threshold
// grows linearly with the mean depth and is evaluated in double precision, then truncated (:26).
static bool orc_triangle_ok(const unsigned short *d, long a, long b, long c)
{
	const long at[3] = {a, b, c};
	const int v[3] = {d[a], d[b], d[c]};
	if (v[0] == 0 || v[1] == 0 || v[2] == 0) return false;
	const int thr = (int)((v[0] + v[1] + v[2]) / 3.0 * 0.00272 + 7.273);
	for (int e = 0; e < 3; e++) {
		const int i1 = e, i2 = (e + 1) % 3;
		const int v1 = v[i1], v2 = v[i2];
		if (std::abs(v1 - v2) < thr) continue;
		const long step = at[i2] - at[i1];
		const int fwd = d[at[i2] + step];
		if (fwd != 0 && std::abs(v2 - v1 - (fwd - v2)) < thr) continue;
		const int bwd = d[at[i1] - step];
		if (bwd != 0 && std::abs(v2 - v1 - (v1 - bwd)) < thr) continue;
		return false;
	}
	return true;
}

// MeshGenerator::generateTrianglesGradients (+Region), meshGenerator.cpp:76-181: the four row bands the reference
// hands to its threads tile [0,h) and are concatenated in order, so the result is one raster scan over
// y in [2, h-2), x in [1, w-2) (:86-89).  For a pixel that owns a vertex, the quad {p, p-w, p-w+1, p+1} is split along
// one diagonal (triangles 0,1) or — only when neither of those passes — the other (2,3) (:113-121); a triangle is
// emitted when all three of its corners own vertices, with the corner order of triangles_shifts (:100-103).
// tri receives 3 ints per triangle (per-sensor vertex indices + vertex_base); returns the triangle count.
int orc_triangles_one(const unsigned short *depth, const int *depth_to_vertices, int w, int h, int vertex_base, int *tri)
{
	int n = 0;
	const long W = w;
	const long corner[4][3] = {{1, -W, 0}, {1, -W + 1, -W}, {0, -W + 1, -W}, {0, 1, -W + 1}};
	for (int y = 2; y < h - 2; y++)
		for (int x = 1; x < w - 2; x++) {
			const long p = (long)y * w + x;
			if (depth_to_vertices[p] == -1) continue;
			bool ok[4] = {false, false, false, false};
			ok[0] = orc_triangle_ok(depth, p, p - W, p + 1);
			ok[1] = orc_triangle_ok(depth, p + 1, p - W, p - W + 1);
			if (!ok[0] && !ok[1]) {
				ok[2] = orc_triangle_ok(depth, p, p - W, p - W + 1);
				ok[3] = orc_triangle_ok(depth, p, p - W + 1, p + 1);
			}
			for (int t = 0; t < 4; t++) {
				if (!ok[t]) continue;
				const int m0 = depth_to_vertices[p + corner[t][0]], m1 = depth_to_vertices[p + corner[t][1]], m2 = depth_to_vertices[p + corner[t][2]];
				if (m0 == -1 || m1 == -1 || m2 == -1) continue;
				tri[3 * n] = m0 + vertex_base; tri[3 * n + 1] = m1 + vertex_base; tri[3 * n + 2] = m2 + vertex_base;
				n++;
			}
		}
	return n;
}

// The bcolor_transfer = bgenerate_triangles = false branch of generateMeshFromDepthMaps (depthprocessing.cpp:1715-1792):
// createVertices per sensor, generateTriangles (always executed, :1786) on the raw depth copy and the pixel->vertex map,
// formMesh (:1578-1629) concatenating vertices and triangles in sensor order with the indices rebased.
// out_triangles must hold 6*sum(w*h) ints.  Returns the vertex count; *n_triangles receives the triangle count.
int orc_generate_mesh_triangles(int n_maps, const unsigned char *depth_maps, const unsigned char *depth_colors,
	const int *widths, const int *heights, const float *intr_params, const float *wtransform_params,
	const float *bounds6, void *out_vertices, int *out_triangles, int *n_triangles, int *per_map_counts, int *per_map_triangles)
{
	Vtx *out = (Vtx*)out_vertices;
	size_t depth_pos = 0, colors_pos = 0;
	int total = 0, ntri = 0;
	for (int i = 0; i < n_maps; i++) {
		const size_t npx = (size_t)widths[i] * heights[i];
		std::vector<float> xyz(3 * npx);
		std::vector<unsigned char> rgb(3 * npx);
		std::vector<int> d2v(npx);
		const unsigned short *dm = (const unsigned short*)(depth_maps + depth_pos);
		int n = orc_create_vertices(dm, depth_colors + colors_pos, widths[i], heights[i], intr_params + 7 * i, wtransform_params + 12 * i, bounds6,
			xyz.data(), rgb.data(), d2v.data(), nullptr);
		for (int j = 0; j < n; j++) {
			Vtx &v = out[total + j];
			v.R = rgb[3 * j]; v.G = rgb[3 * j + 1]; v.B = rgb[3 * j + 2]; v.A = 255;
			v.X = xyz[3 * j]; v.Y = xyz[3 * j + 1]; v.Z = xyz[3 * j + 2];
		}
		const int nt = orc_triangles_one(dm, d2v.data(), widths[i], heights[i], total, out_triangles + 3 * (size_t)ntri);
		if (per_map_counts) per_map_counts[i] = n;
		if (per_map_triangles) per_map_triangles[i] = nt;
		total += n;
		ntri += nt;
		depth_pos += npx * 2;
		colors_pos += npx * 3;
	}
	*n_triangles = ntri;
	return total;
}

// ---------------------------------------------------------------------------------------------------
// Radial-distortion correction of the depth/colour maps
// ---------------------------------------------------------------------------------------------------

// depthMapAndColorRadialCorrection, src/NativeUtils/depthprocessing.cpp:191-261, one sensor, in place.
//  1. forward warp (:201-220): every non-zero pixel moves to (int)(u*d*fx + cx), (int)(v*d*fy + cy) with
//     u = (x-cx)/fx, v = (y-cy)/fy (note: y-cy, not cy-y), r = u*u + v*v, d = 1 - r2*r - r4*r*r - r6*r*r*r, all fp32,
//     C truncation; sources are visited in raster order and simply overwrite, so the LAST source in raster order wins.
//  2. hole closing (:226-257), raster order over the interior, IN PLACE: a zero pixel whose 8 neighbours (order
//     NW,N,NE,W,E,SW,S,SE) contain more than 4 "consistent" non-zero depths — each within 30 of the previously accepted
//     one — becomes their integer mean (colours likewise).  Because it is in place, neighbours NW,N,NE,W may already
//     hold values filled earlier in the same pass.
void orc_radial_correction_one(unsigned short *depth, unsigned char *colors, int w, int h, const float *intr7)
{
	const float cx = intr7[0], cy = intr7[1], fx = intr7[2], fy = intr7[3], r2 = intr7[4], r4 = intr7[5], r6 = intr7[6];
	std::vector<unsigned short> out((size_t)w * h, 0);
	std::vector<unsigned char> outc((size_t)w * h * 3, 0);
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			const size_t src = (size_t)x + (size_t)y * w;
			if (depth[src] == 0) continue;
			const float u = (x - cx) / fx;
			const float v = (y - cy) / fy;
			const float r = u * u + v * v;
			const float d = 1 - r2 * r - r4 * r * r - r6 * r * r * r;
			const float fxc = u * d * fx + cx, fyc = v * d * fy + cy;
			// (int) of an out-of-range or NaN float is the x86 "integer indefinite" value INT_MIN, which fails the >= 0 test below
			const int xc = (fxc > -2147483904.0f && fxc < 2147483648.0f) ? (int)fxc : INT32_MIN;
			const int yc = (fyc > -2147483904.0f && fyc < 2147483648.0f) ? (int)fyc : INT32_MIN;
			if (xc >= 0 && yc >= 0 && xc < w && yc < h) {
				const size_t dst = (size_t)xc + (size_t)yc * w;
				out[dst] = depth[src];
				memcpy(&outc[dst * 3], colors + src * 3, 3);
			}
		}
	const int nb[8] = {-w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1};
	for (int y = 1; y < h - 1; y++)
		for (int x = 1; x < w - 1; x++) {
			const long pos = (long)x + (long)y * w;
			if (out[pos] != 0) continue;
			int n = 0, sum = 0, sr = 0, sg = 0, sb = 0, prev = -1;
			for (int i = 0; i < 8; i++) {
				const int val = out[pos + nb[i]];
				if (val > 0 && (prev == -1 || std::abs(val - prev) < 30)) {
					prev = val; n++; sum += val;
					sr += outc[(pos + nb[i]) * 3]; sg += outc[(pos + nb[i]) * 3 + 1]; sb += outc[(pos + nb[i]) * 3 + 2];
				}
			}
			if (n > 4) {
				out[pos] = (unsigned short)(sum / n);
				outc[pos * 3] = (unsigned char)(sr / n); outc[pos * 3 + 1] = (unsigned char)(sg / n); outc[pos * 3 + 2] = (unsigned char)(sb / n);
			}
		}
	memcpy(depth, out.data(), (size_t)w * h * 2);
	memcpy(colors, outc.data(), (size_t)w * h * 3);
}

// depthMapAndColorSetRadialCorrection, depthprocessing.cpp:1794-1815: every sensor of the packed frame, independently.
void orc_radial_correction(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, const int *widths, const int *heights, const float *intr_params)
{
	size_t dpos = 0, cpos = 0;
	for (int i = 0; i < n_maps; i++) {
		const size_t npx = (size_t)widths[i] * heights[i];
		orc_radial_correction_one((unsigned short*)(depth_maps + dpos), depth_colors + cpos, widths[i], heights[i], intr_params + 7 * i);
		dpos += npx * 2;
		cpos += npx * 3;
	}
}

// KinectCapture::filterFlyingPixels(int neighbourhoodSize, float thr, int maxNonFittingNeighbours),
// src/LiveScanClient/kinectCapture.cpp:132-174.  PARITY UNPINNED: kinectCapture.cpp needs the Kinect SDK
// headers and cannot be compiled here, so this restatement is checked against the source text only.
// A pixel of the interior [k, w-k) x [k, h-k) is zeroed when more than nNeighbours/2 of its (2k+1)^2-1 neighbours differ
// from it by more than thr (int difference compared with the float threshold, :163); the caller's
// maxNonFittingNeighbours is overwritten (:150); removals are collected first and applied afterwards (:169-172).
void orc_filter_flying_pixels(unsigned short *depth, int w, int h, int k, float thr, int maxNonFittingNeighbours)
{
	const int n_nb = (2 * k + 1) * (2 * k + 1) - 1;
	std::vector<int> shifts;
	for (int a = -k; a <= k; a++)
		for (int b = -k; b <= k; b++)
			if (a != 0 || b != 0) shifts.push_back(a * w + b);                  // :141-147 (x * width + y)
	maxNonFittingNeighbours = n_nb / 2;
	std::vector<int> remove;
	for (int y = k; y < h - k; y++)
		for (int x = k; x < w - k; x++) {
			const int pos = y * w + x, val = depth[pos];
			int n_diff = 0;
			for (int sft : shifts) {
				const int diff = std::abs((int)depth[pos + sft] - val);
				if (diff > thr) n_diff++;
			}
			if (n_diff > maxNonFittingNeighbours) remove.push_back(pos);
		}
	for (int pos : remove) depth[pos] = 0;
}

// ---------------------------------------------------------------------------------------------------
// Nearest neighbours
// ---------------------------------------------------------------------------------------------------

// FindClosestPointForEach, src/NativeUtils/icp.cpp:18-32: kd-tree on verts1, 1-NN of every verts2 point.
void orc_find_closest(const float *verts1, int n1, const float *verts2, int n2, unsigned long long *indices, float *dists)
{
	KdTree tree;
	tree.build((const P3*)verts1, (size_t)n1);
#pragma omp parallel for
	for (int i = 0; i < n2; i++) {
		size_t idx = 0; float d = 0;
		KnnSet rs; rs.init(&idx, &d, 1);
		find_neighbors(tree, rs, verts2 + 3 * (size_t)i);
		indices[i] = idx; dists[i] = d;
	}
}

// Brute-force 1-NN with the same fp32 metric (independent check of the tree; O(n1*n2), small inputs only).
// Ties resolve to the lowest index.
void orc_find_closest_brute(const float *verts1, int n1, const float *verts2, int n2, unsigned long long *indices, float *dists)
{
	const P3 *p = (const P3*)verts1;
#pragma omp parallel for
	for (int i = 0; i < n2; i++) {
		const float *q = verts2 + 3 * (size_t)i;
		float best = FLT_MAX; size_t bi = 0;
		for (int j = 0; j < n1; j++) {
			const float d0 = q[0] - p[j].X, d1 = q[1] - p[j].Y, d2 = q[2] - p[j].Z;
			const float d = d0 * d0 + d1 * d1 + d2 * d2;
			if (d < best) { best = d; bi = j; }
		}
		indices[i] = bi; dists[i] = best;
	}
}

// KNNeighbors, src/LiveScanClient/filter.cpp:19-34: squared distance from each point to its k-th nearest
// neighbour in its own cloud, self included; FLT_MAX when the cloud has fewer than k points
// (KNNResultSet::init, nanoflann.h:93).
void orc_knn_kdist(const float *verts, int n, int k, float *kdist)
{
	KdTree tree;
	tree.build((const P3*)verts, (size_t)n);
#pragma omp parallel for
	for (int i = 0; i < n; i++) {
		std::vector<size_t> idx(k); std::vector<float> d(k);
		KnnSet rs; rs.init(idx.data(), d.data(), (size_t)k);
		find_neighbors(tree, rs, verts + 3 * (size_t)i);
		kdist[i] = d[k - 1];
	}
}

// Brute-force k-th neighbour distance (independent check; O(n^2)).
void orc_knn_kdist_brute(const float *verts, int n, int k, float *kdist)
{
	const P3 *p = (const P3*)verts;
#pragma omp parallel for
	for (int i = 0; i < n; i++) {
		std::vector<float> d(n);
		const float *q = verts + 3 * (size_t)i;
		for (int j = 0; j < n; j++) {
			const float d0 = q[0] - p[j].X, d1 = q[1] - p[j].Y, d2 = q[2] - p[j].Z;
			d[j] = d0 * d0 + d1 * d1 + d2 * d2;
		}
		if (n < k) { kdist[i] = FLT_MAX; continue; }
		std::nth_element(d.begin(), d.begin() + (k - 1), d.end());
		kdist[i] = d[k - 1];
	}
}

// ---------------------------------------------------------------------------------------------------
// Filter
// ---------------------------------------------------------------------------------------------------

// filter, src/LiveScanClient/filter.cpp:36-81.  verts: n*3 f32, colors: n*4 u8 (RGB struct, utils.h:105-111),
// both compacted in place (stable); old_to_new[i] = new index or -1.  Returns the surviving count.
// k<=0 || maxDist<=0: nothing is touched, old_to_new is filled with -2 ("no entry", the reference returns a
// map holding only the {-1:-1} sentinel) and n is returned.
int orc_filter(float *verts, unsigned char *colors, int n, int k, float maxDist, int *old_to_new)
{
	if (k <= 0 || maxDist <= 0) {
		for (int i = 0; i < n; i++) old_to_new[i] = -2;
		return n;
	}
	std::vector<float> kd(n);
	if (n > 0) orc_knn_kdist(verts, n, k, kd.data());
	const float distThreshold = (float)std::pow((double)maxDist, 2.0);   // filter.cpp:52: float = pow(float, int)
	int last = 0;
	for (int i = 0; i < n; i++) {
		if (kd[i] > distThreshold) { old_to_new[i] = -1; continue; }
		memmove(verts + 3 * (size_t)last, verts + 3 * (size_t)i, 12);
		memmove(colors + 4 * (size_t)last, colors + 4 * (size_t)i, 4);
		old_to_new[i] = last;
		last++;
	}
	return last;
}

// ---------------------------------------------------------------------------------------------------
// ICP
// ---------------------------------------------------------------------------------------------------

// Per-iteration trace of the quantities the GPU path is compared on at stage level (all optional).
struct OrcIcpTrace {
	int n_matched;        // one-to-one matches before rejection (icp.cpp:95-126)
	int n_accepted;       // after RejectOutlierMatches (icp.cpp:56-73)
	float sigma;          // GetStandardDeviation of squared distances (icp.cpp:34-54)
	float T[3];           // tempT (icp.cpp:141)
	float Rk[9];          // tempR (icp.cpp:153-163)
};

// ICP, src/NativeUtils/icp.cpp:75-177.  verts2, R (row-major 3x3) and t are updated in place; returns 1.0f.
// trace (maxIter entries) may be NULL.
float orc_icp_trace(const float *verts1, float *verts2, int nVerts1, int nVerts2, float *R, float *t, int maxIter, OrcIcpTrace *trace)
{
	const P3 *v1 = (const P3*)verts1;
	std::vector<P3> v2((P3*)verts2, (P3*)verts2 + nVerts2);
	float matR[9]; memcpy(matR, R, sizeof(matR));
	float error = 1;
	KdTree tree;

	for (int iter = 0; iter < maxIter; iter++) {
		std::vector<P3> matched1, matched2;
		std::vector<float> distances(nVerts2);
		std::vector<size_t> indices(nVerts2);
		// FindClosestPointForEach (icp.cpp:18-32): the tree is rebuilt every iteration
		tree.build(v1, (size_t)nVerts1);
#pragma omp parallel for
		for (int i = 0; i < nVerts2; i++) {
			KnnSet rs; rs.init(&indices[i], &distances[i], 1);
			find_neighbors(tree, rs, &v2[i].X);
		}
		// one match per target point, smaller d2 wins, later source index wins ties (icp.cpp:95-126)
		std::vector<float> matchDistances;
		std::vector<int> matchIdxs(nVerts1, -1);
		for (int i = 0; i < nVerts2; i++) {
			int pos = matchIdxs[indices[i]];
			if (pos != -1 && matchDistances[pos] < distances[i]) continue;
			if (pos == -1) {
				matched1.push_back(v1[indices[i]]);
				matched2.push_back(v2[i]);
				matchDistances.push_back(distances[i]);
				matchIdxs[indices[i]] = (int)matched1.size() - 1;
			} else {
				matched2[pos] = v2[i];
				matchDistances[pos] = distances[i];
			}
		}
		const int n_matched = (int)matched1.size();
		// GetStandardDeviation (icp.cpp:34-54): fp32 running sums, squares taken in double
		float mean = 0;
		for (size_t i = 0; i < matchDistances.size(); i++) mean += matchDistances[i];
		mean /= matchDistances.size();
		float sd = 0;
		for (size_t i = 0; i < matchDistances.size(); i++) sd += pow(matchDistances[i] - mean, 2);
		sd /= matchDistances.size();
		sd = sqrt(sd);
		// RejectOutlierMatches (icp.cpp:56-73), maxStdDev = 2.5
		{
			std::vector<P3> f1, f2;
			for (size_t i = 0; i < matched1.size(); i++) {
				if (matchDistances[i] > 2.5f * sd) continue;
				f1.push_back(matched1[i]); f2.push_back(matched2[i]);
			}
			matched1 = f1; matched2 = f2;
		}
		const int m = (int)matched1.size();
		// tempT = reduce(matched1 - matched2, AVG) (icp.cpp:141): fp32 differences, sequential fp32 column sums
		float tempT[3] = {0, 0, 0};
		if (m > 0) {
			float buf[3] = { matched1[0].X - matched2[0].X, matched1[0].Y - matched2[0].Y, matched1[0].Z - matched2[0].Z };
			for (int i = 1; i < m; i++) {
				buf[0] = buf[0] + (matched1[i].X - matched2[i].X);
				buf[1] = buf[1] + (matched1[i].Y - matched2[i].Y);
				buf[2] = buf[2] + (matched1[i].Z - matched2[i].Z);
			}
			const float scale = (float)(1.0 / m);
			for (int c = 0; c < 3; c++) tempT[c] = buf[c] * scale;
		}
		// verts2 += tempT ; matched2 += tempT (icp.cpp:143-150)
		for (int i = 0; i < nVerts2; i++) { v2[i].X = v2[i].X + tempT[0]; v2[i].Y = v2[i].Y + tempT[1]; v2[i].Z = v2[i].Z + tempT[2]; }
		for (int i = 0; i < m; i++) { matched2[i].X = matched2[i].X + tempT[0]; matched2[i].Y = matched2[i].Y + tempT[1]; matched2[i].Z = matched2[i].Z + tempT[2]; }
		// M = matched2^T * matched1 (icp.cpp:152): gemm general path, fp64 accumulators
		double acc[9] = {0};
		for (int i = 0; i < m; i++) {
			const double q[3] = { matched2[i].X, matched2[i].Y, matched2[i].Z };
			const double p[3] = { matched1[i].X, matched1[i].Y, matched1[i].Z };
			for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) acc[a * 3 + b] += q[a] * p[b];
		}
		float M[9]; for (int a = 0; a < 9; a++) M[a] = (float)acc[a];
		// tempR = U * Vt, reflection fix (icp.cpp:153-163)
		float U[9], Vt[9], tempR[9];
		svd3(M, U, Vt);
		mul33_f32(U, Vt, tempR);
		if ((double)det33_f32(tempR) < 0) {
			float D[9] = {1, 0, 0, 0, 1, 0, 0, 0, -1}, UD[9];
			mul33_f32(U, D, UD);
			mul33_f32(UD, Vt, tempR);
		}
		// verts2 = verts2 * tempR (icp.cpp:165): gemm small-matrix path, fp32 left to right
		for (int i = 0; i < nVerts2; i++) {
			const float a0 = v2[i].X, a1 = v2[i].Y, a2 = v2[i].Z;
			v2[i].X = a0 * tempR[0] + a1 * tempR[3] + a2 * tempR[6];
			v2[i].Y = a0 * tempR[1] + a1 * tempR[4] + a2 * tempR[7];
			v2[i].Z = a0 * tempR[2] + a1 * tempR[5] + a2 * tempR[8];
		}
		// matT += tempT * matR^T (icp.cpp:167): general path (fp64 accumulate), then fp32 +=
		for (int j = 0; j < 3; j++) {
			double s = 0;
			for (int k2 = 0; k2 < 3; k2++) s += (double)tempT[k2] * (double)matR[j * 3 + k2];
			t[j] = t[j] + (float)s;
		}
		// matR = matR * tempR (icp.cpp:168): small-matrix path
		float newR[9];
		mul33_f32(matR, tempR, newR);
		memcpy(matR, newR, sizeof(matR));

		if (trace) {
			trace[iter].n_matched = n_matched; trace[iter].n_accepted = m; trace[iter].sigma = sd;
			memcpy(trace[iter].T, tempT, sizeof(tempT)); memcpy(trace[iter].Rk, tempR, sizeof(tempR));
		}
	}
	memcpy(verts2, v2.data(), (size_t)nVerts2 * sizeof(P3));
	memcpy(R, matR, sizeof(matR));
	return error;
}

float orc_icp(const float *verts1, float *verts2, int nVerts1, int nVerts2, float *R, float *t, int maxIter)
{
	return orc_icp_trace(verts1, verts2, nVerts1, nVerts2, R, t, maxIter, nullptr);
}

// One-to-one matching of icp.cpp:95-126 on precomputed NN results, for stage-level GPU parity tests.
// winner[j] = source index matched to target j, or -1.
void orc_dedupe(const unsigned long long *indices, const float *dists, int n2, int n1, int *winner)
{
	std::vector<float> best(n1, 0.0f);
	for (int j = 0; j < n1; j++) winner[j] = -1;
	for (int i = 0; i < n2; i++) {
		size_t j = indices[i];
		if (winner[j] != -1 && best[j] < dists[i]) continue;
		winner[j] = i; best[j] = dists[i];
	}
}

int orc_sizeof_trace(void) { return (int)sizeof(OrcIcpTrace); }

}  // extern "C"
