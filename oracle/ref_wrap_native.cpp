// Flat C entry points over the reference's C++-only internals (test infrastructure only).
// Compiled against the reference's own headers and linked with its own objects into
// oracle/_ref/libls3d_ref_native.so; see oracle/Makefile.  Nothing here restates an algorithm: every
// function forwards to the reference symbol named in its comment.
#include "depthprocessing.h"
#include "meshGenerator.h"
#include "opencv\cv.h"

// reference internals (src/NativeUtils/depthprocessing.cpp:708, :1578, :1659; icp.cpp:18)
void generateVerticesFromDepthMaps(unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	vector<WorldTranformation> &world_transforms, vector<IntrinsicCameraParameters> &intrinsic_params,
	vector<VerticesWithDepthColorMaps> &vertices_with_maps,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int map_index);
void formMesh(Mesh *out_mesh, vector<VerticesWithDepthColorMaps> &vertices_with_maps, vector<vector<TriangleIndexes>> &triangle_indexes);
int generateTriangles(vector<VerticesWithDepthColorMaps> &vertices_with_maps, int *heights, int *widths,
	vector<vector<TriangleIndexes>> &triangle_indexes);
void storeAllFramesInformation(string filename, int n_maps, unsigned char* depth_maps,
	unsigned char *depth_colors, int *widths, int *heights, float *intr_params, float *wtransform_params);
void loadAllFramesInformation(string filename, int &n_maps, unsigned char** depth_maps,
	unsigned char **depth_colors, int **widths, int **heights, float **intr_params, float **wtransform_params);
void FindClosestPointForEach(PointCloud &sourceCloud, cv::Mat &destPoints, vector<float> &distances, vector<size_t> &indices);

extern "C" {

// The bcolor_transfer=false / bgenerate_triangles=false branch of generateMeshFromDepthMaps
// (depthprocessing.cpp:1715-1792) without its LOAD_FRAMES_INFORMATION file override (:16,:1726-1730):
// generateVerticesFromDepthMaps (thread per sensor) -> [generateTriangles] -> formMesh.
// per_map_counts (may be NULL) receives each sensor's vertex count.
void ref_generate_mesh(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int with_triangles, int *per_map_counts)
{
	vector<VerticesWithDepthColorMaps> vertices_with_maps(n_maps);
	vector<IntrinsicCameraParameters> intrinsic_params(n_maps);
	vector<WorldTranformation> world_transforms(n_maps);
	vector<vector<TriangleIndexes>> triangle_indexes(n_maps);
	for (int i = 0; i < n_maps; i++) {
		intrinsic_params[i] = IntrinsicCameraParameters(intr_params + i * 7);
		world_transforms[i] = WorldTranformation(wtransform_params + i * 12);
	}
	generateVerticesFromDepthMaps(depth_maps, depth_colors, widths, heights, world_transforms, intrinsic_params,
		vertices_with_maps, minX, minY, minZ, maxX, maxY, maxZ, -1);
	if (with_triangles)
		generateTriangles(vertices_with_maps, heights, widths, triangle_indexes);
	if (per_map_counts)
		for (int i = 0; i < n_maps; i++) per_map_counts[i] = (int)vertices_with_maps[i].vertices.size();
	formMesh(out_mesh, vertices_with_maps, triangle_indexes);
}

// createVertices' side outputs for one sensor (depthprocessing.cpp:122-187): the two pixel<->vertex maps.
int ref_vertex_maps(unsigned char *depth_map, unsigned char *depth_colors, int w, int h, float *intr7, float *wt12,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int *depth_to_vertices, int *vertices_to_depth)
{
	vector<VerticesWithDepthColorMaps> v(1);
	vector<IntrinsicCameraParameters> ip(1, IntrinsicCameraParameters(intr7));
	vector<WorldTranformation> wt(1, WorldTranformation(wt12));
	generateVerticesFromDepthMaps(depth_map, depth_colors, &w, &h, wt, ip, v, minX, minY, minZ, maxX, maxY, maxZ, -1);
	memcpy(depth_to_vertices, v[0].depth_to_vertices_map.data(), sizeof(int) * (size_t)w * h);
	memcpy(vertices_to_depth, v[0].vertices_to_depth_map.data(), sizeof(int) * v[0].vertices_to_depth_map.size());
	return (int)v[0].vertices.size();
}

// FindClosestPointForEach (icp.cpp:18-32): kd-tree on verts1, 1-NN of every verts2 row.
void ref_find_closest(float *verts1, int n1, float *verts2, int n2, unsigned long long *indices, float *dists)
{
	PointCloud cloud1;
	cloud1.pts = vector<Point3f>((Point3f*)verts1, (Point3f*)verts1 + n1);
	cv::Mat verts2Mat(n2, 3, CV_32F, verts2);
	vector<float> d(n2);
	vector<size_t> idx(n2);
	FindClosestPointForEach(cloud1, verts2Mat, d, idx);
	for (int i = 0; i < n2; i++) { indices[i] = idx[i]; dists[i] = d[i]; }
}

// storeAllFramesInformation / loadAllFramesInformation (depthprocessing.cpp:1316-1385): the frames dump, written and read by
// the reference's own code.  ref_load_frames copies what the reference allocated into the caller's arrays (sized by the caller
// from the file it wrote itself) and returns n_maps.
void ref_store_frames(const char *filename, int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params)
{
	storeAllFramesInformation(filename, n_maps, depth_maps, depth_colors, widths, heights, intr_params, wtransform_params);
}

int ref_load_frames(const char *filename, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params)
{
	int n = 0;
	unsigned char *d = nullptr, *c = nullptr;
	int *w = nullptr, *h = nullptr;
	float *ip = nullptr, *wt = nullptr;
	loadAllFramesInformation(filename, n, &d, &c, &w, &h, &ip, &wt);
	size_t pd = 0, pc = 0;
	for (int i = 0; i < n; i++) { pd += (size_t)w[i] * h[i] * 2; pc += (size_t)w[i] * h[i] * 3; }
	memcpy(depth_maps, d, pd); memcpy(depth_colors, c, pc);
	memcpy(widths, w, sizeof(int) * n); memcpy(heights, h, sizeof(int) * n);
	memcpy(intr_params, ip, sizeof(float) * 7 * n); memcpy(wtransform_params, wt, sizeof(float) * 12 * n);
	delete[] d; delete[] c; delete[] w; delete[] h; delete[] ip; delete[] wt;
	return n;
}

}  // extern "C"
